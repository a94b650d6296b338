// Batched symmetric eigendecomposition (two-sided parallel Jacobi, eig.cu).
#pragma once
#include "common.cuh"

#define HYP_SYEVJ_SMEM_LIMIT (227 * 1024 - 256)
#define HYP_SYEVJ_MAX_SIDE 512

// doubles of global scratch per matrix needed when the problem does not fit in shared memory
int64_t hyp_syevj_gwork_doubles(int max_side);
// Eigen-decomposition of nmat symmetric matrices: matrix c is the side[c] x side[c] block at
// Ain + in_off[c] (leading dimension side rounded up to even, both triangles filled), optionally
// divided by divv[div_off[c] + div_idx].  Eigenvalues ascending -> lam + lam_off[c]; eigenvectors
// (columns, same layout as Ain) -> Vout + in_off[c] when Vout != nullptr.  Ain is not modified.
void hyp_syevj_batched(hyp_ctx* ctx, int nmat, int max_side, const int* d_sides, const int64_t* d_in_off,
                       const double* Ain, double* Vout, const int64_t* d_lam_off, double* lam,
                       const double* divv, const int64_t* d_div_off, int div_idx, double* gwork);

// Matrix-domain cones (PosSemidefTri, HypoPerLogdetTri, HypoRootdetTri): batched per-cone state
// and congruence products.  Implemented in cones_mat.cu.
#pragma once
#include "common.cuh"

void hyp_mat_alloc_group(hyp_ctx* ctx, ConeGroup& g);
void hyp_mat_update_state(hyp_ctx* ctx, ConeGroup& g);
void hyp_mat_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols,
                  int64_t ld_prod, int64_t ld_arr, int mode, int64_t row_shift);
void hyp_mat_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir);

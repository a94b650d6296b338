// Matrix-domain cones (PosSemidefTri, HypoPerLogdetTri, HypoRootdetTri): batched per-cone state
// and congruence products.  Implemented in cones_mat.cu.
#pragma once
#include "common.cuh"

void hyp_mat_alloc_group(hyp_ctx* ctx, ConeGroup& g);
void hyp_mat_update_state(hyp_ctx* ctx, ConeGroup& g);
void hyp_mat_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols,
                  int64_t ld_prod, int64_t ld_arr, int mode, int64_t row_shift);
void hyp_mat_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir);

// shared with cones_spec.cu
void hyp_mat_ensure_work(hyp_ctx* ctx, int64_t doubles);   // grows ctx->d_matwork
// Y_j = X' M_j X in place for the cc matrices (d x d, ld lde, stride lde*lde) of Mall; C1: (lde*cc) x d scratch
void hyp_mat_congruence(hyp_ctx* ctx, const double* X, int d, int lde, double* Mall, int64_t cc, double* C1,
                        int64_t ldc1);

// EpiPerSepSpectral{MatrixCSqr} (cones_spec.cu)
void hyp_spec_alloc_group(hyp_ctx* ctx, ConeGroup& g);
void hyp_spec_update_state(hyp_ctx* ctx, ConeGroup& g);
void hyp_spec_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols,
                   int64_t ld_prod, int64_t ld_arr, int mode, int64_t row_shift);
void hyp_spec_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir);

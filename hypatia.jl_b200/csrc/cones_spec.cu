// EpiPerSepSpectral{MatrixCSqr} on the device: batched per-cone state (Cholesky gate, Jacobi
// eigendecomposition, spectral scalars) and Hessian-type products as congruences around an
// elementwise kernel.  Kernels: eig_kernels.cuh, cones_spec_kernels.cuh, cones_mat_kernels.cuh.
//
// reference: src/Cones/epipersepspectral/matrixcsqr.jl:91-564 (update_feas, is_dual_feas, update_grad,
// update_hess_aux, hess_prod!, update_inv_hess_aux, inv_hess_prod!, update_dder3_aux, dder3),
// sepspectralfun.jl:17-116, src/linearalgebra/dense.jl:69 (update_eigen!).
// ConeGroup fields reused: d_W = smat(w) (then smat(dual w)), d_U = V (eigenvectors of W / v),
// d_Ut = V', d_Wi = theta, d_Ui = Dh, d_Uit = scratch of the Cholesky gate.
// The reference rotates column by column with four small GEMMs (matrixcsqr.jl:292-318); here a
// whole chunk of columns is rotated by two TMA + DMMA GEMM launches on each side of the middle kernel.
#include "cones_mat.cuh"
#include "cones_mat_kernels.cuh"
#include "cones_spec_kernels.cuh"
#include "eig.cuh"
#include <cstdlib>

using hypdev::pack_cols_kernel;
using hypdev::spec_dder3_kernel;
using hypdev::spec_dualfeas_kernel;
using hypdev::spec_mid_kernel;
using hypdev::spec_post_kernel;
using hypdev::unpack_cols_kernel;
using hypdev::unpack_state_kernel;

namespace {

template <typename T>
T* upload_vec(const std::vector<T>& v) {
    T* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace

void hyp_spec_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    hyp_mat_alloc_group(ctx, g);   // six d x d matrices per cone, even leading dimension
    g.h_voff.assign(g.count, 0);
    std::vector<int64_t> voff7(g.count, 0);
    int64_t tot = 0;
    for (int i = 0; i < g.count; i++) {
        g.h_voff[i] = tot;
        voff7[i] = tot + 7 * (int64_t)g.h_side[i];
        tot += 8 * (int64_t)g.h_side[i];
        g.h_hkind.push_back(ctx->h_cone_hkind[g.h_kidx[i]]);
        g.h_hparam.push_back(ctx->h_cone_hparam[g.h_kidx[i]]);
    }
    g.d_voff = upload_vec(g.h_voff);
    g.d_voff7 = upload_vec(voff7);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_hparam = upload_vec(g.h_hparam);
    CUDA_TRY(cudaMalloc(&g.d_vecs, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    CUDA_TRY(cudaMemset(g.d_vecs, 0, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
}

void hyp_spec_update_state(hyp_ctx* ctx, ConeGroup& g) {
    const int64_t maxlen = (int64_t)g.max_side * (g.max_side + 1) / 2;
    dim3 ugrid(g.count, (unsigned)std::max<int64_t>(1, std::min<int64_t>((maxlen + 255) / 256, 64)));
    // work buffer: [global scratch of the eigensolver (large sides only) | inverse written by the dual gate]
    const bool need_gwork = hyp_syevj_gwork_doubles(g.max_side) * 8 > HYP_SYEVJ_SMEM_LIMIT;
    const int64_t gw_total = need_gwork ? hyp_syevj_gwork_doubles(g.max_side) * g.count : 0;
    hyp_mat_ensure_work(ctx, gw_total + g.mat_total + 16);
    double* gwork = need_gwork ? ctx->d_matwork : nullptr;
    double* inv_scratch = ctx->d_matwork + gw_total;
    // ---- primal: Cholesky gate on a copy, eigendecomposition of W / v (matrixcsqr.jl:91-115) ----
    unpack_state_kernel<<<ugrid, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_moff, 2, ctx->d_point,
                                                        g.d_W, g.d_Uit);
    ctx->launches++;
    // sides above 128 skip the gate (hyp_chol_batched ignores them): their feasibility rests on the eigenvalues
    hyp_chol_batched(ctx, g.count, g.d_side, g.d_moff, g.d_kidx, g.d_Uit, g.d_Ui, ctx->d_feas);
    hyp_syevj_batched(ctx, g.count, g.max_side, g.d_side, g.d_moff, g.d_W, g.d_U, g.d_voff, g.d_vecs, ctx->d_point,
                      g.d_off, 1, gwork);
    spec_post_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_moff, g.d_voff, g.d_kidx,
                                                      g.d_hkind, g.d_hparam, ctx->d_point, g.d_U, g.d_Ut, g.d_Wi,
                                                      g.d_Ui, g.d_vecs, g.d_scal, ctx->d_grad, ctx->d_feas);
    ctx->launches++;
    // ---- dual feasibility (matrixcsqr.jl:119-138): Cholesky flag + eigenvalues of dual W / dual u ----
    CUDA_TRY(cudaMemsetAsync(ctx->d_tmpflag, 1, ctx->K, ctx->stream));
    unpack_state_kernel<<<ugrid, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_moff, 2, ctx->d_dual,
                                                        g.d_W, g.d_Uit);
    ctx->launches++;
    hyp_chol_batched(ctx, g.count, g.d_side, g.d_moff, g.d_kidx, g.d_Uit, inv_scratch, ctx->d_tmpflag);
    hyp_syevj_batched(ctx, g.count, g.max_side, g.d_side, g.d_moff, g.d_W, nullptr, g.d_voff7, g.d_vecs, ctx->d_dual,
                      g.d_off, 0, gwork);
    spec_dualfeas_kernel<<<g.count, 128, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_voff7, g.d_kidx, g.d_hkind,
                                                          g.d_hparam, ctx->d_dual, g.d_vecs, ctx->d_tmpflag,
                                                          ctx->d_dual_feas);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

void hyp_spec_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                   int64_t ld_arr, int mode, int64_t row_shift) {
    if (mode == HYP_PROD_SQRT_HESS || mode == HYP_PROD_INV_SQRT_HESS)
        throw HypError{"sqrt_hess_prod is not defined for EpiPerSepSpectral"};
    // few columns, sides that fit in shared memory: one fused launch for the whole group
    {
        static int max_small_cols = -1;
        if (max_small_cols < 0) {
            const char* e = getenv("HYP_MAT_SMALL_MAXCOLS");
            max_small_cols = e ? atoi(e) : 8;
        }
        const int64_t smem = (int64_t)2 * g.max_side * (g.max_side | 1) * sizeof(double);
        if (ncols <= max_small_cols && g.count > 0 && smem <= 226 * 1024) {
            static bool attr = false;
            if (!attr) {
                CUDA_TRY(cudaFuncSetAttribute(hypdev::spec_small_prod_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                attr = true;
            }
            dim3 grid(g.count, (unsigned)ncols);
            hypdev::spec_small_prod_kernel<<<grid, 256, smem, ctx->stream>>>(
                mode, g.count, g.d_off, g.d_side, g.d_moff, g.d_voff, g.d_dual, g.d_U, g.d_Ut, g.d_Wi, g.d_Ui,
                g.d_vecs, g.d_scal, arr, ld_arr, prod, ld_prod, row_shift);
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            return;
        }
    }
    const int64_t budget = (int64_t)48 << 20;   // doubles per workspace matrix (384 MB)
    for (int i = 0; i < g.count; i++) {
        const int d = g.h_side[i], lde = (d + 1) & ~1;
        const int64_t len = (int64_t)d * (d + 1) / 2;
        int m = mode;
        if (m == HYP_PROD_BLOCK) m = g.h_dual[i] ? HYP_PROD_INV_HESS : HYP_PROD_HESS;
        if (m == HYP_PROD_BLOCK_INV) m = g.h_dual[i] ? HYP_PROD_HESS : HYP_PROD_INV_HESS;
        const int inverse = (m == HYP_PROD_INV_HESS) ? 1 : 0;
        const int64_t per_col = (int64_t)lde * lde;
        const int64_t cmax = std::max<int64_t>(1, std::min<int64_t>(ncols, budget / per_col));
        const int64_t ldc1 = (int64_t)lde * cmax;
        hyp_mat_ensure_work(ctx, per_col * cmax + ldc1 * d + 16);
        double* Mall = ctx->d_matwork;
        double* C1 = Mall + per_col * cmax;
        const int64_t row0 = g.h_off[i] - row_shift;
        const double* V = g.d_U + g.h_moff[i];
        const double* Vt = g.d_Ut + g.h_moff[i];
        for (int64_t j0 = 0; j0 < ncols; j0 += cmax) {
            const int64_t cc = std::min(cmax, ncols - j0);
            const double* a0 = arr + row0 + j0 * ld_arr;
            double* p0 = prod + row0 + j0 * ld_prod;
            dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((len + 255) / 256, cc > 64 ? 8 : 64)),
                      (unsigned)std::min<int64_t>(cc, 65535));
            unpack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, a0 + 2, ld_arr, cc, Mall);
            hyp_mat_congruence(ctx, V, d, lde, Mall, cc, C1, ldc1);                     // R_j = V' M_j V
            spec_mid_kernel<<<(unsigned)std::min<int64_t>(cc, 4 * (int64_t)ctx->sm_count), 256, 0, ctx->stream>>>(
                inverse, d, lde, g.d_scal + 8 * i, g.d_vecs + g.h_voff[i], g.d_Wi + g.h_moff[i],
                g.d_Ui + g.h_moff[i], Mall, a0, ld_arr, p0, ld_prod, cc);
            hyp_mat_congruence(ctx, Vt, d, lde, Mall, cc, C1, ldc1);                    // V (.) V'
            pack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, Mall, cc, nullptr, nullptr, nullptr,
                                                            p0 + 2, ld_prod);
            ctx->launches += 3;
        }
    }
    CUDA_TRY(cudaGetLastError());
}

void hyp_spec_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir) {
    for (int i = 0; i < g.count; i++) {
        const int d = g.h_side[i], lde = (d + 1) & ~1;
        const int64_t len = (int64_t)d * (d + 1) / 2;
        const int64_t per = (int64_t)lde * lde;
        hyp_mat_ensure_work(ctx, 3 * per + (int64_t)lde * d + 16);
        double* R = ctx->d_matwork;
        double* X = R + per;
        double* OUT = X + per;
        double* C1 = OUT + per;
        const int64_t o = g.h_off[i];
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((len + 255) / 256, 64)), 1);
        unpack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, dir + o + 2, ctx->q, 1, R);
        hyp_mat_congruence(ctx, g.d_U + g.h_moff[i], d, lde, R, 1, C1, lde);
        spec_dder3_kernel<<<1, 256, 0, ctx->stream>>>(d, lde, g.d_scal + 8 * i, g.d_vecs + g.h_voff[i],
                                                     g.d_Ui + g.h_moff[i], R, X, OUT, dir + o, out + o);
        hyp_mat_congruence(ctx, g.d_Ut + g.h_moff[i], d, lde, OUT, 1, C1, lde);
        pack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, OUT, 1, nullptr, nullptr, nullptr, out + o + 2,
                                                        ctx->q);
        ctx->launches += 3;
    }
    CUDA_TRY(cudaGetLastError());
}

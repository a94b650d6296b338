// Host side of the batched symmetric eigensolver (kernel: eig_kernels.cuh).
#include "common.cuh"
#include "eig.cuh"
#include "eig_kernels.cuh"

using hypdev::syevj_batched_kernel;
using hypdev::syevj_work_doubles;

int64_t hyp_syevj_gwork_doubles(int max_side) {
    // the global-memory variant always reserves room for V (the rotation table sits behind it)
    return syevj_work_doubles(max_side, true);
}
void hyp_syevj_batched(hyp_ctx* ctx, int nmat, int max_side, const int* d_sides, const int64_t* d_in_off,
                       const double* Ain, double* Vout, const int64_t* d_lam_off, double* lam,
                       const double* divv, const int64_t* d_div_off, int div_idx, double* gwork) {
    if (nmat <= 0) return;
    const bool wantv = Vout != nullptr;
    const int64_t smem = syevj_work_doubles(max_side, wantv) * 8;
    if (smem <= HYP_SYEVJ_SMEM_LIMIT) {
        if (wantv) {
            auto k = syevj_batched_kernel<true, true>;
            CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<nmat, HYP_SYEVJ_THREADS, smem, ctx->stream>>>(nmat, d_sides, d_in_off, Ain, Vout, d_lam_off, lam, divv, d_div_off,
                                               div_idx, nullptr, 0);
        } else {
            auto k = syevj_batched_kernel<true, false>;
            CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<nmat, HYP_SYEVJ_THREADS, smem, ctx->stream>>>(nmat, d_sides, d_in_off, Ain, nullptr, d_lam_off, lam, divv,
                                               d_div_off, div_idx, nullptr, 0);
        }
    } else {
        if (max_side > HYP_SYEVJ_MAX_SIDE) throw HypError{"batched Jacobi eigensolver: matrix side above 512 is not supported"};
        if (!gwork) throw HypError{"batched Jacobi eigensolver: global workspace missing"};
        const int64_t stride = hyp_syevj_gwork_doubles(max_side);
        if (wantv)
            syevj_batched_kernel<false, true><<<nmat, HYP_SYEVJ_THREADS, 0, ctx->stream>>>(
                nmat, d_sides, d_in_off, Ain, Vout, d_lam_off, lam, divv, d_div_off, div_idx, gwork, stride);
        else
            syevj_batched_kernel<false, false><<<nmat, HYP_SYEVJ_THREADS, 0, ctx->stream>>>(
                nmat, d_sides, d_in_off, Ain, nullptr, d_lam_off, lam, divv, d_div_off, div_idx, gwork, stride);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

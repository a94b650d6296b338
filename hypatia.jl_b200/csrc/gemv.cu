// Tall-skinny GEMV / GEMV' over the row-sharded G panel (K6 in SURVEY.md section 2.3).
// reference call sites: mul!(.., G', z) qrchol.jl:52; mul!(Gx, G, x) qrchol.jl:73;
// mul!(sol.s, G, sol.x, -1, true) common.jl:144; residual GEMVs common.jl:91,94.
//
// Both kernels are HBM-bound (8*rows*ncols algorithmic bytes per call): coalesced 16-byte
// loads along the contiguous (row) direction of the column-major panel, several independent
// loads in flight per thread, and deterministic fixed-order reductions (no atomics).
#include "common.cuh"
#include <cstdlib>
#include "gemv_kernels.cuh"

void hyp_gemv_t(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld,
                const double* x, double alpha, double beta, double* y) {
    if (ncols <= 0) return;
    if (rows <= 0) {
        hyp_axpby(ctx, ncols, 0.0, y, beta, y);
        return;
    }
    hyp_time_begin(ctx, T_GEMV);
    bool vec = (ld % 2 == 0) && ((uintptr_t)M % 16 == 0) && ((uintptr_t)x % 16 == 0);
    if (rows >= 4096) {
        int grid = (int)std::min<int64_t>(ncols, (int64_t)ctx->sm_count * 16);
        if (vec)
            hypdev::gemv_t_cta_kernel<true><<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x, alpha, beta, y);
        else
            hypdev::gemv_t_cta_kernel<false><<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x, alpha, beta, y);
    } else {
        int grid = (int)std::min<int64_t>((ncols + 7) / 8, (int64_t)ctx->sm_count * 8);
        hypdev::gemv_t_warp_kernel<<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x, alpha, beta, y);
    }
    ctx->launches++;
    hyp_time_end(ctx, T_GEMV);
}

void hyp_gemv_n(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld,
                const double* x, double alpha, double beta, double* y) {
    if (rows <= 0) return;
    if (ncols <= 0) {
        hyp_axpby(ctx, rows, 0.0, y, beta, y);
        return;
    }
    hyp_time_begin(ctx, T_GEMV);
    bool vec = (ld % 2 == 0) && ((uintptr_t)M % 16 == 0);
    int row_blocks = ceil_div(rows, 256);
    int64_t target = (int64_t)ctx->sm_count * 8;
    int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(target / row_blocks, 64));
    nchunks = (int)std::min<int64_t>(nchunks, (ncols + 31) / 32);
    while ((int64_t)nchunks * rows > ctx->partial_doubles && nchunks > 1) nchunks--;
    int64_t cpc = (ncols + nchunks - 1) / nchunks;
    nchunks = ceil_div(ncols, cpc);
    dim3 grid(row_blocks, nchunks);
    if (vec)
        hypdev::gemv_n_kernel<true><<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, x, cpc, ctx->d_partial);
    else
        hypdev::gemv_n_kernel<false><<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, x, cpc, ctx->d_partial);
    int rgrid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)ctx->sm_count * 8);
    hypdev::gemv_n_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial, alpha, beta, y);
    ctx->launches += 2;
    hyp_time_end(ctx, T_GEMV);
}

// w[0:rows] = alphaN * M x + betaN * w  AND  y[0:ncols] = alphaT * M' z + betaT * y  in ONE pass over M
// (gemv_kernels.cuh).  Falls back to the two separate passes when M is not 16-byte aligned with an even leading
// dimension.  The T partials ((rows / 256) * 4 x ncols doubles) live in ctx->d_partial2.
void hyp_gemv_nt(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x,
                 const double* z, double alphaN, double betaN, double* w, double alphaT, double betaT, double* y) {
    const bool vec = (ld % 2 == 0) && ((uintptr_t)M % 16 == 0);
    static int fused = -1;
    if (fused < 0) fused = getenv("HYP_NO_FUSED_GEMV") ? 0 : 1;
    if (!vec || !fused || rows < 4096 || ncols < 64) {
        hyp_gemv_n(ctx, rows, ncols, M, ld, x, alphaN, betaN, w);
        hyp_gemv_t(ctx, rows, ncols, M, ld, z, alphaT, betaT, y);
        return;
    }
    hyp_time_begin(ctx, T_GEMV);
    const int row_blocks = ceil_div(rows, 256);
    const int64_t target = (int64_t)ctx->sm_count * 8;
    int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(target / row_blocks, 64));
    nchunks = (int)std::min<int64_t>(nchunks, (ncols + 63) / 64);
    while ((int64_t)nchunks * rows > ctx->partial_doubles && nchunks > 1) nchunks--;
    int64_t cpc = round_up((ncols + nchunks - 1) / nchunks, 8);
    nchunks = ceil_div(ncols, cpc);
    const int64_t needT = (int64_t)row_blocks * 4 * ncols;
    if (needT > ctx->partial2_doubles) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_partial2) cudaFree(ctx->d_partial2);
        ctx->d_partial2 = nullptr;
        CUDA_TRY(cudaMalloc(&ctx->d_partial2, (size_t)needT * sizeof(double)));
        ctx->partial2_doubles = needT;
    }
    dim3 grid(row_blocks, nchunks);
    hypdev::gemv_nt_kernel<<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, x, z, cpc, ctx->d_partial,
                                                         ctx->d_partial2);
    const int rgrid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)ctx->sm_count * 8);
    hypdev::gemv_n_reduce2_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial, alphaN, betaN, w);
    const int tgrid = (int)std::min<int64_t>(ceil_div(ncols, 256), (int64_t)ctx->sm_count * 8);
    hypdev::gemv_t_reduce_kernel<<<tgrid, 256, 0, ctx->stream>>>(ncols, row_blocks * 4, ctx->d_partial2, alphaT,
                                                                betaT, y);
    ctx->launches += 3;
    hyp_time_end(ctx, T_GEMV);
    CUDA_TRY(cudaGetLastError());
}

// ---- two right-hand sides per pass (gemv_kernels.cuh); callers check hyp_gemv2_ok first ----
bool hyp_gemv2_ok(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* a, const double* b) {
    (void)ctx;
    return rows >= 4096 && ncols >= 64 && (ld % 2 == 0) && ((uintptr_t)M % 16 == 0) && ((uintptr_t)a % 16 == 0) &&
           ((uintptr_t)b % 16 == 0);
}

void hyp_gemv_t2(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x0,
                 const double* x1, double alpha, double beta, double* y0, double* y1) {
    hyp_time_begin(ctx, T_GEMV);
    int grid = (int)std::min<int64_t>(ncols, (int64_t)ctx->sm_count * 16);
    hypdev::gemv_t2_cta_kernel<<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x0, x1, alpha, beta, y0, y1);
    ctx->launches++;
    hyp_time_end(ctx, T_GEMV);
}

static void ensure_partial(hyp_ctx* ctx, double** buf, int64_t* have, int64_t need) {
    if (need <= *have) return;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    CUDA_TRY(cudaMalloc(buf, (size_t)need * sizeof(double)));
    *have = need;
}

void hyp_gemv_n2(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x0,
                 const double* x1, double alpha, double beta, double* y0, double* y1) {
    hyp_time_begin(ctx, T_GEMV);
    // the same chunking as hyp_gemv_n, so that each column is bit-identical to the single-vector product
    int row_blocks = ceil_div(rows, 256);
    int64_t target = (int64_t)ctx->sm_count * 8;
    int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(target / row_blocks, 64));
    nchunks = (int)std::min<int64_t>(nchunks, (ncols + 31) / 32);
    while ((int64_t)nchunks * rows > ctx->partial_doubles && nchunks > 1) nchunks--;
    int64_t cpc = (ncols + nchunks - 1) / nchunks;
    nchunks = ceil_div(ncols, cpc);
    const int64_t pstride = (int64_t)nchunks * rows;
    ensure_partial(ctx, &ctx->d_partial3, &ctx->partial3_doubles, 2 * pstride);
    dim3 grid(row_blocks, nchunks);
    hypdev::gemv_n2_kernel<<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, x0, x1, cpc, ctx->d_partial3, pstride);
    int rgrid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)ctx->sm_count * 8);
    hypdev::gemv_n_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial3, alpha, beta, y0);
    hypdev::gemv_n_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial3 + pstride, alpha, beta, y1);
    ctx->launches += 3;
    hyp_time_end(ctx, T_GEMV);
}

// w_v = alphaN * M x_v + betaN * w_v  and  y_v = alphaT * M' z_v + betaT * y_v  for v = 0, 1 in ONE pass over M
void hyp_gemv_nt2(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* xa,
                  const double* xb, const double* za, const double* zb, double alphaN, double betaN, double* wa,
                  double* wb, double alphaT, double betaT, double* ya, double* yb) {
    hyp_time_begin(ctx, T_GEMV);
    const int row_blocks = ceil_div(rows, 256);
    const int64_t target = (int64_t)ctx->sm_count * 8;
    int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(target / row_blocks, 64));
    nchunks = (int)std::min<int64_t>(nchunks, (ncols + 63) / 64);
    while ((int64_t)nchunks * rows > ctx->partial_doubles && nchunks > 1) nchunks--;
    int64_t cpc = round_up((ncols + nchunks - 1) / nchunks, 8);
    nchunks = ceil_div(ncols, cpc);
    const int64_t psN = (int64_t)nchunks * rows, psT = (int64_t)row_blocks * 4 * ncols;
    ensure_partial(ctx, &ctx->d_partial3, &ctx->partial3_doubles, 2 * psN);
    ensure_partial(ctx, &ctx->d_partial4, &ctx->partial4_doubles, 2 * psT);
    dim3 grid(row_blocks, nchunks);
    hypdev::gemv_nt2_kernel<<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, xa, xb, za, zb, cpc, ctx->d_partial3, psN,
                                                          ctx->d_partial4, psT);
    const int rgrid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)ctx->sm_count * 8);
    const int tgrid = (int)std::min<int64_t>(ceil_div(ncols, 256), (int64_t)ctx->sm_count * 8);
    hypdev::gemv_n_reduce2_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial3, alphaN, betaN, wa);
    hypdev::gemv_n_reduce2_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial3 + psN, alphaN, betaN, wb);
    hypdev::gemv_t_reduce_kernel<<<tgrid, 256, 0, ctx->stream>>>(ncols, row_blocks * 4, ctx->d_partial4, alphaT, betaT, ya);
    hypdev::gemv_t_reduce_kernel<<<tgrid, 256, 0, ctx->stream>>>(ncols, row_blocks * 4, ctx->d_partial4 + psT, alphaT, betaT, yb);
    ctx->launches += 5;
    hyp_time_end(ctx, T_GEMV);
    CUDA_TRY(cudaGetLastError());
}

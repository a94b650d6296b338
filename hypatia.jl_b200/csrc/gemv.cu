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

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[j] = alpha * dot(M[:, j], x) + beta * y[j];  one CTA per column.
template <bool VEC>
__global__ void __launch_bounds__(256) gemv_t_cta_kernel(int64_t rows, int64_t ncols,
                                                         const double* __restrict__ M, int64_t ld,
                                                         const double* __restrict__ x, double alpha,
                                                         double beta, double* __restrict__ y) {
    __shared__ double sm[8];
    for (int64_t j = blockIdx.x; j < ncols; j += gridDim.x) {
        const double* col = M + j * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (VEC) {
            const double2* c2 = reinterpret_cast<const double2*>(col);
            const double2* x2 = reinterpret_cast<const double2*>(x);
            int64_t n2 = rows >> 1;
            int64_t i = threadIdx.x;
            for (; i + 3 * 256 < n2; i += 4 * 256) {
                double2 m0 = __ldg(c2 + i), m1 = __ldg(c2 + i + 256), m2 = __ldg(c2 + i + 512),
                        m3 = __ldg(c2 + i + 768);
                double2 v0 = x2[i], v1 = x2[i + 256], v2 = x2[i + 512], v3 = x2[i + 768];
                a0 += m0.x * v0.x + m0.y * v0.y;
                a1 += m1.x * v1.x + m1.y * v1.y;
                a2 += m2.x * v2.x + m2.y * v2.y;
                a3 += m3.x * v3.x + m3.y * v3.y;
            }
            for (; i < n2; i += 256) {
                double2 m0 = __ldg(c2 + i);
                double2 v0 = x2[i];
                a0 += m0.x * v0.x + m0.y * v0.y;
            }
            if ((rows & 1) && threadIdx.x == 0) a1 += col[rows - 1] * x[rows - 1];
        } else {
            for (int64_t i = threadIdx.x; i < rows; i += 256) a0 += col[i] * x[i];
        }
        double acc = warp_sum((a0 + a1) + (a2 + a3));
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) t += sm[w];
            y[j] = alpha * t + (beta == 0.0 ? 0.0 : beta * y[j]);
        }
        __syncthreads();
    }
}

// one warp per column (short columns)
__global__ void __launch_bounds__(256) gemv_t_warp_kernel(int64_t rows, int64_t ncols,
                                                          const double* __restrict__ M, int64_t ld,
                                                          const double* __restrict__ x, double alpha,
                                                          double beta, double* __restrict__ y) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < ncols; j += nwarps) {
        const double* col = M + j * ld;
        double acc = 0;
        for (int64_t i = lane; i < rows; i += 32) acc += col[i] * x[i];
        acc = warp_sum(acc);
        if (lane == 0) y[j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[j]);
    }
}

// partial[chunk][r] = sum_{j in chunk} M[r, j] x[j]; thread owns 2 consecutive rows.
template <bool VEC>
__global__ void __launch_bounds__(128) gemv_n_kernel(int64_t rows, int64_t ncols,
                                                     const double* __restrict__ M, int64_t ld,
                                                     const double* __restrict__ x, int64_t cols_per_chunk,
                                                     double* __restrict__ partial) {
    int64_t r = (blockIdx.x * 128 + threadIdx.x) * 2;
    int64_t j0 = blockIdx.y * cols_per_chunk;
    int64_t j1 = min(ncols, j0 + cols_per_chunk);
    if (r >= rows) return;
    double ax = 0, ay = 0, bx = 0, by = 0;
    if (VEC && r + 1 < rows) {
        const double* base = M + r;
        int64_t j = j0;
        for (; j + 7 < j1; j += 8) {
            double2 m[8];
#pragma unroll
            for (int u = 0; u < 8; u++) m[u] = __ldg(reinterpret_cast<const double2*>(base + (j + u) * ld));
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                double x0 = x[j + u], x1 = x[j + u + 1];
                ax += m[u].x * x0;
                ay += m[u].y * x0;
                bx += m[u + 1].x * x1;
                by += m[u + 1].y * x1;
            }
        }
        for (; j < j1; j++) {
            double2 m0 = __ldg(reinterpret_cast<const double2*>(base + j * ld));
            double x0 = x[j];
            ax += m0.x * x0;
            ay += m0.y * x0;
        }
        double* out = partial + blockIdx.y * rows + r;
        out[0] = ax + bx;
        out[1] = ay + by;
    } else {
        for (int64_t j = j0; j < j1; j++) {
            double x0 = x[j];
            ax += M[r + j * ld] * x0;
            if (r + 1 < rows) ay += M[r + 1 + j * ld] * x0;
        }
        double* out = partial + blockIdx.y * rows + r;
        out[0] = ax;
        if (r + 1 < rows) out[1] = ay;
    }
}

__global__ void gemv_n_reduce_kernel(int64_t rows, int nchunks, const double* __restrict__ partial,
                                     double alpha, double beta, double* __restrict__ y) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0;
        for (int c = 0; c < nchunks; c++) acc += partial[(int64_t)c * rows + r];
        y[r] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[r]);
    }
}

}  // namespace

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
            gemv_t_cta_kernel<true><<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x, alpha, beta, y);
        else
            gemv_t_cta_kernel<false><<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x, alpha, beta, y);
    } else {
        int grid = (int)std::min<int64_t>((ncols + 7) / 8, (int64_t)ctx->sm_count * 8);
        gemv_t_warp_kernel<<<grid, 256, 0, ctx->stream>>>(rows, ncols, M, ld, x, alpha, beta, y);
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
        gemv_n_kernel<true><<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, x, cpc, ctx->d_partial);
    else
        gemv_n_kernel<false><<<grid, 128, 0, ctx->stream>>>(rows, ncols, M, ld, x, cpc, ctx->d_partial);
    int rgrid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)ctx->sm_count * 8);
    gemv_n_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(rows, nchunks, ctx->d_partial, alpha, beta, y);
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

// Dense Cholesky factorisation and triangular solves of the Schur complement (K3 / K5 of
// SURVEY.md section 2.3).
//
// reference call sites: posdef_fact!(A) = cholesky!(Symmetric(A, :U), check=false)
// (src/linearalgebra/dense.jl:191-192, LAPACK dpotrf 'U') called from update_lhs_fact
// (qrchol.jl:249-250); ldiv!(x, fact, rhs) (qrchol.jl:68, LAPACK dpotrs).
//
// potrf: right-looking blocked upper Cholesky, block size 128.
//   per block column k:  (1) one-CTA panel kernel: factor the 128 x 128 diagonal block in
//   registers (1024 threads x 4 x 4 cyclic sub-blocks, one barrier per pivot) and invert the
//   triangular factor in shared memory;  (2) U12 = U11^-T A12 as a TN GEMM with the inverted
//   block (syrk.cu, TMA + DMMA);  (3) trailing update A22 -= U12' U12 on the upper tiles with the
//   same TMA + DMMA kernel as the Schur SYRK.  Bound: tensor (FP64 DMMA); m^3/3 flops.
//   The inverted diagonal blocks are kept: the triangular solves use them.
// trsv: one persistent kernel per triangular solve; CTAs take block columns in dependency order
//   from a ticket counter and publish finished 128-blocks of the solution through release /
//   acquire flags, so a solve is one launch instead of 2 * m / 128.  Bound: HBM (reads the
//   triangle once: 4 m^2 bytes) - in practice latency of the block dependency chain.
#include "common.cuh"
#include "chol_kernels.cuh"

using namespace hypdev;   // NB, LDU, PT, panel_kernel, chol_batched_kernel

namespace {

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- triangular solve with the blocked factor -------------------------------------------
// flags[0] = ticket counter (zeroed by the host before the launch), flags[1 + k] = epoch when
// block k of the solution is final.  CTA for block k: the inverted diagonal block goes to shared
// memory up front, and every off-diagonal tile is loaded into registers BEFORE the CTA waits for
// the solution block it multiplies, so the dependency chain only sees: flag -> 1 KB vector load ->
// 64 FMAs per thread -> reduction -> 128 x 128 matvec from shared memory -> publish.
constexpr int TRSV_SMEM = NB * NB * 8;

template <bool TRANS>
__global__ void __launch_bounds__(256, 1)
trsv_kernel(const double* __restrict__ F, int64_t ldf, int64_t m, const double* __restrict__ dinv,
            double* x, int* flags, int nblk, int epoch) {
    extern __shared__ double sD[];           // Dinv_k, 128 x 128 col-major
    __shared__ double sv[2][NB];
    __shared__ double sacc[2][NB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) s_ticket = atomicAdd(&flags[0], 1);
        __syncthreads();
        const int t = s_ticket;
        __syncthreads();
        if (t >= nblk) return;
        const int k = TRANS ? t : nblk - 1 - t;
        const int64_t c0 = (int64_t)k * NB;
        {
            const double* Dk = dinv + (int64_t)k * NB * NB;
#pragma unroll 8
            for (int idx = tid; idx < NB * NB; idx += 256) sD[idx] = Dk[idx];
        }

        if (TRANS) {
            // y_k = Dinv_k' (b_k - sum_{j<k} U[j-block, k-block]' y_j); warp w owns 16 columns,
            // lane l rows l, l+32, l+64, l+96 of every tile
            double pacc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) pacc[i] = 0.0;
            const int64_t cw = c0 + warp * 16;
            for (int j = 0; j < k; j++) {
                double tl[16][4];
                const double* Ut = F + (int64_t)j * NB + cw * ldf;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const bool ok = cw + i < m;
                    const double* col = Ut + (int64_t)i * ldf;
#pragma unroll
                    for (int h = 0; h < 4; h++) tl[i][h] = ok ? col[lane + 32 * h] : 0.0;
                }
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                double* svj = sv[j & 1];
                if (tid < NB) svj[tid] = __ldcg(x + (int64_t)j * NB + tid);
                __syncthreads();
                const double v0 = svj[lane], v1 = svj[lane + 32], v2 = svj[lane + 64], v3 = svj[lane + 96];
#pragma unroll
                for (int i = 0; i < 16; i++)
                    pacc[i] += tl[i][0] * v0 + tl[i][1] * v1 + tl[i][2] * v2 + tl[i][3] * v3;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                double a = pacc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                pacc[i] = a;
            }
            __syncthreads();
            double* vb = sv[0];
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    int64_t c = cw + i;
                    vb[warp * 16 + i] = (c < m) ? (x[c] - pacc[i]) : 0.0;
                }
            }
            __syncthreads();
            // y[c] = sum_{r <= c} Dinv[r, c] v[r]
            const double v0 = vb[lane], v1 = vb[lane + 32], v2 = vb[lane + 64], v3 = vb[lane + 96];
            double res[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const double* col = sD + (warp * 16 + i) * NB;
                double a = col[lane] * v0 + col[lane + 32] * v1 + col[lane + 64] * v2 + col[lane + 96] * v3;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                res[i] = a;
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    int64_t c = cw + i;
                    if (c < m) x[c] = res[i];
                }
            }
        } else {
            // x_k = Dinv_k (y_k - sum_{j>k} U[k-block, j-block] x_j); thread owns a row, the two
            // halves of the CTA split the 128 columns of a tile
            const int r = tid & (NB - 1), half = tid >> 7;
            const int64_t grow = c0 + r;
            double acc = 0.0;
            for (int j = nblk - 1; j > k; j--) {
                double tl[64];
                const int64_t cb = (int64_t)j * NB + half * 64;
                const double* Ut = F + grow + cb * ldf;
#pragma unroll
                for (int c = 0; c < 64; c++) tl[c] = (grow < m && cb + c < m) ? Ut[(int64_t)c * ldf] : 0.0;
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                double* svj = sv[j & 1];
                if (tid < NB) {
                    int64_t c = (int64_t)j * NB + tid;
                    svj[tid] = (c < m) ? __ldcg(x + c) : 0.0;
                }
                __syncthreads();
                const double* svh = svj + half * 64;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    a0 += tl[c] * svh[c];
                    a1 += tl[c + 1] * svh[c + 1];
                }
                acc += a0 + a1;
            }
            sacc[half][r] = acc;
            __syncthreads();
            double* vb = sv[0];
            if (tid < NB) vb[tid] = (c0 + tid < m) ? (x[c0 + tid] - sacc[0][tid] - sacc[1][tid]) : 0.0;
            __syncthreads();
            // x[r] = sum_{c >= r} Dinv[r, c] v[c]
            double a0 = 0.0, a1 = 0.0;
            {
                const double* row = sD + r + (half * 64) * NB;
                const double* svh = vb + half * 64;
#pragma unroll 16
                for (int c = 0; c < 64; c += 2) {
                    a0 += row[c * NB] * svh[c];
                    a1 += row[(c + 1) * NB] * svh[c + 1];
                }
            }
            __syncthreads();
            sacc[half][r] = a0 + a1;
            __syncthreads();
            if (tid < NB && c0 + tid < m) x[c0 + tid] = sacc[0][tid] + sacc[1][tid];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&flags[1 + k], epoch);
    }
}

bool g_panel_attr_set = false;
void set_panel_attr() {
    if (g_panel_attr_set) return;
    CUDA_TRY(cudaFuncSetAttribute(panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  NB * LDU * 8));
    CUDA_TRY(cudaFuncSetAttribute(panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  NB * LDU * 8));
    CUDA_TRY(cudaFuncSetAttribute(chol_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  NB * LDU * 8));
    CUDA_TRY(cudaFuncSetAttribute(trsv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    g_panel_attr_set = true;
}

}  // namespace

void hyp_potrf_upper(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info) {
    if (m <= 0) return;
    TimeScope ts(ctx, T_POTRF);
    set_panel_attr();
    cudaStream_t bulk = ctx->stream, chain = ctx->stream2;
    CUDA_TRY(cudaMemsetAsync(d_info, 0, sizeof(int), bulk));
    // Two-level blocking with look-ahead.  Outer blocks of OB = 512 columns; inside an outer block four
    // 128-wide panels.
    //   chain stream: factors the OB x OB diagonal block (panel kernel + TRSM / update restricted to that
    //                 block: a handful of tiles) - the latency-bound part;
    //   bulk stream:  block row U12 = U11^-T A12 for the columns right of the block, then the depth-OB
    //                 update of the trailing matrix - the throughput-bound part.  The first thing it
    //                 updates is the NEXT diagonal block, after which the chain stream starts on it while
    //                 the bulk stream finishes the update (its persistent grids leave 8 SMs free).
    const int64_t OB = 512;
    const bool lookahead = m > 2 * OB;
    auto on = [&](cudaStream_t s, int cap) {
        ctx->launch_stream = s;
        ctx->grid_cap = cap;
    };
    const int cap = lookahead ? std::max(1, ctx->sm_count - 8) : 0;
    int nblk_outer = 0;
    // the chain may start once the input matrix is in place (everything before us on the bulk stream)
    CUDA_TRY(cudaEventRecord(ctx->ev_bulk[0], bulk));
    CUDA_TRY(cudaStreamWaitEvent(chain, ctx->ev_bulk[0], 0));
    for (int64_t K0 = 0; K0 < m; K0 += OB, nblk_outer++) {
        const int64_t Kend = std::min(K0 + OB, m);
        const int64_t rest2 = m - Kend;
        // ---- chain: diagonal block ----
        on(chain, 0);
        for (int64_t k0 = K0; k0 < Kend; k0 += NB) {
            const int64_t nb = std::min<int64_t>(NB, m - k0);
            const int64_t k = k0 / NB;
            panel_kernel<true><<<1, PT, NB * LDU * 8, chain>>>(A, lda, m, k, d_dinv, d_info);
            ctx->launches++;
            const int64_t rin = Kend - (k0 + nb);       // columns / rows of this outer block right of / below the panel
            if (rin > 0) {
                double* A12 = A + k0 + (k0 + nb) * lda;
                double* A22 = A + (k0 + nb) + (k0 + nb) * lda;
                const double* Dk = d_dinv + k * NB * NB;
                hyp_gemm_tn(ctx, Dk, NB, A12, lda, nb, nb, rin, A12, lda, 1.0, 0.0);
                hyp_atb_upper(ctx, A12, lda, A12, lda, nb, rin, A22, lda, -1.0, 1.0);
            }
        }
        CUDA_TRY(cudaEventRecord(ctx->ev_chain[nblk_outer & 1], chain));
        if (rest2 <= 0) break;
        // ---- bulk: block row right of the outer block, panel by panel ----
        CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_chain[nblk_outer & 1], 0));
        on(bulk, cap);
        for (int64_t k0 = K0; k0 < Kend; k0 += NB) {
            const int64_t nb = std::min<int64_t>(NB, m - k0);
            const int64_t k = k0 / NB;
            double* A12r = A + k0 + Kend * lda;                 // rows of the panel, columns >= Kend
            const double* Dk = d_dinv + k * NB * NB;
            hyp_gemm_tn(ctx, Dk, NB, A12r, lda, nb, nb, rest2, A12r, lda, 1.0, 0.0);
            const int64_t rin = Kend - (k0 + nb);
            if (rin > 0) {
                // rows below the panel inside the outer block: A[k0+nb:Kend, Kend:] -= U[panel, k0+nb:Kend]' U12r
                const double* Pu = A + k0 + (k0 + nb) * lda;
                hyp_gemm_tn(ctx, Pu, lda, A12r, lda, nb, rin, rest2, A + (k0 + nb) + Kend * lda, lda, -1.0, 1.0);
            }
        }
        // ---- bulk: depth-OB trailing update; the next diagonal block first ----
        double* P = A + K0 + Kend * lda;
        double* T = A + Kend + Kend * lda;
        const int64_t kd = Kend - K0;
        const int64_t dn = std::min<int64_t>(OB, rest2);
        hyp_atb_upper(ctx, P, lda, P, lda, kd, dn, T, lda, -1.0, 1.0);
        CUDA_TRY(cudaEventRecord(ctx->ev_bulk[nblk_outer & 1], bulk));
        CUDA_TRY(cudaStreamWaitEvent(chain, ctx->ev_bulk[nblk_outer & 1], 0));
        if (rest2 > dn) {
            // rows of the next diagonal block x the columns right of it, then everything below
            hyp_gemm_tn(ctx, P, lda, P + dn * lda, lda, kd, dn, rest2 - dn, T + dn * lda, lda, -1.0, 1.0);
            hyp_atb_upper(ctx, P + dn * lda, lda, P + dn * lda, lda, kd, rest2 - dn, T + dn + dn * lda, lda, -1.0, 1.0);
        }
    }
    // the caller's stream continues after the last diagonal block
    CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_chain[nblk_outer & 1], 0));
    on(nullptr, 0);
    CUDA_TRY(cudaGetLastError());
}

void hyp_trtri_diag(hyp_ctx* ctx, const double* U, int64_t ldu, int64_t m, double* d_dinv) {
    if (m <= 0) return;
    set_panel_attr();
    int nblk = ceil_div(m, NB);
    panel_kernel<false><<<nblk, PT, NB * LDU * 8, ctx->stream>>>(const_cast<double*>(U), ldu, m, 0,
                                                                d_dinv, nullptr);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

void hyp_trsv_upper(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, const double* d_dinv,
                    double* x, bool trans) {
    if (m <= 0) return;
    TimeScope ts(ctx, T_TRSV);
    int nblk = ceil_div(m, NB);
    CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int epoch = ++ctx->trsv_epoch;
    set_panel_attr();
    int grid = std::min(nblk, ctx->sm_count);
    if (trans)
        trsv_kernel<true><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, nblk, epoch);
    else
        trsv_kernel<false><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, nblk, epoch);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

void hyp_chol_batched(hyp_ctx* ctx, int ncones, const int* d_sides, const int64_t* d_moff,
                      const int* d_kidx, double* U, double* Ui, uint8_t* d_flag) {
    if (ncones <= 0) return;
    set_panel_attr();
    chol_batched_kernel<<<ncones, PT, NB * LDU * 8, ctx->stream>>>(ncones, d_sides, d_moff, d_kidx, U, Ui,
                                                                    d_flag);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

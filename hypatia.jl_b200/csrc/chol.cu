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
#include <cstring>
#include <cstdlib>

using namespace hypdev;   // NB, LDU, PT, panel_kernel, chol_batched_kernel

namespace {

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
    CUDA_TRY(cudaFuncSetAttribute(trsv_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    g_panel_attr_set = true;
}

}  // namespace

void hyp_potrf_upper(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info) {
    if (m <= 0) return;
    // default: the task-graph kernel (chol_dag.cu), one launch for the whole factorisation;
    // HYP_POTRF=stream keeps the two-stream version below (also the fallback for unaligned input)
    {
        const char* pv = getenv("HYP_POTRF");
        const bool want_stream = pv && !strcmp(pv, "stream");
        if (!want_stream && hyp_potrf_upper_dag(ctx, A, lda, m, d_dinv, d_info)) return;
    }
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
    // profiling aid (tools/potrf_probe.py): 1 = only the latency chain, 2 = only the bulk GEMMs (results are garbage)
    const char* pm = getenv("HYP_POTRF_MODE");
    const int probe = pm ? atoi(pm) : 0;
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
            if (probe != 2) {
                panel_kernel<true><<<1, PT, NB * LDU * 8, chain>>>(A, lda, m, k, d_dinv, d_info);
                ctx->launches++;
            }
            const int64_t rin = Kend - (k0 + nb);       // columns / rows of this outer block right of / below the panel
            if (rin > 0 && probe != 2) {
                double* A12 = A + k0 + (k0 + nb) * lda;
                double* A22 = A + (k0 + nb) + (k0 + nb) * lda;
                const double* Dk = d_dinv + k * NB * NB;
                hyp_gemm_tn(ctx, Dk, NB, A12, lda, nb, nb, rin, A12, lda, 1.0, 0.0);
                hyp_atb_upper(ctx, A12, lda, A12, lda, nb, rin, A22, lda, -1.0, 1.0);
            }
        }
        CUDA_TRY(cudaEventRecord(ctx->ev_chain[nblk_outer & 1], chain));
        if (rest2 <= 0) break;
        if (probe == 1) continue;
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

// two right-hand sides x and x + xstride per sweep: every tile of the factor is read once for both
void hyp_trsv_upper2(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, const double* d_dinv, double* x,
                     int64_t xstride, bool trans) {
    if (m <= 0) return;
    TimeScope ts(ctx, T_TRSV);
    int nblk = ceil_div(m, NB);
    CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int epoch = ++ctx->trsv_epoch;
    set_panel_attr();
    int grid = std::min(nblk, ctx->sm_count);
    if (trans)
        trsv_kernel<true, 2><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, nblk, epoch, xstride);
    else
        trsv_kernel<false, 2><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, nblk, epoch, xstride);
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

// phase timestamps of the last panel_kernel<true> launch (tools/panel_probe.py): cycles relative to the start
extern "C" int hyp_test_panel_clocks(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* cycles16) {
    if (!ctx || m > NB) return -1;
    try {
        set_panel_attr();
        double *dA = nullptr, *dD = nullptr;
        int* dI = nullptr;
        CUDA_TRY(cudaMalloc(&dA, (size_t)lda * m * 8));
        CUDA_TRY(cudaMalloc(&dD, (size_t)NB * NB * 8));
        CUDA_TRY(cudaMalloc(&dI, 8));
        long long clk[16];
        {
            const char* pf = getenv("HYP_PANEL_FLAGS");
            int flags = pf ? atoi(pf) : 0;
            CUDA_TRY(cudaMemcpyToSymbol(g_panel_flags, &flags, sizeof(int)));
        }
        for (int rep = 0; rep < 3; rep++) {
            CUDA_TRY(cudaMemcpyAsync(dA, A, (size_t)lda * m * 8, cudaMemcpyDefault, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(dI, 0, 8, ctx->stream));
            panel_kernel<true><<<1, PT, NB * LDU * 8, ctx->stream>>>(dA, lda, m, 0, dD, dI);
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        CUDA_TRY(cudaMemcpyFromSymbol(clk, g_panel_clk, sizeof(clk)));
        for (int i = 0; i < 16; i++) cycles16[i] = (double)(clk[i] - clk[0]);
        cudaFree(dA);
        cudaFree(dD);
        cudaFree(dI);
        return 0;
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        return -1;
    }
}

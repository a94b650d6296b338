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
#include "trsv_tasks.h"
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
    CUDA_TRY(cudaFuncSetAttribute(trsv_pkt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_pkt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_pkt_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_pkt_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_seg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_seg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_seg_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(trsv_seg_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_SMEM));
    g_panel_attr_set = true;
}

// ticket lists of the segmented triangular solve, per (device, nblk, sweep direction), and its partial-sum scratch
struct TrsvLists {
    int device, nblk, trans, ntasks, maxseg;
    TrsvTask* d_tasks;
};
std::vector<TrsvLists> g_trsv_lists;
int trsv_seg_len() {
    static int seg = 0;
    if (!seg) {
        const char* e = getenv("HYP_TRSV_SEG");           // 0 (default) = one ticket per block column; n > 0: segments of n tiles
        seg = e ? atoi(e) : 0;
        if (seg < 0) seg = 8;
        if (seg == 0) seg = -1;
    }
    return seg;
}
const TrsvLists& trsv_lists(hyp_ctx* ctx, int nblk, bool trans) {
    for (auto& e : g_trsv_lists)
        if (e.device == ctx->device && e.nblk == nblk && e.trans == (trans ? 1 : 0)) return e;
    int maxseg = 1;
    std::vector<TrsvTask> t = trsv_build_tasks(nblk, trans, trsv_seg_len(), &maxseg);
    TrsvLists e{ctx->device, nblk, trans ? 1 : 0, (int)t.size(), maxseg, nullptr};
    CUDA_TRY(cudaMalloc((void**)&e.d_tasks, t.size() * sizeof(TrsvTask)));
    CUDA_TRY(cudaMemcpyAsync(e.d_tasks, t.data(), t.size() * sizeof(TrsvTask), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    g_trsv_lists.push_back(e);
    return g_trsv_lists.back();
}
// partial sums: nblk x maxseg x (up to 2 right-hand sides) x 128 doubles, per context
// packet variant: opt-in (HYP_TRSV_PKT=1 or hyp_test_set_trsv_pkt).  Measured on C3 / C2 (profiles/r02_bench_trsv_packets_ab.md):
// 3.69 vs 3.67 ms and 1.21 vs 1.28 ms per step - the sweeps are bound by the 128 KB tile every CTA pulls into ONE SM per
// dependency step (~42 B/clk per SM from L2), not by the flag / fence protocol the packets remove.
int g_trsv_pkt = -1;
bool trsv_use_pkt() {
    if (g_trsv_pkt < 0) {
        const char* e = getenv("HYP_TRSV_PKT");
        g_trsv_pkt = (e && atoi(e) == 1) ? 1 : 0;
    }
    return g_trsv_pkt == 1;
}

// packet buffer of trsv_pkt_kernel: 2 words per entry, nblk blocks of 128 entries, nrhs right-hand sides; zeroed when
// (re)allocated - epochs of the context start at 1 and only grow, so a zero word never looks published
unsigned long long* trsv_pkt(hyp_ctx* ctx, int nblk, int nrhs) {
    const int64_t need = (int64_t)nrhs * nblk * NB * 2;
    if (ctx->trsv_pkt_words < need) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_trsv_pkt) cudaFree(ctx->d_trsv_pkt);
        ctx->d_trsv_pkt = nullptr;
        ctx->trsv_pkt_words = 0;
        CUDA_TRY(cudaMalloc((void**)&ctx->d_trsv_pkt, (size_t)need * 8));
        CUDA_TRY(cudaMemset(ctx->d_trsv_pkt, 0, (size_t)need * 8));
        ctx->trsv_pkt_words = need;
    }
    return ctx->d_trsv_pkt;
}

double* trsv_part(hyp_ctx* ctx, int nblk, int maxseg) {
    const int64_t need = (int64_t)nblk * maxseg * 2 * NB;
    if (ctx->trsv_part_len < need) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_trsv_part) cudaFree(ctx->d_trsv_part);
        ctx->d_trsv_part = nullptr;
        CUDA_TRY(cudaMalloc((void**)&ctx->d_trsv_part, (size_t)need * sizeof(double)));
        ctx->trsv_part_len = need;
    }
    return ctx->d_trsv_part;
}

}  // namespace

// ---- blocked Cholesky whose depth-512 trailing updates run on the 5th-generation tensor cores --------------------------
// The FP64 DMMA pipe bounds the task-graph kernel (chol_dag.cu): 333 GFLOP at 35 TFLOP/s are 9.5 ms for m = 10000 before
// any dependency stall (measured 16.9 ms).  81 % of those flops are the depth-512 updates of the trailing matrix
// A22 -= U12' U12, a SYRK - exactly what the digit-sliced tcgen05 kernel of the Schur assembly (ozaki.cu) computes at
// 2.5 x the DMMA rate: the finished block row U12 (512 x rest) is cut into seven int8 digit slices (one pass, 75 MB) and
// ozaki_syrk_pair64_kernel subtracts the 28 exact digit-pair products from the trailing tiles (alpha = -1, beta = 1).
// What stays on CUDA cores / DMMA is the latency chain: the 128-wide panels of the 512 x 512 diagonal block and the
// triangular solves of the block row (products with the inverted diagonal blocks).
// Two streams.  chain: diag(b) -> block row over the NEXT block's columns -> DMMA update of the next diagonal block ->
// diag(b + 1) ...; bulk: block row over the far columns -> slicing -> SYRK part a (tile rows of the next two blocks, the
// chain waits for it before its near block row) -> SYRK part b (the rest).  The bulk grids leave 8 SMs to the chain.
static bool potrf_upper_i8(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info) {
    // outer block = depth of the trailing updates: 512 (default) or 1024 (HYP_POTRF_OB=1024: half as many passes over the
    // trailing tiles, each twice as deep, against a longer diagonal-block chain)
    static const int ob_env = getenv("HYP_POTRF_OB") ? atoi(getenv("HYP_POTRF_OB")) : 0;
    // 1024 pays once the bulk stream dominates (m = 20000: 69 -> 55 ms); below that the deeper update of the next diagonal
    // block (128 x 128 tiles, on the chain) costs more than the trailing passes save
    const int64_t OB = (ob_env == 512 || ob_env == 1024) ? ob_env : (m >= 16384 ? 1024 : 512);
    if (m <= 1024 || !hyp_ozaki_pair64_ready(ctx)) return false;
    TimeScope ts(ctx, T_POTRF);
    set_panel_attr();
    cudaStream_t bulk = ctx->stream, chain = ctx->stream2;
    // digit slices of one block row: 7 x OB x m bytes, and the column scales
    if (ctx->chol_digits_cols < m) {
        CUDA_TRY(cudaStreamSynchronize(bulk));
        if (ctx->d_chol_digits) cudaFree(ctx->d_chol_digits);
        if (ctx->d_chol_dscale) cudaFree(ctx->d_chol_dscale);
        ctx->d_chol_digits = nullptr;
        ctx->d_chol_dscale = nullptr;
        CUDA_TRY(cudaMalloc((void**)&ctx->d_chol_digits, (size_t)7 * 1024 * m));       // sized for the deeper outer block
        CUDA_TRY(cudaMalloc((void**)&ctx->d_chol_dscale, (size_t)m * sizeof(double)));
        ctx->chol_digits_cols = m;
    }
    const int64_t ldd = OB, sstride = OB * ctx->chol_digits_cols;
    CUDA_TRY(cudaMemsetAsync(d_info, 0, sizeof(int), bulk));
    struct LaunchStateGuard {                      // the launch helpers go back to the context's stream on every exit path
        hyp_ctx* c;
        ~LaunchStateGuard() {
            c->launch_stream = nullptr;
            c->grid_cap = 0;
            c->small_tiles = false;
        }
    } guard{ctx};
    // the chain stream's products are small (128 x <= 512 outputs): latency tile shape unless HYP_POTRF_TILES=big
    // HYP_POTRF_TILES = big: 128 x 128 tiles everywhere; chain / bulk: the latency shape only on that stream (debugging)
    static const char* tiles_env = getenv("HYP_POTRF_TILES");
    static const bool narrow_chain = !(tiles_env && (!strcmp(tiles_env, "big") || !strcmp(tiles_env, "bulk")));
    static const bool narrow_rows = !(tiles_env && (!strcmp(tiles_env, "big") || !strcmp(tiles_env, "chain")));
    // debugging: which launches may use the latency shape (1 in-block TRSM, 2 in-block update, 4 near block row,
    // 8 update of the next diagonal block, 16 far block row)
    // Default 21 = the triangular-solve-type launches (1, 4, 16).  The UPDATE-type launches (2, 8: C -= P'P with P = R) stay on
    // 128 x 128 tiles: with the latency shape, repeated factorisations of one matrix stopped being bit-identical whenever
    // such a launch overlapped the slicing kernel of the other stream (tools/potrf_race.py; profiles/r02_potrf_race.md:
    // up to 78 of 79 repetitions on some boxes, none on others; relative error of the factor ~ 1e-3).  Every read / write
    // set of the two streams is disjoint and the event order is complete, so the cause is not understood - the shape is
    // therefore only used where 60 - 80 repetitions on an affected box showed no difference.
    static const int nmask = getenv("HYP_POTRF_NARROW_MASK") ? atoi(getenv("HYP_POTRF_NARROW_MASK")) : 21;
    auto on = [&](cudaStream_t s, int cap) {
        ctx->launch_stream = s;
        ctx->grid_cap = cap;
        ctx->small_tiles = narrow_chain && s == chain;
    };
    const int cap = std::max(2, (ctx->sm_count - 8) & ~1);
    const int pa_rows = (int)(2 * OB / 256);       // SYRK part a: the 256-row tile pairs of the next two outer blocks
    // profiling aid (tools/potrf_probe.py; results are garbage): 1 = no digit-sliced updates, 2 = chain stream only,
    // 3 = bulk stream only (no panels / in-block products)
    const char* pm = getenv("HYP_POTRF_MODE");
    const int probe = pm ? atoi(pm) : 0;
    // debugging aid: device-wide synchronisation points (bit 0: end of every outer block, 1: after the chain's near part,
    // 2: after SYRK part a, 3: after the diagonal block)
    const char* psy = getenv("HYP_POTRF_SYNC");
    const int dbg_sync = psy ? atoi(psy) : 0;
    // block row of outer block [K0, Kend) over the columns [c0, c0 + nc): U = U11^-T A, panel by panel.  2 * OB / 128 - 1
    // dependent launches of depth 128: the latency tile shape (128 x 32) unless the panel is wide enough to fill the
    // chip several times over with 128 x 128 tiles
    auto block_row = [&](int64_t K0, int64_t Kend, int64_t c0, int64_t nc) {
        const bool saved = ctx->small_tiles;
        if (narrow_rows && (nmask & 16) && (ctx->launch_stream == bulk) && nc <= (int64_t)4 * 128 * ctx->sm_count) ctx->small_tiles = true;
        struct Restore {
            hyp_ctx* c;
            bool v;
            ~Restore() { c->small_tiles = v; }
        } restore{ctx, saved};
        for (int64_t k0 = K0; k0 < Kend; k0 += NB) {
            const int64_t nb = std::min<int64_t>(NB, m - k0);
            double* Ar = A + k0 + c0 * lda;                       // rows of the panel
            hyp_gemm_tn(ctx, d_dinv + (k0 / NB) * NB * NB, NB, Ar, lda, nb, nb, nc, Ar, lda, 1.0, 0.0);
            const int64_t rin = Kend - (k0 + nb);
            if (rin > 0)                                          // rows below the panel inside the outer block
                hyp_gemm_tn(ctx, A + k0 + (k0 + nb) * lda, lda, Ar, lda, nb, rin, nc, A + (k0 + nb) + c0 * lda, lda, -1.0, 1.0);
        }
    };
    CUDA_TRY(cudaEventRecord(ctx->ev_bulk[0], bulk));
    CUDA_TRY(cudaStreamWaitEvent(chain, ctx->ev_bulk[0], 0));
    int b = 0;
    for (int64_t K0 = 0; K0 < m; K0 += OB, b++) {
        const int64_t Kend = std::min(K0 + OB, m);
        const int64_t rest = m - Kend, kd = Kend - K0;
        // ---- chain: the diagonal block ----
        on(chain, 0);
        for (int64_t k0 = K0; k0 < Kend && probe != 3; k0 += NB) {
            const int64_t nb = std::min<int64_t>(NB, m - k0);
            panel_kernel<true><<<1, PT, NB * LDU * 8, chain>>>(A, lda, m, k0 / NB, d_dinv, d_info);
            ctx->launches++;
            const int64_t rin = Kend - (k0 + nb);
            if (rin > 0) {
                double* A12 = A + k0 + (k0 + nb) * lda;
                ctx->small_tiles = narrow_chain && (nmask & 1);
                hyp_gemm_tn(ctx, d_dinv + (k0 / NB) * NB * NB, NB, A12, lda, nb, nb, rin, A12, lda, 1.0, 0.0);
                ctx->small_tiles = narrow_chain && (nmask & 2);
                hyp_atb_upper(ctx, A12, lda, A12, lda, nb, rin, A + (k0 + nb) + (k0 + nb) * lda, lda, -1.0, 1.0);
            }
        }
        CUDA_TRY(cudaEventRecord(ctx->ev_chain[b & 1], chain));
        if (dbg_sync & 8) CUDA_TRY(cudaDeviceSynchronize());
        if (rest <= 0) break;
        const int64_t near = std::min(OB, rest);
        double* P = A + K0 + Kend * lda;                           // the block row right of the block (kd x rest)
        double* T = A + Kend + Kend * lda;                         // the trailing matrix
        // ---- chain: block row over the next block's columns (they carry block b - 1's update from SYRK part a) ----
        if (b >= 1) CUDA_TRY(cudaStreamWaitEvent(chain, ctx->ev_bulk[(b - 1) & 1], 0));
        ctx->small_tiles = narrow_chain && (nmask & 4);
        if (probe != 3) block_row(K0, Kend, Kend, near);
        static const bool near_event_early = getenv("HYP_POTRF_NEAR_EVENT") && !strcmp(getenv("HYP_POTRF_NEAR_EVENT"), "early");
        if (near_event_early) CUDA_TRY(cudaEventRecord(ctx->ev_near[b & 1], chain));
        ctx->small_tiles = narrow_chain && (nmask & 8);
        if (probe != 3) hyp_atb_upper(ctx, P, lda, P, lda, kd, near, T, lda, -1.0, 1.0);
        // the bulk stream's slicing kernel starts only after the update of the next diagonal block: with the event in
        // front of it the two overlapped, and repeated factorisations of one matrix stopped being bit-identical
        // (tools/potrf_race.py, profiles/r02_potrf_race_*.jsonl)
        if (!near_event_early) CUDA_TRY(cudaEventRecord(ctx->ev_near[b & 1], chain));
        if (dbg_sync & 2) CUDA_TRY(cudaDeviceSynchronize());
        // ---- bulk: far block row, slicing, digit-sliced trailing update ----
        on(bulk, cap);
        CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_chain[b & 1], 0));
        if (rest > near && probe != 2) {
            if (probe != 5) block_row(K0, Kend, Kend + near, rest - near);        // 4: far block row only, 5: slicing only
            CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_near[b & 1], 0));
            if (probe != 4) hyp_ozaki_slice_short(ctx, P, lda, kd, rest, ctx->d_chol_digits, ldd, sstride, ctx->d_chol_dscale);
            if (probe != 1 && probe != 4 && probe != 5)
                hyp_ozaki_syrk_rows(ctx, ctx->d_chol_digits, ldd, sstride, ctx->d_chol_dscale, kd, rest, T, lda, -1.0, 1.0, 0, pa_rows,
                                    (int)(near / NB));
        }
        CUDA_TRY(cudaEventRecord(ctx->ev_bulk[b & 1], bulk));
        if (dbg_sync & 4) CUDA_TRY(cudaDeviceSynchronize());
        if (rest > near && probe != 2 && probe != 1 && probe != 4 && probe != 5)
            hyp_ozaki_syrk_rows(ctx, ctx->d_chol_digits, ldd, sstride, ctx->d_chol_dscale, kd, rest, T, lda, -1.0, 1.0, pa_rows, -1, 0);
        if (dbg_sync & 1) CUDA_TRY(cudaDeviceSynchronize());
    }
    // the caller's stream continues after the last diagonal block
    CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_chain[b & 1], 0));
    on(nullptr, 0);
    ctx->small_tiles = false;
    CUDA_TRY(cudaGetLastError());
    return true;
}

void hyp_potrf_upper(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info) {
    if (m <= 0) return;
    // default for large matrices: tcgen05 trailing updates (potrf_upper_i8 above); otherwise, and with HYP_POTRF=dag, the
    // task-graph kernel (chol_dag.cu), one launch for the whole factorisation; HYP_POTRF=stream keeps the two-stream
    // FP64-DMMA version below (also the fallback for unaligned input)
    {
        const char* pv = getenv("HYP_POTRF");
        const bool want_stream = pv && !strcmp(pv, "stream");
        const bool want_dag = pv && !strcmp(pv, "dag");
        if (!want_stream && !want_dag && potrf_upper_i8(ctx, A, lda, m, d_dinv, d_info)) return;
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
    int epoch = ++ctx->trsv_epoch;
    set_panel_attr();
    if (trsv_seg_len() > 0) {
        // flags: ticket, nblk block flags, nblk segment counters (see hyp_load_model / hyp_test_potrs for the size)
        const TrsvLists& L = trsv_lists(ctx, nblk, trans);
        double* part = trsv_part(ctx, nblk, L.maxseg);
        CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, (size_t)(1 + 2 * nblk) * sizeof(int), ctx->stream));
        const int grid = std::min(L.ntasks, ctx->sm_count);
        if (trans)
            trsv_seg_kernel<true><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, L.d_tasks, L.ntasks, nblk,
                                                                         epoch, part, L.maxseg);
        else
            trsv_seg_kernel<false><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, L.d_tasks, L.ntasks, nblk,
                                                                          epoch, part, L.maxseg);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        return;
    }
    unsigned long long* pkt = trsv_use_pkt() ? trsv_pkt(ctx, nblk, 1) : nullptr;
    CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int grid = std::min(nblk, ctx->sm_count);
    if (pkt && trans)
        trsv_pkt_kernel<true><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, pkt, nblk, epoch);
    else if (pkt)
        trsv_pkt_kernel<false><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, pkt, nblk, epoch);
    else if (trans)
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
    int epoch = ++ctx->trsv_epoch;
    set_panel_attr();
    if (trsv_seg_len() > 0) {
        const TrsvLists& L = trsv_lists(ctx, nblk, trans);
        double* part = trsv_part(ctx, nblk, L.maxseg);
        CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, (size_t)(1 + 2 * nblk) * sizeof(int), ctx->stream));
        const int grid = std::min(L.ntasks, ctx->sm_count);
        if (trans)
            trsv_seg_kernel<true, 2><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, L.d_tasks, L.ntasks,
                                                                            nblk, epoch, part, L.maxseg, xstride);
        else
            trsv_seg_kernel<false, 2><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, L.d_tasks, L.ntasks,
                                                                             nblk, epoch, part, L.maxseg, xstride);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        return;
    }
    unsigned long long* pkt = trsv_use_pkt() ? trsv_pkt(ctx, nblk, 2) : nullptr;
    CUDA_TRY(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream));
    int grid = std::min(nblk, ctx->sm_count);
    if (pkt && trans)
        trsv_pkt_kernel<true, 2><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, pkt, nblk, epoch, xstride);
    else if (pkt)
        trsv_pkt_kernel<false, 2><<<grid, 256, TRSV_SMEM, ctx->stream>>>(F, ldf, m, d_dinv, x, ctx->d_flags, pkt, nblk, epoch, xstride);
    else if (trans)
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

void hyp_trsv_set_pkt(int on) { g_trsv_pkt = on ? 1 : 0; }

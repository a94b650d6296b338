// Factor-and-invert kernel of one diagonal block (at most 128 x 128) of an upper Cholesky factorisation: the
// latency-critical piece of the blocked dpotrf (K3: dense.jl:191-192 via qrchol.jl:249-250) and the per-cone
// cholesky! + inv_fact! of the matrix cones and of the generic Hessian factorisation (K9: Cones.jl:239-259).
// Launched from chol.cu (panel_kernel inside hyp_potrf_upper / hyp_trtri_diag, chol_batched_kernel for cone groups);
// compiled for the host by tests/emu/ (CPU-tier tests: tests/test_emu_chol.py).
#pragma once
#include "devdefs.cuh"

namespace hypdev {

constexpr int NB = 128;
constexpr int LDU = NB + 1;   // padded leading dimension of the shared-memory block

// Factor (FACTOR) and invert one upper-triangular diagonal block of at most 128 x 128, by one CTA
// of 256 threads, entirely in shared memory (sA: 128 x 128, ld LDU; the upper triangle holds
// A -> U, the strictly lower triangle receives X' where X = U^-1, diagX the diagonal of X).
// Blocked right-looking algorithm with 32 x 32 sub-blocks:
//   (a) warp 0 factors and inverts the diagonal sub-block in REGISTERS (lane = column, pivots
//       broadcast with shuffles: no shared-memory round trips or block barriers on the pivot chain),
//   (b) all warps form the block row U_bc = X_bb' A_bc and (c) apply the rank-32 trailing update,
//       both as 4 x 4 register-tiled products;
//   (d) afterwards the off-diagonal block columns of X follow from
//       X[0:c, c] = -X[0:c, 0:c] U[0:c, c] X[c, c].
// Ab: the block in global memory (upper part is read; U is written back, optionally with zeros
// below the diagonal); Db receives X (dn x dn entries, leading dim ldd; rows / columns past nb are
// identity padding).  Returns the 1-based index of the first non-positive pivot or 0.
constexpr int SB = 32;
constexpr int PT = 256;          // threads of a panel CTA

// phase timestamps of the LAST panel_body call of a launch (clock64 of thread 0; read back by hyp_test_panel_clocks):
// [0] start, [1] tile loaded, then per 32-column sub-block b: [2 + 3 b] diagonal sub-block factored + inverted,
// [3 + 3 b] block row, [4 + 3 b] trailing update; [14] inverse assembled, [15] results stored
#ifndef HYP_EMU
static __device__ long long g_panel_clk[16];
static __device__ int g_panel_flags;      // experiments (tools/panel_probe.py): bit 0 = no early stores in A_3
#define HYP_PANEL_CLK(i) do { if (threadIdx.x == 0) g_panel_clk[i] = clock64(); } while (0)
#else
#define HYP_PANEL_CLK(i) do { } while (0)
#endif
constexpr int LDX = SB + 1;
constexpr int LDT = 3 * SB + 1;  // sT holds up to 96 x 32

template <bool FACTOR>
__device__ __forceinline__ void base_block(double* sA, double* diagX, double* sX, int o, int lane,
                                           int* s_bad) {
    // Lane c holds column c of the 32 x 32 block (col) and column c of W (wcol), where W starts as
    // the identity and receives the same row operations as the factorisation, so that it ends as
    // U^-T = X': the inverse comes out of the pivot loop instead of a second substitution sweep.
    const unsigned FULL = 0xffffffffu;
    double col[SB], wcol[SB];
#pragma unroll
    for (int r = 0; r < SB; r++) {
        col[r] = sA[(o + r) + (o + lane) * LDU];
        wcol[r] = (r == lane) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int j = 0; j < SB; j++) {
        const double d = __shfl_sync(FULL, col[j], j);
        double rinv = 0.0;
        if (FACTOR) {
            if (d > 0.0) {
                rinv = rsqrt(d);
                // one Newton step on the reciprocal square root: keeps sd * sd = d to the last bits
                rinv = rinv + 0.5 * rinv * (1.0 - d * rinv * rinv);
            } else if (lane == 0 && *s_bad == 0) {
                *s_bad = o + j + 1;
            }
        } else {
            rinv = (d != 0.0) ? 1.0 / d : 0.0;
        }
        double u;
        if (FACTOR) {
            u = (lane > j) ? col[j] * rinv : 0.0;
            if (lane == j) col[j] = d * rinv;
            else if (lane > j) col[j] = u;
        } else {
            u = (lane > j) ? col[j] : 0.0;     // U is given: row j of U, not rescaled
        }
        const double wj = wcol[j] * rinv;      // row j of W (nonzero for lanes <= j)
        wcol[j] = wj;
        if (lane == j) diagX[o + j] = rinv;
#pragma unroll
        for (int r = j + 1; r < SB; r++) {
            const double ur = __shfl_sync(FULL, u, r);     // U[j, r]
            if (FACTOR) col[r] = fma(-ur, u, col[r]);
            wcol[r] = fma(-ur, wj, wcol[r]);
        }
    }
    if (FACTOR) {
#pragma unroll
        for (int r = 0; r < SB; r++)
            if (r <= lane) sA[(o + r) + (o + lane) * LDU] = col[r];
    }
    // lane k holds row k of X: X[k, i] = W[i, k] = wcol[i] (i >= k).  X' goes below the diagonal
    // of the block, and a dense copy of X' (sX[i + k * LDX] = X[k, i]) serves the block row.
#pragma unroll
    for (int i = 0; i < SB; i++) {
        if (i > lane) sA[(o + i) + (o + lane) * LDU] = wcol[i];
        sX[i + lane * LDX] = (i >= lane) ? wcol[i] : 0.0;
    }
    __syncwarp();
}

// BARID = 0: the CTA is exactly the PT threads of the panel (block barrier).  BARID > 0: the PT panel threads are a
// subset of a larger CTA (the task-graph Cholesky of chol_dag.cu keeps a TMA producer warp beside them) and
// synchronise on named barrier BARID.  LDCG: read the block through L2 (another CTA of the same launch wrote it).
template <int BARID>
__device__ __forceinline__ void panel_sync() {
#ifdef HYP_EMU
    __syncthreads();
#else
    if (BARID == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BARID), "n"(PT) : "memory");
#endif
}

template <bool FACTOR, int BARID = 0, bool LDCG = false>
__device__ __forceinline__ int panel_body(double* __restrict__ Ab, int64_t lda, int nb,
                                          double* __restrict__ Db, int ldd, int dn, bool zero_lower,
                                          double* sA, double* diagX, double* sX, double* sT, int* s_bad) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) *s_bad = 0;
    HYP_PANEL_CLK(0);
    // 64 elements per thread, 16 global loads in flight at a time
    for (int base = 0; base < NB * NB; base += PT * 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int idx = base + u * PT + tid;
            int r = idx & (NB - 1), c = idx >> 7;
            double x = 0.0;
            if (r < nb && c < nb) {
                if (r <= c) {
#ifdef HYP_EMU
                    x = Ab[r + (int64_t)c * lda];
#else
                    x = LDCG ? __ldcg(Ab + r + (int64_t)c * lda) : Ab[r + (int64_t)c * lda];
#endif
                }
            } else if (r == c) {
                x = 1.0;
            }
            v[u] = x;
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int idx = base + u * PT + tid;
            sA[(idx & (NB - 1)) + (idx >> 7) * LDU] = v[u];
        }
    }
    panel_sync<BARID>();
    HYP_PANEL_CLK(1);

    for (int b = 0; b < NB / SB; b++) {
        const int o = b * SB;
        if (warp == 0) base_block<FACTOR>(sA, diagX, sX, o, lane, s_bad);
        panel_sync<BARID>();
        HYP_PANEL_CLK(2 + 3 * b);
        if (FACTOR && b < NB / SB - 1) {
            const int t0 = o + SB;
            const int ncol = NB - t0;
            // (b) block row U[o + i, c] = sum_k X[k, i] A[o + k, c]: 4 x 4 tiles, 8 row tiles
            {
                const int ti = tid & 7, tj = tid >> 3;          // tj: 0..31 column tiles
                const bool act = tj * 4 < ncol;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
                if (act) {
                    const double* xp = sX + ti * 4;
                    const double* ap = sA + o + (t0 + tj * 4) * LDU;
#pragma unroll 4
                    for (int k = 0; k < SB; k++) {
                        double xa[4], av[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) xa[a] = xp[a + k * LDX];
#pragma unroll
                        for (int q = 0; q < 4; q++) av[q] = ap[k + q * LDU];
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[a][q] = fma(xa[a], av[q], acc[a][q]);
                    }
                }
                panel_sync<BARID>();
                if (act) {
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) sA[(o + ti * 4 + a) + (t0 + tj * 4 + q) * LDU] = acc[a][q];
                }
                panel_sync<BARID>();
                HYP_PANEL_CLK(3 + 3 * b);
            }
            // (c) trailing update A[r, c] -= sum_k U[o + k, r] U[o + k, c] on 4 x 4 tiles with r-tile <= c-tile
            {
                const int nt4 = ncol / 4;
                const int ntiles = nt4 * (nt4 + 1) / 2;
                for (int t = tid; t < ntiles; t += PT) {
                    // tile (tr, tc) with tr <= tc from the linear index t = tc (tc + 1) / 2 + tr
                    int tc = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                    while ((tc + 1) * (tc + 2) / 2 <= t) tc++;
                    while (tc * (tc + 1) / 2 > t) tc--;
                    const int tr = t - tc * (tc + 1) / 2;
                    const double* ur = sA + o + (t0 + tr * 4) * LDU;
                    const double* uc = sA + o + (t0 + tc * 4) * LDU;
                    double acc[4][4];
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
#pragma unroll 4
                    for (int k = 0; k < SB; k++) {
                        double rv[4], cv[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) rv[a] = ur[k + a * LDU];
#pragma unroll
                        for (int q = 0; q < 4; q++) cv[q] = uc[k + q * LDU];
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[a][q] = fma(rv[a], cv[q], acc[a][q]);
                    }
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int r = t0 + tr * 4 + a, c = t0 + tc * 4 + q;
                            if (r <= c) sA[r + c * LDU] -= acc[a][q];
                        }
                }
                panel_sync<BARID>();
                HYP_PANEL_CLK(4 + 3 * b);
            }
        }
    }

    // (d) off-diagonal block columns of X.  For cb = 1..3 with h = 32 cb:
    //   T = X[0:h, 0:h] U[0:h, h:h+32]   (X upper triangular: X[i, k] for k > i sits at sA[k + i LDU])
    //   X[0:h, h:h+32] = -T X[cb, cb]
    for (int cb = 1; cb < NB / SB; cb++) {
        const int h = cb * SB;
        // dense copy of X[cb, cb] (upper triangular): sX[k + j * LDX] = X[h + k, h + j]
        for (int idx = tid; idx < SB * SB; idx += PT) {
            int k = idx & 31, j = idx >> 5;
            double v = 0.0;
            if (k < j) v = sA[(h + j) + (h + k) * LDU];
            else if (k == j) v = diagX[h + j];
            sX[k + j * LDX] = v;
        }
        // T tiles: rows i0..i0+3 (i0 = 4 ti), cols 4 tj..; h/4 x 8 tiles
        const int ntile = (h / 4) * 8;
        for (int t = tid; t < ntile; t += PT) {
            const int ti = t % (h / 4), tj = t / (h / 4);
            const int i0 = ti * 4;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
            const double* up = sA + (h + tj * 4) * LDU;     // up[k + q * LDU] = U[k, h + 4 tj + q]
            // diagonal 4 x 4 piece: k in i0 .. i0 + 3 with the triangular structure
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int k = i0 + kk;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    double xv = 0.0;
                    if (kk > a) xv = sA[k + (i0 + a) * LDU];
                    else if (kk == a) xv = diagX[k];
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[a][q] = fma(xv, up[k + q * LDU], acc[a][q]);
                }
            }
            for (int k = i0 + 4; k < h; k++) {
                double xv[4], uv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) xv[a] = sA[k + (i0 + a) * LDU];
#pragma unroll
                for (int q = 0; q < 4; q++) uv[q] = up[k + q * LDU];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[a][q] = fma(xv[a], uv[q], acc[a][q]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) sT[(i0 + a) + (tj * 4 + q) * LDT] = acc[a][q];
        }
        panel_sync<BARID>();
        // X[i, h + j] = -sum_{k <= j} T[i, k] X[cb, cb][k, j]; stored at sA[(h + j) + i * LDU]
        for (int t = tid; t < ntile; t += PT) {
            const int ti = t % (h / 4), tj = t / (h / 4);
            const int i0 = ti * 4, j0 = tj * 4;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
            for (int k = 0; k < j0 + 4; k++) {
                double tv[4], xv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) tv[a] = sT[(i0 + a) + k * LDT];
#pragma unroll
                for (int q = 0; q < 4; q++) xv[q] = sX[k + (j0 + q) * LDX];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[a][q] = fma(tv[a], xv[q], acc[a][q]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) sA[(h + j0 + q) + (i0 + a) * LDU] = -acc[a][q];
        }
        panel_sync<BARID>();
    }

    HYP_PANEL_CLK(14);
    if (FACTOR) {
        for (int idx = tid; idx < NB * NB; idx += PT) {
            int r = idx & (NB - 1), c = idx >> 7;
            if (r < nb && c < nb) {
                if (r <= c) Ab[r + (int64_t)c * lda] = sA[r + c * LDU];
                else if (zero_lower) Ab[r + (int64_t)c * lda] = 0.0;
            }
        }
    }
    for (int idx = tid; idx < dn * dn; idx += PT) {
        int r = idx % dn, c = idx / dn;
        double x = 0.0;
        if (r < c) x = sA[c + r * LDU];
        else if (r == c) x = diagX[r];
        Db[r + (int64_t)c * ldd] = x;
    }
    HYP_PANEL_CLK(15);
    return *s_bad;
}

// ---- software-pipelined factor-and-invert (the latency-critical PANEL of the blocked Cholesky) -------------------------
// Same arithmetic as panel_body<true> (same base_block, same 4 x 4 register-tiled products, same results up to the
// order of independent updates), re-scheduled so that ONLY the chain of 128 dependent pivots (warp 0: 4 diagonal
// sub-blocks of 32) and the short hand-offs between them are on the critical path.  Measured phase costs of the
// bulk-synchronous version (clock64, tools/panel_probe.py, 147.5 k cycles): load 9.3 k, 4 x diag 13.2 k, block rows
// 11.1 k, trailing updates 23.2 k, inverse assembly 42.3 k, stores 8.5 k.  Here, between two block barriers:
//   A_b: warp 0 factors + inverts diagonal sub-block b in registers  ||  warps 1-7: the work nobody waits for yet -
//        loading the rest of the tile (b = 0), the trailing update of row b - 1 except the next diagonal sub-block,
//        T_b = X[0:h, 0:h] U[0:h, b] (the expensive half of the inverse's block column b), and, for b = 3, the global
//        stores of everything that is already final;
//   B_b: all warps: X[0:h, b] = -T_b X_bb, block row U[b, b+1:], barrier, update of the next diagonal sub-block.
// sT must hold 96 x 32 (ld LDT), sX 32 x 32 (ld LDX).
template <int BARID, bool LDCG>
__device__ __forceinline__ double panel_ld(const double* p) {
#ifdef HYP_EMU
    return *p;
#else
    return LDCG ? __ldcg(p) : *p;
#endif
}

// Register tiles are 4 x 4 with INTERLEAVED indices: tile t of a 32-wide block owns the indices t, t + 8, t + 16, t + 24
// of that block.  Threads of a warp then read consecutive columns of the (odd-pitch) shared-memory block - conflict-free -
// where contiguous 4 x 4 tiles put them 4 columns = 8 banks apart (2- to 4-way conflicts; measured: the 4 x 4 products
// of the bulk-synchronous version ran at ~ 200 cycles per k-step instead of ~ 64).
__device__ __forceinline__ int il_idx(int t8, int a) { return t8 + 8 * a; }

template <int BARID = 0, bool LDCG = false>
__device__ __forceinline__ int panel_body_fast(double* __restrict__ Ab, int64_t lda, int nb, double* __restrict__ Db,
                                               int ldd, int dn, bool zero_lower, double* sA, double* diagX, double* sX,
                                               double* sT, int* s_bad) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PT - 32;                       // worker threads beside the pivot warp
    // the pivot warp is the LAST warp: the SM's issue arbiter prefers the highest warp id, and the chain of dependent
    // pivots must not queue behind the workers' FMAs
    constexpr int PW = PT / 32 - 1;
    if (tid == 0) *s_bad = 0;
    HYP_PANEL_CLK(0);
    // ---- I0: the first diagonal sub-block ----
    for (int idx = tid; idx < SB * SB; idx += PT) {
        const int r = idx & (SB - 1), c = idx >> 5;
        double x = 0.0;
        if (r < nb && c < nb) {
            if (r <= c) x = panel_ld<BARID, LDCG>(Ab + r + (int64_t)c * lda);
        } else if (r == c) {
            x = 1.0;
        }
        sA[r + c * LDU] = x;
    }
    panel_sync<BARID>();
    HYP_PANEL_CLK(1);

    // one copy of the loop body: unrolled over b the kernel grows to 240 KB of code and every phase runs on a cold
    // instruction cache (measured: the 192-tile X column of b = 3 took 34 k cycles, 10 x its arithmetic)
#pragma unroll 1
    for (int b = 0; b < NB / SB; b++) {
        const int o = b * SB;           // this diagonal sub-block
        const int h = o;                // rows above it
        // ================= A_b =================
        if (warp == PW) {
            base_block<true>(sA, diagX, sX, o, lane, s_bad);
        } else {
            const int wt = tid;
            if (b == 0) {
                // rest of the tile: everything but the (0, 0) sub-block; 16 loads in flight per thread
                for (int base = 0; base < NB * NB; base += NW * 16) {
                    double v[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int idx = base + u * NW + wt;
                        const int r = idx & (NB - 1), c = idx >> 7;
                        double x = 0.0;
                        if (idx < NB * NB && !(r < SB && c < SB)) {
                            if (r < nb && c < nb) {
                                if (r <= c) x = panel_ld<BARID, LDCG>(Ab + r + (int64_t)c * lda);
                            } else if (r == c) {
                                x = 1.0;
                            }
                        }
                        v[u] = x;
                    }
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        const int idx = base + u * NW + wt;
                        const int r = idx & (NB - 1), c = idx >> 7;
                        if (idx < NB * NB && !(r < SB && c < SB)) sA[r + c * LDU] = v[u];
                    }
                }
            } else {
                // (1) trailing update of row b - 1 on everything right / below the sub-block (b, b), which the
                //     previous B interval already updated:  A[r, c] -= sum_k U[op + k, r] U[op + k, c]
                //     32 x 32 blocks (br <= bc) of the trailing region, 8 x 8 interleaved tiles each
                const int op = o - SB;
                const int nblk = (NB - o) / SB;
                const int npair = nblk * (nblk + 1) / 2;
                for (int t = wt; t < npair * 64; t += NW) {
                    const int pr = t >> 6;
                    if (pr == 0) continue;                               // block (0, 0) = sub-block (b, b): done
                    // pairs in the order (0,0) (0,1) (1,1) (0,2) (1,2) (2,2)
                    const int bc = pr >= 3 ? 2 : (pr >= 1 ? 1 : 0);
                    const int br = pr - bc * (bc + 1) / 2;
                    const int tr8 = t & 7, tc8 = (t >> 3) & 7;
                    const double* ur = sA + op + (o + 32 * br + tr8) * LDU;
                    const double* uc = sA + op + (o + 32 * bc + tc8) * LDU;
                    double acc[4][4];
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
#pragma unroll 2
                    for (int k = 0; k < SB; k++) {
                        double rv[4], cv[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) rv[a] = ur[k + 8 * a * LDU];
#pragma unroll
                        for (int q = 0; q < 4; q++) cv[q] = uc[k + 8 * q * LDU];
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[a][q] = fma(rv[a], cv[q], acc[a][q]);
                    }
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int r = o + 32 * br + il_idx(tr8, a), c = o + 32 * bc + il_idx(tc8, q);
                            if (r <= c) sA[r + c * LDU] -= acc[a][q];
                        }
                }
                // (2) T_b = X[0:h, 0:h] U[0:h, o:o+32]   (X[i, k] for k > i sits at sA[k + i LDU], diagonal in diagX)
                for (int t = wt; t < b * 64; t += NW) {
                    const int ib = t >> 6, tr8 = t & 7, tc8 = (t >> 3) & 7;
                    const int ibase = 32 * ib + tr8;
                    double acc[4][4];
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
                    const double* up = sA + (o + tc8) * LDU;         // up[k + 8 q LDU] = U[k, o + tc8 + 8 q]
                    const double* xp = sA + ibase * LDU;             // xp[k + 8 a LDU] = X[ibase + 8 a, k] for k > row
                    // k runs over whole 32-blocks, the same k in every thread of the warp (conflict-free operand loads):
                    // inside the tile's own block the entries left of the diagonal are masked (the upper triangle of sA
                    // holds U there), right of it X is dense
                    const int kb0 = 32 * ib;
#pragma unroll 2
                    for (int k = kb0; k < kb0 + 32; k++) {
                        double xv[4], uv[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) {
                            const int row = ibase + 8 * a;
                            const double off = xp[k + 8 * a * LDU];
                            xv[a] = k > row ? off : (k == row ? diagX[k] : 0.0);
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) uv[q] = up[k + 8 * q * LDU];
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[a][q] = fma(xv[a], uv[q], acc[a][q]);
                    }
#pragma unroll 2
                    for (int k = kb0 + 32; k < h; k++) {
                        double xv[4], uv[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) xv[a] = xp[k + 8 * a * LDU];
#pragma unroll
                        for (int q = 0; q < 4; q++) uv[q] = up[k + 8 * q * LDU];
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[a][q] = fma(xv[a], uv[q], acc[a][q]);
                    }
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) sT[(ibase + 8 * a) + (tc8 + 8 * q) * LDT] = acc[a][q];
                }
#ifndef HYP_EMU
                const bool early = !(g_panel_flags & 1);
#else
                const bool early = true;
#endif
                if (b == NB / SB - 1 && early) {
                    // (3) everything above / left of the last sub-block is final: store it while warp 0 finishes the chain
                    // four independent shared-memory loads in flight per thread, then the four stores
                    const int rmax = o < nb ? o : nb;
                    for (int c = warp; c < nb; c += PW) {
                        double v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int r = lane + 32 * u;
                            v[u] = (r <= c && r < rmax) ? sA[r + c * LDU] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int r = lane + 32 * u;
                            if (r < rmax) {
                                if (r <= c) Ab[r + (int64_t)c * lda] = v[u];
                                else if (zero_lower) Ab[r + (int64_t)c * lda] = 0.0;
                            } else if (zero_lower && c < o && r < nb) {
                                Ab[r + (int64_t)c * lda] = 0.0;            // below the diagonal, left of the last sub-block
                            }
                        }
                    }
                    const int cmax = o < dn ? o : dn;
                    for (int c = warp; c < cmax; c += PW) {             // X columns 0 .. o-1
                        double v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int r = lane + 32 * u;
                            v[u] = r < c ? sA[c + r * LDU] : (r == c ? diagX[r] : 0.0);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int r = lane + 32 * u;
                            if (r < dn) Db[r + (int64_t)c * ldd] = v[u];
                        }
                    }
                }
            }
        }
        panel_sync<BARID>();
        HYP_PANEL_CLK(2 + 3 * b);
        // ================= B_b =================
        // X[0:h, o + j] = -sum_{k <= j} T[i, k] X_bb[k, j]   (X_bb[k, j] = sX[j + k LDX], zero for k > j);
        // stored at sA[(o + j) + i LDU]
        for (int t = tid; t < b * 64; t += PT) {
            const int ib = t >> 6, tr8 = t & 7, tc8 = (t >> 3) & 7;
            const int ibase = 32 * ib + tr8;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
            const int kend = tc8 + 24;                                   // largest column of this tile
#pragma unroll 2
            for (int k = 0; k <= kend; k++) {
                double tv[4], xv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) tv[a] = sT[(ibase + 8 * a) + k * LDT];
#pragma unroll
                for (int q = 0; q < 4; q++) xv[q] = sX[(tc8 + 8 * q) + k * LDX];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[a][q] = fma(tv[a], xv[q], acc[a][q]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) sA[(o + tc8 + 8 * q) + (ibase + 8 * a) * LDU] = -acc[a][q];
        }
        if (b == NB / SB - 1) HYP_PANEL_CLK(12);       // thread 0's own share of the last X column
        if (b < NB / SB - 1) {
            const int t0 = o + SB;
            const int ncol = NB - t0;
            // block row U[o + i, c] = sum_k X_bb[k, i] A[o + k, c]: 8 row tiles x ncol / 4 column tiles
            {
                const int tr8 = tid & 7, tj = tid >> 3;
                const bool act = tj * 4 < ncol;
                const int cbase = t0 + 32 * (tj >> 3) + (tj & 7);
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
                if (act) {
                    const double* xp = sX + tr8;
                    const double* ap = sA + o + cbase * LDU;
#pragma unroll 2
                    for (int k = 0; k < SB; k++) {
                        double xa[4], av[4];
#pragma unroll
                        for (int a = 0; a < 4; a++) xa[a] = xp[8 * a + k * LDX];
#pragma unroll
                        for (int q = 0; q < 4; q++) av[q] = ap[k + 8 * q * LDU];
#pragma unroll
                        for (int a = 0; a < 4; a++)
#pragma unroll
                            for (int q = 0; q < 4; q++) acc[a][q] = fma(xa[a], av[q], acc[a][q]);
                    }
                }
                panel_sync<BARID>();
                if (act) {
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int q = 0; q < 4; q++) sA[(o + tr8 + 8 * a) + (cbase + 8 * q) * LDU] = acc[a][q];
                }
                panel_sync<BARID>();
            }
            HYP_PANEL_CLK(3 + 3 * b);
            // the next diagonal sub-block: A[r, c] -= sum_k U[o + k, r] U[o + k, c], r, c in [t0, t0 + 32): all 256
            // threads, 2 x 2 interleaved tiles (rows tr, tr + 16; columns tc, tc + 16)
            {
                const int tr = tid & 15, tc = tid >> 4;
                const double* ur = sA + o + (t0 + tr) * LDU;
                const double* uc = sA + o + (t0 + tc) * LDU;
                double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
#pragma unroll 8
                for (int k = 0; k < SB; k++) {
                    const double r0 = ur[k], r1 = ur[k + 16 * LDU], c0 = uc[k], c1 = uc[k + 16 * LDU];
                    a00 = fma(r0, c0, a00);
                    a01 = fma(r0, c1, a01);
                    a10 = fma(r1, c0, a10);
                    a11 = fma(r1, c1, a11);
                }
                const int r_0 = t0 + tr, r_1 = r_0 + 16, c_0 = t0 + tc, c_1 = c_0 + 16;
                if (r_0 <= c_0) sA[r_0 + c_0 * LDU] -= a00;
                if (r_0 <= c_1) sA[r_0 + c_1 * LDU] -= a01;
                if (r_1 <= c_0) sA[r_1 + c_0 * LDU] -= a10;
                if (r_1 <= c_1) sA[r_1 + c_1 * LDU] -= a11;
            }
            panel_sync<BARID>();
            HYP_PANEL_CLK(4 + 3 * b);
        } else {
            panel_sync<BARID>();
        }
    }
    HYP_PANEL_CLK(14);
    // ---- what is left: U rows 96 .. 127 and the last block column of X ----
    {
#ifndef HYP_EMU
        const int o = (g_panel_flags & 1) ? 0 : NB - SB;
#else
        const int o = NB - SB;
#endif
        for (int c = warp; c < nb; c += PT / 32) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int r = lane + 32 * u;
                v[u] = (r >= o && r <= c) ? sA[r + c * LDU] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int r = lane + 32 * u;
                if (r >= o && r < nb) {
                    if (r <= c) Ab[r + (int64_t)c * lda] = v[u];
                    else if (zero_lower && c >= o) Ab[r + (int64_t)c * lda] = 0.0;
                }
            }
        }
        for (int c = o + warp; c < dn; c += PT / 32) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int r = lane + 32 * u;
                v[u] = r < c ? sA[c + r * LDU] : (r == c ? diagX[r] : 0.0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int r = lane + 32 * u;
                if (r < dn) Db[r + (int64_t)c * ldd] = v[u];
            }
        }
    }
    HYP_PANEL_CLK(15);
    return *s_bad;
}

//   FACTOR:  grid = 1; diagonal block blk0 of the m x m matrix A; a bad pivot sets *info.
//   !FACTOR: grid = number of diagonal blocks of the m x m triangular matrix A; inverts only.
template <bool FACTOR>
__global__ void __launch_bounds__(PT, 1)
panel_kernel(double* __restrict__ A, int64_t lda, int64_t m, int64_t blk0, double* __restrict__ dinv,
             int* __restrict__ info) {
    HYP_DYN_SMEM(double, sU);                 // NB x NB col-major, ld = LDU
    __shared__ double diagX[NB];
    __shared__ double sX[SB * LDX];
    __shared__ double sT[SB * LDT];
    __shared__ int s_bad;
    const int64_t blk = FACTOR ? blk0 : (int64_t)blockIdx.x;
    const int64_t k0 = blk * NB;
    const int nb = (int)((int64_t)NB < m - k0 ? (int64_t)NB : m - k0);
    int bad;
    if (FACTOR)
        // the block is read through L2 (LDCG): it was written by kernels of the same stream that ran on OTHER SMs, while
        // kernels of the second stream keep every SM busy - see the note on L1 in syrk.cu (epilogue of atb_upper_kernel)
        bad = panel_body_fast<0, true>(A + k0 + k0 * lda, lda, nb, dinv + blk * (int64_t)NB * NB, NB, NB, false, sU, diagX,
                                       sX, sT, &s_bad);
    else
        bad = panel_body<false>(A + k0 + k0 * lda, lda, nb, dinv + blk * (int64_t)NB * NB, NB, NB, false, sU, diagX, sX,
                                sT, &s_bad);
    if (FACTOR && bad && threadIdx.x == 0) atomicCAS(info, 0, (int)(k0 + bad));
}

// Batched variant for the matrix cones: CTA c factors the side x side matrix at U + moff[c]
// (leading dim = side rounded up to even) in place and writes its inverse to Ui + moff[c].
// A failed factorisation clears flag[kidx[c]].  Cones with side > 128 are skipped (blocked path).
static __global__ void __launch_bounds__(PT, 1)
chol_batched_kernel(int ncones, const int* __restrict__ sides, const int64_t* __restrict__ moff,
                    const int* __restrict__ kidx, double* __restrict__ U, double* __restrict__ Ui,
                    uint8_t* __restrict__ flag) {
    HYP_DYN_SMEM(double, sU);
    __shared__ double diagX[NB];
    __shared__ double sX[SB * LDX];
    __shared__ double sT[SB * LDT];
    __shared__ int s_bad;
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int side = sides[c];
    if (side > NB) return;
    const int lde = (side + 1) & ~1;
    int bad = panel_body_fast<0, false>(U + moff[c], lde, side, Ui + moff[c], lde, side, true, sU, diagX, sX, sT, &s_bad);
    if (bad && threadIdx.x == 0) flag[kidx[c]] = 0;
}

#ifdef HYP_EMU
// host emulation (tests/emu/): blocks run one after the other, so the first CTA takes every ticket in dependency order
// and the flag protocol degenerates to plain loads / stores; the arithmetic of the solve is what gets checked
__device__ __forceinline__ int ld_acquire(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#endif

// ---- triangular solve with the blocked factor -------------------------------------------
// flags[0] = ticket counter (zeroed by the host before the launch), flags[1 + k] = epoch when
// block k of the solution is final.  CTA for block k: the inverted diagonal block goes to shared
// memory up front, and every off-diagonal tile is loaded into registers BEFORE the CTA waits for
// the solution block it multiplies, so the dependency chain only sees: flag -> 1 KB vector load ->
// 64 FMAs per thread -> reduction -> 128 x 128 matvec from shared memory -> publish.
constexpr int TRSV_SMEM = NB * NB * 8;

// NRHS right-hand sides share every tile of the factor (hyp_solve_system_multi): x_v = x + v * xstride.  Per right-hand
// side the arithmetic and its order are those of the single-vector solve.
// pull the 128 x 128 tile at `tile` (leading dimension ldf) into L2: 1024 lines of 128 B, four per thread of a 256-thread CTA.
// The sweeps read every tile exactly once, from DRAM; issued one tile ahead this turns the DRAM latency of the register
// loads into an L2 hit.
__device__ __forceinline__ void trsv_prefetch_tile(const double* tile, int64_t ldf, int64_t rows_left, int64_t cols_left) {
#ifndef HYP_EMU
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int L = (int)threadIdx.x + 256 * q;
        const int c = L >> 3, r = (L & 7) * 16;
        if (c < cols_left && r < rows_left) asm volatile("prefetch.global.L2 [%0];" ::"l"(tile + r + (int64_t)c * ldf));
    }
#else
    (void)tile; (void)ldf; (void)rows_left; (void)cols_left;
#endif
}

template <bool TRANS, int NRHS = 1>
__global__ void __launch_bounds__(256, 1)
trsv_kernel(const double* __restrict__ F, int64_t ldf, int64_t m, const double* __restrict__ dinv,
            double* x, int* flags, int nblk, int epoch, int64_t xstride = 0) {
    HYP_DYN_SMEM(double, sD);                // Dinv_k, 128 x 128 col-major
    __shared__ double sv[NRHS][2][NB];
    __shared__ double sacc[NRHS][2][NB];
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
            double pacc[NRHS][16];
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) pacc[v][i] = 0.0;
            const int64_t cw = c0 + warp * 16;
            // my own block of the right-hand side does not depend on the chain: fetched now, used after the last tile
            double xown[NRHS][16];
            if (lane == 0) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
#pragma unroll
                    for (int i = 0; i < 16; i++) xown[v][i] = (cw + i < m) ? x[v * xstride + cw + i] : 0.0;
            }
            for (int j = 0; j < k; j++) {
                double tl[16][4];
                const double* Ut = F + (int64_t)j * NB + cw * ldf;
                if (j + 1 < k) trsv_prefetch_tile(F + (int64_t)(j + 1) * NB + c0 * ldf, ldf, NB, m - c0);
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
                if (tid < NB) {
#pragma unroll
                    for (int v = 0; v < NRHS; v++) sv[v][j & 1][tid] = __ldcg(x + v * xstride + (int64_t)j * NB + tid);
                }
                __syncthreads();
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double* svj = sv[v][j & 1];
                    const double v0 = svj[lane], v1 = svj[lane + 32], v2 = svj[lane + 64], v3 = svj[lane + 96];
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        pacc[v][i] += tl[i][0] * v0 + tl[i][1] * v1 + tl[i][2] * v2 + tl[i][3] * v3;
                }
            }
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    double a = pacc[v][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    pacc[v][i] = a;
                }
            __syncthreads();
            if (lane == 0) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        int64_t c = cw + i;
                        sv[v][0][warp * 16 + i] = (c < m) ? (xown[v][i] - pacc[v][i]) : 0.0;
                    }
            }
            __syncthreads();
            // y[c] = sum_{r <= c} Dinv[r, c] v[r]
#pragma unroll
            for (int v = 0; v < NRHS; v++) {
                const double* vb = sv[v][0];
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
                        if (c < m) x[v * xstride + c] = res[i];
                    }
                }
            }
        } else {
            // x_k = Dinv_k (y_k - sum_{j>k} U[k-block, j-block] x_j); thread owns a row, the two
            // halves of the CTA split the 128 columns of a tile
            const int r = tid & (NB - 1), half = tid >> 7;
            const int64_t grow = c0 + r;
            double acc[NRHS];
#pragma unroll
            for (int v = 0; v < NRHS; v++) acc[v] = 0.0;
            double xown[NRHS];               // my own entry of the right-hand side (threads 0 .. 127), fetched off the chain
#pragma unroll
            for (int v = 0; v < NRHS; v++) xown[v] = (tid < NB && c0 + tid < m) ? x[v * xstride + c0 + tid] : 0.0;
            for (int j = nblk - 1; j > k; j--) {
                double tl[64];
                const int64_t cb = (int64_t)j * NB + half * 64;
                if (j - 1 > k) trsv_prefetch_tile(F + c0 + (int64_t)(j - 1) * NB * ldf, ldf, m - c0, NB);
                const double* Ut = F + grow + cb * ldf;
#pragma unroll
                for (int c = 0; c < 64; c++) tl[c] = (grow < m && cb + c < m) ? Ut[(int64_t)c * ldf] : 0.0;
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                if (tid < NB) {
                    int64_t c = (int64_t)j * NB + tid;
#pragma unroll
                    for (int v = 0; v < NRHS; v++) sv[v][j & 1][tid] = (c < m) ? __ldcg(x + v * xstride + c) : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double* svh = sv[v][j & 1] + half * 64;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 64; c += 2) {
                        a0 += tl[c] * svh[c];
                        a1 += tl[c + 1] * svh[c + 1];
                    }
                    acc[v] += a0 + a1;
                }
            }
#pragma unroll
            for (int v = 0; v < NRHS; v++) sacc[v][half][r] = acc[v];
            __syncthreads();
            if (tid < NB) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
                    sv[v][0][tid] = (c0 + tid < m) ? (xown[v] - sacc[v][0][tid] - sacc[v][1][tid]) : 0.0;
            }
            __syncthreads();
            // x[r] = sum_{c >= r} Dinv[r, c] v[c]
            double o0[NRHS];
#pragma unroll
            for (int v = 0; v < NRHS; v++) {
                double a0 = 0.0, a1 = 0.0;
                const double* row = sD + r + (half * 64) * NB;
                const double* svh = sv[v][0] + half * 64;
#pragma unroll 16
                for (int c = 0; c < 64; c += 2) {
                    a0 += row[c * NB] * svh[c];
                    a1 += row[(c + 1) * NB] * svh[c + 1];
                }
                o0[v] = a0 + a1;
            }
            __syncthreads();
#pragma unroll
            for (int v = 0; v < NRHS; v++) sacc[v][half][r] = o0[v];
            __syncthreads();
            if (tid < NB && c0 + tid < m) {
#pragma unroll
                for (int v = 0; v < NRHS; v++) x[v * xstride + c0 + tid] = sacc[v][0][tid] + sacc[v][1][tid];
            }
        }
        // only the threads that stored entries of x need the fence in front of the barrier; thread 0 then publishes
        if (TRANS ? (lane == 0) : (tid < NB)) __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&flags[1 + k], epoch);
    }
}

// ---- packet variant of the triangular solve (default) ----------------------------------------------------------
// The dependency step of trsv_kernel is: producer stores x_k, __threadfence, st.release of the block flag; consumer
// polls the flag (one L2 round trip per poll), barrier, THEN loads the 128 values (a second round trip).  Here every
// value travels as two 8-byte words {32 bits of the double, epoch} (the LL idea of NCCL: an aligned 8-byte access is
// single-copy atomic, so a word whose epoch matches carries valid data - no fence, no separate flag), and the consumer
// threads poll the very words they need: one round trip per dependency step, and the producer's fence + flag store
// leave the chain.  pkt: 2 words per entry, entry (v, k, r) at ((v * nblk + k) * 128 + r) * 2; zeroed at allocation,
// epochs start at 1 and grow with every sweep of the context.  Arithmetic and its order are those of trsv_kernel.
#ifdef HYP_EMU
__device__ __forceinline__ void trsv_pkt_post(unsigned long long* p, double val, int epoch) {
    unsigned long long b;
    memcpy(&b, &val, 8);
    const unsigned long long e = (unsigned long long)(unsigned)epoch << 32;
    __atomic_store_n(p, (b & 0xffffffffull) | e, __ATOMIC_RELEASE);
    __atomic_store_n(p + 1, (b >> 32) | e, __ATOMIC_RELEASE);
}
__device__ __forceinline__ double trsv_pkt_wait(const unsigned long long* p, int epoch) {
    unsigned long long w0, w1;
    do {
        w0 = __atomic_load_n(p, __ATOMIC_ACQUIRE);
        w1 = __atomic_load_n(p + 1, __ATOMIC_ACQUIRE);
    } while ((unsigned)(w0 >> 32) != (unsigned)epoch || (unsigned)(w1 >> 32) != (unsigned)epoch);
    const unsigned long long b = (w0 & 0xffffffffull) | (w1 << 32);
    double val;
    memcpy(&val, &b, 8);
    return val;
}
#else
__device__ __forceinline__ void trsv_pkt_post(unsigned long long* p, double val, int epoch) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(val);
    const unsigned long long e = (unsigned long long)(unsigned)epoch << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | e), "l"((b >> 32) | e)
                 : "memory");
}
__device__ __forceinline__ double trsv_pkt_wait(const unsigned long long* p, int epoch) {
    unsigned long long w0, w1;
    do {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    } while ((unsigned)(w0 >> 32) != (unsigned)epoch || (unsigned)(w1 >> 32) != (unsigned)epoch);
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
#endif

template <bool TRANS, int NRHS = 1>
__global__ void __launch_bounds__(256, 1)
trsv_pkt_kernel(const double* __restrict__ F, int64_t ldf, int64_t m, const double* __restrict__ dinv,
                double* x, int* flags, unsigned long long* pkt, int nblk, int epoch, int64_t xstride = 0) {
    HYP_DYN_SMEM(double, sD);                // Dinv_k, 128 x 128 col-major
    __shared__ double sv[NRHS][2][NB];
    __shared__ double sacc[NRHS][2][NB];
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
            double pacc[NRHS][16];
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) pacc[v][i] = 0.0;
            const int64_t cw = c0 + warp * 16;
            // my own block of the right-hand side does not depend on the chain: fetched now, used after the last tile
            double xown[NRHS][16];
            if (lane == 0) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
#pragma unroll
                    for (int i = 0; i < 16; i++) xown[v][i] = (cw + i < m) ? x[v * xstride + cw + i] : 0.0;
            }
            for (int j = 0; j < k; j++) {
                double tl[16][4];
                const double* Ut = F + (int64_t)j * NB + cw * ldf;
                if (j + 1 < k) trsv_prefetch_tile(F + (int64_t)(j + 1) * NB + c0 * ldf, ldf, NB, m - c0);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const bool ok = cw + i < m;
                    const double* col = Ut + (int64_t)i * ldf;
#pragma unroll
                    for (int h = 0; h < 4; h++) tl[i][h] = ok ? col[lane + 32 * h] : 0.0;
                }
                // one L2 round trip: every thread polls the packet of the entry it stages (no flag, no second load)
                if (tid < NB * NRHS) {
                    const int v = tid >> 7, r = tid & (NB - 1);
                    sv[v][j & 1][r] = trsv_pkt_wait(pkt + (((int64_t)v * nblk + j) * NB + r) * 2, epoch);
                }
                __syncthreads();
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double* svj = sv[v][j & 1];
                    const double v0 = svj[lane], v1 = svj[lane + 32], v2 = svj[lane + 64], v3 = svj[lane + 96];
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        pacc[v][i] += tl[i][0] * v0 + tl[i][1] * v1 + tl[i][2] * v2 + tl[i][3] * v3;
                }
            }
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    double a = pacc[v][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    pacc[v][i] = a;
                }
            __syncthreads();
            if (lane == 0) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        int64_t c = cw + i;
                        sv[v][0][warp * 16 + i] = (c < m) ? (xown[v][i] - pacc[v][i]) : 0.0;
                    }
            }
            __syncthreads();
            // y[c] = sum_{r <= c} Dinv[r, c] v[r]
#pragma unroll
            for (int v = 0; v < NRHS; v++) {
                const double* vb = sv[v][0];
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
                        if (c < m) x[v * xstride + c] = res[i];
                        // entries past m of the last block are published as zeros (their factor entries are read as zeros)
                        trsv_pkt_post(pkt + (((int64_t)v * nblk + k) * NB + warp * 16 + i) * 2, (c < m) ? res[i] : 0.0, epoch);
                    }
                }
            }
        } else {
            // x_k = Dinv_k (y_k - sum_{j>k} U[k-block, j-block] x_j); thread owns a row, the two
            // halves of the CTA split the 128 columns of a tile
            const int r = tid & (NB - 1), half = tid >> 7;
            const int64_t grow = c0 + r;
            double acc[NRHS];
#pragma unroll
            for (int v = 0; v < NRHS; v++) acc[v] = 0.0;
            double xown[NRHS];               // my own entry of the right-hand side (threads 0 .. 127), fetched off the chain
#pragma unroll
            for (int v = 0; v < NRHS; v++) xown[v] = (tid < NB && c0 + tid < m) ? x[v * xstride + c0 + tid] : 0.0;
            for (int j = nblk - 1; j > k; j--) {
                double tl[64];
                const int64_t cb = (int64_t)j * NB + half * 64;
                if (j - 1 > k) trsv_prefetch_tile(F + c0 + (int64_t)(j - 1) * NB * ldf, ldf, m - c0, NB);
                const double* Ut = F + grow + cb * ldf;
#pragma unroll
                for (int c = 0; c < 64; c++) tl[c] = (grow < m && cb + c < m) ? Ut[(int64_t)c * ldf] : 0.0;
                // one L2 round trip: every thread polls the packet of the entry it stages (no flag, no second load)
                if (tid < NB * NRHS) {
                    const int v = tid >> 7, r = tid & (NB - 1);
                    sv[v][j & 1][r] = trsv_pkt_wait(pkt + (((int64_t)v * nblk + j) * NB + r) * 2, epoch);
                }
                __syncthreads();
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double* svh = sv[v][j & 1] + half * 64;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 64; c += 2) {
                        a0 += tl[c] * svh[c];
                        a1 += tl[c + 1] * svh[c + 1];
                    }
                    acc[v] += a0 + a1;
                }
            }
#pragma unroll
            for (int v = 0; v < NRHS; v++) sacc[v][half][r] = acc[v];
            __syncthreads();
            if (tid < NB) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
                    sv[v][0][tid] = (c0 + tid < m) ? (xown[v] - sacc[v][0][tid] - sacc[v][1][tid]) : 0.0;
            }
            __syncthreads();
            // x[r] = sum_{c >= r} Dinv[r, c] v[c]
            double o0[NRHS];
#pragma unroll
            for (int v = 0; v < NRHS; v++) {
                double a0 = 0.0, a1 = 0.0;
                const double* row = sD + r + (half * 64) * NB;
                const double* svh = sv[v][0] + half * 64;
#pragma unroll 16
                for (int c = 0; c < 64; c += 2) {
                    a0 += row[c * NB] * svh[c];
                    a1 += row[(c + 1) * NB] * svh[c + 1];
                }
                o0[v] = a0 + a1;
            }
            __syncthreads();
#pragma unroll
            for (int v = 0; v < NRHS; v++) sacc[v][half][r] = o0[v];
            __syncthreads();
            if (tid < NB) {
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double xv = (c0 + tid < m) ? sacc[v][0][tid] + sacc[v][1][tid] : 0.0;
                    if (c0 + tid < m) x[v * xstride + c0 + tid] = xv;
                    trsv_pkt_post(pkt + (((int64_t)v * nblk + k) * NB + tid) * 2, xv, epoch);
                }
            }
        }
        // nothing to fence or to flag: a packet carries its own epoch.  The barrier only protects the shared buffers
        // (sD, sv, sacc) against the next ticket of this CTA.
        __syncthreads();
    }
}

struct TrsvTask {
    int k, j0, nj, w;     // block column, first tile (TRANS: ascending from j0, else descending), tiles, nseg << 16 | seg << 1 | final
};

// Segmented variant (default): a block column is cut into segments of at most SEG tiles; every segment is a ticket.
// Non-final segments write their partial sums to global scratch and bump a counter, the block's FINAL segment (the one
// that ends at the block just solved) adds them in segment order, applies the inverted diagonal block and publishes.
// With one ticket per block column the late columns were bound by their own serial loop (78 tiles x ~4.7 us per sweep at
// m = 10000), not by the dependency chain (~2 us per block); here only <= SEG tiles of a column sit in front of the chain.
// Ticket order = the order in which a ticket's last input becomes available (host: trsv_tasks), a topological order of the
// dependency graph, so the earliest unfinished ticket is always held by a resident CTA: no deadlock.
template <bool TRANS, int NRHS = 1>
__global__ void __launch_bounds__(256, 1)
trsv_seg_kernel(const double* __restrict__ F, int64_t ldf, int64_t m, const double* __restrict__ dinv,
                double* x, int* flags, const TrsvTask* __restrict__ tasks, int ntasks, int nblk, int epoch,
                double* __restrict__ part, int maxseg, int64_t xstride = 0) {
    int* cnt = flags + 1 + nblk;             // finished non-final segments per block column (zeroed per solve)
    HYP_DYN_SMEM(double, sD);                // Dinv_k, 128 x 128 col-major
    __shared__ double sv[NRHS][2][NB];
    __shared__ double sacc[NRHS][2][NB];
    __shared__ int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) s_ticket = atomicAdd(&flags[0], 1);
        __syncthreads();
        const int t = s_ticket;
        __syncthreads();
        if (t >= ntasks) return;
        const TrsvTask task = tasks[t];
        const int k = task.k, j0 = task.j0, nj = task.nj;
        const bool final = task.w & 1;
        const int seg = (task.w >> 1) & 0x7fff, nseg = task.w >> 16;
        const int64_t c0 = (int64_t)k * NB;
        if (final) {
            const double* Dk = dinv + (int64_t)k * NB * NB;
#pragma unroll 8
            for (int idx = tid; idx < NB * NB; idx += 256) sD[idx] = Dk[idx];
        }

        if (TRANS) {
            // y_k = Dinv_k' (b_k - sum_{j<k} U[j-block, k-block]' y_j); warp w owns 16 columns,
            // lane l rows l, l+32, l+64, l+96 of every tile
            double pacc[NRHS][16];
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) pacc[v][i] = 0.0;
            const int64_t cw = c0 + warp * 16;
            double psum[NRHS][16];            // lane 0: partial sums of the earlier segments (final ticket only)
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) psum[v][i] = 0.0;
            for (int jj = 0; jj < nj; jj++) {
                const int j = j0 + jj;
                double tl[16][4];
                const double* Ut = F + (int64_t)j * NB + cw * ldf;
                if (jj + 1 < nj) trsv_prefetch_tile(F + (int64_t)(j + 1) * NB + c0 * ldf, ldf, NB, m - c0);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const bool ok = cw + i < m;
                    const double* col = Ut + (int64_t)i * ldf;
#pragma unroll
                    for (int h = 0; h < 4; h++) tl[i][h] = ok ? col[lane + 32 * h] : 0.0;
                }
                if (final && jj == nj - 1 && nseg > 1) {
                    // the earlier segments of this column finished several chain steps ago: their sums are fetched here,
                    // behind the tile loads and in front of the wait for the last block - not on the chain
                    if (tid == 0) {
                        while (ld_acquire(&cnt[k]) < nseg - 1) {
                        }
                    }
                    __syncthreads();
                    if (lane == 0) {
                        for (int sg = 0; sg < nseg - 1; sg++)
#pragma unroll
                            for (int v = 0; v < NRHS; v++)
#pragma unroll
                                for (int i = 0; i < 16; i++)
                                    psum[v][i] += __ldcg(part + (((int64_t)k * maxseg + sg) * NRHS + v) * NB + warp * 16 + i);
                    }
                }
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                if (tid < NB) {
#pragma unroll
                    for (int v = 0; v < NRHS; v++) sv[v][j & 1][tid] = __ldcg(x + v * xstride + (int64_t)j * NB + tid);
                }
                __syncthreads();
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double* svj = sv[v][j & 1];
                    const double v0 = svj[lane], v1 = svj[lane + 32], v2 = svj[lane + 64], v3 = svj[lane + 96];
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        pacc[v][i] += tl[i][0] * v0 + tl[i][1] * v1 + tl[i][2] * v2 + tl[i][3] * v3;
                }
            }
#pragma unroll
            for (int v = 0; v < NRHS; v++)
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    double a = pacc[v][i];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    pacc[v][i] = a;
                }
            __syncthreads();
            if (!final) {
                // partial sums of this segment -> global scratch; the block's final segment adds them up in segment order
                if (lane == 0) {
#pragma unroll
                    for (int v = 0; v < NRHS; v++)
#pragma unroll
                        for (int i = 0; i < 16; i++)
                            part[(((int64_t)k * maxseg + seg) * NRHS + v) * NB + warp * 16 + i] = pacc[v][i];
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) atomicAdd(&cnt[k], 1);
                continue;
            }
            if (lane == 0) {
#pragma unroll
                for (int v = 0; v < NRHS; v++)
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        int64_t c = cw + i;
                        sv[v][0][warp * 16 + i] = (c < m) ? (x[v * xstride + c] - (pacc[v][i] + psum[v][i])) : 0.0;
                    }
            }
            __syncthreads();
            // y[c] = sum_{r <= c} Dinv[r, c] v[r]
#pragma unroll
            for (int v = 0; v < NRHS; v++) {
                const double* vb = sv[v][0];
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
                        if (c < m) x[v * xstride + c] = res[i];
                    }
                }
            }
        } else {
            // x_k = Dinv_k (y_k - sum_{j>k} U[k-block, j-block] x_j); thread owns a row, the two
            // halves of the CTA split the 128 columns of a tile
            const int r = tid & (NB - 1), half = tid >> 7;
            const int64_t grow = c0 + r;
            double acc[NRHS];
#pragma unroll
            for (int v = 0; v < NRHS; v++) acc[v] = 0.0;
            double psum[NRHS];                // threads 0 .. 127: partial sums of the earlier segments (final ticket only)
#pragma unroll
            for (int v = 0; v < NRHS; v++) psum[v] = 0.0;
            for (int jj = 0; jj < nj; jj++) {
                const int j = j0 - jj;
                double tl[64];
                const int64_t cb = (int64_t)j * NB + half * 64;
                if (jj + 1 < nj) trsv_prefetch_tile(F + c0 + (int64_t)(j - 1) * NB * ldf, ldf, m - c0, NB);
                const double* Ut = F + grow + cb * ldf;
#pragma unroll
                for (int c = 0; c < 64; c++) tl[c] = (grow < m && cb + c < m) ? Ut[(int64_t)c * ldf] : 0.0;
                if (final && jj == nj - 1 && nseg > 1) {
                    if (tid == 0) {
                        while (ld_acquire(&cnt[k]) < nseg - 1) {
                        }
                    }
                    __syncthreads();
                    if (tid < NB) {
                        for (int sg = 0; sg < nseg - 1; sg++)
#pragma unroll
                            for (int v = 0; v < NRHS; v++) psum[v] += __ldcg(part + (((int64_t)k * maxseg + sg) * NRHS + v) * NB + tid);
                    }
                }
                if (tid == 0) {
                    while (ld_acquire(&flags[1 + j]) != epoch) {
                    }
                }
                __syncthreads();
                if (tid < NB) {
                    int64_t c = (int64_t)j * NB + tid;
#pragma unroll
                    for (int v = 0; v < NRHS; v++) sv[v][j & 1][tid] = (c < m) ? __ldcg(x + v * xstride + c) : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double* svh = sv[v][j & 1] + half * 64;
                    double a0 = 0.0, a1 = 0.0;
#pragma unroll
                    for (int c = 0; c < 64; c += 2) {
                        a0 += tl[c] * svh[c];
                        a1 += tl[c + 1] * svh[c + 1];
                    }
                    acc[v] += a0 + a1;
                }
            }
#pragma unroll
            for (int v = 0; v < NRHS; v++) sacc[v][half][r] = acc[v];
            __syncthreads();
            if (!final) {
                if (tid < NB) {
#pragma unroll
                    for (int v = 0; v < NRHS; v++)
                        part[(((int64_t)k * maxseg + seg) * NRHS + v) * NB + tid] = sacc[v][0][tid] + sacc[v][1][tid];
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) atomicAdd(&cnt[k], 1);
                continue;
            }
            if (tid < NB) {
#pragma unroll
                for (int v = 0; v < NRHS; v++) {
                    const double a = sacc[v][0][tid] + sacc[v][1][tid] + psum[v];
                    sv[v][0][tid] = (c0 + tid < m) ? (x[v * xstride + c0 + tid] - a) : 0.0;
                }
            }
            __syncthreads();
            // x[r] = sum_{c >= r} Dinv[r, c] v[c]
            double o0[NRHS];
#pragma unroll
            for (int v = 0; v < NRHS; v++) {
                double a0 = 0.0, a1 = 0.0;
                const double* row = sD + r + (half * 64) * NB;
                const double* svh = sv[v][0] + half * 64;
#pragma unroll 16
                for (int c = 0; c < 64; c += 2) {
                    a0 += row[c * NB] * svh[c];
                    a1 += row[(c + 1) * NB] * svh[c + 1];
                }
                o0[v] = a0 + a1;
            }
            __syncthreads();
#pragma unroll
            for (int v = 0; v < NRHS; v++) sacc[v][half][r] = o0[v];
            __syncthreads();
            if (tid < NB && c0 + tid < m) {
#pragma unroll
                for (int v = 0; v < NRHS; v++) x[v * xstride + c0 + tid] = sacc[v][0][tid] + sacc[v][1][tid];
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&flags[1 + k], epoch);
    }
}

}  // namespace hypdev

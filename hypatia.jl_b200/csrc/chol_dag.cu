// Task-graph Cholesky: the whole blocked factorisation of the Schur complement (K3 of SURVEY.md
// section 2.3) as ONE persistent kernel.
//
// reference call site: posdef_fact!(A) = cholesky!(Symmetric(A, :U), check=false)
// (src/linearalgebra/dense.jl:191-192, LAPACK dpotrf 'U') called from update_lhs_fact
// (src/Solvers/systemsolvers/qrchol.jl:249-250).
//
// Why: the stream-level version (chol.cu) issues ~400 dependent launches for m = 10000 and spends
// 24.5 ms where the FP64 tensor roofline allows 9.5 ms - 15.5 % of the 1-GPU iteration and half of the
// 8-GPU iteration (replicated factor).  Here every CTA stays resident and the dependencies between
// 128 x 128 tile tasks are version counters in global memory:
//
//   PANEL(k)           factor + invert diagonal tile (k, k)               (panel_body, CUDA cores)
//   TRSM(k, j)         U_kj = Dinv_k' A_kj                 j > k          (TMA + DMMA, depth 128)
//   UPD(k0, nk, i, j)  A_ij -= sum_{k0 <= k < k0 + nk} U_ki' U_kj         (TMA + DMMA, depth 128 nk)
//
// Two-level blocking as before: inside an outer block of 4 tile rows the updates have depth 128
// (nk = 1, only the rows of that block), the rest of the trailing matrix is updated once per outer
// block with depth 512 (nk = 4).  Every tile (i, j) therefore has a fixed sequence of writers -
// UPD of outer block 0, 1, .., own block's rows, then its TRSM / PANEL - and ver[i * nt + j] counts
// how many of them have finished: a writer waits for ver == its sequence number (this also
// serialises the read-modify-write of the tile), a reader waits for ver == fin(row) (tile final).
//
// Two ticket queues, both subsequences of the sequential right-looking order (so the earliest
// unfinished task is always held by a live CTA and its inputs are complete: no deadlock):
//   chain queue (CTA 0):  PANEL(k), then TRSM(k, k+1) and UPD(k, 1, k+1, k+1) when k+1 is in the
//                         same outer block - the latency-critical diagonal band;
//   bulk queue (all other CTAs): everything else, tiles of the next outer block first.
// In each CTA one producer warp takes tickets, waits for the task's inputs (ld.acquire), and streams
// the operand tiles through a 6-stage TMA ring; 8 consumer warps run DMMA.8x8x4 and the epilogue, or
// panel_body for PANEL tasks (its scratch overlays the idle ring).  The producer runs ahead of the
// consumers by up to two tasks, which hides the dependency check and the TMA latency of depth-128 tasks.
// Bound: FP64 tensor pipe for the bulk, latency of 128 dependent pivots per PANEL for the chain.
#include "common.cuh"
#include "chol_kernels.cuh"

using namespace hypdev;

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int STAGES = 6;
constexpr int TILE_BYTES = BM * BK * 8;            // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int NCW = 8;                             // consumer warps
constexpr int NTHREADS = (NCW + 1) * 32;
constexpr int FIFO = 2;
constexpr int OBT = 4;                             // tiles per outer block
constexpr int RING_BYTES = STAGES * STAGE_BYTES;   // 192 KB; panel scratch overlays it
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 512;
// panel scratch inside the ring region
constexpr int PANEL_SA = NB * LDU * 8;
constexpr int PANEL_DIAGX = NB * 8;
constexpr int PANEL_SX = SB * LDX * 8;
constexpr int PANEL_ST = SB * LDT * 8;
static_assert(PANEL_SA + PANEL_DIAGX + PANEL_SX + PANEL_ST + 64 <= RING_BYTES, "panel scratch must fit in the ring");

enum { TASK_PANEL = 0, TASK_TRSM = 1, TASK_UPD = 2, TASK_STOP = 3 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DAG_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DAG_WAIT_DONE;\n"
        "bra DAG_WAIT_LOOP;\n"
        "DAG_WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ int ld_acq(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_ver(const int* p, int expect) {
    while (ld_acq(p) < expect) __nanosleep(40);
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// ---- version arithmetic (see the header comment) ----
// Tile (k, j) INSIDE the diagonal 4 x 4 tile block of its outer block is written by: the depth-512 update of every earlier
// outer block, one depth-128 update per earlier row of its own block (right-looking, chain queue), then its TRSM /
// PANEL.  A tile to the RIGHT of that block gets the updates of its own block's earlier rows as ONE merged update of
// depth 128 r (left-looking, inside a row group of the bulk queue).
__device__ __forceinline__ int t_blk(int k) { return k / OBT; }
__device__ __forceinline__ int t_bend(int k, int nt) { return min((k / OBT + 1) * OBT, nt); }
// version of tile (k, j) once it is final
__device__ __forceinline__ int t_fin(int k, int j, int nt) {
    const int r = k % OBT;
    return t_blk(k) + (j < t_bend(k, nt) ? r : (r > 0 ? 1 : 0)) + 1;
}
// version tile (i, j) must have before UPD(k0, nk, i, j) writes it
__device__ __forceinline__ int t_upd_exp(int k0, int i, int j, int nt) {
    if (t_blk(k0) < t_blk(i)) return t_blk(k0);                       // update by an earlier outer block
    return t_blk(i) + (j < t_bend(i, nt) ? k0 - OBT * t_blk(i) : 0);   // own block: right-looking step / merged update
}

__global__ void __launch_bounds__(NTHREADS, 1)
potrf_dag_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapD,
                 const int4* __restrict__ chain_tasks, const int2* __restrict__ chain_groups, int n_chain,
                 const int4* __restrict__ bulk_tasks, const int2* __restrict__ bulk_groups, int n_bulk,
                 double* __restrict__ A, int64_t lda, int64_t m, int nt, double* __restrict__ dinv,
                 int* __restrict__ ver, int* __restrict__ tickets, int* __restrict__ info, int n_chain_ctas,
                 unsigned long long* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t ring = (base + 1023u) & ~1023u;
    uint8_t* ring_ptr = smem_raw + (ring - base);
    const uint32_t bar_full = ring + RING_BYTES;
    const uint32_t bar_empty = bar_full + STAGES * 8;
    const uint32_t fifo_full = bar_empty + STAGES * 8;
    const uint32_t fifo_empty = fifo_full + FIFO * 8;
    int4* fifo_desc = reinterpret_cast<int4*>(ring_ptr + RING_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool is_chain = (int)blockIdx.x < n_chain_ctas;
    const int4* __restrict__ tasks = is_chain ? chain_tasks : bulk_tasks;
    const int2* __restrict__ groups = is_chain ? chain_groups : bulk_groups;
    const int n_groups = is_chain ? n_chain : n_bulk;
    int* ticket = tickets + (is_chain ? 0 : 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, NCW);
        }
        for (int f = 0; f < FIFO; f++) {
            mbar_init(fifo_full + f * 8, 1);
            mbar_init(fifo_empty + f * 8, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        // ===== producer: tickets, dependencies, TMA =====
        if (lane != 0) return;
        int stage = 0, slot = 0;
        uint32_t phase = 0, fphase = 0;
        const int* pending_ver = nullptr;     // a PANEL of this CTA whose scratch still occupies the ring
        int pending_fin = 0;
        unsigned long long t_begin = 0, t_wait = 0;
        if (dbg) t_begin = gtime();
        bool stop = false;
        while (!stop) {
            // a ticket is a GROUP of tasks run back to back by this CTA (single tasks, or the row group of one tile column)
            const int t = atomicAdd(ticket, 1);
            const int2 grp = (t < n_groups) ? groups[t] : make_int2(0, 1);
            for (int sub = 0; sub < grp.y; sub++) {
                unsigned long long tw0 = 0;
                if (dbg) tw0 = gtime();
                const int4 task = (t < n_groups) ? tasks[grp.x + sub] : make_int4(TASK_STOP, 0, 0, 0);
                const int type = task.x & 0xff, nk = task.x >> 8, k0 = task.y, ti = task.z, tj = task.w;
                if (type != TASK_STOP) {
                    // inputs: the C tile must have seen all earlier writers, the operand tiles must be final
                    const int cexp = (type == TASK_UPD) ? t_upd_exp(k0, ti, tj, nt) : t_fin(ti, tj, nt) - 1;
                    wait_ver(ver + (int64_t)ti * nt + tj, cexp);
                    if (type == TASK_TRSM) {
                        wait_ver(ver + (int64_t)ti * nt + ti, t_fin(ti, ti, nt));
                    } else if (type == TASK_UPD) {
                        for (int r = 0; r < nk; r++) {
                            const int k = k0 + r;
                            wait_ver(ver + (int64_t)k * nt + ti, t_fin(k, ti, nt));
                            if (tj != ti) wait_ver(ver + (int64_t)k * nt + tj, t_fin(k, tj, nt));
                        }
                    }
                }
                if (dbg) t_wait += gtime() - tw0;
                // hand the task to the consumers
                mbar_wait(fifo_empty + slot * 8, fphase ^ 1u);
                fifo_desc[slot] = task;
                mbar_arrive(fifo_full + slot * 8);
                if (++slot == FIFO) {
                    slot = 0;
                    fphase ^= 1u;
                }
                if (type == TASK_STOP) {
                    if (dbg) {
                        dbg[blockIdx.x * 8 + 0] = t_wait;
                        dbg[blockIdx.x * 8 + 1] = gtime() - t_begin;
                    }
                    stop = true;
                    break;
                }
                if (type == TASK_PANEL) {
                    pending_ver = ver + (int64_t)ti * nt + ti;
                    pending_fin = t_fin(ti, ti, nt);
                    continue;
                }
                if (pending_ver) {
                    wait_ver(pending_ver, pending_fin);      // the ring is free again
                    pending_ver = nullptr;
                }
                // data written through the generic proxy (by other CTAs, or by this CTA's consumers earlier in the
                // group) is read by the async proxy (TMA) below
                asm volatile("fence.proxy.async;" ::: "memory");
                const bool single = (type == TASK_UPD) && (ti == tj);
                const int nkb = nk * (BM / BK);
                const CUtensorMap* mp = (type == TASK_TRSM) ? &mapD : &mapA;
                const int pk0 = (type == TASK_TRSM) ? 0 : k0 * BM;
                const int pc0 = ti * BM;                    // TRSM: column block ti of the Dinv strip
                const int rk0 = (type == TASK_TRSM) ? ti * BM : k0 * BM;
                const int rc0 = tj * BM;
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                    const uint32_t full = bar_full + stage * 8;
                    mbar_expect_tx(full, single ? TILE_BYTES : STAGE_BYTES);
                    const uint32_t dst = ring + stage * STAGE_BYTES;
                    tma_load_2d(dst, mp, pk0 + kb * BK, pc0, full);
                    if (!single) tma_load_2d(dst + TILE_BYTES, &mapA, rk0 + kb * BK, rc0, full);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int warp_m = warp >> 2;
    const int warp_n = warp & 3;
    const int g = lane >> 2;
    const int t4 = lane & 3;
    uint32_t koff[4];
#pragma unroll
    for (int s = 0; s < 4; s++) koff[s] = (uint32_t)((((2 * s + (t4 >> 1)) ^ g) << 4) + ((t4 & 1) << 3));
    const uint32_t offA = (uint32_t)((warp_m * 64 + g) * 128);
    const uint32_t offB = (uint32_t)((warp_n * 32 + g) * 128);

    int stage = 0, slot = 0;
    uint32_t phase = 0, fphase = 0;
    unsigned long long c_wait = 0, c_busy = 0, c_panel = 0, c_tasks = 0;
    while (true) {
        unsigned long long c0 = 0;
        if (dbg) c0 = gtime();
        mbar_wait(fifo_full + slot * 8, fphase);
        unsigned long long c1 = 0;
        if (dbg) {
            c1 = gtime();
            c_wait += c1 - c0;
        }
        const int4 task = fifo_desc[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(fifo_empty + slot * 8);
        if (++slot == FIFO) {
            slot = 0;
            fphase ^= 1u;
        }
        const int type = task.x & 0xff, nk = task.x >> 8, k0 = task.y, ti = task.z, tj = task.w;
        if (type == TASK_STOP) {
            if (dbg && threadIdx.x == 0) {
                dbg[blockIdx.x * 8 + 2] = c_wait;
                dbg[blockIdx.x * 8 + 3] = c_busy;
                dbg[blockIdx.x * 8 + 4] = c_tasks;
                dbg[blockIdx.x * 8 + 5] = c_panel;
            }
            break;
        }
        int* cver = ver + (int64_t)ti * nt + tj;
        int cnext;
        if (type == TASK_PANEL) {
            double* sA = reinterpret_cast<double*>(ring_ptr);
            double* diagX = reinterpret_cast<double*>(ring_ptr + PANEL_SA);
            double* sX = reinterpret_cast<double*>(ring_ptr + PANEL_SA + PANEL_DIAGX);
            double* sT = reinterpret_cast<double*>(ring_ptr + PANEL_SA + PANEL_DIAGX + PANEL_SX);
            int* s_bad = reinterpret_cast<int*>(ring_ptr + PANEL_SA + PANEL_DIAGX + PANEL_SX + PANEL_ST);
            const int64_t r0 = (int64_t)ti * NB;
            const int nb = (int)((int64_t)NB < m - r0 ? (int64_t)NB : m - r0);
            const int bad = panel_body_fast<1, true>(A + r0 + r0 * lda, lda, nb, dinv + (int64_t)ti * NB * NB, NB, NB,
                                                    false, sA, diagX, sX, sT, s_bad);
            if (bad && threadIdx.x == 0) atomicCAS(info, 0, (int)(r0 + bad));
            cnext = t_fin(ti, ti, nt);
        } else {
            const bool single = (type == TASK_UPD) && (ti == tj);
            const int nkb = nk * (BM / BK);
            double acc[8][4][2];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int kb = 0; kb < nkb; kb++) {
                mbar_wait(bar_full + stage * 8, phase);
                const uint32_t sAa = ring + stage * STAGE_BYTES;
                const uint32_t sBb = single ? sAa : sAa + TILE_BYTES;
                const uint32_t pa = sAa + offA, pb = sBb + offB;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    double a[8], b[4];
#pragma unroll
                    for (int i = 0; i < 8; i++) a[i] = lds64(pa + i * 1024 + koff[s]);
#pragma unroll
                    for (int j = 0; j < 4; j++) b[j] = lds64(pb + j * 1024 + koff[s]);
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + stage * 8);
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            // epilogue: TRSM overwrites the tile, UPD subtracts.  The tile is read through L2 (other CTAs wrote it), ALL
            // loads of a half first and then the stores: interleaved they serialise on the L2 latency (64 round trips)
            const bool upd = type == TASK_UPD;
            const int64_t row0 = (int64_t)ti * BM + warp_m * 64 + g;
            const int64_t col0 = (int64_t)tj * BM + warp_n * 32 + 2 * t4;
#pragma unroll
            for (int jh = 0; jh < 4; jh += 2) {
                double cv[2][2][8];
                if (upd) {
#pragma unroll
                    for (int j = 0; j < 2; j++)
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int64_t col = col0 + 8 * (jh + j) + e;
                            const double* cp = A + col * lda;
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const int64_t row = row0 + 8 * i;
                                cv[j][e][i] = (col < m && row < m) ? __ldcg(cp + row) : 0.0;
                            }
                        }
                }
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int64_t col = col0 + 8 * (jh + j) + e;
                        if (col >= m) continue;
                        double* cp = A + col * lda;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int64_t row = row0 + 8 * i;
                            if (row < m) cp[row] = upd ? (cv[j][e][i] - acc[i][jh + j][e]) : acc[i][jh + j][e];
                        }
                    }
            }
            cnext = (upd ? t_upd_exp(k0, ti, tj, nt) : t_fin(ti, tj, nt) - 1) + 1;
        }
        // publish: every consumer thread's stores, then the version counter
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) st_rel(cver, cnext);
        if (dbg) {
            const unsigned long long c2 = gtime();
            c_busy += c2 - c1;
            if (type == TASK_PANEL) c_panel += c2 - c1;
            c_tasks++;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) throw HypError{"cuTensorMapEncodeTiled entry point not available"};
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

void make_map(CUtensorMap* map, const double* basep, int64_t klen, int64_t ncols, int64_t ld) {
    if (((uintptr_t)basep & 15) || (ld & 1)) throw HypError{"potrf: matrix must be 16-byte aligned with an even leading dimension"};
    cuuint64_t dims[2] = {(cuuint64_t)klen, (cuuint64_t)ncols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {BK, BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)basep, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw HypError{"potrf: cuTensorMapEncodeTiled failed"};
}

struct DagLists {
    int device, nt;
    int4 *d_chain, *d_bulk;
    int2 *d_chain_groups, *d_bulk_groups;
    int n_chain, n_bulk;          // number of GROUPS (tickets)
};
std::vector<DagLists> g_lists;

inline int4 mk(int type, int nk, int k0, int i, int j) { return make_int4(type | (nk << 8), k0, i, j); }

struct Queue {
    std::vector<int4> tasks;
    std::vector<int2> groups;
    void one(int4 t) {
        groups.push_back(make_int2((int)tasks.size(), 1));
        tasks.push_back(t);
    }
    void group(const std::vector<int4>& g) {
        groups.push_back(make_int2((int)tasks.size(), (int)g.size()));
        tasks.insert(tasks.end(), g.begin(), g.end());
    }
};

// Task lists in an order that is a topological order of the task graph (see the header comment).
// chain queue: everything inside the diagonal 4 x 4 tile block of an outer block - the four PANELs and the TRSM /
//   depth-128 UPD tasks between them (right-looking) - so that the chain team can factor outer block B + 1 while the
//   bulk CTAs are still applying outer block B to the far part of the trailing matrix (look-ahead of one block);
// bulk queue: per outer block B
//   * one ROW GROUP per tile column j right of the block, run by one CTA back to back (left-looking inside the block):
//       TRSM(k0, j); UPD(k0, 1, k0+1, j); TRSM(k0+1, j); UPD(k0, 2, k0+2, j); TRSM(k0+2, j); UPD(k0, 3, k0+3, j); TRSM(k0+3, j)
//     - all dependencies inside a group are the CTA's own, so the 70-odd groups of a block are independent of each other
//     (the separate TRSM / UPD phases of a right-looking block row left half of the CTAs idle at every phase boundary);
//   * the depth-512 update of the trailing matrix: first the tiles in the rows of the NEXT outer block (its chain and
//     row groups wait for them), the far rest is interleaved 1 : 1 with the row groups of the next block.
void build_lists(int nt, Queue& chain, Queue& bulk) {
    std::vector<int4> far_prev;
    for (int B0 = 0; B0 < nt; B0 += OBT) {
        const int Bend = std::min(B0 + OBT, nt);          // exclusive
        for (int k = B0; k < Bend; k++) {
            chain.one(mk(TASK_PANEL, 0, k, k, k));
            for (int j = k + 1; j < Bend; j++) chain.one(mk(TASK_TRSM, 1, k, k, j));
            for (int i = k + 1; i < Bend; i++)
                for (int j = i; j < Bend; j++) chain.one(mk(TASK_UPD, 1, k, i, j));
        }
        // row groups of this block interleaved with the far updates of the previous one
        size_t fp = 0;
        for (int j = Bend; j < nt; j++) {
            std::vector<int4> g;
            g.push_back(mk(TASK_TRSM, 1, B0, B0, j));
            for (int r = 1; B0 + r < Bend; r++) {
                g.push_back(mk(TASK_UPD, r, B0, B0 + r, j));
                g.push_back(mk(TASK_TRSM, 1, B0 + r, B0 + r, j));
            }
            bulk.group(g);
            if (fp < far_prev.size()) bulk.one(far_prev[fp++]);
        }
        for (; fp < far_prev.size(); fp++) bulk.one(far_prev[fp]);
        far_prev.clear();
        if (Bend >= nt) break;
        const int nk = Bend - B0;                          // == OBT here
        const int Nend = std::min(Bend + OBT, nt);
        for (int i = Bend; i < Nend; i++)
            for (int j = i; j < Nend; j++) bulk.one(mk(TASK_UPD, nk, B0, i, j));
        for (int i = Bend; i < Nend; i++)
            for (int j = Nend; j < nt; j++) bulk.one(mk(TASK_UPD, nk, B0, i, j));
        for (int i = Nend; i < nt; i++)
            for (int j = i; j < nt; j++) far_prev.push_back(mk(TASK_UPD, nk, B0, i, j));
    }
}

template <typename T>
T* to_device(hyp_ctx* ctx, const std::vector<T>& v) {
    T* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return d;
}

DagLists& get_lists(hyp_ctx* ctx, int nt) {
    for (auto& e : g_lists)
        if (e.device == ctx->device && e.nt == nt) return e;
    Queue chain, bulk;
    build_lists(nt, chain, bulk);
    DagLists e{ctx->device, nt, to_device(ctx, chain.tasks), to_device(ctx, bulk.tasks), to_device(ctx, chain.groups),
               to_device(ctx, bulk.groups), (int)chain.groups.size(), (int)bulk.groups.size()};
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    g_lists.push_back(e);
    return g_lists.back();
}

}  // namespace

// dinv must hold ceil(m / 128) blocks of 128 x 128 whose strictly lower parts are zero (they are: dalloc zero-fills
// and the panel writes zeros there).  Returns false when the task-graph kernel does not apply (alignment).
bool hyp_potrf_upper_dag(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info) {
    if (m <= 0) return true;
    if (((uintptr_t)A & 15) || (lda & 1) || ((uintptr_t)d_dinv & 15)) return false;
    TimeScope ts(ctx, T_POTRF);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(potrf_dag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    const int nt = ceil_div(m, BM);
    DagLists& L = get_lists(ctx, nt);
    CUtensorMap mapA, mapD;
    make_map(&mapA, A, m, m, lda);
    make_map(&mapD, d_dinv, NB, (int64_t)nt * NB, NB);
    cudaStream_t s = ctx->stream;
    // version counters (nt * nt) + the two ticket counters, per context
    const int64_t nver = (int64_t)nt * nt + 2;
    if (ctx->dag_ver_len < nver) {
        CUDA_TRY(cudaStreamSynchronize(s));
        if (ctx->d_dag_ver) cudaFree(ctx->d_dag_ver);
        ctx->d_dag_ver = nullptr;
        CUDA_TRY(cudaMalloc((void**)&ctx->d_dag_ver, (size_t)nver * sizeof(int)));
        ctx->dag_ver_len = nver;
    }
    int* d_ver = ctx->d_dag_ver;
    CUDA_TRY(cudaMemsetAsync(d_info, 0, sizeof(int), s));
    CUDA_TRY(cudaMemsetAsync(d_ver, 0, (size_t)nver * sizeof(int), s));
    static int chain_team = -1;
    if (chain_team < 0) {
        const char* e = getenv("HYP_POTRF_CHAIN_CTAS");
        chain_team = e ? std::max(1, atoi(e)) : 8;
    }
    const int n_chain_ctas = std::min(chain_team, L.n_chain);
    const int grid = n_chain_ctas + std::max(0, std::min(ctx->sm_count - n_chain_ctas, L.n_bulk));
    unsigned long long* dbg = nullptr;
    if (getenv("HYP_POTRF_DEBUG")) {
        if (!ctx->d_dag_dbg) CUDA_TRY(cudaMalloc((void**)&ctx->d_dag_dbg, 256 * 8 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemsetAsync(ctx->d_dag_dbg, 0, 256 * 8 * sizeof(unsigned long long), s));
        dbg = ctx->d_dag_dbg;
    }
    potrf_dag_kernel<<<grid, NTHREADS, SMEM_BYTES, s>>>(mapA, mapD, L.d_chain, L.d_chain_groups, L.n_chain, L.d_bulk,
                                                        L.d_bulk_groups, L.n_bulk, A, lda, m, nt,
                                                        d_dinv, d_ver, d_ver + (size_t)nt * nt, d_info, n_chain_ctas, dbg);
    if (dbg) {
        std::vector<unsigned long long> h(256 * 8);
        CUDA_TRY(cudaMemcpyAsync(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        double sums[2][8] = {{0}};
        int cnt[2] = {0, 0};
        for (int b = 0; b < grid; b++) {
            const int c = b < n_chain_ctas ? 0 : 1;
            cnt[c]++;
            for (int q = 0; q < 8; q++) sums[c][q] += (double)h[b * 8 + q];
        }
        for (int c = 0; c < 2; c++)
            if (cnt[c])
                fprintf(stderr,
                        "[potrf_dag m=%lld] %s CTAs=%d  avg per CTA: producer dep-wait %.3f ms, producer total %.3f ms, consumer "
                        "fifo-wait %.3f ms, consumer busy %.3f ms (panel %.3f ms), tasks %.1f\n",
                        (long long)m, c == 0 ? "chain" : "bulk", cnt[c], sums[c][0] / cnt[c] * 1e-6, sums[c][1] / cnt[c] * 1e-6,
                        sums[c][2] / cnt[c] * 1e-6, sums[c][3] / cnt[c] * 1e-6, sums[c][5] / cnt[c] * 1e-6, sums[c][4] / cnt[c]);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return true;
}

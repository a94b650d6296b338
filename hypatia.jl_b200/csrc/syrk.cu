// Upper-triangular A'B contraction in FP64:  C = alpha * P' R + beta * C  (K1/K2 of SURVEY.md
// section 2.3).  With P == R it is the Schur-complement SYRK of the reference
// (outer_prod!(HGQ2, lhs, true, false) = BLAS.syrk!('U','T'), qrchol.jl:234 / dense.jl:80-86);
// with P = G_k, R = H_k G_k it is the mul!(lhs, G_k', H_k G_k, true, true) branch
// (qrchol.jl:245) restricted to the upper triangle; with alpha = -1, beta = 1 on a row panel
// of the factor it is the trailing update of the blocked Cholesky (chol.cu).
//
// Design (sm_100a):
//   * both operands are K-major (the contraction index runs down the contiguous columns of the
//     column-major panels), so a 16(k) x 128(cols) box is one TMA tile; TMA (UTMALDG) with the
//     128-byte swizzle stages tiles into shared memory through a 6-deep mbarrier ring;
//     out-of-range rows/columns are zero-filled by the TMA unit, so ragged q and m need no
//     special casing in the main loop;
//   * one producer warp issues the TMA loads, 8 consumer warps (2 x 4, 64 x 32 accumulators
//     each) run DMMA.8x8x4 (mma.sync.m8n8k4.f64 - tcgen05.mma has no f64 kind) on fragments read
//     conflict-free from the swizzled tiles (8 rows x 4 k of 8 bytes = 2 wavefronts, the
//     minimum for 256 bytes);
//   * persistent grid of one CTA per SM walking a precomputed list of upper-triangular
//     128 x 128 tiles ordered in 8-tile row groups, so the CTAs of a wave share row/column
//     panels in L2; diagonal tiles of a SYRK load one operand tile instead of two.
// Bound: tensor (FP64 DMMA) pipe.  Algorithmic flops per launch = klen * ncols * (ncols + 1).
#include "common.cuh"

namespace {

constexpr int BM = 128;           // tile rows of C (columns of P)
constexpr int BN = 128;           // tile cols of C (columns of R)
constexpr int BK = 16;            // k per stage: 16 doubles = 128 B = one swizzle row
constexpr int STAGES = 6;
constexpr int TILE_BYTES = BM * BK * 8;           // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;       // P tile + R tile
constexpr int NUM_CONSUMER_WARPS = 8;
constexpr int NUM_THREADS = (NUM_CONSUMER_WARPS + 1) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
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

// Template: the 8 consumer warps form a WM x WN grid, each warp owns MI x NJ blocks of 8 x 8: tile = (64 WM MI / 8 ...)
//   <2, 4, 8, 4>: 128 x 128 tiles (throughput shape: the Schur SYRK, the bulk GEMMs);
//   <4, 2, 4, 2>: 128 x 32 tiles  (latency shape for the small products on the Cholesky's critical chain: a quarter of
//                 the work per CTA, four times the CTAs; the 128-row P tile keeps in-place TRSM products safe - every
//                 CTA reads all 128 rows of its column block before it writes them).
template <int WM, int WN, int MI, int NJ>
__global__ void __launch_bounds__(NUM_THREADS, 1)
atb_upper_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapR,
                 const int4* __restrict__ tiles, int n_tiles, int nkb, int same_operand,
                 int64_t mrows, int64_t ncols, double* __restrict__ Cbase, int64_t ldc,
                 int64_t c_group_stride, double alpha, double beta) {
    static_assert(WM * WN == NUM_CONSUMER_WARPS, "8 consumer warps");
    constexpr int TSM = WM * MI * 8, TSN = WN * NJ * 8;          // tile rows / cols of C
    constexpr int P_BYTES = TSM * BK * 8, R_BYTES = TSN * BK * 8;
    constexpr int STAGE_BYTES = P_BYTES + R_BYTES;
    static_assert(TSM == BM && (TSN == BN || TSN == 32), "tile shapes the host knows");
    extern __shared__ uint8_t smem_raw[];
    uint32_t base = smem_u32(smem_raw);
    uint32_t tiles_smem = (base + 1023u) & ~1023u;
    uint32_t bar_full = tiles_smem + STAGES * STAGE_BYTES;
    uint32_t bar_empty = bar_full + STAGES * 8;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, NUM_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NUM_CONSUMER_WARPS) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int4 tile = tiles[t];
                bool single = TSM == TSN && same_operand && (tile.x == tile.y);
                for (int kb = 0; kb < nkb; kb++) {
                    mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                    uint32_t full = bar_full + stage * 8;
                    mbar_expect_tx(full, single ? P_BYTES : STAGE_BYTES);
                    uint32_t dst = tiles_smem + stage * STAGE_BYTES;
                    tma_load_2d(dst, &mapP, kb * BK, tile.x * TSM, full);
                    if (!single) tma_load_2d(dst + P_BYTES, &mapR, tile.z + kb * BK, tile.y * TSN, full);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        return;
    }

    // ===== DMMA consumers =====
    const int warp_m = warp / WN;      // MI * 8 rows of the tile
    const int warp_n = warp % WN;      // NJ * 8 cols of the tile
    const int g = lane >> 2;           // 0..7
    const int t4 = lane & 3;
    uint32_t koff[4];
#pragma unroll
    for (int s = 0; s < 4; s++) koff[s] = (uint32_t)((((2 * s + (t4 >> 1)) ^ g) << 4) + ((t4 & 1) << 3));
    const uint32_t offA = (uint32_t)((warp_m * MI * 8 + g) * 128);
    const uint32_t offB = (uint32_t)((warp_n * NJ * 8 + g) * 128);

    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int4 tile = tiles[t];
        bool single = TSM == TSN && same_operand && (tile.x == tile.y);
        double* __restrict__ C = Cbase + (int64_t)tile.w * c_group_stride;
        double acc[MI][NJ][2];
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kb = 0; kb < nkb; kb++) {
            mbar_wait(bar_full + stage * 8, phase);
            uint32_t sA = tiles_smem + stage * STAGE_BYTES;
            uint32_t sB = single ? sA : sA + P_BYTES;
            uint32_t pa = sA + offA, pb = sB + offB;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; i++) a[i] = lds64(pa + i * 1024 + koff[s]);
#pragma unroll
                for (int j = 0; j < NJ; j++) b[j] = lds64(pb + j * 1024 + koff[s]);
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + stage * 8);
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        }

        // epilogue: C[row, col] = alpha * acc + beta * C  (upper tiles; masked at the edges)
        const int64_t row0 = (int64_t)tile.x * TSM + warp_m * MI * 8 + g;
        const int64_t col0 = (int64_t)tile.y * TSN + warp_n * NJ * 8 + 2 * t4;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int64_t col = col0 + 8 * j + e;
                if (col >= ncols) continue;
                double* cp = C + col * ldc;
#pragma unroll
                for (int i = 0; i < MI; i++) {
                    int64_t row = row0 + 8 * i;
                    if (row < mrows) {
                        double v = alpha * acc[i][j][e];
                        if (beta != 0.0) v += beta * __ldcg(cp + row);      // through L2: see chol.cu (two concurrent streams)
                        cp[row] = v;
                    }
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess)
            throw HypError{"cuTensorMapEncodeTiled entry point not available"};
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

void make_map(CUtensorMap* map, const double* base, int64_t klen, int64_t ncols, int64_t ld, int box_rows = BM) {
    if (((uintptr_t)base & 15) || (ld & 1))
        throw HypError{"atb: operand must be 16-byte aligned with an even leading dimension"};
    cuuint64_t dims[2] = {(cuuint64_t)klen, (cuuint64_t)ncols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides,
                                 box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d) klen=%lld ncols=%lld ld=%lld",
                 (int)r, (long long)klen, (long long)ncols, (long long)ld);
        throw HypError{buf};
    }
}

// Tile = {row tile, col tile, k offset of the R operand, output group}.
// Upper-triangular tile list for an nt x nt tile grid, in 8-tile row groups so that the CTAs
// of one wave share row / column panels in L2.
std::vector<int4> build_upper_tiles(int nt) {
    std::vector<int4> tiles;
    tiles.reserve((size_t)nt * (nt + 1) / 2);
    const int GROUP = 8;
    for (int gi = 0; gi < nt; gi += GROUP)
        for (int tj = gi; tj < nt; tj++)
            for (int ti = gi; ti < std::min(gi + GROUP, tj + 1); ti++) tiles.push_back(make_int4(ti, tj, 0, 0));
    return tiles;
}

// full mt x nt grids for `ngroups` independent products; group g reads R at k offset g * kstride
std::vector<int4> build_full_tiles(int mt, int nt, int ngroups, int kstride) {
    std::vector<int4> tiles;
    tiles.reserve((size_t)mt * nt * ngroups);
    const int GROUP = 8;
    for (int g = 0; g < ngroups; g++)
        for (int gi = 0; gi < mt; gi += GROUP)
            for (int tj = 0; tj < nt; tj++)
                for (int ti = gi; ti < std::min(gi + GROUP, mt); ti++)
                    tiles.push_back(make_int4(ti, tj, g * kstride, g));
    return tiles;
}

// narrow (128 x 32) tiles of the latency shape: upper = every tile with an element on or above the diagonal
std::vector<int4> build_narrow_tiles(int kind, int mt, int nt32) {
    std::vector<int4> tiles;
    for (int tj = 0; tj < nt32; tj++)
        for (int ti = 0; ti < mt; ti++)
            if (kind != 0 || ti <= tj / 4) tiles.push_back(make_int4(ti, tj, 0, 0));
    return tiles;
}

struct TileCacheEntry {
    int device, kind, mt, nt, ngroups, kstride;
    int4* d_tiles;
    int n_tiles;
};
std::vector<TileCacheEntry> g_tile_cache;

// device tile lists are cached per (device, shape); they are tiny (16 B per tile)
void get_tiles(hyp_ctx* ctx, int kind, int mt, int nt, int ngroups, int kstride, int4** d_tiles,
               int* n_tiles) {
    for (auto& e : g_tile_cache)
        if (e.device == ctx->device && e.kind == kind && e.mt == mt && e.nt == nt &&
            e.ngroups == ngroups && e.kstride == kstride) {
            *d_tiles = e.d_tiles;
            *n_tiles = e.n_tiles;
            return;
        }
    std::vector<int4> tiles = kind >= 2 ? build_narrow_tiles(kind - 2, mt, nt)
                              : (kind == 0 ? build_upper_tiles(nt) : build_full_tiles(mt, nt, ngroups, kstride));
    TileCacheEntry e{ctx->device, kind, mt, nt, ngroups, kstride, nullptr, (int)tiles.size()};
    CUDA_TRY(cudaMalloc(&e.d_tiles, tiles.size() * sizeof(int4)));
    CUDA_TRY(cudaMemcpyAsync(e.d_tiles, tiles.data(), tiles.size() * sizeof(int4),
                             cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    g_tile_cache.push_back(e);
    *d_tiles = e.d_tiles;
    *n_tiles = e.n_tiles;
}

void launch_atb(hyp_ctx* ctx, int kind, const double* P, int64_t ldp, const double* R, int64_t ldr,
                int64_t klen, int64_t mrows, int64_t ncols, double* C, int64_t ldc, double alpha,
                double beta, int ngroups = 1, int64_t r_kstride = 0, int64_t c_group_stride = 0) {
    if (ncols <= 0 || mrows <= 0 || klen <= 0 || ngroups <= 0) return;
    if (ngroups > 1 && (r_kstride & 1)) throw HypError{"grouped GEMM: the k stride must be even (TMA coordinates are 16-byte aligned)"};
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(atb_upper_kernel<2, 4, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(atb_upper_kernel<4, 2, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    // latency shape (128 x 32 tiles) for the small products of the Cholesky's chain stream (chol.cu sets small_tiles)
    const bool narrow = ctx->small_tiles && ngroups == 1;
    int4* d_tiles = nullptr;
    int n_tiles = 0;
    if (narrow) get_tiles(ctx, kind + 2, ceil_div(mrows, BM), ceil_div(ncols, 32), 1, 0, &d_tiles, &n_tiles);
    else get_tiles(ctx, kind, ceil_div(mrows, BM), ceil_div(ncols, BN), ngroups, (int)r_kstride, &d_tiles, &n_tiles);
    CUtensorMap mapP, mapR;
    make_map(&mapP, P, klen, mrows, ldp);
    // grouped products read R at k offsets g * r_kstride; rows past a group's klen meet the
    // zero-filled out-of-range rows of P, so they contribute nothing
    make_map(&mapR, R, ngroups > 1 ? r_kstride * (ngroups - 1) + klen : klen, ncols, ldr, narrow ? 32 : BN);
    int nkb = ceil_div(klen, BK);
    int grid = std::min(n_tiles, ctx->sm_count);
    if (ctx->grid_cap > 0) grid = std::min(grid, ctx->grid_cap);
    int same = (kind == 0 && P == R && ldp == ldr) ? 1 : 0;
    cudaStream_t st = ctx->launch_stream ? ctx->launch_stream : ctx->stream;
    if (narrow)
        atb_upper_kernel<4, 2, 4, 2><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mapP, mapR, d_tiles, n_tiles, nkb, same, mrows, ncols, C,
                                                                            ldc, c_group_stride, alpha, beta);
    else
        atb_upper_kernel<2, 4, 8, 4><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mapP, mapR, d_tiles, n_tiles, nkb, same, mrows, ncols, C,
                                                                            ldc, c_group_stride, alpha, beta);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

// ---- simple CUDA-core GEMM (one-off products at load time, e.g. G * Ap_Q, qrchol.jl:154) ----
constexpr int ST = 32;
__global__ void __launch_bounds__(ST * ST / 4)
gemm_simple_kernel(int transA, int transB, int64_t M, int64_t N, int64_t Kd,
                   const double* __restrict__ A, int64_t lda, const double* __restrict__ B,
                   int64_t ldb, double* __restrict__ C, int64_t ldc) {
    __shared__ double sA[ST][ST + 1], sB[ST][ST + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
    const int64_t i0 = (int64_t)blockIdx.x * ST, j0 = (int64_t)blockIdx.y * ST;
    double acc[4] = {0, 0, 0, 0};
    for (int64_t k0 = 0; k0 < Kd; k0 += ST) {
        for (int r = ty; r < ST; r += 8) {
            // sA[m][k], sB[k][n]
            int64_t gi = i0 + (transA ? r : tx), gk = k0 + (transA ? tx : r);
            double va = 0.0;
            if (gi < M && gk < Kd) va = transA ? A[gk + gi * lda] : A[gi + gk * lda];
            if (transA) sA[r][tx] = va; else sA[tx][r] = va;
            int64_t gkb = k0 + (transB ? r : tx), gj = j0 + (transB ? tx : r);
            double vb = 0.0;
            if (gkb < Kd && gj < N) vb = transB ? B[gj + gkb * ldb] : B[gkb + gj * ldb];
            if (transB) sB[r][tx] = vb; else sB[tx][r] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < ST; k++) {
            double a = sA[tx][k];
#pragma unroll
            for (int u = 0; u < 4; u++) acc[u] += a * sB[k][ty + 8 * u];
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        int64_t gi = i0 + tx, gj = j0 + ty + 8 * u;
        if (gi < M && gj < N) C[gi + gj * ldc] = acc[u];
    }
}

}  // namespace

void hyp_atb_upper(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                   int64_t klen, int64_t ncols, double* C, int64_t ldc, double alpha, double beta) {
    launch_atb(ctx, 0, P, ldp, R, ldr, klen, ncols, ncols, C, ldc, alpha, beta);
}

void hyp_gemm_tn(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                 int64_t klen, int64_t mrows, int64_t ncols, double* C, int64_t ldc, double alpha,
                 double beta) {
    launch_atb(ctx, 1, P, ldp, R, ldr, klen, mrows, ncols, C, ldc, alpha, beta);
}

void hyp_gemm_tn_grouped(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                         int64_t klen, int64_t mrows, int64_t ncols, int ngroups, int64_t r_kstride,
                         double* C, int64_t ldc, int64_t c_group_stride, double alpha, double beta) {
    launch_atb(ctx, 1, P, ldp, R, ldr, klen, mrows, ncols, C, ldc, alpha, beta, ngroups, r_kstride,
               c_group_stride);
}

void hyp_gemm_simple(hyp_ctx* ctx, bool transA, bool transB, int64_t M, int64_t N, int64_t Kd,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                     int64_t ldc) {
    if (M <= 0 || N <= 0) return;
    dim3 grid(ceil_div(M, ST), ceil_div(N, ST));
    gemm_simple_kernel<<<grid, ST * ST / 4, 0, ctx->stream>>>(transA ? 1 : 0, transB ? 1 : 0, M, N, Kd,
                                                              A, lda, B, ldb, C, ldc);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

// FP64-accurate Schur SYRK on the 5th-generation tensor cores (tcgen05, kind::i8) by error-free
// slicing ("Ozaki scheme").  The default Schur SYRK of the library; syrk.cu (FP64 DMMA) is the
// alternative for two-operand products and for HYP_SCHUR_SYRK=dmma.
//
// Why: tcgen05.mma has no f64 kind, and the FP64 DMMA pipe is saturated by syrk.cu (96 % active,
// 34 TFLOP/s).  The only way past that roofline is to run the contraction on the int8 pipe
// (4.5 POP/s dense): every column of the K-major operand is scaled by a power of two and cut into
// signed 8-bit digits (int8 matrices).  Default (radix 256): S = 7 balanced digits
//     a = 2^e * sum_s 2^-(7 + 8 s) d_s,  d_s in [-128, 127]  (a carry pass removes the +128 rint can give),
// the 28 digit-pair products with s + t <= 6 are exact in int32 (|sum| <= 7 K 2^14, K <= 18688 rows per launch)
// and are recombined in FP64:  C_ij = 2^(e_i + e_j) * sum_d 2^-(14 + 8 d) * sum_{s+t=d} (D_s' D_t)_ij.
// HYP_OZAKI_RADIX=128 keeps the first scheme: S = 8 digits in [-64, 64], a = 2^e sum_s 2^-(6 + 7 s) d_s,
// 36 pair products with s + t <= 7, K <= 65472 rows per launch.  Both carry 56 bits below the column
// maximum; the chip runs this kernel at its power cap, so the 22 % fewer MMAs of radix 256 are what
// makes it faster (measured 101 -> 91 ms on C3), not memory traffic (the quad kernel below cuts the
// L2 -> SM bytes by a third and is not faster).  The order of the tile list decides which operand panel stays in
// L2: row-major (default) keeps the 256-row A panel resident and streams the B panels (134 GB of DRAM traffic per C3
// SYRK instead of 197 GB, 84.8 instead of 86.6 ms).
//
// This file: (1) the slicing kernels are in ozaki_slice_kernels.cuh, (2) a TMA + tcgen05 int8 TN GEMM
// (C_int32 = A' B, both operands K-major, SWIZZLE_128B, accumulator in TMEM), (3) a reference
// recombination used by the tests.
#include "common.cuh"
#include "ozaki_slice_kernels.cuh"

namespace {

constexpr int TM = 128, TN = 128;          // output tile
constexpr int TKB = 128;                   // k bytes per stage (one 128-byte swizzle row)
constexpr int I8_STAGES = 4;
constexpr int I8_TILE_BYTES = TM * TKB;    // 16 KB
constexpr int I8_STAGE_BYTES = 2 * I8_TILE_BYTES;
constexpr int I8_THREADS = 6 * 32;         // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int I8_SMEM = I8_STAGES * I8_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte
// offset (8-row group pitch = 1024 B) >> 4 in [32,46), version 1 in [46,48), layout type
// SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor, kind::i8: D = S32, A = B = signed int8, both K-major, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

// C (int32, ldc) [tile] = A' B over k in [0, K): A: K x M, B: K x N int8, K-major.
// One CTA per 128 x 128 output tile.
__global__ void __launch_bounds__(I8_THREADS, 1)
i8_gemm_tn_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int nkb,
                  int64_t M, int64_t N, int32_t* __restrict__ C, int64_t ldc) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t tiles = (base + 1023u) & ~1023u;
    const uint32_t bar_full = tiles + I8_STAGES * I8_STAGE_BYTES;
    const uint32_t bar_empty = bar_full + I8_STAGES * 8;
    const uint32_t bar_done = bar_empty + I8_STAGES * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tm = blockIdx.x, tn = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int s = 0; s < I8_STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // 128 TMEM columns for the 128 x 128 int32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_acc = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; kb++) {
                mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                const uint32_t full = bar_full + stage * 8;
                mbar_expect_tx(full, I8_STAGE_BYTES);
                const uint32_t dst = tiles + stage * I8_STAGE_BYTES;
                tma_load_2d(dst, &mapA, kb * TKB, tm * TM, full);
                tma_load_2d(dst + I8_TILE_BYTES, &mapB, kb * TKB, tn * TN, full);
                if (++stage == I8_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_i8(TM, TN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; kb++) {
                mbar_wait(bar_full + stage * 8, phase);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t sa = tiles + stage * I8_STAGE_BYTES;
                const uint32_t sb = sa + I8_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < TKB / 32; k++) {
                    // advance 32 bytes along K inside the 128-byte swizzle row
                    const uint64_t ad = make_desc_sw128(sa + k * 32);
                    const uint64_t bd = make_desc_sw128(sb + k * 32);
                    umma_i8(tmem_acc, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_empty + stage * 8);     // frees the stage when the MMAs retire
                if (++stage == I8_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            umma_commit(bar_done);
        }
    } else {
        // epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31
        mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int lg = warp & 3;
        const int64_t row = (int64_t)tm * TM + lg * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                  "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                  "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                  "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int64_t col = (int64_t)tn * TN + c0 + j;
                    if (col < N) C[row + col * ldc] = (int32_t)v[j];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(128));
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
        if (!p || qres != cudaDriverEntryPointSuccess) throw HypError{"cuTensorMapEncodeTiled not available"};
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

void make_map_i8(CUtensorMap* map, const int8_t* base, int64_t K, int64_t cols, int64_t ld) {
    if (((uintptr_t)base & 15) || (ld & 15)) throw HypError{"int8 operand must be 16-byte aligned with ld % 16 == 0"};
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {TKB, TM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw HypError{"cuTensorMapEncodeTiled (int8) failed"};
}

// ---- slicing: column exponents and the signed digit matrices: ozaki_slice_kernels.cuh (hypdev::colmax_kernel,
// slice_kernel, slice256_kernel), shared with the CPU-tier emulation ----
}  // namespace
namespace {
// =====================================================================================================
// Fused FP64-by-slicing SYRK:  C(upper 128-tiles) = alpha * A' A + beta * C,  A given by its digit slices.
// Per output tile two passes over K, each with four int32 accumulators (128 x 128 each) filling the
// 512 TMEM columns: pass 0 collects the digit pairs with s + t = 0..3, pass 1 those with s + t = 4..7.
// Stage = one MMA K step (32 bytes of k): up to 16 SWIZZLE_32B tiles of 128 rows x 32 B.
// =====================================================================================================
constexpr int OZ_S = 8;                       // digit slices
constexpr int OZ_KB = 32;                     // k bytes per stage = MMA K
constexpr int OZ_TILE = TM * OZ_KB;           // 4 KB
constexpr int OZ_STAGE = 2 * OZ_S * OZ_TILE;  // 64 KB: A-side slices then B-side slices
constexpr int OZ_STAGES = 3;
constexpr int OZ_SMEM = OZ_STAGES * OZ_STAGE + 1024 + 256;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// K-major SWIZZLE_32B descriptor: rows of 32 B, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t make_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}

__global__ void __launch_bounds__(I8_THREADS, 1)
ozaki_syrk_kernel(const __grid_constant__ CUtensorMap mapD4, const __grid_constant__ CUtensorMap mapD8,
                  const int4* __restrict__ tiles, int n_tiles,
                  int k0, int nkb, const int* __restrict__ expo, int64_t ncols, double* __restrict__ C, int64_t ldc,
                  double alpha, double beta, int dbg_no_tma) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    const uint32_t bar_full = stg + OZ_STAGES * OZ_STAGE;
    const uint32_t bar_empty = bar_full + OZ_STAGES * 8;
    const uint32_t bar_tfull = bar_empty + OZ_STAGES * 8;
    const uint32_t bar_tempty = bar_tfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < OZ_STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int4 tile = tiles[t];
                const bool single = tile.x == tile.y;
                for (int pass = 0; pass < 2; pass++) {
                    // pass 0: a stage carries TWO k steps of the 4 leading slices (2 x 32 KB);
                    // pass 1: one k step of all 8 slices (64 KB).  One TMA box per side and k step.
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const uint32_t bytes = pass == 0 ? (single ? 2u : 4u) * 4 * OZ_TILE : (single ? 1u : 2u) * OZ_S * OZ_TILE;
                    for (int st = 0; st < nst; st++) {
                        mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                        const uint32_t full = bar_full + stage * 8;
                        if (dbg_no_tma) {   // micro-benchmark of the MMA rate: no loads, stale operands
                            mbar_arrive(full);
                            if (++stage == OZ_STAGES) {
                                stage = 0;
                                phase ^= 1u;
                            }
                            continue;
                        }
                        mbar_expect_tx(full, bytes);
                        const uint32_t dst = stg + stage * OZ_STAGE;
                        if (pass == 0) {
                            for (int h = 0; h < 2; h++) {
                                const int kc = k0 + (2 * st + h) * OZ_KB;
                                tma_load_3d(dst + h * (OZ_STAGE / 2), &mapD4, kc, tile.x * TM, 0, full);
                                if (!single)
                                    tma_load_3d(dst + h * (OZ_STAGE / 2) + 4 * OZ_TILE, &mapD4, kc, tile.y * TN, 0, full);
                            }
                        } else {
                            const int kc = k0 + st * OZ_KB;
                            tma_load_3d(dst, &mapD8, kc, tile.x * TM, 0, full);
                            if (!single) tma_load_3d(dst + OZ_S * OZ_TILE, &mapD8, kc, tile.y * TN, 0, full);
                        }
                        if (++stage == OZ_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc_i8(TM, TN);
            int stage = 0;
            uint32_t phase = 0, item = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int4 tile = tiles[t];
                const bool single = tile.x == tile.y;
                for (int pass = 0; pass < 2; pass++, item++) {
                    // the epilogue must have drained the accumulators of the previous item
                    if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const int d0 = pass * 4;
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const int nhalf = pass == 0 ? 2 : 1;
                    const int smax = pass == 0 ? 3 : OZ_S - 1;
                    const uint32_t b_off = pass == 0 ? 4 * OZ_TILE : OZ_S * OZ_TILE;
                    for (int st = 0; st < nst; st++) {
                        mbar_wait(bar_full + stage * 8, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        for (int h = 0; h < nhalf; h++) {
                            const uint32_t sa = stg + stage * OZ_STAGE + h * (OZ_STAGE / 2);
                            const uint32_t sb = single ? sa : sa + b_off;
                            for (int s = 0; s <= smax; s++) {
                                const uint64_t ad = make_desc_sw32(sa + s * OZ_TILE);
                                // t ranges over d0 - s .. d0 + 3 - s, clipped to [0, 7]
                                int tlo = d0 - s, thi = d0 + 3 - s;
                                if (tlo < 0) tlo = 0;
                                if (thi > OZ_S - 1) thi = OZ_S - 1;
                                for (int tt = tlo; tt <= thi; tt++) {
                                    const uint64_t bd = make_desc_sw32(sb + tt * OZ_TILE);
                                    const int g = s + tt - d0;
                                    // every group of a pass is first touched by the pairs with s = 0
                                    const uint32_t accum = (st > 0 || h > 0 || s > 0) ? 1u : 0u;
                                    umma_i8(tmem0 + (uint32_t)(g * TN), ad, bd, idesc, accum);
                                }
                            }
                        }
                        umma_commit(bar_empty + stage * 8);
                        if (++stage == OZ_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit(bar_tfull);
                }
            }
        }
    } else {
        // ===== epilogue: int32 groups -> FP64, scaled, into C =====
        const int lg = warp & 3;
        uint32_t item = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int4 tile = tiles[t];
            const int64_t row = (int64_t)tile.x * TM + lg * 32 + lane;
            const double rs = (row < ncols) ? alpha * ldexp(1.0, expo[row]) : 0.0;
            for (int pass = 0; pass < 2; pass++, item++) {
                mbar_wait(bar_tfull, item & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int d0 = pass * 4;
                const double g0 = ldexp(1.0, -(12 + 7 * d0)), g1 = ldexp(1.0, -(12 + 7 * (d0 + 1))),
                             g2 = ldexp(1.0, -(12 + 7 * (d0 + 2))), g3 = ldexp(1.0, -(12 + 7 * (d0 + 3)));
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 16) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t taddr = tmem0 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * TN + c0);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]),
                              "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]),
                              "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]),
                              "=r"(v[g][15])
                            : "r"(taddr));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (row < ncols) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tile.y * TN + c0 + j;
                            if (col < ncols) {
                                // smallest terms first
                                double x = (double)(int32_t)v[3][j] * g3;
                                x += (double)(int32_t)v[2][j] * g2;
                                x += (double)(int32_t)v[1][j] * g1;
                                x += (double)(int32_t)v[0][j] * g0;
                                x *= rs * ldexp(1.0, expo[col]);
                                double* cp = C + row + col * ldc;
                                if (pass == 0) *cp = (beta == 0.0) ? x : (x + beta * *cp);
                                else *cp += x;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty);
            }
        }
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// ---- 2 x 2 cluster variant: TMA multicast halves the L2 -> shared-memory traffic ------------------
// The single-CTA kernel is bound by L2 bandwidth (every CTA streams all digit slices of its row and
// column panels: ~7.5 TB/s measured).  Here a cluster of 4 CTAs computes the 2 x 2 block of output
// tiles (2 SI + ri, 2 SJ + rj): the A-side slices of tile row I are needed by both CTAs of that row,
// the B-side slices of tile column J by both CTAs of that column, so every CTA loads HALF of the
// slices of each side and multicasts them to its row / column peer.  Stage release is cluster-wide:
// a CTA refills a stage only after itself, its row peer and its column peer have retired the MMAs
// that read it (tcgen05.commit multicast onto the three empty barriers).
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                               uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(I8_THREADS, 1)
ozaki_syrk_cluster_kernel(const __grid_constant__ CUtensorMap mapD2, const __grid_constant__ CUtensorMap mapD4,
                          const int2* __restrict__ supers, int n_super, int k0, int nkb,
                          const double* __restrict__ dscale, int64_t ncols,
                          double* __restrict__ C, int64_t ldc, double alpha, double beta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    const uint32_t bar_full = stg + OZ_STAGES * OZ_STAGE;
    const uint32_t bar_empty = bar_full + OZ_STAGES * 8;
    const uint32_t bar_tfull = bar_empty + OZ_STAGES * 8;
    const uint32_t bar_tempty = bar_tfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank, cid, ncl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
    const int ri = (int)(crank >> 1), rj = (int)(crank & 1);
    const uint16_t mask_row = (uint16_t)((1u << (ri * 2)) | (1u << (ri * 2 + 1)));        // same tile row
    const uint16_t mask_col = (uint16_t)((1u << rj) | (1u << (2 + rj)));                  // same tile column
    const uint16_t mask_rel = (uint16_t)(mask_row | mask_col);                            // me + both peers

    if (threadIdx.x == 0) {
        for (int s = 0; s < OZ_STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, 3);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();                       // every peer's barriers exist before any remote arrival
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;

    // super tile st -> (SI, SJ), SI <= SJ, from the host-built list (square blocks for L2 locality)
    auto super_coords = [&](int st, int& SI, int& SJ) {
        const int2 c = supers[st];
        SI = c.x;
        SJ = c.y;
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int st = (int)cid; st < n_super; st += (int)ncl) {
                int SI, SJ;
                super_coords(st, SI, SJ);
                const int tI = 2 * SI + ri, tJ = 2 * SJ + rj;
                for (int pass = 0; pass < 2; pass++) {
                    const int ns = pass == 0 ? 4 : OZ_S;
                    const int hs = ns / 2;                                 // slices this CTA loads per side
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const uint32_t bytes = (pass == 0 ? 2u : 1u) * 2u * ns * OZ_TILE;
                    const CUtensorMap* mp = pass == 0 ? &mapD2 : &mapD4;
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                        const uint32_t full = bar_full + stage * 8;
                        mbar_expect_tx(full, bytes);
                        const uint32_t dst = stg + stage * OZ_STAGE;
                        const int nh = pass == 0 ? 2 : 1;
                        for (int h = 0; h < nh; h++) {
                            const int kc = k0 + (pass == 0 ? (2 * it + h) : it) * OZ_KB;
                            const uint32_t d0s = dst + h * (OZ_STAGE / 2);
                            // A side: my half of the slices of tile row tI -> me and my row peer
                            tma_load_3d_mc(d0s + (rj * hs) * OZ_TILE, mp, kc, tI * TM, rj * hs, full, mask_row);
                            // B side: my half of the slices of tile column tJ -> me and my column peer
                            tma_load_3d_mc(d0s + (ns + ri * hs) * OZ_TILE, mp, kc, tJ * TN, ri * hs, full, mask_col);
                        }
                        if (++stage == OZ_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_i8(TM, TN);
            int stage = 0;
            uint32_t phase = 0, item = 0;
            for (int st = (int)cid; st < n_super; st += (int)ncl) {
                for (int pass = 0; pass < 2; pass++, item++) {
                    if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const int d0 = pass * 4;
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const int nhalf = pass == 0 ? 2 : 1;
                    const int smax = pass == 0 ? 3 : OZ_S - 1;
                    const uint32_t b_off = pass == 0 ? 4 * OZ_TILE : OZ_S * OZ_TILE;
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_full + stage * 8, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        for (int h = 0; h < nhalf; h++) {
                            const uint32_t sa = stg + stage * OZ_STAGE + h * (OZ_STAGE / 2);
                            const uint32_t sb = sa + b_off;
                            for (int s = 0; s <= smax; s++) {
                                const uint64_t ad = make_desc_sw32(sa + s * OZ_TILE);
                                int tlo = d0 - s, thi = d0 + 3 - s;
                                if (tlo < 0) tlo = 0;
                                if (thi > OZ_S - 1) thi = OZ_S - 1;
                                for (int tt = tlo; tt <= thi; tt++) {
                                    const uint64_t bd = make_desc_sw32(sb + tt * OZ_TILE);
                                    const int g = s + tt - d0;
                                    const uint32_t accum = (it > 0 || h > 0 || s > 0) ? 1u : 0u;
                                    umma_i8(tmem0 + (uint32_t)(g * TN), ad, bd, idesc, accum);
                                }
                            }
                        }
                        umma_commit_mc(bar_empty + stage * 8, mask_rel);     // release in all three CTAs
                        if (++stage == OZ_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit(bar_tfull);
                }
            }
        }
    } else {
        const int lg = warp & 3;
        uint32_t item = 0;
        for (int st = (int)cid; st < n_super; st += (int)ncl) {
            int SI, SJ;
            super_coords(st, SI, SJ);
            const int tI = 2 * SI + ri, tJ = 2 * SJ + rj;
            const int64_t row = (int64_t)tI * TM + lg * 32 + lane;
            const bool store = tI <= tJ;                 // the lower tile of a diagonal super tile is not needed
            const double rs = (row < ncols) ? alpha * dscale[row] : 0.0;
            for (int pass = 0; pass < 2; pass++, item++) {
                mbar_wait(bar_tfull, item & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int d0 = pass * 4;
                const double g0 = ldexp(1.0, -(12 + 7 * d0)), g1 = ldexp(1.0, -(12 + 7 * (d0 + 1))),
                             g2 = ldexp(1.0, -(12 + 7 * (d0 + 2))), g3 = ldexp(1.0, -(12 + 7 * (d0 + 3)));
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 16) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t taddr = tmem0 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * TN + c0);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]),
                              "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]),
                              "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]),
                              "=r"(v[g][15])
                            : "r"(taddr));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (store && row < ncols) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + j;
                            if (col < ncols) {
                                double x = (double)(int32_t)v[3][j] * g3;
                                x += (double)(int32_t)v[2][j] * g2;
                                x += (double)(int32_t)v[1][j] * g1;
                                x += (double)(int32_t)v[0][j] * g0;
                                x *= rs * dscale[col];
                                double* cp = C + row + col * ldc;
                                if (pass == 0) *cp = (beta == 0.0) ? x : (x + beta * *cp);
                                else *cp += x;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();                       // no CTA leaves while a peer may still signal its barriers
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// ---- CTA-pair variant (cta_group::2, M = 256): the production kernel --------------------------------
// The 128 x 128 x 32 int8 MMA with both operands in shared memory reads 8 KB per 64 cycles - exactly
// the 128 B/cycle shared-memory port, which (together with the TMA writes) bounds the kernels above.
// A CTA pair sharing one tile column J issues M = 256 MMAs instead: each SM reads its own 128-row A
// tile but only HALF (64 rows) of the B tile, and each CTA loads only that half: 25 % fewer operand
// reads and TMA writes per MAC, and a 48 KB stage (4 stages instead of 3).
// CTA r of the pair owns output tile (2 P + r, J).  Both CTAs run a TMA producer whose loads signal
// the LEADER's full barrier; the leader's MMA thread issues tcgen05.mma.cta_group::2 and commits
// (multicast) to both CTAs' empty / accumulator-full barriers; both CTAs run the epilogue on their own
// TMEM half and report back to the leader's accumulator-empty barrier.
constexpr int OZP_STAGE = OZ_S * OZ_TILE + OZ_S * (OZ_TILE / 2);   // 48 KB: A slices, then B half slices
constexpr int OZP_STAGES = 4;
constexpr int OZP_SMEM = OZP_STAGES * OZP_STAGE + 1024 + 256;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                uint32_t leader_bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// ---- MMA issue loop of the CTA-pair kernel for seven radix-256 digits (28 pair products) --------------------------
// The first version ran the generic (s, t) loops in ONE thread: 35 SASS instructions per tcgen05.mma (descriptor
// construction, R2UR moves into the uniform registers UTCIMMA reads), i.e. 143 cycles per MMA against the 64-cycle
// floor of a 256 x 128 x 32 int8 MMA - the issuing thread, not the tensor pipe or the operand traffic, bounded the kernel
// (profiles/r02_syrk_probe_mma_vs_tma.json: the MMA stream alone needed 88 % of the kernel time).  Here the WHOLE warp
// runs the loop, so every value is warp-uniform and stays in uniform registers; the digit pairs of a pass are unrolled
// at compile time, each descriptor is the stage's base descriptor plus a constant, and one elected lane issues.
template <int PASS>
__device__ __forceinline__ void issue_pass_r256(uint64_t dbase, uint32_t tmem0, uint32_t idesc, int nst, uint32_t bar_full,
                                                uint32_t bar_empty, uint32_t bar_tfull, int& stage, uint32_t& phase,
                                                bool issuer, uint16_t empty_mask = 3, uint16_t tfull_mask = 3) {
    constexpr int NSL = 7;
    constexpr int D0 = PASS * 4;
    constexpr int NS = PASS == 0 ? 4 : OZ_S;          // A slice slots in front of the B half slices
    constexpr int NHALF = PASS == 0 ? 2 : 1;
    constexpr int SMAX = PASS == 0 ? 3 : NSL - 1;
    for (int it = 0; it < nst; it++) {
        mbar_wait(bar_full + stage * 8, phase);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t sd = dbase + (uint64_t)((uint32_t)(stage * OZP_STAGE) >> 4);
        const uint32_t first = it > 0 ? 1u : 0u;
#pragma unroll
        for (int h = 0; h < NHALF; h++) {
#pragma unroll
            for (int sl = 0; sl <= SMAX; sl++) {
                constexpr int dummy = 0;
                (void)dummy;
                const int tlo = D0 - sl > 0 ? D0 - sl : 0;
                const int thi = (D0 + 3 - sl) < (NSL - 1 - sl) ? (D0 + 3 - sl) : (NSL - 1 - sl);
#pragma unroll
                for (int tt = 0; tt < NSL; tt++) {
                    if (tt < tlo || tt > thi) continue;
                    const uint32_t offA = (uint32_t)(h * (OZP_STAGE / 2) + sl * OZ_TILE) >> 4;
                    const uint32_t offB = (uint32_t)(h * (OZP_STAGE / 2) + NS * OZ_TILE + tt * (OZ_TILE / 2)) >> 4;
                    const uint32_t accum = (h > 0 || sl > 0) ? 1u : first;
                    if (issuer) umma_i8_2sm(tmem0 + (uint32_t)((sl + tt - D0) * TN), sd + offA, sd + offB, idesc, accum);
                }
            }
        }
        if (issuer) umma_commit_2sm(bar_empty + stage * 8, empty_mask);       // frees the stage in the CTAs that filled it
        if (++stage == OZP_STAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
    if (issuer) umma_commit_2sm(bar_tfull, tfull_mask);                        // accumulators ready in both CTAs of the pair
}

__global__ void __launch_bounds__(I8_THREADS, 1)
ozaki_syrk_pair_kernel(const __grid_constant__ CUtensorMap mapA4, const __grid_constant__ CUtensorMap mapA8,
                       const __grid_constant__ CUtensorMap mapB4, const __grid_constant__ CUtensorMap mapB8,
                       const int2* __restrict__ pairs, int n_pairs, int k0, int nkb,
                       const double* __restrict__ dscale, int64_t ncols, double* __restrict__ C, int64_t ldc,
                       double alpha, double beta, int nsl, int wbits_probe) {
    // bits 8.. of the last argument: profiling probes (HYP_OZAKI_PROBE; results are garbage, timing only):
    //   1 = no TMA loads (the MMA issuer runs on stale shared memory), 2 = no MMAs (loads + epilogue only)
    const int wbits = wbits_probe & 0xff, probe = (wbits_probe >> 8) & 3;
    // nsl = number of digit slices that enter the product (7 or 8): pairs with s + t <= nsl - 1;
    // wbits = bits per digit (7: radix 128, 8: radix 256): digit s has weight 2^-(wbits - 1 + wbits s)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    const uint32_t bar_full = stg + OZP_STAGES * OZP_STAGE;
    const uint32_t bar_empty = bar_full + OZP_STAGES * 8;
    const uint32_t bar_tfull = bar_empty + OZP_STAGES * 8;
    const uint32_t bar_tempty = bar_tfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank, cid, ncl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < OZP_STAGES; s++) {
            mbar_init(bar_full + s * 8, 1);      // leader: its own expect_tx arrival (+ the bytes of both CTAs)
            mbar_init(bar_empty + s * 8, 1);     // one multicast commit from the leader
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);                // 4 epilogue warps of each CTA (used in the leader only)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my A tile and my half of the B tile =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
                const int2 pr = pairs[pi];
                const int tI = 2 * pr.x + (int)crank, tJ = pr.y;
                for (int pass = 0; pass < 2; pass++) {
                    const int ns = pass == 0 ? 4 : OZ_S;
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const int nh = pass == 0 ? 2 : 1;
                    const CUtensorMap* ma = pass == 0 ? &mapA4 : &mapA8;
                    const CUtensorMap* mb = pass == 0 ? &mapB4 : &mapB8;
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                        const uint32_t full_leader = mapa_u32(bar_full + stage * 8, 0);
                        // pass 1 loads only the nsl slices that enter the product (host passes nsl-slice boxes)
                        if (probe == 1) {
                            if (leader) mbar_arrive(bar_full + stage * 8);
                            if (++stage == OZP_STAGES) {
                                stage = 0;
                                phase ^= 1u;
                            }
                            continue;
                        }
                        if (leader)
                            mbar_expect_tx(bar_full + stage * 8,
                                           pass == 0 ? 2u * OZP_STAGE : 2u * (uint32_t)nsl * (OZ_TILE + OZ_TILE / 2));
                        const uint32_t dst = stg + stage * OZP_STAGE;
                        for (int h = 0; h < nh; h++) {
                            const int kc = k0 + (pass == 0 ? (2 * it + h) : it) * OZ_KB;
                            const uint32_t d0s = dst + h * (OZP_STAGE / 2);
                            tma_load_3d_2sm(d0s, ma, kc, tI * TM, 0, full_leader);
                            tma_load_3d_2sm(d0s + ns * OZ_TILE, mb, kc, tJ * TN + (int)crank * (TN / 2), 0, full_leader);
                        }
                        if (++stage == OZP_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        static_assert(true, "");
        if (leader && nsl == 7 && probe != 2 && !(wbits_probe & 0x4000)) {
            // radix-256 digits: warp-uniform issue loop, one elected lane issues (issue_pass_r256)
            const uint32_t idesc = make_idesc_i8(2 * TM, TN);
            const uint64_t dbase = make_desc_sw32(stg);
            const bool issuer = elect_one_sync();
            int stage = 0;
            uint32_t phase = 0, item = 0;
            for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
                for (int pass = 0; pass < 2; pass++, item++) {
                    if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (pass == 0)
                        issue_pass_r256<0>(dbase, tmem0, idesc, (nkb + 1) / 2, bar_full, bar_empty, bar_tfull, stage, phase, issuer);
                    else
                        issue_pass_r256<1>(dbase, tmem0, idesc, nkb, bar_full, bar_empty, bar_tfull, stage, phase, issuer);
                }
            }
        } else
        if (leader && lane == 0) {
            const uint32_t idesc = make_idesc_i8(2 * TM, TN);
            int stage = 0;
            uint32_t phase = 0, item = 0;
            for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
                for (int pass = 0; pass < 2; pass++, item++) {
                    if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const int d0 = pass * 4;
                    const int ns = pass == 0 ? 4 : OZ_S;
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const int nhalf = pass == 0 ? 2 : 1;
                    const int smax = pass == 0 ? 3 : nsl - 1;
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_full + stage * 8, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        for (int h = 0; h < (probe == 2 ? 0 : nhalf); h++) {
                            const uint32_t sa = stg + stage * OZP_STAGE + h * (OZP_STAGE / 2);
                            const uint32_t sb = sa + ns * OZ_TILE;
                            for (int s = 0; s <= smax; s++) {
                                const uint64_t ad = make_desc_sw32(sa + s * OZ_TILE);
                                int tlo = d0 - s, thi = d0 + 3 - s;
                                if (tlo < 0) tlo = 0;
                                if (thi > nsl - 1 - s) thi = nsl - 1 - s;     // s + t <= nsl - 1
                                for (int tt = tlo; tt <= thi; tt++) {
                                    const uint64_t bd = make_desc_sw32(sb + tt * (OZ_TILE / 2));
                                    const int g = s + tt - d0;
                                    const uint32_t accum = (it > 0 || h > 0 || s > 0) ? 1u : 0u;
                                    umma_i8_2sm(tmem0 + (uint32_t)(g * TN), ad, bd, idesc, accum);
                                }
                            }
                        }
                        umma_commit_2sm(bar_empty + stage * 8, 3);       // frees the stage in both CTAs
                        if (++stage == OZP_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit_2sm(bar_tfull, 3);                        // accumulators ready in both CTAs
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs, own TMEM half) =====
        const int lg = warp & 3;
        const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
        uint32_t item = 0;
        for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
            const int2 pr = pairs[pi];
            const int tI = 2 * pr.x + (int)crank, tJ = pr.y;
            const int64_t row = (int64_t)tI * TM + lg * 32 + lane;
            const bool store = tI <= tJ;
            const double rs = (row < ncols) ? alpha * dscale[row] : 0.0;
            for (int pass = 0; pass < 2; pass++, item++) {
                mbar_wait(bar_tfull, item & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int d0 = pass * 4;
                // a digit-sum group beyond the last slice pair (d0 + 3 > nsl - 1) was never written: weight 0
                const int wb0 = 2 * (wbits - 1);
                const double g0 = ldexp(1.0, -(wb0 + wbits * d0)), g1 = ldexp(1.0, -(wb0 + wbits * (d0 + 1))),
                             g2 = ldexp(1.0, -(wb0 + wbits * (d0 + 2))),
                             g3 = (d0 + 3 <= nsl - 1) ? ldexp(1.0, -(wb0 + wbits * (d0 + 3))) : 0.0;
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 16) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t taddr = tmem0 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * TN + c0);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]),
                              "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]),
                              "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]),
                              "=r"(v[g][15])
                            : "r"(taddr));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (store && row < ncols) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + j;
                            if (col < ncols) {
                                double x = (double)(int32_t)v[3][j] * g3;
                                x += (double)(int32_t)v[2][j] * g2;
                                x += (double)(int32_t)v[1][j] * g1;
                                x += (double)(int32_t)v[0][j] * g0;
                                x *= rs * dscale[col];
                                double* cp = C + row + col * ldc;
                                if (pass == 0) *cp = (beta == 0.0) ? x : (x + beta * *cp);
                                else *cp += x;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// ---- quad variant: two CTA pairs of a 4-CTA cluster share their A tiles ---------------------------------------
// Quad (P, Jq): tile rows 2P, 2P + 1 and tile columns 2Jq, 2Jq + 1.  Pair p of the cluster (ranks 2p, 2p + 1) owns
// column 2Jq + p and runs exactly the pair kernel above (cta_group::2, M = 256), but the A tile of row 2P + r is
// needed by CTA r of BOTH pairs: each of the two loads half of its slices and TMA-multicasts them to the other,
// which cuts the L2 -> SM bytes per MMA by a third (A is two thirds of a stage).  Barriers: the multicast signals
// the full barrier of each destination CTA itself, so the second CTA of a pair forwards "my A tile has landed"
// to its leader with a remote arrive; a stage is refilled only after BOTH pair leaders have retired the MMAs
// that read it (their commits are multicast to all four CTAs).
__global__ void __launch_bounds__(I8_THREADS, 1)
ozaki_syrk_quad_kernel(const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA4,
                       const __grid_constant__ CUtensorMap mapB4, const __grid_constant__ CUtensorMap mapB8,
                       const int2* __restrict__ quads, int n_quads, int k0, int nkb,
                       const double* __restrict__ dscale, int64_t ncols, double* __restrict__ C, int64_t ldc,
                       double alpha, double beta, int nsl, int wbits_flags) {
    const int wbits = wbits_flags & 0xff;
    // nsl = number of digit slices that enter the product (7 or 8): pairs with s + t <= nsl - 1;
    // wbits = bits per digit (7: radix 128, 8: radix 256): digit s has weight 2^-(wbits - 1 + wbits s)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    const uint32_t bar_full = stg + OZP_STAGES * OZP_STAGE;
    const uint32_t bar_empty = bar_full + OZP_STAGES * 8;
    const uint32_t bar_tfull = bar_empty + OZP_STAGES * 8;
    const uint32_t bar_tempty = bar_tfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank, cid, ncl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
    const uint32_t pr = crank >> 1, r = crank & 1u;        // pair inside the quad, CTA inside the pair
    const bool leader = r == 0;
    const uint32_t leader_rank = crank & ~1u;
    const uint16_t mask_a = (uint16_t)((1u << r) | (1u << (2 + r)));      // the CTAs that hold tile row 2P + r
    const uint16_t mask_pair = (uint16_t)(3u << (2 * pr));                // my CTA pair

    if (threadIdx.x == 0) {
        for (int s = 0; s < OZP_STAGES; s++) {
            // leader: its own expect_tx arrival + the peer's forwarded "my A tile has landed";
            // peer: its own expect_tx arrival (A bytes only; its B half signals the leader directly)
            mbar_init(bar_full + s * 8, leader ? 2 : 1);
            mbar_init(bar_empty + s * 8, 2);     // one multicast commit from each of the two pair leaders
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);                // 4 epilogue warps of each CTA (used in the leader only)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my A tile and my half of the B tile =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
                const int2 qd = quads[qi];
                const int tI = 2 * qd.x + (int)r, tJ = 2 * qd.y + (int)pr;
                for (int pass = 0; pass < 2; pass++) {
                    const int ns = pass == 0 ? 4 : OZ_S;
                    const int hs = ns / 2;                                 // A slices this CTA loads (and multicasts)
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const int nh = pass == 0 ? 2 : 1;
                    const CUtensorMap* ma = pass == 0 ? &mapA2 : &mapA4;
                    const CUtensorMap* mb = pass == 0 ? &mapB4 : &mapB8;
                    const uint32_t bytes_a = (uint32_t)(nh * ns) * OZ_TILE;
                    const uint32_t bytes_bh = (uint32_t)(nh * ns) * (OZ_TILE / 2);
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_empty + stage * 8, phase ^ 1u);
                        const uint32_t full = bar_full + stage * 8;
                        const uint32_t full_leader = mapa_u32(full, leader_rank);
                        mbar_expect_tx(full, leader ? bytes_a + 2u * bytes_bh : bytes_a);
                        const uint32_t dst = stg + stage * OZP_STAGE;
                        for (int h = 0; h < nh; h++) {
                            const int kc = k0 + (pass == 0 ? (2 * it + h) : it) * OZ_KB;
                            const uint32_t d0s = dst + h * (OZP_STAGE / 2);
                            // A: my half of the slices of tile row tI -> me and the same-row CTA of the other pair
                            tma_load_3d_mc(d0s + ((int)pr * hs) * OZ_TILE, ma, kc, tI * TM, (int)pr * hs, full, mask_a);
                            // B: my 64-row half of tile column tJ (private to my pair), signals the pair leader
                            tma_load_3d_2sm(d0s + ns * OZ_TILE, mb, kc, tJ * TN + (int)r * (TN / 2), 0, full_leader);
                        }
                        if (++stage == OZP_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && !leader) {
        // ===== forwarder (second CTA of a pair): tells the pair leader when my A tile of a stage has landed =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
                for (int pass = 0; pass < 2; pass++) {
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_full + stage * 8, phase);
                        mbar_arrive_cluster(mapa_u32(bar_full + stage * 8, leader_rank));
                        if (++stage == OZP_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && nsl == 7 && !(wbits_flags & 0x4000)) {
        // ===== MMA issuer (leader CTA of each pair), radix-256 digits: warp-uniform loop, elected lane issues =====
        const uint32_t idesc = make_idesc_i8(2 * TM, TN);
        const uint64_t dbase = make_desc_sw32(stg);
        const bool issuer = elect_one_sync();
        int stage = 0;
        uint32_t phase = 0, item = 0;
        for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
            for (int pass = 0; pass < 2; pass++, item++) {
                if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (pass == 0)
                    issue_pass_r256<0>(dbase, tmem0, idesc, (nkb + 1) / 2, bar_full, bar_empty, bar_tfull, stage, phase, issuer, 0xF,
                                       mask_pair);
                else
                    issue_pass_r256<1>(dbase, tmem0, idesc, nkb, bar_full, bar_empty, bar_tfull, stage, phase, issuer, 0xF, mask_pair);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA of each pair) =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc_i8(2 * TM, TN);
            int stage = 0;
            uint32_t phase = 0, item = 0;
            for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
                for (int pass = 0; pass < 2; pass++, item++) {
                    if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const int d0 = pass * 4;
                    const int ns = pass == 0 ? 4 : OZ_S;
                    const int nst = pass == 0 ? (nkb + 1) / 2 : nkb;
                    const int nhalf = pass == 0 ? 2 : 1;
                    const int smax = pass == 0 ? 3 : nsl - 1;
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bar_full + stage * 8, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        for (int h = 0; h < nhalf; h++) {
                            const uint32_t sa = stg + stage * OZP_STAGE + h * (OZP_STAGE / 2);
                            const uint32_t sb = sa + ns * OZ_TILE;
                            for (int s = 0; s <= smax; s++) {
                                const uint64_t ad = make_desc_sw32(sa + s * OZ_TILE);
                                int tlo = d0 - s, thi = d0 + 3 - s;
                                if (tlo < 0) tlo = 0;
                                if (thi > nsl - 1 - s) thi = nsl - 1 - s;     // s + t <= nsl - 1
                                for (int tt = tlo; tt <= thi; tt++) {
                                    const uint64_t bd = make_desc_sw32(sb + tt * (OZ_TILE / 2));
                                    const int g = s + tt - d0;
                                    const uint32_t accum = (it > 0 || h > 0 || s > 0) ? 1u : 0u;
                                    umma_i8_2sm(tmem0 + (uint32_t)(g * TN), ad, bd, idesc, accum);
                                }
                            }
                        }
                        umma_commit_2sm(bar_empty + stage * 8, 0xF);     // one of the two releases of this stage, in all four CTAs
                        if (++stage == OZP_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit_2sm(bar_tfull, mask_pair);                // accumulators ready in both CTAs of my pair
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs, own TMEM half) =====
        const int lg = warp & 3;
        const uint32_t tempty_leader = mapa_u32(bar_tempty, leader_rank);
        uint32_t item = 0;
        for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
            const int2 qd = quads[qi];
            const int tI = 2 * qd.x + (int)r, tJ = 2 * qd.y + (int)pr;
            const int64_t row = (int64_t)tI * TM + lg * 32 + lane;
            const bool store = tI <= tJ;
            const double rs = (row < ncols) ? alpha * dscale[row] : 0.0;
            for (int pass = 0; pass < 2; pass++, item++) {
                mbar_wait(bar_tfull, item & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int d0 = pass * 4;
                // a digit-sum group beyond the last slice pair (d0 + 3 > nsl - 1) was never written: weight 0
                const int wb0 = 2 * (wbits - 1);
                const double g0 = ldexp(1.0, -(wb0 + wbits * d0)), g1 = ldexp(1.0, -(wb0 + wbits * (d0 + 1))),
                             g2 = ldexp(1.0, -(wb0 + wbits * (d0 + 2))),
                             g3 = (d0 + 3 <= nsl - 1) ? ldexp(1.0, -(wb0 + wbits * (d0 + 3))) : 0.0;
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 16) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t taddr = tmem0 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * TN + c0);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]),
                              "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]),
                              "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]),
                              "=r"(v[g][15])
                            : "r"(taddr));
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (store && row < ncols) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + j;
                            if (col < ncols) {
                                double x = (double)(int32_t)v[3][j] * g3;
                                x += (double)(int32_t)v[2][j] * g2;
                                x += (double)(int32_t)v[1][j] * g1;
                                x += (double)(int32_t)v[0][j] * g0;
                                x *= rs * dscale[col];
                                double* cp = C + row + col * ldc;
                                if (pass == 0) *cp = (beta == 0.0) ? x : (x + beta * *cp);
                                else *cp += x;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// ---- CTA-pair kernel with 64-byte k rows (SWIZZLE_64B): the production kernel for radix-256 digits ---------------
// The pair kernel above moves its operands in 32-byte rows (one MMA K step per row, SWIZZLE_32B): every TMA request is
// one 32-byte sector of a different column.  Its loads alone (HYP_OZAKI_PROBE=2: no MMAs) need 70 ms on C3 = 28 B per
// clock per SM, i.e. about one request per clock per SM, while its MMA stream alone (HYP_OZAKI_PROBE=1) needs 40 ms
// (74 cycles per 256 x 128 x 32 MMA): the request rate of the TMA unit, not the L2 slices (6300 B/clk chip-wide =
// 42 B/clk/SM) and not the tensor pipe, bounds it.  Here a stage holds TWO MMA K steps in 64-byte rows, so every
// request carries two sectors of one line: half the requests per byte.  Shared memory: pass 0 (digit sums 0..3, slices
// 0..3 of both operands) 4 stages x 48 KB, pass 1 (digit sums 4..6, all 7 slices) 2 stages x 84 KB, both rings in the
// same 192 KB; the producer starts the ring of the next pass when the accumulator-full barrier of the previous one
// has fired (all MMAs that read the other ring have retired).  Everything else (cta_group::2, M = 256, four int32
// accumulators in the 512 TMEM columns, warp-uniform unrolled issue loop, epilogue) is the pair kernel.
constexpr int P64_KB = 64;                                  // k bytes per shared-memory row = 2 MMA K steps
constexpr int P64_TA = TM * P64_KB;                         // 8 KB: one digit slice of my 128-row A tile
constexpr int P64_TB = (TN / 2) * P64_KB;                   // 4 KB: one digit slice of my 64-row half of the B tile
constexpr int P64_NSL = 7;
// L1 = number of digit sums collected by pass 0 (template parameter of the kernel): pass 0 needs the slices 0 .. L1 - 1 of
// both operands, pass 1 (digit sums L1 .. 6, at most four accumulators: L1 >= 3) all seven.  L1 = 3 (default) moves
// 3 + 7 = 10 slices per k step and output tile, L1 = 4 (the first version, HYP_OZAKI_SPLIT=4) 4 + 7 = 11.
constexpr int P64_UNIT = P64_TA + P64_TB;                   // 12 KB: one slice of my A tile and of my B half tile
constexpr int P64_STAGE1 = P64_NSL * P64_UNIT;              // 84 KB
constexpr int P64_STAGES1 = 2;
constexpr int P64_RING = 192 * 1024;
constexpr int P64_MAXST0 = 5;                               // barrier slots reserved for the pass-0 ring
constexpr int P64_SMEM = P64_RING + 1024 + 256;
__host__ __device__ constexpr int p64_stage0(int L1) { return L1 * P64_UNIT; }              // 36 / 48 KB
__host__ __device__ constexpr int p64_stages0(int L1) { return P64_RING / (L1 * P64_UNIT) > P64_MAXST0 ? P64_MAXST0 : P64_RING / (L1 * P64_UNIT); }
static_assert(P64_STAGES1 * P64_STAGE1 <= P64_RING, "pass-1 ring");

// K-major SWIZZLE_64B descriptor: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

template <int PASS, int L1>
__device__ __forceinline__ void issue_pass64(uint64_t dbase, uint32_t tmem0, uint32_t idesc, int nst, uint32_t bar_full,
                                             uint32_t bar_empty, uint32_t bar_tfull, int& stage, uint32_t& phase,
                                             bool issuer, bool no_mma) {
    constexpr int NSL = P64_NSL;
    constexpr int D0 = PASS * L1;                       // first digit sum of the pass
    constexpr int NACC = PASS == 0 ? L1 : NSL - L1;     // accumulators (digit sums) of the pass
    constexpr int NS = PASS == 0 ? L1 : NSL;            // A slices in front of the B half slices
    constexpr int SMAX = NS - 1;
    constexpr int STAGE = PASS == 0 ? p64_stage0(L1) : P64_STAGE1;
    constexpr int NSTAGES = PASS == 0 ? p64_stages0(L1) : P64_STAGES1;
    static_assert(NACC <= 4, "four 128-column accumulators fill the tensor memory");
    for (int it = 0; it < nst; it++) {
        mbar_wait(bar_full + stage * 8, phase);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t sd = dbase + (uint64_t)((uint32_t)(stage * STAGE) >> 4);
        const uint32_t first = it > 0 ? 1u : 0u;
        if (!no_mma) {
#pragma unroll
            for (int h = 0; h < 2; h++) {                // the two 32-byte K steps inside the 64-byte rows
#pragma unroll
                for (int sl = 0; sl <= SMAX; sl++) {
                    const int tlo = D0 - sl > 0 ? D0 - sl : 0;
                    const int thi = (D0 + NACC - 1 - sl) < (NSL - 1 - sl) ? (D0 + NACC - 1 - sl) : (NSL - 1 - sl);
#pragma unroll
                    for (int tt = 0; tt < NSL; tt++) {
                        if (tt < tlo || tt > thi) continue;
                        const uint32_t offA = (uint32_t)(sl * P64_TA + h * 32) >> 4;
                        const uint32_t offB = (uint32_t)(NS * P64_TA + tt * P64_TB + h * 32) >> 4;
                        const uint32_t accum = (h > 0 || sl > 0) ? 1u : first;
                        if (issuer) umma_i8_2sm(tmem0 + (uint32_t)((sl + tt - D0) * TN), sd + offA, sd + offB, idesc, accum);
                    }
                }
            }
        }
        if (issuer) umma_commit_2sm(bar_empty + stage * 8, 3);              // frees the stage in both CTAs
        if (++stage == NSTAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
    if (issuer) umma_commit_2sm(bar_tfull, 3);                               // accumulators ready in both CTAs
}

template <int L1>
__global__ void __launch_bounds__(I8_THREADS, 1)
ozaki_syrk_pair64_kernel(const __grid_constant__ CUtensorMap mapA4, const __grid_constant__ CUtensorMap mapA7,
                         const __grid_constant__ CUtensorMap mapB4, const __grid_constant__ CUtensorMap mapB7,
                         const int4* __restrict__ pairs, int n_pairs, int k0, int nst,
                         const double* __restrict__ dscale, const double* __restrict__ dscale_col, int64_t mrows,
                         int64_t ncols, double* __restrict__ C, int64_t ldc, int64_t c_group_stride, double alpha,
                         double beta, int probe_full) {
    // pairs[i] = {P, J, k offset of the B-side operand, output group}: tile rows 2P, 2P + 1 of tile column J; the last two
    // are 0 except for grouped products (C_g = A' B[g * kstride .. , :], the congruences of the matrix cones).
    // probe_full bit 8: store every tile (general product C = A' B) instead of the upper ones (SYRK / P'(HG)).
    const int probe = probe_full & 0xff;
    const bool full = (probe_full & 0x100) != 0;
    // dscale: column scales of the A-side operand (rows of C), dscale_col: of the B-side operand (columns of C); the
    // same array for the SYRK, two arrays for the two-operand product C = P' R (mixed / log-det models, qrchol.jl:245)
    // probe (HYP_OZAKI_PROBE; results are garbage, timing only): 1 = no TMA loads, 2 = no MMAs (loads + epilogue only)
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    constexpr int P64_STAGES0 = p64_stages0(L1), P64_STAGE0 = p64_stage0(L1);
    const uint32_t bar_full0 = stg + P64_RING;
    const uint32_t bar_empty0 = bar_full0 + P64_MAXST0 * 8;
    const uint32_t bar_full1 = bar_empty0 + P64_MAXST0 * 8;
    const uint32_t bar_empty1 = bar_full1 + P64_STAGES1 * 8;
    const uint32_t bar_tfull = bar_empty1 + P64_STAGES1 * 8;
    const uint32_t bar_tempty = bar_tfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank, cid, ncl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P64_STAGES0; s++) {
            mbar_init(bar_full0 + s * 8, 1);     // leader: its own expect_tx arrival (+ the bytes of both CTAs)
            mbar_init(bar_empty0 + s * 8, 1);    // one multicast commit from the leader
        }
        for (int s = 0; s < P64_STAGES1; s++) {
            mbar_init(bar_full1 + s * 8, 1);
            mbar_init(bar_empty1 + s * 8, 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);                // 4 epilogue warps of each CTA (used in the leader only)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): my A tile and my half of the B tile =====
        if (lane == 0) {
            int stage[2] = {0, 0};
            uint32_t phase[2] = {0, 0};
            uint32_t item = 0;
            for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
                const int4 pr = pairs[pi];
                const int tI = 2 * pr.x + (int)crank, tJ = pr.y, kboff = pr.z;
                for (int pass = 0; pass < 2; pass++, item++) {
                    // the two rings share their shared memory: the previous item's MMAs must all have retired
                    if (item > 0) mbar_wait(bar_tfull, (item - 1) & 1u);
                    const int ns = pass == 0 ? L1 : P64_NSL;
                    const int nstages = pass == 0 ? P64_STAGES0 : P64_STAGES1;
                    const uint32_t sbytes = pass == 0 ? P64_STAGE0 : P64_STAGE1;
                    const uint32_t bfull = pass == 0 ? bar_full0 : bar_full1;
                    const uint32_t bempty = pass == 0 ? bar_empty0 : bar_empty1;
                    const CUtensorMap* ma = pass == 0 ? &mapA4 : &mapA7;
                    const CUtensorMap* mb = pass == 0 ? &mapB4 : &mapB7;
                    int st = stage[pass];
                    uint32_t ph = phase[pass];
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bempty + st * 8, ph ^ 1u);
                        if (probe == 1) {
                            if (leader) mbar_arrive(bfull + st * 8);
                        } else {
                            const uint32_t full_leader = mapa_u32(bfull + st * 8, 0);
                            if (leader) mbar_expect_tx(bfull + st * 8, 2u * sbytes);
                            const uint32_t dst = stg + st * sbytes;
                            const int kc = k0 + it * P64_KB;
                            tma_load_3d_2sm(dst, ma, kc, tI * TM, 0, full_leader);
                            tma_load_3d_2sm(dst + ns * P64_TA, mb, kc + kboff, tJ * TN + (int)crank * (TN / 2), 0, full_leader);
                        }
                        if (++st == nstages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                    stage[pass] = st;
                    phase[pass] = ph;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only): warp-uniform issue loop, one elected lane issues =====
        if (leader) {
            const uint32_t idesc = make_idesc_i8(2 * TM, TN);
            const uint64_t dbase = make_desc_sw64(stg);
            const bool issuer = elect_one_sync();
            int stage0 = 0, stage1 = 0;
            uint32_t phase0 = 0, phase1 = 0, item = 0;
            for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
                for (int pass = 0; pass < 2; pass++, item++) {
                    if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (pass == 0)
                        issue_pass64<0, L1>(dbase, tmem0, idesc, nst, bar_full0, bar_empty0, bar_tfull, stage0, phase0, issuer, probe == 2);
                    else
                        issue_pass64<1, L1>(dbase, tmem0, idesc, nst, bar_full1, bar_empty1, bar_tfull, stage1, phase1, issuer, probe == 2);
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs, own TMEM half) =====
        const int lg = warp & 3;
        const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
        uint32_t item = 0;
        double* const C0 = C;
        const int64_t ncols_rows = mrows;
        for (int pi = (int)cid; pi < n_pairs; pi += (int)ncl) {
            const int4 pr = pairs[pi];
            const int tI = 2 * pr.x + (int)crank, tJ = pr.y;
            double* C = C0 + (int64_t)pr.w * c_group_stride;
            const int64_t row = (int64_t)tI * TM + lg * 32 + lane;
            const bool store = full || tI <= tJ;
            const double rs = (row < ncols_rows) ? alpha * dscale[row] : 0.0;
            for (int pass = 0; pass < 2; pass++, item++) {
                if (pass == 0 && beta != 0.0 && store && row < ncols_rows) {
                    // C += ...: pull my rows of the tile into L2 while the MMAs of this pass run (the tile comes from DRAM)
#pragma unroll 8
                    for (int j = lane & 1; j < TN; j += 2) {              // 16 lanes share a 128-byte line: two lanes per line suffice
                        const int64_t col = (int64_t)tJ * TN + j;
                        if (col < ncols && (lane & 15) < 2)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(C + row + col * ldc));
                    }
                }
                // the old values of C (second pass, or beta != 0) are loaded one 16-column chunk AHEAD of their use: the
                // first chunk before the accumulators are even ready, chunk c + 1 while chunk c is converted and stored
                // (written as 16 load / store pairs they serialise on 16 L2 or DRAM round trips per chunk)
                const bool rmw = (pass == 1 || beta != 0.0) && store && row < ncols_rows;
                const double bt = pass == 1 ? 1.0 : beta;
                double cold[16], cnext[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int64_t col = (int64_t)tJ * TN + j;
                    cold[j] = (rmw && col < ncols) ? __ldcg(C + row + col * ldc) : 0.0;
                }
                mbar_wait(bar_tfull, item & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int d0 = pass * L1, nacc = pass == 0 ? L1 : P64_NSL - L1;
                // digit sum d has weight 2^-(14 + 8 d); accumulators the pass never wrote get weight 0
                const double g0 = ldexp(1.0, -(14 + 8 * d0)), g1 = ldexp(1.0, -(14 + 8 * (d0 + 1))),
                             g2 = ldexp(1.0, -(14 + 8 * (d0 + 2))), g3 = nacc > 3 ? ldexp(1.0, -(14 + 8 * (d0 + 3))) : 0.0;
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 16) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t taddr = tmem0 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * TN + c0);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]),
                              "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]),
                              "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]),
                              "=r"(v[g][15])
                            : "r"(taddr));
                    }
                    if (c0 + 16 < TN) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + 16 + j;
                            cnext[j] = (rmw && col < ncols) ? __ldcg(C + row + col * ldc) : 0.0;
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (store && row < ncols_rows) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + j;
                            if (col < ncols) {
                                double x = (double)(int32_t)v[3][j] * g3;
                                x += (double)(int32_t)v[2][j] * g2;
                                x += (double)(int32_t)v[1][j] * g1;
                                x += (double)(int32_t)v[0][j] * g0;
                                x *= rs * dscale_col[col];
                                C[row + col * ldc] = rmw ? (x + bt * cold[j]) : x;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; j++) cold[j] = cnext[j];
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// ---- quad variant of the 64-byte-row kernel: two CTA pairs of a 4-CTA cluster share their A tiles ---------------------
// Quad (P, Jq): tile rows 2P, 2P + 1 and tile columns 2Jq, 2Jq + 1.  Pair p (cluster ranks 2p, 2p + 1) owns column
// 2Jq + p and runs the pair kernel above; the A tile of row 2P + r is needed by CTA r of BOTH pairs, so each of the two
// loads half of its slices (one TMA per slice) and multicasts them to the other: a third fewer bytes leave the L2 slices
// per MMA, and the L2 -> SM path is what bounds the pair kernel once its requests are 64 bytes wide (loads alone 48 ms,
// MMAs alone 39 ms on C3).  Barriers: a multicast signals the full barrier of each destination CTA itself, so the second
// CTA of a pair forwards "my A tile has landed" to its leader; a stage is refilled only after BOTH pair leaders have
// retired the MMAs that read it (commits multicast to all four CTAs); the ring of the next pass is started when both
// leaders have committed the end of the previous pass (drain barrier).
template <int PASS, int L1>
__device__ __forceinline__ void issue_pass64q(uint64_t dbase, uint32_t tmem0, uint32_t idesc, int nst, uint32_t bar_full,
                                              uint32_t bar_empty, uint32_t bar_tfull, uint32_t bar_drain, int& stage,
                                              uint32_t& phase, bool issuer, uint16_t mask_pair) {
    constexpr int NSL = P64_NSL;
    constexpr int D0 = PASS * L1;
    constexpr int NACC = PASS == 0 ? L1 : NSL - L1;
    constexpr int NS = PASS == 0 ? L1 : NSL;
    constexpr int STAGE = PASS == 0 ? p64_stage0(L1) : P64_STAGE1;
    constexpr int NSTAGES = PASS == 0 ? p64_stages0(L1) : P64_STAGES1;
    for (int it = 0; it < nst; it++) {
        mbar_wait(bar_full + stage * 8, phase);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint64_t sd = dbase + (uint64_t)((uint32_t)(stage * STAGE) >> 4);
        const uint32_t first = it > 0 ? 1u : 0u;
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int sl = 0; sl < NS; sl++) {
                const int tlo = D0 - sl > 0 ? D0 - sl : 0;
                const int thi = (D0 + NACC - 1 - sl) < (NSL - 1 - sl) ? (D0 + NACC - 1 - sl) : (NSL - 1 - sl);
#pragma unroll
                for (int tt = 0; tt < NSL; tt++) {
                    if (tt < tlo || tt > thi) continue;
                    const uint32_t offA = (uint32_t)(sl * P64_TA + h * 32) >> 4;
                    const uint32_t offB = (uint32_t)(NS * P64_TA + tt * P64_TB + h * 32) >> 4;
                    const uint32_t accum = (h > 0 || sl > 0) ? 1u : first;
                    if (issuer) umma_i8_2sm(tmem0 + (uint32_t)((sl + tt - D0) * TN), sd + offA, sd + offB, idesc, accum);
                }
            }
        }
        if (issuer) umma_commit_2sm(bar_empty + stage * 8, 0xF);     // one of the two releases of this stage, in all four CTAs
        if (++stage == NSTAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
    if (issuer) {
        umma_commit_2sm(bar_tfull, mask_pair);                       // accumulators ready in both CTAs of my pair
        umma_commit_2sm(bar_drain, 0xF);                             // my pair no longer reads the ring of this pass
    }
}

template <int L1>
__global__ void __launch_bounds__(I8_THREADS, 1)
ozaki_syrk_quad64_kernel(const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB0,
                         const __grid_constant__ CUtensorMap mapB1, const int2* __restrict__ quads, int n_quads, int k0,
                         int nst, const double* __restrict__ dscale, int64_t ncols, double* __restrict__ C, int64_t ldc,
                         double alpha, double beta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    constexpr int P64_STAGES0 = p64_stages0(L1), P64_STAGE0 = p64_stage0(L1);
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    const uint32_t bar_full0 = stg + P64_RING;
    const uint32_t bar_empty0 = bar_full0 + P64_MAXST0 * 8;
    const uint32_t bar_full1 = bar_empty0 + P64_MAXST0 * 8;
    const uint32_t bar_empty1 = bar_full1 + P64_STAGES1 * 8;
    const uint32_t bar_tfull = bar_empty1 + P64_STAGES1 * 8;
    const uint32_t bar_tempty = bar_tfull + 8;
    const uint32_t bar_drain = bar_tempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank, cid, ncl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(ncl));
    const uint32_t pr = crank >> 1, r = crank & 1u;        // pair inside the quad, CTA inside the pair
    const bool leader = r == 0;
    const uint32_t leader_rank = crank & ~1u;
    const uint16_t mask_a = (uint16_t)((1u << r) | (1u << (2 + r)));      // the CTAs that hold tile row 2P + r
    const uint16_t mask_pair = (uint16_t)(3u << (2 * pr));                // my CTA pair

    if (threadIdx.x == 0) {
        for (int s = 0; s < P64_STAGES0; s++) {
            // leader: its own expect_tx arrival + the peer's forwarded "my A tile has landed";
            // peer: its own expect_tx arrival (A bytes only; its B half signals the leader directly)
            mbar_init(bar_full0 + s * 8, leader ? 2 : 1);
            mbar_init(bar_empty0 + s * 8, 2);    // one multicast commit from each of the two pair leaders
        }
        for (int s = 0; s < P64_STAGES1; s++) {
            mbar_init(bar_full1 + s * 8, leader ? 2 : 1);
            mbar_init(bar_empty1 + s * 8, 2);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 8);                // 4 epilogue warps of each CTA of the pair (used in the leader only)
        mbar_init(bar_drain, 2);                 // both pair leaders
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;

    if (warp == 0) {
        // ===== TMA producer (all four CTAs): my share of the slices of my A tile (multicast) and my half of the B tile =====
        if (lane == 0) {
            int stage[2] = {0, 0};
            uint32_t phase[2] = {0, 0};
            uint32_t item = 0;
            for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
                const int2 qd = quads[qi];
                const int tI = 2 * qd.x + (int)r, tJ = 2 * qd.y + (int)pr;
                for (int pass = 0; pass < 2; pass++, item++) {
                    // the two rings share their shared memory: BOTH pairs must have retired the previous item's MMAs
                    if (item > 0) mbar_wait(bar_drain, (item - 1) & 1u);
                    const int ns = pass == 0 ? L1 : P64_NSL;
                    const int s_lo = pr == 0 ? 0 : (ns + 1) / 2, s_hi = pr == 0 ? (ns + 1) / 2 : ns;   // my share of the A slices
                    const int nstages = pass == 0 ? P64_STAGES0 : P64_STAGES1;
                    const uint32_t sbytes = pass == 0 ? P64_STAGE0 : P64_STAGE1;
                    const uint32_t bfull = pass == 0 ? bar_full0 : bar_full1;
                    const uint32_t bempty = pass == 0 ? bar_empty0 : bar_empty1;
                    const CUtensorMap* mb = pass == 0 ? &mapB0 : &mapB1;
                    const uint32_t bytes_a = (uint32_t)ns * P64_TA, bytes_bh = (uint32_t)ns * P64_TB;
                    int st = stage[pass];
                    uint32_t ph = phase[pass];
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bempty + st * 8, ph ^ 1u);
                        const uint32_t full = bfull + st * 8;
                        const uint32_t full_leader = mapa_u32(full, leader_rank);
                        mbar_expect_tx(full, leader ? bytes_a + 2u * bytes_bh : bytes_a);
                        const uint32_t dst = stg + st * sbytes;
                        const int kc = k0 + it * P64_KB;
                        for (int sl = s_lo; sl < s_hi; sl++)
                            tma_load_3d_mc(dst + sl * P64_TA, &mapA1, kc, tI * TM, sl, full, mask_a);
                        tma_load_3d_2sm(dst + ns * P64_TA, mb, kc, tJ * TN + (int)r * (TN / 2), 0, full_leader);
                        if (++st == nstages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                    stage[pass] = st;
                    phase[pass] = ph;
                }
            }
        }
    } else if (warp == 1 && !leader) {
        // ===== forwarder (second CTA of a pair): tells the pair leader when my A tile of a stage has landed =====
        if (lane == 0) {
            int stage[2] = {0, 0};
            uint32_t phase[2] = {0, 0};
            for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
                for (int pass = 0; pass < 2; pass++) {
                    const int nstages = pass == 0 ? P64_STAGES0 : P64_STAGES1;
                    const uint32_t bfull = pass == 0 ? bar_full0 : bar_full1;
                    int st = stage[pass];
                    uint32_t ph = phase[pass];
                    for (int it = 0; it < nst; it++) {
                        mbar_wait(bfull + st * 8, ph);
                        mbar_arrive_cluster(mapa_u32(bfull + st * 8, leader_rank));
                        if (++st == nstages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                    stage[pass] = st;
                    phase[pass] = ph;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA of each pair): warp-uniform loop, one elected lane issues =====
        const uint32_t idesc = make_idesc_i8(2 * TM, TN);
        const uint64_t dbase = make_desc_sw64(stg);
        const bool issuer = elect_one_sync();
        int stage0 = 0, stage1 = 0;
        uint32_t phase0 = 0, phase1 = 0, item = 0;
        for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
            for (int pass = 0; pass < 2; pass++, item++) {
                if (item > 0) mbar_wait(bar_tempty, (item - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (pass == 0)
                    issue_pass64q<0, L1>(dbase, tmem0, idesc, nst, bar_full0, bar_empty0, bar_tfull, bar_drain, stage0, phase0, issuer,
                                         mask_pair);
                else
                    issue_pass64q<1, L1>(dbase, tmem0, idesc, nst, bar_full1, bar_empty1, bar_tfull, bar_drain, stage1, phase1, issuer,
                                         mask_pair);
            }
        }
    } else {
        // ===== epilogue (both CTAs, own TMEM half) =====
        const int lg = warp & 3;
        const uint32_t tempty_leader = mapa_u32(bar_tempty, leader_rank);
        uint32_t item = 0;
        for (int qi = (int)cid; qi < n_quads; qi += (int)ncl) {
            const int2 qd = quads[qi];
            const int tI = 2 * qd.x + (int)r, tJ = 2 * qd.y + (int)pr;
            const int64_t row = (int64_t)tI * TM + lg * 32 + lane;
            const bool store = tI <= tJ;
            const double rs = (row < ncols) ? alpha * dscale[row] : 0.0;
            for (int pass = 0; pass < 2; pass++, item++) {
                if (pass == 0 && beta != 0.0 && store && row < ncols) {
                    // C += ...: pull my rows of the tile into L2 while the MMAs of this pass run (the tile comes from DRAM)
#pragma unroll 8
                    for (int j = lane & 1; j < TN; j += 2) {              // 16 lanes share a 128-byte line: two lanes per line suffice
                        const int64_t col = (int64_t)tJ * TN + j;
                        if (col < ncols && (lane & 15) < 2)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(C + row + col * ldc));
                    }
                }
                mbar_wait(bar_tfull, item & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int d0 = pass * L1, nacc = pass == 0 ? L1 : P64_NSL - L1;
                // digit sum d has weight 2^-(14 + 8 d); accumulators the pass never wrote get weight 0
                const double g0 = ldexp(1.0, -(14 + 8 * d0)), g1 = ldexp(1.0, -(14 + 8 * (d0 + 1))),
                             g2 = ldexp(1.0, -(14 + 8 * (d0 + 2))), g3 = nacc > 3 ? ldexp(1.0, -(14 + 8 * (d0 + 3))) : 0.0;
#pragma unroll 1
                for (int c0 = 0; c0 < TN; c0 += 16) {
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t taddr = tmem0 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(g * TN + c0);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(v[g][0]), "=r"(v[g][1]), "=r"(v[g][2]), "=r"(v[g][3]), "=r"(v[g][4]),
                              "=r"(v[g][5]), "=r"(v[g][6]), "=r"(v[g][7]), "=r"(v[g][8]), "=r"(v[g][9]),
                              "=r"(v[g][10]), "=r"(v[g][11]), "=r"(v[g][12]), "=r"(v[g][13]), "=r"(v[g][14]),
                              "=r"(v[g][15])
                            : "r"(taddr));
                    }
                    // the old values of C (second pass, or beta != 0): all 16 loads in flight BEFORE the TMEM wait and the
                    // stores - written as 16 load / store pairs they serialise on 16 L2 (or DRAM) round trips per chunk
                    double cold[16];
                    const bool rmw = pass == 1 || beta != 0.0;
                    const double bt = pass == 1 ? 1.0 : beta;
                    if (store && row < ncols && rmw) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + j;
                            cold[j] = (col < ncols) ? __ldcg(C + row + col * ldc) : 0.0;
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (store && row < ncols) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int64_t col = (int64_t)tJ * TN + c0 + j;
                            if (col < ncols) {
                                double x = (double)(int32_t)v[3][j] * g3;
                                x += (double)(int32_t)v[2][j] * g2;
                                x += (double)(int32_t)v[1][j] * g1;
                                x += (double)(int32_t)v[0][j] * g0;
                                x *= rs * dscale[col];
                                C[row + col * ldc] = rmw ? (x + bt * cold[j]) : x;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// digit-slice tensor map with 64-byte k rows (SWIZZLE_64B) for the kernel above
void make_map_digits64(CUtensorMap* map, const int8_t* base, int64_t K, int64_t cols, int64_t ldd,
                       int64_t slice_stride, int nslices, int box_slices, int box_rows) {
    if (((uintptr_t)base & 15) || (ldd & 15) || (slice_stride & 15))
        throw HypError{"digit slices must be 16-byte aligned with ld % 16 == 0"};
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)cols, (cuuint64_t)nslices};
    cuuint64_t strides[2] = {(cuuint64_t)ldd, (cuuint64_t)slice_stride};
    cuuint32_t box[3] = {P64_KB, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw HypError{"cuTensorMapEncodeTiled (digit slices, 64-byte rows) failed"};
}

// pass split of the 64-byte-row kernel: digit sums 0 .. L1 - 1 in pass 0 (3 by default, HYP_OZAKI_SPLIT=4: 4)
int p64_split() {
    static int l1 = 0;
    if (!l1) {
        const char* e = getenv("HYP_OZAKI_SPLIT");
        l1 = (e && e[0] == '4') ? 4 : 3;
        CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_pair64_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, P64_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_pair64_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, P64_SMEM));
    }
    return l1;
}

// one launch of the 64-byte-row CTA-pair kernel over a pair list (cfg carries grid, stream and the cluster attribute)
// One launch of the 64-byte-row CTA-pair kernel.  A side: digits / dscale of the K x mrows operand (tile rows of C);
// B side (digitsB .. ; nullptr = the A side: SYRK): digits / dscaleB of the Kb x ncols operand (tile columns of C).
struct P64Operand {
    const int8_t* digits;
    int64_t K, cols, ldd, slice_stride;
    int nslices_alloc;
    const double* dscale;
};
void launch_pair64_ex(hyp_ctx* ctx, cudaLaunchConfig_t* cfg, const P64Operand& A, const P64Operand& B, const int4* d_pairs,
                      int n_pairs, int k0, int64_t klen, double* C, int64_t ldc, int64_t c_group_stride, double alpha,
                      double beta, int probe, bool full) {
    const int l1 = p64_split();
    CUtensorMap mA0, mA1, mB0, mB1;
    make_map_digits64(&mA0, A.digits, A.K, A.cols, A.ldd, A.slice_stride, A.nslices_alloc, l1, TM);
    make_map_digits64(&mA1, A.digits, A.K, A.cols, A.ldd, A.slice_stride, A.nslices_alloc, P64_NSL, TM);
    make_map_digits64(&mB0, B.digits, B.K, B.cols, B.ldd, B.slice_stride, B.nslices_alloc, l1, TN / 2);
    make_map_digits64(&mB1, B.digits, B.K, B.cols, B.ldd, B.slice_stride, B.nslices_alloc, P64_NSL, TN / 2);
    const int nst = (int)ceil_div(klen, P64_KB);
    const int pf = (probe & 0xff) | (full ? 0x100 : 0);
    if (l1 == 3)
        CUDA_TRY(cudaLaunchKernelEx(cfg, ozaki_syrk_pair64_kernel<3>, mA0, mA1, mB0, mB1, d_pairs, n_pairs, k0, nst, A.dscale,
                                    B.dscale, A.cols, B.cols, C, ldc, c_group_stride, alpha, beta, pf));
    else
        CUDA_TRY(cudaLaunchKernelEx(cfg, ozaki_syrk_pair64_kernel<4>, mA0, mA1, mB0, mB1, d_pairs, n_pairs, k0, nst, A.dscale,
                                    B.dscale, A.cols, B.cols, C, ldc, c_group_stride, alpha, beta, pf));
    ctx->launches++;
}
// digitsB / dscaleB (same layout as digits): the B-side operand of the two-operand product; nullptr = SYRK
void launch_pair64(hyp_ctx* ctx, cudaLaunchConfig_t* cfg, const int8_t* digits, int64_t K, int64_t ncols, int64_t ldd,
                   int64_t slice_stride, int nslices_alloc, const int4* d_pairs, int n_pairs, int k0, int64_t klen,
                   const double* dscale, double* C, int64_t ldc, double alpha, double beta, int probe,
                   const int8_t* digitsB = nullptr, const double* dscaleB = nullptr) {
    P64Operand A{digits, K, ncols, ldd, slice_stride, nslices_alloc, dscale};
    P64Operand B{digitsB ? digitsB : digits, K, ncols, ldd, slice_stride, nslices_alloc, dscaleB ? dscaleB : dscale};
    launch_pair64_ex(ctx, cfg, A, B, d_pairs, n_pairs, k0, klen, C, ldc, 0, alpha, beta, probe, false);
}

void make_map_digits(CUtensorMap* map, const int8_t* base, int64_t K, int64_t cols, int64_t ldd,
                     int64_t slice_stride, int nslices, int box_slices, int box_rows = TM) {
    if (((uintptr_t)base & 15) || (ldd & 15) || (slice_stride & 15))
        throw HypError{"digit slices must be 16-byte aligned with ld % 16 == 0"};
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)cols, (cuuint64_t)nslices};
    cuuint64_t strides[2] = {(cuuint64_t)ldd, (cuuint64_t)slice_stride};
    cuuint32_t box[3] = {OZ_KB, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw HypError{"cuTensorMapEncodeTiled (digit slices) failed"};
}

}  // namespace

// ---- unit-test entry points ----------------------------------------------------------------------
extern "C" int hyp_test_i8_gemm_tn(hyp_ctx* ctx, const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb,
                                   int64_t K, int64_t M, int64_t N, int32_t* C, int64_t ldc) {
    if (!ctx) return -1;
    try {
        CUDA_TRY(cudaSetDevice(ctx->device));
        // device copies with 16-byte aligned leading dimensions
        const int64_t la = round_up(std::max<int64_t>(K, 16), 16), lb = la;
        int8_t *dA = nullptr, *dB = nullptr;
        int32_t* dC = nullptr;
        CUDA_TRY(cudaMalloc(&dA, (size_t)la * M));
        CUDA_TRY(cudaMalloc(&dB, (size_t)lb * N));
        CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 4));
        CUDA_TRY(cudaMemsetAsync(dA, 0, (size_t)la * M, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(dB, 0, (size_t)lb * N, ctx->stream));
        CUDA_TRY(cudaMemcpy2DAsync(dA, la, A, lda, K, M, cudaMemcpyDefault, ctx->stream));
        CUDA_TRY(cudaMemcpy2DAsync(dB, lb, B, ldb, K, N, cudaMemcpyDefault, ctx->stream));
        CUtensorMap mapA, mapB;
        make_map_i8(&mapA, dA, K, M, la);
        make_map_i8(&mapB, dB, K, N, lb);
        static bool attr = false;
        if (!attr) {
            CUDA_TRY(cudaFuncSetAttribute(i8_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
            attr = true;
        }
        dim3 grid(ceil_div(M, TM), ceil_div(N, TN));
        i8_gemm_tn_kernel<<<grid, I8_THREADS, I8_SMEM, ctx->stream>>>(mapA, mapB, ceil_div(K, TKB), M, N, dC, M);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy2DAsync(C, ldc * 4, dC, M * 4, M * 4, N, cudaMemcpyDefault, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(dA);
        cudaFree(dB);
        cudaFree(dC);
        return 0;
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        cudaGetLastError();
        return -1;
    }
}

// digits (nslices x K x ncols int8, host or device) and exponents of a K x ncols FP64 matrix
extern "C" int hyp_test_ozaki_slices(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols,
                                     int nslices, int8_t* digits, int* expo) {
    if (!ctx) return -1;
    // nslices < 0: the radix-256 slicer with -nslices digits
    const bool r256 = nslices < 0;
    if (r256) nslices = -nslices;
    try {
        CUDA_TRY(cudaSetDevice(ctx->device));
        double* dA = nullptr;
        int8_t* dD = nullptr;
        int* dE = nullptr;
        CUDA_TRY(cudaMalloc(&dA, (size_t)K * ncols * 8));
        const int64_t ldd = round_up(std::max<int64_t>(K, 16), 16);
        CUDA_TRY(cudaMalloc(&dD, (size_t)nslices * ldd * ncols));
        CUDA_TRY(cudaMalloc(&dE, (size_t)ncols * 4));
        CUDA_TRY(cudaMemcpy2DAsync(dA, K * 8, A, lda * 8, K * 8, ncols, cudaMemcpyDefault, ctx->stream));
        hypdev::colmax_kernel<<<(unsigned)ncols, 256, 0, ctx->stream>>>(K, ncols, dA, K, dE, nullptr, r256 ? 1 : 0);
        dim3 grid(std::max(1, std::min(ceil_div(K, 2048), 64)), (unsigned)std::min<int64_t>(ncols, 65535));
        if (r256)
            hypdev::slice256_kernel<<<grid, 256, 0, ctx->stream>>>(K, ncols, dA, K, dE, nslices, dD, ldd, ldd * ncols);
        else
            hypdev::slice_kernel<<<grid, 256, 0, ctx->stream>>>(K, ncols, dA, K, dE, nslices, dD, ldd, ldd * ncols);
        ctx->launches += 2;
        CUDA_TRY(cudaGetLastError());
        for (int s = 0; s < nslices; s++)
            CUDA_TRY(cudaMemcpy2DAsync(digits + (size_t)s * K * ncols, K, dD + (size_t)s * ldd * ncols, ldd, K, ncols,
                                       cudaMemcpyDefault, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(expo, dE, (size_t)ncols * 4, cudaMemcpyDefault, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(dA);
        cudaFree(dD);
        cudaFree(dE);
        return 0;
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        cudaGetLastError();
        return -1;
    }
}

// ---- product entry points -------------------------------------------------------------------------
// digits / exponents of the K x ncols FP64 matrix A (device) into caller-provided device buffers
// digit radix of the slicing: 256 (default: seven balanced 8-bit digits, 28 pair products, 56 bits; CTA-pair / quad
// kernels) or 128 (HYP_OZAKI_RADIX=128, and whenever HYP_OZAKI_CLUSTER selects the single-CTA / 2 x 2 cluster
// kernels, which only know the radix-128 weights: eight 7-bit digits, 36 pair products, the same 56 bits)
static int ozaki_radix() {
    static int r = 0;
    if (!r) {
        const char* e = getenv("HYP_OZAKI_RADIX");
        const char* c = getenv("HYP_OZAKI_CLUSTER");
        r = ((e && atoi(e) == 128) || (c && (c[0] == '0' || c[0] == '1'))) ? 128 : 256;
    }
    return r;
}

int hyp_ozaki_radix() { return ozaki_radix(); }

void hyp_ozaki_slice(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols, int8_t* digits,
                     int64_t ldd, int64_t slice_stride, int* expo, double* dscale, bool have_expo) {
    if (K <= 0 || ncols <= 0) return;
    const bool r256 = ozaki_radix() == 256;
    if (!have_expo)
        hypdev::colmax_kernel<<<(unsigned)ncols, 256, 0, ctx->stream>>>(K, ncols, A, lda, expo, dscale, r256 ? 1 : 0);
    dim3 grid(std::max(1, std::min(ceil_div(K, 2048), 32)), (unsigned)std::min<int64_t>(ncols, 65535));
    if (r256)
        hypdev::slice256_kernel<<<grid, 256, 0, ctx->stream>>>(K, ncols, A, lda, expo, 7, digits, ldd, slice_stride);
    else
        hypdev::slice_kernel<<<grid, 256, 0, ctx->stream>>>(K, ncols, A, lda, expo, OZ_S, digits, ldd, slice_stride);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
}

// radix-128 scheme only (HYP_OZAKI_RADIX=128): digit slices entering the product of the CTA-pair kernel, 8 or 7
// (HYP_OZAKI_SLICES=7: 28 instead of 36 pair products; truncation error <= 2^-48 |a_i|_max |a_j|_max K instead of
// 2^-55).  The default radix-256 scheme always uses its 7 digits.
static int ozaki_slices() {
    static int n = 0;
    if (!n) {
        const char* e = getenv("HYP_OZAKI_SLICES");
        n = (e && e[0] == '7') ? 7 : 8;
    }
    return n;
}

// C(upper 128-tiles) = alpha * A' A + beta * C from the digit slices of A
void hyp_ozaki_syrk(hyp_ctx* ctx, const int8_t* digits, int64_t ldd, int64_t slice_stride, const int* expo,
                    const double* dscale, int64_t K, int64_t ncols, double* C, int64_t ldc, double alpha,
                    double beta, const int8_t* digitsB, const double* dscaleB) {
    if (K <= 0 || ncols <= 0) return;
    if (digitsB && !hyp_ozaki_pair64_ready(ctx))
        throw HypError{"two-operand digit-sliced product needs the 64-byte-row CTA-pair kernel"};
    static bool attr = false;
    if (!attr) {
        CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
        attr = true;
    }
    // upper-triangular tile list (8-tile row groups), cached per size
    static std::vector<std::pair<int, std::pair<int4*, int>>> cache;
    const int nt = ceil_div(ncols, TM);
    int4* d_tiles = nullptr;
    int n_tiles = 0;
    for (auto& e : cache)
        if (e.first == nt) {
            d_tiles = e.second.first;
            n_tiles = e.second.second;
        }
    if (!d_tiles) {
        std::vector<int4> tl;
        const int GROUP = 8;
        for (int gi = 0; gi < nt; gi += GROUP)
            for (int tj = gi; tj < nt; tj++)
                for (int ti = gi; ti < std::min(gi + GROUP, tj + 1); ti++) tl.push_back(make_int4(ti, tj, 0, 0));
        n_tiles = (int)tl.size();
        CUDA_TRY(cudaMalloc(&d_tiles, tl.size() * sizeof(int4)));
        CUDA_TRY(cudaMemcpyAsync(d_tiles, tl.data(), tl.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cache.push_back({nt, {d_tiles, n_tiles}});
    }
    CUtensorMap mapD2, mapD4, mapD8;
    make_map_digits(&mapD2, digits, K, ncols, ldd, slice_stride, OZ_S, 2);
    make_map_digits(&mapD4, digits, K, ncols, ldd, slice_stride, OZ_S, 4);
    make_map_digits(&mapD8, digits, K, ncols, ldd, slice_stride, OZ_S, 8);
    static int use_cluster = -1;
    static int max_clusters = 0;
    if (use_cluster < 0) {
        const char* e = getenv("HYP_OZAKI_CLUSTER");
        // 0: one CTA per tile; 1: 2 x 2 clusters with TMA multicast; 2 (default): CTA pairs, cta_group::2;
        // 3: quads = two CTA pairs sharing their A tiles by multicast
        // 4 (default with radix-256 digits): CTA pairs with 64-byte k rows
        use_cluster = e ? (e[0] - '0') : (ozaki_radix() == 256 ? 4 : 2);
        // 5: quads of the 64-byte-row kernel (two CTA pairs share their A tiles by multicast)
        if (use_cluster < 0 || use_cluster > 5) use_cluster = 2;
        if (use_cluster >= 4 && ozaki_radix() != 256) use_cluster = 2;
        if (use_cluster == 5) {
            p64_split();
            CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_quad64_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, P64_SMEM));
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(4 * 64);
            q.blockDim = dim3(I8_THREADS);
            q.dynamicSmemBytes = P64_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            q.attrs = at;
            q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_quad64_kernel<3>, &q) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                use_cluster = 4;
            }
        }
        if (use_cluster == 4) {
            p64_split();
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(2 * 128);
            q.blockDim = dim3(I8_THREADS);
            q.dynamicSmemBytes = P64_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            q.attrs = at;
            q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_pair64_kernel<3>, &q) != cudaSuccess ||
                max_clusters < 1) {
                cudaGetLastError();
                use_cluster = 2;
            }
        }
        if (use_cluster == 3) {
            CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZP_SMEM));
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(4 * 64);
            q.blockDim = dim3(I8_THREADS);
            q.dynamicSmemBytes = OZP_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            q.attrs = at;
            q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_quad_kernel, &q) != cudaSuccess ||
                max_clusters < 1) {
                cudaGetLastError();
                use_cluster = 2;
            }
        }
        if (use_cluster == 2) {
            CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZP_SMEM));
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(2 * 128);
            q.blockDim = dim3(I8_THREADS);
            q.dynamicSmemBytes = OZP_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            q.attrs = at;
            q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_pair_kernel, &q) != cudaSuccess ||
                max_clusters < 1) {
                cudaGetLastError();
                use_cluster = 1;
            }
        }
        if (use_cluster == 1) {
            CUDA_TRY(cudaFuncSetAttribute(ozaki_syrk_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(4 * 64);
            q.blockDim = dim3(I8_THREADS);
            q.dynamicSmemBytes = OZ_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            q.attrs = at;
            q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_cluster_kernel, &q) != cudaSuccess ||
                max_clusters < 1) {
                cudaGetLastError();
                use_cluster = 0;
            }
        }
    }
    const bool r256 = ozaki_radix() == 256;
    if (r256 && use_cluster < 2)
        throw HypError{"radix-256 digit slices need the CTA-pair or quad SYRK kernel (set HYP_OZAKI_RADIX=128)"};
    const int nsl_eff = r256 ? 7 : ozaki_slices();
    const int wbits = r256 ? 8 : 7;
    // exactness of the int32 accumulators: the largest digit-sum group has 7 pairs of |d| <= 128 (radix 256) or
    // 8 pairs of |d| <= 64 (radix 128): rows per launch <= (2^31 - 1) / (7 * 2^14) = 18724 or / (8 * 2^12) = 65535.
    // K is cut into the fewest equal chunks below that bound (every launch pays two epilogue passes over C).
    // Chunks are multiples of 64 rows: pass 0 packs two 32-row k steps per stage, and the second one of an odd
    // last stage must fall beyond K (zero-filled by TMA), never into the next chunk.
    const int64_t CHUNK_MAX = r256 ? 18688 : 65472;
    const int64_t nchunks = (K + CHUNK_MAX - 1) / CHUNK_MAX;
    const int64_t CHUNK = round_up((K + nchunks - 1) / nchunks, 64);
    const int grid = std::min(n_tiles, ctx->sm_count);
    for (int64_t k0 = 0; k0 < K; k0 += CHUNK) {
        const int64_t klen = std::min(CHUNK, K - k0);
        if (use_cluster == 5) {
            // (P, Jq): tile rows 2P, 2P+1 x tile columns 2Jq, 2Jq+1 for P <= Jq, row pair by row pair (the A panel of
            // the row pair stays in L2 while the B panels stream, as in the pair kernel's default order)
            static std::vector<std::pair<int, std::pair<int2*, int>>> q64cache;
            int2* d_quads = nullptr;
            int n_quads = 0;
            for (auto& e : q64cache)
                if (e.first == nt) {
                    d_quads = e.second.first;
                    n_quads = e.second.second;
                }
            if (!d_quads) {
                std::vector<int2> ql;
                for (int pp = 0; 2 * pp < nt; pp++)
                    for (int jq = pp; 2 * jq < nt; jq++) ql.push_back(make_int2(pp, jq));
                n_quads = (int)ql.size();
                CUDA_TRY(cudaMalloc(&d_quads, ql.size() * sizeof(int2)));
                CUDA_TRY(cudaMemcpyAsync(d_quads, ql.data(), ql.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                q64cache.push_back({nt, {d_quads, n_quads}});
            }
            CUtensorMap mA1, mB0, mB1;
            make_map_digits64(&mA1, digits, K, ncols, ldd, slice_stride, OZ_S, 1, TM);
            make_map_digits64(&mB0, digits, K, ncols, ldd, slice_stride, OZ_S, 3, TN / 2);
            make_map_digits64(&mB1, digits, K, ncols, ldd, slice_stride, OZ_S, P64_NSL, TN / 2);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(4 * std::min(n_quads, max_clusters));
            cfg.blockDim = dim3(I8_THREADS);
            cfg.dynamicSmemBytes = P64_SMEM;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, ozaki_syrk_quad64_kernel<3>, mA1, mB0, mB1, (const int2*)d_quads, n_quads, (int)k0,
                                        (int)ceil_div(klen, P64_KB), dscale, ncols, C, ldc, alpha, k0 == 0 ? beta : 1.0));
            ctx->launches++;
            continue;
        }
        if (use_cluster == 3) {
            // (P, Jq): tile rows 2P, 2P+1 x tile columns 2Jq, 2Jq+1, for 2P <= 2Jq+1, column pair by column pair
            static std::vector<std::pair<int, std::pair<int2*, int>>> qcache;
            int2* d_quads = nullptr;
            int n_quads = 0;
            for (auto& e : qcache)
                if (e.first == nt) {
                    d_quads = e.second.first;
                    n_quads = e.second.second;
                }
            if (!d_quads) {
                std::vector<int2> ql;
                for (int jq = 0; 2 * jq < nt; jq++)
                    for (int pp = 0; 2 * pp <= 2 * jq + 1 && 2 * pp < nt; pp++) ql.push_back(make_int2(pp, jq));
                n_quads = (int)ql.size();
                CUDA_TRY(cudaMalloc(&d_quads, ql.size() * sizeof(int2)));
                CUDA_TRY(cudaMemcpyAsync(d_quads, ql.data(), ql.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                qcache.push_back({nt, {d_quads, n_quads}});
            }
            CUtensorMap mapB4, mapB8;
            make_map_digits(&mapB4, digits, K, ncols, ldd, slice_stride, OZ_S, 4, TN / 2);
            make_map_digits(&mapB8, digits, K, ncols, ldd, slice_stride, OZ_S, 8, TN / 2);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(4 * std::min(n_quads, max_clusters));
            cfg.blockDim = dim3(I8_THREADS);
            cfg.dynamicSmemBytes = OZP_SMEM;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, ozaki_syrk_quad_kernel, mapD2, mapD4, mapB4, mapB8, (const int2*)d_quads,
                                        n_quads, (int)k0, (int)ceil_div(klen, OZ_KB), dscale, ncols, C, ldc, alpha,
                                        k0 == 0 ? beta : 1.0, nsl_eff, wbits | (getenv("HYP_OZAKI_OLD_ISSUE") ? 0x4000 : 0)));
            ctx->launches++;
            continue;
        }
        if (use_cluster == 2 || use_cluster == 4) {
            // (P, J): tile rows 2P, 2P+1 of tile column J, for 2P <= J.  Orders of the list (the clusters take
            // consecutive entries, so the order decides which operand panels stay in L2):
            //   row by row (HYP_OZAKI_ORDER=row, the default until the end of round 2): the 256-row A panel of row pair P
            //     (30 MB per 16 k rows) is L2-resident and the 128-column B panels (15 MB each) stream from DRAM once per (P, J);
            //   column by column (HYP_OZAKI_ORDER=col, the first version): the B panel is resident and the A panels, twice
            //     the size, stream.  Measured on C3 (same box): 134 vs 197 GB of DRAM traffic per SYRK, L2 hit rate
            //     72 vs 65 %, 84.8 vs 86.6 ms (profiles/r01_ozaki_pair_row_order_ncu.txt).
            //   HYP_OZAKI_ORDER=row2 (experimental, not measured yet): two row pairs resident (60 MB), their tiles of one
            //     tile column adjacent in the list, so that one B panel stream can serve both.
            //   HYP_OZAKI_ORDER=row<N> (N = 2 .. 16; N = 6 is the default): N row pairs resident, their tiles of one tile
            //     column adjacent in the list.  The clusters stride through the list (entry cid, cid + ncl, ...), so the 74
            //     clusters of a round then cover an N x 74/N block of tiles: every B panel serves N clusters and every A
            //     panel 74/N instead of one B panel stream per cluster.  ncu on C3 (profiles/r02_ozaki_pair64_tile_order_ncu.md):
            //     DRAM traffic per SYRK 117.6 -> 56.6 GB, L2 hit rate 68 -> 80 %, 53.8 -> 50.9 ms for the three launches
            //     (timed alone); bit-identical results (the order does not touch the arithmetic of a tile).
            static const int row_order = [] {
                const char* e = getenv("HYP_OZAKI_ORDER");
                if (e && e[0] == 'c') return 0;
                if (e && !strncmp(e, "row", 3)) return e[3] ? std::max(1, std::min(16, atoi(e + 3))) : 1;
                return 6;
            }();
            static std::vector<std::pair<int, std::pair<int2*, int>>> pcache;
            int2* d_pairs = nullptr;
            int n_pairs = 0;
            for (auto& e : pcache)
                if (e.first == nt) {
                    d_pairs = e.second.first;
                    n_pairs = e.second.second;
                }
            if (!d_pairs) {
                std::vector<int2> pl;
                if (row_order >= 2) {
                    for (int pp = 0; 2 * pp < nt; pp += row_order)
                        for (int tj = 2 * pp; tj < nt; tj++)
                            for (int r = 0; r < row_order; r++)
                                if (2 * (pp + r) <= tj && 2 * (pp + r) < nt) pl.push_back(make_int2(pp + r, tj));
                } else if (row_order == 1) {
                    for (int pp = 0; 2 * pp < nt; pp++)
                        for (int tj = 2 * pp; tj < nt; tj++) pl.push_back(make_int2(pp, tj));
                } else {
                    for (int tj = 0; tj < nt; tj++)
                        for (int pp = 0; 2 * pp <= tj; pp++) pl.push_back(make_int2(pp, tj));
                }
                n_pairs = (int)pl.size();
                CUDA_TRY(cudaMalloc(&d_pairs, pl.size() * sizeof(int2)));
                CUDA_TRY(cudaMemcpyAsync(d_pairs, pl.data(), pl.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                pcache.push_back({nt, {d_pairs, n_pairs}});
            }
            if (use_cluster == 4) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(2 * std::min(n_pairs, max_clusters));
                cfg.blockDim = dim3(I8_THREADS);
                cfg.dynamicSmemBytes = P64_SMEM;
                cfg.stream = ctx->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                const int probe = getenv("HYP_OZAKI_PROBE") ? atoi(getenv("HYP_OZAKI_PROBE")) : 0;   // tools/syrk_probe.py
                // the 64-byte-row kernel takes {P, J, k offset of B, group} entries
                static std::vector<std::pair<int, int4*>> p4cache;
                int4* d_pairs4 = nullptr;
                for (auto& e : p4cache)
                    if (e.first == nt) d_pairs4 = e.second;
                if (!d_pairs4) {
                    std::vector<int2> h2((size_t)n_pairs);
                    CUDA_TRY(cudaMemcpyAsync(h2.data(), d_pairs, h2.size() * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
                    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                    std::vector<int4> h4((size_t)n_pairs);
                    for (int i = 0; i < n_pairs; i++) h4[i] = make_int4(h2[i].x, h2[i].y, 0, 0);
                    CUDA_TRY(cudaMalloc(&d_pairs4, std::max<size_t>(h4.size(), 1) * sizeof(int4)));
                    CUDA_TRY(cudaMemcpyAsync(d_pairs4, h4.data(), h4.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
                    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                    p4cache.push_back({nt, d_pairs4});
                }
                launch_pair64(ctx, &cfg, digits, K, ncols, ldd, slice_stride, OZ_S, d_pairs4, n_pairs, (int)k0, klen, dscale, C, ldc,
                              alpha, k0 == 0 ? beta : 1.0, probe, digitsB, dscaleB);
                continue;
            }
            CUtensorMap mapB4, mapB8, mapA8;
            make_map_digits(&mapB4, digits, K, ncols, ldd, slice_stride, OZ_S, 4, TN / 2);
            make_map_digits(&mapB8, digits, K, ncols, ldd, slice_stride, OZ_S, nsl_eff, TN / 2);   // pass 1: nsl slices
            make_map_digits(&mapA8, digits, K, ncols, ldd, slice_stride, OZ_S, nsl_eff);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * std::min(n_pairs, max_clusters));
            cfg.blockDim = dim3(I8_THREADS);
            cfg.dynamicSmemBytes = OZP_SMEM;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, ozaki_syrk_pair_kernel, mapD4, mapA8, mapB4, mapB8, (const int2*)d_pairs,
                                        n_pairs, (int)k0, (int)ceil_div(klen, OZ_KB), dscale, ncols, C, ldc, alpha,
                                        k0 == 0 ? beta : 1.0, nsl_eff,
                                        wbits | ((getenv("HYP_OZAKI_PROBE") ? atoi(getenv("HYP_OZAKI_PROBE")) : 0) << 8) |
                                            (getenv("HYP_OZAKI_OLD_ISSUE") ? 0x4000 : 0)));
            ctx->launches++;
            continue;
        }
        if (use_cluster == 1) {
            const int nsr = (nt + 1) / 2;
            const int n_super = nsr * (nsr + 1) / 2;
            // super tiles column by column over the upper triangle (the clusters running at the same time
            // share one column panel; a 6 x 6 blocked order was measured: 14 % MORE DRAM traffic, slower)
            static std::vector<std::pair<int, int2*>> scache;
            int2* d_supers = nullptr;
            for (auto& e : scache)
                if (e.first == nsr) d_supers = e.second;
            if (!d_supers) {
                std::vector<int2> sl;
                for (int sj = 0; sj < nsr; sj++)
                    for (int si = 0; si <= sj; si++) sl.push_back(make_int2(si, sj));
                CUDA_TRY(cudaMalloc(&d_supers, sl.size() * sizeof(int2)));
                CUDA_TRY(cudaMemcpyAsync(d_supers, sl.data(), sl.size() * sizeof(int2), cudaMemcpyHostToDevice,
                                         ctx->stream));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                scache.push_back({nsr, d_supers});
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(4 * std::min(n_super, max_clusters));
            cfg.blockDim = dim3(I8_THREADS);
            cfg.dynamicSmemBytes = OZ_SMEM;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, ozaki_syrk_cluster_kernel, mapD2, mapD4, (const int2*)d_supers, n_super, (int)k0,
                                        (int)ceil_div(klen, OZ_KB), dscale, ncols, C, ldc, alpha, k0 == 0 ? beta : 1.0));
            ctx->launches++;
            continue;
        }
        ozaki_syrk_kernel<<<grid, I8_THREADS, OZ_SMEM, ctx->stream>>>(mapD4, mapD8, d_tiles, n_tiles, (int)k0,
                                                                     ceil_div(klen, OZ_KB), expo, ncols, C, ldc, alpha,
                                                                     k0 == 0 ? beta : 1.0, getenv("HYP_OZAKI_NO_TMA") ? 1 : 0);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
}

// ---- pieces of the blocked Cholesky (chol.cu): its depth-512 trailing updates run on the digit-sliced kernel ----------
// true when the CTA-pair kernel with 64-byte rows can be launched on this device (radix-256 digits, default kernel choice)
bool hyp_ozaki_pair64_ready(hyp_ctx* ctx) {
    static int ready = -1;
    if (ready < 0) {
        ready = 0;
        const char* c = getenv("HYP_OZAKI_CLUSTER");
        if (ozaki_radix() == 256 && (!c || c[0] == '4')) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(2 * 128);
            q.blockDim = dim3(I8_THREADS);
            q.dynamicSmemBytes = P64_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            q.attrs = at;
            q.numAttrs = 1;
            int mc = 0;
            p64_split();
            if (cudaOccupancyMaxActiveClusters(&mc, ozaki_syrk_pair64_kernel<3>, &q) == cudaSuccess && mc >= 1)
                ready = mc;
            else
                cudaGetLastError();
        }
    }
    (void)ctx;
    return ready > 0;
}

// digit slices + scales of a block row of at most 1024 rows, one kernel, on the launch stream
void hyp_ozaki_slice_short(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols, int8_t* digits, int64_t ldd,
                           int64_t slice_stride, double* dscale) {
    if (K <= 0 || ncols <= 0) return;
    if (K > 1024 || (ldd & 15) || ldd < K) throw HypError{"hyp_ozaki_slice_short: at most 1024 rows, ldd % 16 == 0"};
    cudaStream_t s = ctx->launch_stream ? ctx->launch_stream : ctx->stream;
    const int grid = (int)std::min<int64_t>(ceil_div(ncols, 8), 4 * ctx->sm_count);
    if (K <= 512)
        hypdev::slice256_short_kernel<2><<<grid, 256, 0, s>>>((int)K, ncols, A, lda, dscale, P64_NSL, digits, ldd, slice_stride);
    else
        hypdev::slice256_short_kernel<4><<<grid, 256, 0, s>>>((int)K, ncols, A, lda, dscale, P64_NSL, digits, ldd, slice_stride);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

// C(upper tiles of the tile rows 2 p_lo .. 2 p_hi - 1) = alpha A' A + beta C from the digit slices of A (K <= 18688 rows),
// on the launch stream, with at most grid_cap CTAs (chol.cu leaves SMs to its look-ahead chain).  skip_diag > 0: the
// leading skip_diag x skip_diag TILE block is left alone (the chain already updated the next diagonal block itself).
void hyp_ozaki_syrk_rows(hyp_ctx* ctx, const int8_t* digits, int64_t ldd, int64_t slice_stride, const double* dscale,
                         int64_t K, int64_t ncols, double* C, int64_t ldc, double alpha, double beta, int p_lo, int p_hi,
                         int skip_diag) {
    if (K <= 0 || ncols <= 0) return;
    if (K > 18688 || (K & 63)) throw HypError{"hyp_ozaki_syrk_rows: K must be a multiple of 64, at most 18688"};
    if (!hyp_ozaki_pair64_ready(ctx)) throw HypError{"hyp_ozaki_syrk_rows: CTA-pair kernel not available"};
    const int nt = ceil_div(ncols, TM);
    const int np = (nt + 1) / 2;
    p_hi = std::min(p_hi < 0 ? np : p_hi, np);
    if (p_lo >= p_hi) return;
    struct Entry {
        int device, nt, p_lo, p_hi, skip;
        int4* d_pairs;
        int n_pairs;
    };
    static std::vector<Entry> cache;
    int4* d_pairs = nullptr;
    int n_pairs = 0;
    for (auto& e : cache)
        if (e.device == ctx->device && e.nt == nt && e.p_lo == p_lo && e.p_hi == p_hi && e.skip == skip_diag) {
            d_pairs = e.d_pairs;
            n_pairs = e.n_pairs;
        }
    cudaStream_t s = ctx->launch_stream ? ctx->launch_stream : ctx->stream;
    if (!d_pairs) {
        std::vector<int4> pl;
        for (int pp = p_lo; pp < p_hi; pp++)
            for (int tj = 2 * pp; tj < nt; tj++) {
                if (2 * pp + 1 < skip_diag && tj < skip_diag) continue;      // both tile rows of the pair inside the block
                pl.push_back(make_int4(pp, tj, 0, 0));
            }
        n_pairs = (int)pl.size();
        CUDA_TRY(cudaMalloc(&d_pairs, std::max<size_t>(pl.size(), 1) * sizeof(int4)));
        if (n_pairs) CUDA_TRY(cudaMemcpyAsync(d_pairs, pl.data(), pl.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        cache.push_back({ctx->device, nt, p_lo, p_hi, skip_diag, d_pairs, n_pairs});
    }
    if (!n_pairs) return;
    static int max_clusters = 0;
    if (!max_clusters) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(2 * 128);
        q.blockDim = dim3(I8_THREADS);
        q.dynamicSmemBytes = P64_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        q.attrs = at;
        q.numAttrs = 1;
        CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_pair64_kernel<3>, &q));
    }
    int ncl = std::min(n_pairs, max_clusters);
    if (ctx->grid_cap > 0) ncl = std::min(ncl, std::max(1, ctx->grid_cap / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * ncl);
    cfg.blockDim = dim3(I8_THREADS);
    cfg.dynamicSmemBytes = P64_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    launch_pair64(ctx, &cfg, digits, K, ncols, ldd, slice_stride, P64_NSL, d_pairs, n_pairs, 0, K, dscale, C, ldc, alpha, beta, 0);
    CUDA_TRY(cudaGetLastError());
}

// General product C_g = alpha * P' R_g + beta * C_g on the int8 tensor pipe (the congruences of large matrix cones,
// cones_mat.cu): P is klen x mrows, R holds ngroups blocks of klen rows at row offsets g * r_kstride (ncols columns each),
// C_g = C + g * c_group_stride.  Both operands are cut into digit slices (column scales over ALL rows of a column, i.e.
// over all groups of R) in workspaces of the context; every 128 x 128 tile of every C_g is stored.  Returns false when
// the digit-sliced kernel does not apply (caller falls back to the FP64 DMMA product).
bool hyp_ozaki_gemm_tn(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr, int64_t klen, int64_t mrows,
                       int64_t ncols, double* C, int64_t ldc, double alpha, double beta, int ngroups, int64_t r_kstride,
                       int64_t c_group_stride) {
    if (klen <= 0 || mrows <= 0 || ncols <= 0 || ngroups <= 0) return true;
    if (klen > 18688 || !hyp_ozaki_pair64_ready(ctx) || getenv("HYP_CONG_DMMA")) return false;
    // digit rows of group g start at g * ks64 (a multiple of 64): TMA coordinates must be 16-byte aligned, the row pitch of
    // the FP64 operand (an even number) need not be
    const int64_t Kb = ngroups > 1 ? r_kstride * (ngroups - 1) + klen : klen;           // rows of R (FP64 layout)
    const int64_t ks64 = round_up(r_kstride, 64);
    const int64_t Kd = ngroups > 1 ? ks64 * (ngroups - 1) + klen : klen;                // rows of the digit columns of R
    const int64_t lddA = round_up(std::max<int64_t>(klen, 16), 16), lddB = round_up(std::max<int64_t>(Kd, 16), 16);
    const int64_t sA = lddA * mrows, sB = lddB * ncols;
    auto grow = [&](void** p, int64_t* have, int64_t need) {
        if (*have >= need) return;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (*p) cudaFree(*p);
        *p = nullptr;
        CUDA_TRY(cudaMalloc(p, (size_t)need));
        *have = need;
    };
    grow((void**)&ctx->d_gemm_digA, &ctx->gemm_digA_bytes, P64_NSL * sA);
    grow((void**)&ctx->d_gemm_digB, &ctx->gemm_digB_bytes, P64_NSL * sB);
    grow((void**)&ctx->d_gemm_scal, &ctx->gemm_scal_bytes, (mrows + ncols) * 16);
    double* scaleA = ctx->d_gemm_scal;
    double* scaleB = scaleA + mrows;
    int* expoA = reinterpret_cast<int*>(scaleB + ncols);
    int* expoB = expoA + mrows;
    hyp_ozaki_slice(ctx, P, ldp, klen, mrows, ctx->d_gemm_digA, lddA, sA, expoA, scaleA);
    if (ngroups == 1) {
        hyp_ozaki_slice(ctx, R, ldr, Kb, ncols, ctx->d_gemm_digB, lddB, sB, expoB, scaleB);
    } else {
        // one scale per column over all groups, then the groups' rows to their padded digit offsets
        hypdev::colmax_kernel<<<(unsigned)ncols, 256, 0, ctx->stream>>>(Kb, ncols, R, ldr, expoB, scaleB, 1);
        dim3 grid(std::max(1, std::min(ceil_div(klen, 2048), 32)), (unsigned)std::min<int64_t>(ncols, 65535), (unsigned)ngroups);
        hypdev::slice256_kernel<<<grid, 256, 0, ctx->stream>>>(klen, ncols, R, ldr, expoB, P64_NSL, ctx->d_gemm_digB, lddB, sB,
                                                               r_kstride, ks64);
        ctx->launches += 2;
        CUDA_TRY(cudaGetLastError());
    }
    // pair list: every tile-row pair x tile column of every group
    const int mt = ceil_div(mrows, TM), nt = ceil_div(ncols, TN), np = (mt + 1) / 2;
    struct Entry {
        int device, mt, nt, ng;
        int64_t ks;
        int4* d_pairs;
        int n_pairs;
    };
    static std::vector<Entry> cache;
    int4* d_pairs = nullptr;
    int n_pairs = 0;
    for (auto& e : cache)
        if (e.device == ctx->device && e.mt == mt && e.nt == nt && e.ng == ngroups && e.ks == r_kstride) {
            d_pairs = e.d_pairs;
            n_pairs = e.n_pairs;
        }
    if (!d_pairs) {
        std::vector<int4> pl;
        pl.reserve((size_t)ngroups * np * nt);
        for (int g = 0; g < ngroups; g++)
            for (int pp = 0; pp < np; pp++)
                for (int tj = 0; tj < nt; tj++) pl.push_back(make_int4(pp, tj, (int)(g * ks64), g));
        n_pairs = (int)pl.size();
        CUDA_TRY(cudaMalloc(&d_pairs, pl.size() * sizeof(int4)));
        CUDA_TRY(cudaMemcpyAsync(d_pairs, pl.data(), pl.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cache.push_back({ctx->device, mt, nt, ngroups, r_kstride, d_pairs, n_pairs});
    }
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(I8_THREADS);
    cfg.dynamicSmemBytes = P64_SMEM;
    cfg.stream = ctx->stream;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (!max_clusters) {
        cfg.gridDim = dim3(2 * 128);
        CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, ozaki_syrk_pair64_kernel<3>, &cfg));
    }
    cfg.gridDim = dim3(2 * std::min(n_pairs, max_clusters));
    P64Operand A{ctx->d_gemm_digA, klen, mrows, lddA, sA, P64_NSL, scaleA};
    P64Operand B{ctx->d_gemm_digB, Kd, ncols, lddB, sB, P64_NSL, scaleB};
    launch_pair64_ex(ctx, &cfg, A, B, d_pairs, n_pairs, 0, klen, C, ldc, c_group_stride, alpha, beta, 0, true);
    CUDA_TRY(cudaGetLastError());
    return true;
}

// C = A' A (upper 128-tiles) for a host/device FP64 matrix A, through slicing + tcgen05 (unit test)
extern "C" int hyp_test_ozaki_syrk(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols, double* C,
                                   int64_t ldc) {
    if (!ctx) return -1;
    try {
        CUDA_TRY(cudaSetDevice(ctx->device));
        const int64_t ldd = round_up(std::max<int64_t>(K, 16), 16);
        double *dA = nullptr, *dC = nullptr;
        int8_t* dD = nullptr;
        int* dE = nullptr;
        CUDA_TRY(cudaMalloc(&dA, (size_t)K * ncols * 8));
        CUDA_TRY(cudaMalloc(&dC, (size_t)ncols * ncols * 8));
        CUDA_TRY(cudaMalloc(&dD, (size_t)OZ_S * ldd * ncols));
        CUDA_TRY(cudaMalloc(&dE, (size_t)ncols * 4));
        double* dSc = nullptr;
        CUDA_TRY(cudaMalloc(&dSc, (size_t)ncols * 8));
        CUDA_TRY(cudaMemsetAsync(dD, 0, (size_t)OZ_S * ldd * ncols, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(dC, 0, (size_t)ncols * ncols * 8, ctx->stream));
        CUDA_TRY(cudaMemcpy2DAsync(dA, K * 8, A, lda * 8, K * 8, ncols, cudaMemcpyDefault, ctx->stream));
        hyp_ozaki_slice(ctx, dA, K, K, ncols, dD, ldd, ldd * ncols, dE, dSc);
        {
            TimeScope ts(ctx, T_SYRK);      // tools/syrk_probe.py reads this timer
            hyp_ozaki_syrk(ctx, dD, ldd, ldd * ncols, dE, dSc, K, ncols, dC, ncols, 1.0, 0.0);
        }
        CUDA_TRY(cudaMemcpy2DAsync(C, ldc * 8, dC, ncols * 8, ncols * 8, ncols, cudaMemcpyDefault, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(dA);
        cudaFree(dC);
        cudaFree(dD);
        cudaFree(dSc);
        cudaFree(dE);
        return 0;
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        cudaGetLastError();
        return -1;
    }
}

// ---- tcgen05.mma kind::i8 issue-rate probe (tools/mma_probe.py) -----------------------------------------------------
// How fast can one SM (or an SM pair) retire int8 MMAs whose operands sit in shared memory, as a function of the
// shared-memory layout (SWIZZLE_32B / 64B / 128B K-major), of N (128 / 256) and of cta_group?  No TMA, no epilogue: the
// MMA issuer loops over a few operand tiles in (zeroed) shared memory.  The Schur SYRK's digit products use 32-byte K
// rows (K = 32 int8 per instruction) in SWIZZLE_32B tiles; `profiles/r02_syrk_probe_mma_vs_tma.json` shows its MMA
// stream alone already needs 88 % of the kernel time, so this is the number that bounds it.
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr, int swz) {
    // swz: 0 = SWIZZLE_32B (8-row group pitch 256 B), 1 = 64B (512 B), 2 = 128B (1024 B)
    const uint32_t sbo = 256u << swz;
    const uint64_t layout = swz == 0 ? 6 : (swz == 1 ? 4 : 2);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}
__device__ __forceinline__ void umma_i8_n(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int cg2) {
    if (cg2) umma_i8_2sm(tmem_d, adesc, bdesc, idesc, 1u);
    else umma_i8(tmem_d, adesc, bdesc, idesc, 1u);
}

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int swz, int N, int cg2, int nmma, long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t stg = (base + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank = 0;
    if (cg2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    // zero the operand area: 192 KB
    for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + (stg - base))[i] = make_uint4(0, 0, 0, 0);
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (cg2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    // generic-proxy writes (the zero fill) must be visible to the tensor core's async-proxy reads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (cg2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem0 = s_tmem;
    const int style = swz >> 4;
    swz &= 15;
    if (warp == 1 && crank == 0 && (style == 1 || lane == 0)) {
        // style 0: one thread runs the whole issue loop (operands live in ordinary registers: R2UR per MMA);
        // style 1: the warp runs it uniformly and an elected lane issues (operands can stay in uniform registers)
        const uint32_t idesc = make_idesc_i8(cg2 ? 256 : 128, N);
        const int row_bytes = 32 << swz;                       // bytes of K per shared-memory row
        const int ksub = row_bytes / 32;                       // MMAs (K = 32) per row
        const uint32_t a_tile = 128u * row_bytes;              // A: 128 rows per CTA
        const uint32_t b_rows = cg2 ? N / 2 : N;               // B rows held by each CTA
        const uint32_t b_tile = b_rows * row_bytes;
        const int na = 4, nbt = 4;                             // operand tiles cycled through (like digit slices)
        const uint32_t a0 = stg, b0 = stg + na * a_tile;       // <= 4 * 16 KB + 4 * 32 KB = 192 KB
        const int nacc = 512 / N;
        uint64_t ad[16], bd[16];
        uint32_t td[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int ks = u % ksub;
            const int sa = (u / ksub) % na, sb = (u / 4) % nbt;
            ad[u] = make_desc_kmajor(a0 + sa * a_tile + ks * 32, swz);
            bd[u] = make_desc_kmajor(b0 + sb * b_tile + ks * 32, swz);
            td[u] = tmem0 + (uint32_t)((u % nacc) * N);
        }
        const bool issuer = style == 0 ? true : elect_one_sync();
        const long long t0 = clock64();
        for (int i = 0; i < nmma; i += 16) {
#pragma unroll
            for (int u = 0; u < 16; u++)
                if (issuer) umma_i8_n(td[u], ad[u], bd[u], idesc, cg2);
        }
        if (issuer) {
            if (cg2) umma_commit_2sm(bar, 3);
            else umma_commit(bar);
        }
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && issuer) cycles[0] = t1 - t0;
    } else if (cg2 && crank == 1 && warp == 1 && lane == 0) {
        mbar_wait(bar, 0);            // the multicast commit also arrives here
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (cg2) cluster_sync_all();
    if (warp == 2) {
        if (cg2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512));
    }
}

// out[0] = SM cycles per MMA seen by CTA 0, out[1] = wall-clock ms of the launch (all SMs busy: `ctas` CTAs)
extern "C" int hyp_test_mma_rate(hyp_ctx* ctx, int swz, int N, int cg2, int nmma, int ctas, double* out) {
    if (!ctx) return -1;
    try {
        CUDA_TRY(cudaSetDevice(ctx->device));
        const int smem = 193 * 1024 + 1024;
        CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        long long* d_cyc = nullptr;
        CUDA_TRY(cudaMalloc(&d_cyc, 8));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cg2 ? (ctas & ~1) : ctas);
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cg2 ? 2 : 1;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaEventRecord(e0, ctx->stream));
            CUDA_TRY(cudaLaunchKernelEx(&cfg, mma_rate_kernel, swz, N, cg2, nmma, d_cyc));
            CUDA_TRY(cudaEventRecord(e1, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        long long cyc = 0;
        CUDA_TRY(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        out[0] = (double)cyc / nmma;
        out[1] = ms;
        cudaFree(d_cyc);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return 0;
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        cudaGetLastError();
        return -1;
    }
}

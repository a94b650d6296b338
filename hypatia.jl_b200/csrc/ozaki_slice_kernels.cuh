// Digit slicing of the FP64-accurate tcgen05 SYRK (ozaki.cu): per-column power-of-two scale and the signed 8-bit digit
// matrices.  Pure FP64 / integer arithmetic - every step is exact - so the kernels are compiled for the host as well
// (tests/emu/) and the exactness of the decomposition and of its recombination is checked in the CPU test tier
// (tests/test_emu_ozaki.py).
#pragma once
#include "devdefs.cuh"

namespace hypdev {

// ---- slicing: column exponents and the S signed 7-bit digit matrices ----------------------------
static __global__ void colmax_kernel(int64_t K, int64_t ncols, const double* __restrict__ A, int64_t lda,
                              int* __restrict__ expo, double* __restrict__ dscale, int radix256) {
    // expo[j] = smallest e with max_k |A[k, j]| < 2^e  (0 for an all-zero column); radix-256 digits need
    // max |A| <= (127/128) 2^e so that the leading digit stays below 128 after a carry
    __shared__ double sm[8];
    const int64_t j = blockIdx.x;
    if (j >= ncols) return;
    const double* col = A + j * lda;
    double mx = 0.0;
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) mx = fmax(mx, fabs(col[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, sm[w]);
        int e = 0;
        if (mx > 0.0) {
            const double f = frexp(mx, &e);          // mx = f * 2^e, f in [0.5, 1)  =>  mx < 2^e
            if (radix256 && f > 127.0 / 128.0) e++;
        }
        expo[j] = e;
        if (dscale) dscale[j] = ldexp(1.0, e);
    }
}

// D[s][k + j * ldd] = s-th signed digit of A[k, j] * 2^-expo[j].  A thread cuts 8 consecutive rows
// and stores one packed 8-byte word per slice (ldd is a multiple of 16, so the words are aligned).
static __global__ void slice_kernel(int64_t K, int64_t ncols, const double* __restrict__ A, int64_t lda,
                             const int* __restrict__ expo, int nslices, int8_t* __restrict__ D, int64_t ldd,
                             int64_t slice_stride) {
    const int64_t K8 = (K + 7) / 8;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double sc = ldexp(1.0, 6 - expo[j]);
        const double* col = A + j * lda;
        for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < K8;
             g += (int64_t)gridDim.x * blockDim.x) {
            double r[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int64_t k = g * 8 + u;
                r[u] = (k < K) ? col[k] * sc : 0.0;          // |r| < 64
            }
            for (int s = 0; s < nslices; s++) {
                uint64_t w = 0;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const double d = rint(r[u]);
                    w |= (uint64_t)(uint8_t)(int8_t)(int)d << (8 * u);
                    r[u] = (r[u] - d) * 128.0;               // exact: |r - d| <= 0.5
                }
                *reinterpret_cast<uint64_t*>(D + s * slice_stride + g * 8 + j * ldd) = w;
            }
        }
    }
}


// Eight consecutive rows r[0..7] (already multiplied by 2^(7 - e), |r| <= 127) -> one packed 8-byte word per radix-256
// digit slice (the arithmetic of slice256_kernel below, shared with the fused Schur pre-pass of cones_vec_kernels.cuh).
//
// Byte-parallel formulation.  The balanced digits d_0 .. d_{S-1} in [-128, 127] of r are those of the integer
// X = rint(r 256^(S-1)) = sum_s d_s 256^(S-1-s)  (the top-down scheme d_s = rint(r_s), r_{s+1} = 256 (r_s - d_s) with its
// backward carry pass rounds the same tail to nearest-even - every partial sum it subtracts is a multiple of 256 - and a
// balanced representation is unique).  Adding the bias B = sum_s 128 * 256^s makes every digit an ordinary unsigned
// byte: X + B = sum_s (d_s + 128) 256^(S-1-s), and d_s + 128 -> d_s as a two's-complement byte is an XOR with 0x80.  So
// ONE conversion, one 64-bit add and one XOR per value replace 7 x (rint, subtract, scale, convert) + the carry pass -
// the top-down version kept the FP64 conversion pipe busy for longer than the 7.5 GB of the pass take to move - and
// the 8 x 8 byte transpose into per-slice words is 32 byte permutes.  tests/test_emu_ozaki.py holds golden digits
// of the top-down version (ties at every level, carry chains, range ends): bit-identical.
__device__ __forceinline__ void slice256_pack8(const double (&rr)[8], int nslices, uint64_t (&w)[8]) {
    const int S = nslices < 1 ? 1 : (nslices > 8 ? 8 : nslices);
    const uint64_t bias = 0x8080808080808080ull >> (8 * (8 - S));
    // 256^(S-1) as a double, built from its exponent field (exact; r * up <= 127 * 2^56 < 2^63)
    const double up = __longlong_as_double((long long)(1023 + 8 * (S - 1)) << 52);
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        // digit s of row u in byte 7 - s whatever S is (static register indices below); bytes of unused slices are zero
        const uint64_t y = (((uint64_t)__double2ll_rn(rr[u] * up) + bias) ^ bias) << (8 * (8 - S));
        lo[u] = (uint32_t)y;
        hi[u] = (uint32_t)(y >> 32);
    }
    // 8 x 8 byte transpose: t[b] = byte b of rows 0 .. 7
    uint64_t t[8];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t A[4], B[4];                       // row pair p: A = [r0.b0 r1.b0 r0.b1 r1.b1], B = [r0.b2 r1.b2 r0.b3 r1.b3]
#pragma unroll
        for (int pr = 0; pr < 4; pr++) {
            const uint32_t x0 = h ? hi[2 * pr] : lo[2 * pr], x1 = h ? hi[2 * pr + 1] : lo[2 * pr + 1];
            A[pr] = __byte_perm(x0, x1, 0x5140);
            B[pr] = __byte_perm(x0, x1, 0x7362);
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t* F = (c < 2) ? A : B;
            const uint32_t sel = (c & 1) ? 0x7632 : 0x5410;
            t[4 * h + c] = (uint64_t)__byte_perm(F[0], F[1], sel) | ((uint64_t)__byte_perm(F[2], F[3], sel) << 32);
        }
    }
#pragma unroll
    for (int s = 0; s < 8; s++) w[s] = t[7 - s];
}

// expo / dscale of every column from the bit patterns of the column maxima (non-negative doubles order like their
// bit patterns, so the maxima are collected with atomicMax on 64-bit words); same rule as colmax_kernel
static __global__ void expo_from_bits_kernel(int64_t ncols, const unsigned long long* __restrict__ bits,
                                             int* __restrict__ expo, double* __restrict__ dscale, int radix256) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ncols; j += (int64_t)gridDim.x * blockDim.x) {
        double mx;
        const unsigned long long b = bits[j];
        memcpy(&mx, &b, sizeof(double));
        int e = 0;
        if (mx > 0.0) {
            const double f = frexp(mx, &e);
            if (radix256 && f > 127.0 / 128.0) e++;
        }
        expo[j] = e;
        if (dscale) dscale[j] = ldexp(1.0, e);
    }
}

// Column maximum, exponent and the radix-256 digit slices of a SHORT block row (K <= 256 NH rows, NH = 2 or 4: the depth-512
// / depth-1024 panel of the blocked Cholesky, chol.cu) in one pass: one warp per column keeps its (up to) 8 NH values per
// lane in registers.  Lane l holds rows 256 h + 8 l .. + 7 for h < NH, so the loads are 2 KB contiguous per warp and each
// slice gets one 8-byte word per lane and h (256 B contiguous per warp).  Rows K .. ldd - 1 of the digit columns are zeroed.
template <int NH>
static __global__ void __launch_bounds__(256)
slice256_short_kernel(int K, int64_t ncols, const double* __restrict__ A, int64_t lda, double* __restrict__ dscale,
                      int nslices, int8_t* __restrict__ D, int64_t ldd, int64_t slice_stride) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int64_t j = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); j < ncols; j += (int64_t)gridDim.x * wpb) {
        const double* col = A + j * lda;
        double v[NH][8];
        double mx = 0.0;
#pragma unroll
        for (int h = 0; h < NH; h++)
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int k = h * 256 + lane * 8 + u;
                v[h][u] = (k < K) ? col[k] : 0.0;
                mx = fmax(mx, fabs(v[h][u]));
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int e = 0;
        if (mx > 0.0) {
            const double f = frexp(mx, &e);
            if (f > 127.0 / 128.0) e++;
        }
        if (lane == 0) dscale[j] = ldexp(1.0, e);
        const double sc = ldexp(1.0, 7 - e);
#pragma unroll
        for (int h = 0; h < NH; h++) {
            const int k0 = h * 256 + lane * 8;
            if (k0 >= ldd) continue;
            double rr[8];
#pragma unroll
            for (int u = 0; u < 8; u++) rr[u] = v[h][u] * sc;
            uint64_t w[8];
            slice256_pack8(rr, nslices, w);
            for (int s = 0; s < nslices; s++) *reinterpret_cast<uint64_t*>(D + s * slice_stride + j * ldd + k0) = w[s];
        }
    }
}

// Radix-256 variant: balanced signed digits d_s in [-128, 127], a = 2^e sum_s 2^-(7 + 8 s) d_s.  rint can produce
// +128 (remainder >= 0.498): a backward carry pass turns it into -128 and adds one to the next higher digit; the
// leading digit cannot overflow because |a| 2^(7 - e) <= 127.  Seven such digits carry the same 56 bits as eight
// radix-128 digits, so the product needs 28 digit pairs (s + t <= 6) instead of 36 at the same truncation error.
static __global__ void slice256_kernel(int64_t K, int64_t ncols, const double* __restrict__ A, int64_t lda,
                                const int* __restrict__ expo, int nslices, int8_t* __restrict__ D, int64_t ldd,
                                int64_t slice_stride, int64_t a_gstride = 0, int64_t d_gstride = 0) {
    // gridDim.z > 1: row groups (the grouped products of the matrix-cone congruences): group z reads its K rows at row
    // offset z * a_gstride of every column and writes its digit rows at offset z * d_gstride (a multiple of 64, so that
    // the TMA coordinates of the groups stay 16-byte aligned whatever the row pitch of the FP64 operand is)
    A += (int64_t)blockIdx.z * a_gstride;
    D += (int64_t)blockIdx.z * d_gstride;
    const int64_t K8 = (K + 7) / 8;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double sc = ldexp(1.0, 7 - expo[j]);
        const double* col = A + j * lda;
        for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < K8;
             g += (int64_t)gridDim.x * blockDim.x) {
            double rr[8];
            if (g * 8 + 8 <= K && ((reinterpret_cast<uintptr_t>(col) & 15) == 0)) {
                // 16-byte loads: half the load instructions, and every lane uses half of each 32-byte sector it touches
                const double2* c2 = reinterpret_cast<const double2*>(col + g * 8);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double2 v = c2[u];
                    rr[2 * u] = v.x * sc;
                    rr[2 * u + 1] = v.y * sc;
                }
            } else {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int64_t k = g * 8 + u;
                    rr[u] = (k < K) ? col[k] * sc : 0.0;        // |r| <= 127
                }
            }
            uint64_t w[8];
            slice256_pack8(rr, nslices, w);
            for (int s = 0; s < nslices; s++)
                *reinterpret_cast<uint64_t*>(D + s * slice_stride + g * 8 + j * ldd) = w[s];
        }
    }
}

}  // namespace hypdev

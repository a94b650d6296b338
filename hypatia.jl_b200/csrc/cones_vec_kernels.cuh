// Device oracles of the two vector cones of the benchmarked path - Nonnegative (nonnegative.jl:42-145) and
// EpiNormEucl (epinormeucl.jl:44-228) - and the per-cone reductions of check_numerics / get_proxsqr
// (Cones.jl:273-310).  "Thread per element" for the orthant, "warp per (cone, column)" for second-order cones, plus the
// many-column chunk kernel of the Schur pre-pass.  HBM-bound streaming kernels: 16 * rows * ncols algorithmic bytes per
// product.  Launched from cones.cu; compiled for the host by tests/emu/ (CPU-tier tests: tests/test_emu_vec.py).
#pragma once
#include "devdefs.cuh"
#include "ozaki_slice_kernels.cuh"

// product modes (= HYP_PROD_* of the ABI; cones.cu static_asserts the equality) and the orthant's type code
#define VK_HESS 0
#define VK_INV_HESS 1
#define VK_SQRT_HESS 2
#define VK_INV_SQRT_HESS 3
#define VK_NONNEGATIVE 0

namespace hypdev {

// ------------------------------------------------------------------ Nonnegative
// nonnegative.jl:44-60: is_feas = all(point > eps), is_dual_feas likewise, grad = -1 / point
static __global__ void nn_state_kernel(int64_t nrows, const int* __restrict__ rows,
                                const int* __restrict__ rowcone, const double* __restrict__ point,
                                const double* __restrict__ dual, double* __restrict__ grad,
                                uint8_t* feas, uint8_t* dual_feas) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows;
         i += (int64_t)gridDim.x * blockDim.x) {
        int r = rows[i];
        double s = point[r], z = dual[r];
        grad[r] = -1.0 / s;
        if (!(s > HYP_EPS)) feas[rowcone[i]] = 0;
        if (!(z > HYP_EPS)) dual_feas[rowcone[i]] = 0;
    }
}

// nonnegative.jl:82-120: hess arr/s/s, inv_hess arr*s*s, sqrt arr/s, inv_sqrt arr*s
template <int MODE>
__global__ void nn_prod_kernel(int64_t nrows, const int* __restrict__ rows,
                               const double* __restrict__ point, const double* arr, int64_t ld_arr,
                               double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr - row_shift;
        double* pr = prod + j * ld_prod - row_shift;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows;
             i += (int64_t)gridDim.x * blockDim.x) {
            int r = rows[i];
            double s = point[r], v = a[r];
            double out;
            if (MODE == VK_HESS) out = v / s / s;
            else if (MODE == VK_INV_HESS) out = v * s * s;
            else if (MODE == VK_SQRT_HESS) out = v / s;
            else out = v * s;
            pr[r] = out;
        }
    }
}

// nonnegative.jl:122-125
static __global__ void nn_dder3_kernel(int64_t nrows, const int* __restrict__ rows,
                                const double* __restrict__ point, const double* __restrict__ dir,
                                double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nrows;
         i += (int64_t)gridDim.x * blockDim.x) {
        int r = rows[i];
        double s = point[r], t = dir[r] / s;
        out[r] = t * t / s;
    }
}

// ------------------------------------------------------------------ EpiNormEucl
// one warp per cone.  scal[8*c+0] = dist.  epinormeucl.jl:54-90
static __global__ void soc_state_kernel(int ncones, const int64_t* __restrict__ off,
                                 const int* __restrict__ dim, const int* __restrict__ kidx,
                                 const double* __restrict__ point, const double* __restrict__ dual,
                                 double* __restrict__ grad, double* __restrict__ scal, uint8_t* feas,
                                 uint8_t* dual_feas) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    double sw = 0.0, sdw = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        double w = point[o + i], dw = dual[o + i];
        sw += w * w;
        sdw += dw * dw;
    }
    sw = warp_sum(sw);
    sdw = warp_sum(sdw);
    const double u = point[o], du = dual[o];
    double dist = 0.0;
    bool ok = false;
    if (u > HYP_EPS) {
        dist = (u * u - sw) / 2;
        ok = dist > HYP_EPS;
    }
    bool dok = (du > HYP_EPS) && ((du * du - sdw) > 2 * HYP_EPS);
    for (int i = lane; i < d; i += 32) {
        double v = point[o + i] / dist;
        grad[o + i] = (i == 0) ? -v : v;
    }
    if (lane == 0) {
        scal[8 * c] = dist;
        feas[kidx[c]] = ok ? 1 : 0;
        dual_feas[kidx[c]] = dok ? 1 : 0;
    }
}

// one warp per (cone, column).  epinormeucl.jl:121-206
template <int MODE>
__global__ void __launch_bounds__(256)
soc_prod_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const double* __restrict__ scal, const double* __restrict__ point, const double* arr,
                int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    const double dist = scal[8 * c];
    const double u = point[o];
    const double rt2 = 1.4142135623730951;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        // pass 1: dot(w, wj); the first chunk of the column stays in a register
        double dotw = 0.0;
        const double a_first = (lane < d) ? a[lane] : 0.0;
        if (lane >= 1 && lane < d) dotw = point[o + lane] * a_first;
        for (int i = lane + 32; i < d; i += 32) dotw += point[o + i] * a[i];
        dotw = warp_sum(dotw);
        const double uj = __shfl_sync(0xffffffffu, a_first, 0);
        double c0, cw, cj;   // prod[0] = c0 ; prod[i] = cw * w[i] + cj * arr[i]
        if (MODE == VK_HESS) {
            double ga = (dotw - u * uj) / dist;
            c0 = (-ga * u - uj) / dist;
            cw = ga / dist;
            cj = 1.0 / dist;
        } else if (MODE == VK_INV_HESS) {
            double pa = u * uj + dotw;
            c0 = pa * u - dist * uj;
            cw = pa;
            cj = dist;
        } else if (MODE == VK_SQRT_HESS) {
            double distrt2 = dist * rt2, rtdist = sqrt(dist), urtdist = u + rtdist * rt2;
            c0 = (u * uj - dotw) / distrt2;
            cw = (dotw / urtdist - uj) / distrt2;
            cj = 1.0 / rtdist;
        } else {
            double rtdist = sqrt(dist), urtdist = u + rtdist * rt2;
            c0 = (u * uj + dotw) / rt2;
            cw = (dotw / urtdist + uj) / rt2;
            cj = rtdist;
        }
        if (lane < d) pr[lane] = (lane == 0) ? c0 : (cw * point[o + lane] + cj * a_first);
        for (int i = lane + 32; i < d; i += 32) pr[i] = cw * point[o + i] + cj * a[i];
    }
}

// Many-column variant used by the Schur pre-pass (K8): a CTA stages the rows of a chunk of
// consecutive cones of ONE column in shared memory with fully coalesced loads (a column of the
// panel is contiguous over all cones), one thread per cone then applies the rank-one update
// from shared memory (cone dims are mostly odd, e.g. 25: conflict-free strides), and the result
// goes back with coalesced stores.  chunk table: crow0[b], crows[b] = first row / number of rows
// of chunk b, ccone0[b], ccount[b] = first cone (index in the group) / number of cones.
// OUT selects what happens to the product (the fused Schur pre-pass of the digit-sliced SYRK, ozaki.cu, never
// materialises H^{1/2} G): 0 = store it; 3 = store it AND collect the column maxima (saves the slicer's first pass over
// the stored product); 1 = only the column maxima of |.| (atomicMax of the bit patterns into colbits[j]); 2 = cut it
// into radix-256 digit slices with the column exponents expo[j] and store the digits
// (chunks start at multiples of 8 rows, checked by the host, so the packed 8-byte words are aligned).
// rows per chunk: two staged arrays (the column of the panel and the cones' points) of this many doubles fit the 48 KB
// of dynamic shared memory that need no opt-in; dynamic shared memory of a launch = 2 * SOC_CHUNK_ROWS doubles
constexpr int SOC_CHUNK_ROWS = 3000;

template <int MODE, int OUT = 0>
__global__ void __launch_bounds__(256)
soc_prod_chunk_kernel(const int64_t* __restrict__ crow0, const int* __restrict__ crows,
                      const int* __restrict__ ccone0, const int* __restrict__ ccount,
                      const int64_t* __restrict__ off, const int* __restrict__ dim,
                      const double* __restrict__ scal, const double* __restrict__ point, const double* arr,
                      int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift,
                      unsigned long long* __restrict__ colbits = nullptr, const int* __restrict__ expo = nullptr,
                      int8_t* __restrict__ digits = nullptr, int64_t ldd = 0, int64_t slice_stride = 0, int nslices = 0) {
    HYP_DYN_SMEM(double, srow);
    const int b = blockIdx.x;
    const int64_t r0 = crow0[b];
    const int nr = crows[b], c0 = ccone0[b], nc = ccount[b];
    const double rt2 = 1.4142135623730951;
    // The points w of the chunk's cones, staged once per CTA (a chunk is a run of consecutive rows, so this is one coalesced
    // copy).  Read from global memory by one thread per cone they cost 32 sectors per load instruction (lanes 25 doubles
    // apart), twice per entry and column - about half of the load / store pipe time of a column; the host gives every
    // CTA several columns (gridDim.y < ncols) so the copy is amortised.
    double* sw = srow + SOC_CHUNK_ROWS;
    for (int i = threadIdx.x; i < nr; i += blockDim.x) sw[i] = point[r0 + i];
    // Thread t owns cone c0 + t for every column this CTA handles (a chunk holds at most blockDim.x cones: host), so
    // whatever does not depend on the column is computed here, once: the square root and the reciprocals of the rank-one
    // formulas (epinormeucl.jl:117-205).  The columns then cost multiplications only - the kernel was bound by its
    // instruction count (ncu, profiles/r02_soc_prepass_ncu.md: issue slots 46 % busy at 50 % occupancy, FP64 divisions
    // and the square root being ~30-instruction sequences each).  Products by a reciprocal differ from the quotient by
    // at most an ulp; the one- and two-column kernels above keep the divisions.
    const bool has = (int)threadIdx.x < nc;
    int vo = 0, d = 0;
    double u = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0;
    if (has) {
        const int c = c0 + (int)threadIdx.x;
        const int64_t o = off[c];
        d = dim[c];
        vo = (int)(o - r0);
        const double dist = scal[8 * c];
        u = point[o];
        if (MODE == VK_HESS) {
            q0 = 1.0 / dist;
        } else if (MODE == VK_INV_HESS) {
            q0 = dist;
        } else {
            const double rtdist = sqrt(dist);
            q1 = 1.0 / (u + rtdist * rt2);
            if (MODE == VK_SQRT_HESS) {
                q0 = 1.0 / (dist * rt2);
                q2 = 1.0 / rtdist;
            } else {
                q0 = 1.0 / rt2;
                q2 = rtdist;
            }
        }
    }
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (r0 - row_shift);
        double* pr = prod + j * ld_prod + (r0 - row_shift);
        // eight loads in flight per thread before the first shared-memory store: with one load per loop trip the pass ran at
        // 2.7 TB/s, bound by memory-level parallelism (64 warps x 256 B per SM against ~1 us of DRAM latency), not by HBM
        for (int i0 = threadIdx.x; i0 < nr; i0 += 8 * blockDim.x) {
            double t8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * (int)blockDim.x;
                t8[u] = (i < nr) ? a[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * (int)blockDim.x;
                if (i < nr) srow[i] = t8[u];
            }
        }
        __syncthreads();
        double mx = 0.0;   // OUT 1 / 3: largest |entry| this thread produced for column j
        if (has) {
            double* v = srow + vo;
            const double* w = sw + vo;
            const double uj = v[0];
            double s0 = 0.0, s1 = 0.0;
            int i = 1;
            for (; i + 8 <= d; i += 8) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    s0 += w[i + e] * v[i + e];
                    s1 += w[i + e + 1] * v[i + e + 1];
                }
            }
            for (; i < d; i++) s0 += w[i] * v[i];
            const double dotw = s0 + s1;
            double k0, kw, kj;
            if (MODE == VK_HESS) {
                const double ga = (dotw - u * uj) * q0;
                k0 = (-ga * u - uj) * q0; kw = ga * q0; kj = q0;
            } else if (MODE == VK_INV_HESS) {
                const double pa = u * uj + dotw;
                k0 = pa * u - q0 * uj; kw = pa; kj = q0;
            } else if (MODE == VK_SQRT_HESS) {
                k0 = (u * uj - dotw) * q0; kw = (dotw * q1 - uj) * q0; kj = q2;
            } else {
                k0 = (u * uj + dotw) * q0; kw = (dotw * q1 + uj) * q0; kj = q2;
            }
            v[0] = k0;
            mx = fabs(k0);
            i = 1;
            for (; i + 8 <= d; i += 8) {
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const double r = kw * w[i + e] + kj * v[i + e];
                    v[i + e] = r;
                    mx = fmax(mx, fabs(r));
                }
            }
            for (; i < d; i++) {
                const double r = kw * w[i] + kj * v[i];
                v[i] = r;
                mx = fmax(mx, fabs(r));
            }
        }
        __syncthreads();
        if (OUT == 0 || OUT == 3) {
            for (int i = threadIdx.x; i < nr; i += blockDim.x) pr[i] = srow[i];
        }
        if (OUT == 1 || OUT == 3) {
            __shared__ double smx[8];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, smx[w]);
                unsigned long long b;
                memcpy(&b, &mx, sizeof(double));
                if (b) atomicMax(colbits + j, b);
            }
        } else if (OUT == 2) {
            const double sc = ldexp(1.0, 7 - expo[j]);
            int8_t* dcol = digits + j * ldd + (r0 - row_shift);
            for (int g8 = threadIdx.x; g8 * 8 < nr; g8 += blockDim.x) {
                double rr[8];
#pragma unroll
                for (int u = 0; u < 8; u++) rr[u] = (g8 * 8 + u < nr) ? srow[g8 * 8 + u] * sc : 0.0;
                uint64_t w[8];
                slice256_pack8(rr, nslices, w);
                for (int sl = 0; sl < nslices; sl++)
                    *reinterpret_cast<uint64_t*>(dcol + sl * slice_stride + g8 * 8) = w[sl];
            }
        }
        __syncthreads();
    }
}

// epinormeucl.jl:208-228
static __global__ void soc_dder3_kernel(int ncones, const int64_t* __restrict__ off,
                                 const int* __restrict__ dim, const double* __restrict__ scal,
                                 const double* __restrict__ point, const double* __restrict__ dir,
                                 double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    const double dist = scal[8 * c];
    const double u = point[o], ud = dir[o];
    double sww = 0.0, swd = 0.0, sdd = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        double w = point[o + i], wd = dir[o + i];
        sww += w * w;
        swd += w * wd;
        sdd += wd * wd;
    }
    sww = warp_sum(sww);
    swd = warp_sum(swd);
    sdd = warp_sum(sdd);
    const double jdotpd = u * ud - swd;
    const double ga = (swd - u * ud) / dist;
    const double h0 = (-ga * u - ud) / dist;                 // (H dir)[0]
    // (H dir)[i] = (ga * w_i + wd_i) / dist
    const double dHd = ud * h0 + (ga * swd + sdd) / dist;    // dir' H dir
    const double pHd = u * h0 + (ga * sww + swd) / dist;     // point' H dir
    const double dotdHd = -dHd, dotpHd = pHd;
    const double inv2d = 1.0 / (2 * dist);
    for (int i = lane; i < d; i += 32) {
        double r;
        if (i == 0) {
            r = h0 * jdotpd - dotdHd * u - dotpHd * ud;
        } else {
            double w = point[o + i], wd = dir[o + i];
            r = (ga * w + wd) / dist * jdotpd + dotdHd * w + dotpHd * wd;
        }
        out[o + i] = r * inv2d;
    }
}

// ------------------------------------------------------------------ per-cone reductions
// One CTA per local cone.  check_numerics (Cones.jl:273-290) + get_proxsqr (Cones.jl:294-310;
// nonnegative.jl:137-145 for the orthant).  v1 = irtmu*dual + grad, v2 = Hinv v1, v3 = Hinv grad.
static __global__ void __launch_bounds__(128)
cone_prox_kernel(int cone_lo, const int* __restrict__ ctype, const int64_t* __restrict__ coff,
                 const int64_t* __restrict__ cdim, const double* __restrict__ cnu,
                 const double* __restrict__ point, const double* __restrict__ dual,
                 const double* __restrict__ grad, const double* __restrict__ v1,
                 const double* __restrict__ v2, const double* __restrict__ v3, double irtmu,
                 int use_max, double* __restrict__ proxsqr, uint8_t* __restrict__ num_ok) {
    __shared__ double sm[4][4];
    const int k = cone_lo + blockIdx.x;
    const int64_t o = coff[k], d = cdim[k];
    const int type = ctype[k];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;   // v2.v1, grad.point, v3.grad, nonneg aggregate
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) {
        double g = grad[o + i];
        a0 += v2[o + i] * v1[o + i];
        a1 += g * point[o + i];
        a2 += v3[o + i] * g;
        if (type == VK_NONNEGATIVE) {
            double t = point[o + i] * dual[o + i] * irtmu - 1.0;
            t = t * t;
            a3 = use_max ? fmax(a3, t) : a3 + t;
        }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (use_max) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) a3 = fmax(a3, __shfl_xor_sync(0xffffffffu, a3, s));
    } else {
        a3 = warp_sum(a3);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sm[w][0] = a0;
        sm[w][1] = a1;
        sm[w][2] = a2;
        sm[w][3] = a3;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        for (int i = 0; i < 4; i++) {
            d0 += sm[i][0];
            d1 += sm[i][1];
            d2 += sm[i][2];
            d3 = use_max ? fmax(d3, sm[i][3]) : d3 + sm[i][3];
        }
        const double nu = cnu[k];
        const double gtol = 1.220703125e-4;              // eps^(1/4)
        const double Htol = 10 * 0.011048543456039806;   // 10 * sqrt(gtol)
        bool ok = true;
        if (fabs(1 + d1 / nu) > gtol * (double)d) ok = false;
        if (ok && fabs(1 - d2 / nu) > Htol * (double)d) ok = false;
        if (!(d1 == d1) || !(d2 == d2)) ok = false;
        num_ok[k] = ok ? 1 : 0;
        double prox;
        if (type == VK_NONNEGATIVE) {
            prox = d3;
        } else {
            const double negtol = 1.4901161193847656e-08;   // sqrt(eps)
            prox = (d0 < -negtol * (double)d) ? INFINITY : fabs(d0);
        }
        proxsqr[k] = prox;
    }
}

}  // namespace hypdev

// Kernels of the rook-pivoted symmetric-indefinite factorisation A = L D L' and of its solve (ldlt.cu drives them: one
// five-kernel sequence per column, all pivoting decisions on the device).  reference: bunchkaufman!(A, true) =
// LAPACK dsytrf_rook / dsytrs_rook (src/linearalgebra/dense.jl:164-184), increase_diag! (dense.jl:106-113).
// Compiled for the host by tests/emu/ as well (CPU-tier tests: tests/test_emu_ldlt.py).
#pragma once
#include "devdefs.cuh"

namespace hypdev {

// st[0] = k (cursor), st[1] = kstep of the current column (0 before the first), st[2] = p,
// st[3] = kp, st[4] = info, st[5] = skip (singular column: no elimination)
enum { ST_K = 0, ST_STEP, ST_P, ST_KP, ST_INFO, ST_SKIP, ST_NUM };

static __global__ void symmetrize_kernel(double* __restrict__ A, int64_t lda, int64_t n) {
    // mirror the upper triangle into the lower one
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y)
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < j;
             i += (int64_t)gridDim.x * blockDim.x)
            A[j + i * lda] = A[i + j * lda];
}

static __global__ void increase_diag_kernel(double* __restrict__ A, int64_t lda, int64_t n) {
    // dense.jl:106-113: A_jj = (1 + 1e-5) * max(A_jj, 1000 eps)
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x) {
        double d = A[j + j * lda];
        A[j + j * lda] = (1.0 + 1e-5) * fmax(d, 1000.0 * HYP_EPS);
    }
}

// block-wide argmax of |col[i]| over i in [lo, hi), i != skip; returns value, index via smem
static __device__ void block_argmax(const double* __restrict__ col, int64_t lo, int64_t hi, int64_t skip,
                             double* sval, int* sidx, double& vmax, int& imax) {
    double best = -1.0;
    int bi = -1;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        if (i == skip) continue;
        double v = fabs(col[i]);
        if (v > best || !(v == v)) {
            best = (v == v) ? v : INFINITY;
            bi = (int)i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi >= 0 && (bi < 0 || oi < bi))) {
            best = ov;
            bi = oi;
        }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        sval[threadIdx.x >> 5] = best;
        sidx[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    best = -1.0;
    bi = -1;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        if (sval[w] > best || (sval[w] == best && sidx[w] >= 0 && (bi < 0 || sidx[w] < bi))) {
            best = sval[w];
            bi = sidx[w];
        }
    }
    vmax = best < 0 ? 0.0 : best;
    imax = bi;
}

// rook pivot search for column k (dsytf2_rook, lower variant, on full symmetric storage)
static __global__ void __launch_bounds__(1024)
pivot_kernel(const double* __restrict__ A, int64_t lda, int64_t n, int* __restrict__ st,
             int* __restrict__ ipiv) {
    __shared__ double sval[32];
    __shared__ int sidx[32];
    // advance the cursor past the previous column(s)
    const int64_t k = (int64_t)st[ST_K] + st[ST_STEP];
    __syncthreads();
    if (k >= n) {
        if (threadIdx.x == 0) {
            st[ST_K] = (int)n;
            st[ST_STEP] = 0;
        }
        return;
    }
    const double alpha = (1.0 + sqrt(17.0)) / 8.0;
    const double absakk = fabs(A[k + k * lda]);
    double colmax = 0.0;
    int imax = -1;
    if (k < n - 1) block_argmax(A + k * lda, k + 1, n, -1, sval, sidx, colmax, imax);
    int kstep = 1, p = (int)k, kp = (int)k, skip = 0, info = 0;
    if (!(fmax(absakk, colmax) > 0.0)) {
        info = (int)k + 1;
        skip = 1;
    } else if (!(absakk >= alpha * colmax)) {
        for (int it = 0; it < 4096; it++) {
            double rowmax;
            int jmax;
            block_argmax(A + (int64_t)imax * lda, k, n, imax, sval, sidx, rowmax, jmax);
            if (fabs(A[imax + (int64_t)imax * lda]) >= alpha * rowmax) {
                kp = imax;
                break;
            } else if (p == jmax || rowmax <= colmax) {
                kp = imax;
                kstep = 2;
                break;
            } else {
                p = imax;
                colmax = rowmax;
                imax = jmax;
            }
        }
    }
    if (kstep == 2 && k + 1 >= n) {   // cannot form a 2x2 block at the last column
        kstep = 1;
        kp = (int)k;
    }
    if (threadIdx.x == 0) {
        st[ST_K] = (int)k;
        st[ST_STEP] = kstep;
        st[ST_P] = p;
        st[ST_KP] = kp;
        st[ST_SKIP] = skip;
        if (info && st[ST_INFO] == 0) st[ST_INFO] = info;
        // ipiv[3k] = kstep, ipiv[3k+1] = first interchange partner, ipiv[3k+2] = second
        ipiv[3 * k] = kstep;
        if (kstep == 1) {
            ipiv[3 * k + 1] = kp;
            ipiv[3 * k + 2] = (int)k;
        } else {
            ipiv[3 * k + 1] = p;
            ipiv[3 * k + 2] = kp;
            ipiv[3 * (k + 1)] = 0;
        }
    }
}

// symmetric interchange of rows / columns a and b inside the trailing block A[k:n, k:n]
// which == 0: (k, p) when kstep == 2;  which == 1: (k + kstep - 1, kp)
static __global__ void swap_kernel(double* __restrict__ A, int64_t lda, int64_t n, const int* __restrict__ st,
                            int which) {
    const int64_t k = st[ST_K];
    const int kstep = st[ST_STEP];
    if (k >= n || kstep == 0) return;
    int64_t a, b;
    if (which == 0) {
        if (kstep != 2) return;
        a = k;
        b = st[ST_P];
    } else {
        a = k + kstep - 1;
        b = st[ST_KP];
    }
    if (a == b) return;
    for (int64_t t = k + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (t == a) {
            double x = A[a + a * lda];
            A[a + a * lda] = A[b + b * lda];
            A[b + b * lda] = x;
        } else if (t != b) {
            double x = A[t + a * lda], y = A[t + b * lda];
            A[t + a * lda] = y;
            A[t + b * lda] = x;
            A[a + t * lda] = y;
            A[b + t * lda] = x;
        }
    }
}

// pivot column(s): keep the original entries in work[0:n], work[n:2n], write L into A and into
// work[2n:3n], work[3n:4n]
static __global__ void colprep_kernel(double* __restrict__ A, int64_t lda, int64_t n, const int* __restrict__ st,
                               double* __restrict__ work) {
    const int64_t k = st[ST_K];
    const int kstep = st[ST_STEP];
    if (k >= n || kstep == 0 || st[ST_SKIP]) return;
    if (kstep == 1) {
        const double d = A[k + k * lda];
        for (int64_t i = k + 1 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
             i += (int64_t)gridDim.x * blockDim.x) {
            double a = A[i + k * lda];
            double l = a / d;
            work[i] = a;
            work[2 * n + i] = l;
            A[i + k * lda] = l;
        }
    } else {
        const double d11 = A[k + k * lda], d21 = A[k + 1 + k * lda], d22 = A[k + 1 + (k + 1) * lda];
        const double det = d11 * d22 - d21 * d21;
        for (int64_t i = k + 2 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
             i += (int64_t)gridDim.x * blockDim.x) {
            double a1 = A[i + k * lda], a2 = A[i + (k + 1) * lda];
            double w1 = (d22 * a1 - d21 * a2) / det;
            double w2 = (d11 * a2 - d21 * a1) / det;
            work[i] = a1;
            work[n + i] = a2;
            work[2 * n + i] = w1;
            work[3 * n + i] = w2;
            A[i + k * lda] = w1;
            A[i + (k + 1) * lda] = w2;
        }
    }
}

// trailing update A[i, j] -= a_i l_j (+ a2_i l2_j) for i, j > k + kstep - 1 (both triangles)
static __global__ void __launch_bounds__(256)
update_kernel(double* __restrict__ A, int64_t lda, int64_t n, const int* __restrict__ st,
              const double* __restrict__ work) {
    const int64_t k = st[ST_K];
    const int kstep = st[ST_STEP];
    if (k >= n || kstep == 0 || st[ST_SKIP]) return;
    const int64_t t0 = k + kstep;
    const int64_t len = n - t0;
    if (len <= 0) return;
    // 2-D tiling: blockIdx.y strides over columns, threads over rows
    for (int64_t j = t0 + blockIdx.y; j < n; j += gridDim.y) {
        const double l1 = work[2 * n + j];
        const double l2 = kstep == 2 ? work[3 * n + j] : 0.0;
        double* col = A + j * lda;
        for (int64_t i = t0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
             i += (int64_t)gridDim.x * blockDim.x) {
            double v = col[i] - work[i] * l1;
            if (kstep == 2) v -= work[n + i] * l2;
            col[i] = v;
        }
    }
}

static __global__ void init_state_kernel(int* st) {
    if (threadIdx.x < ST_NUM) st[threadIdx.x] = 0;
}

static __global__ void finish_info_kernel(const int* __restrict__ st, int* __restrict__ info) {
    if (threadIdx.x == 0) info[0] = st[ST_INFO];
}

// x <- A^-1 x with the factorisation above (dsytrs_rook order of operations); one CTA
static __global__ void __launch_bounds__(1024)
ldlt_solve_kernel(const double* __restrict__ A, int64_t lda, int64_t n, const int* __restrict__ ipiv,
                  double* __restrict__ x) {
    __shared__ double sred[2][32];
    __shared__ double sb[2];
    const int tid = threadIdx.x;
    // forward: L D y = P' b
    int64_t k = 0;
    while (k < n) {
        const int kstep = ipiv[3 * k];
        if (tid == 0) {
            int p1 = ipiv[3 * k + 1], p2 = ipiv[3 * k + 2];
            if (kstep == 1) {
                if (p1 != k) { double t = x[k]; x[k] = x[p1]; x[p1] = t; }
            } else {
                if (p1 != k) { double t = x[k]; x[k] = x[p1]; x[p1] = t; }
                if (p2 != k + 1) { double t = x[k + 1]; x[k + 1] = x[p2]; x[p2] = t; }
            }
            sb[0] = x[k];
            if (kstep == 2) sb[1] = x[k + 1];
        }
        __syncthreads();
        if (kstep == 1) {
            const double bk = sb[0];
            for (int64_t i = k + 1 + tid; i < n; i += blockDim.x) x[i] -= A[i + k * lda] * bk;
            if (tid == 0) {
                double d = A[k + k * lda];
                x[k] = (d != 0.0) ? bk / d : bk;
            }
        } else {
            const double b1 = sb[0], b2 = sb[1];
            for (int64_t i = k + 2 + tid; i < n; i += blockDim.x)
                x[i] -= A[i + k * lda] * b1 + A[i + (k + 1) * lda] * b2;
            if (tid == 0) {
                double d11 = A[k + k * lda], d21 = A[k + 1 + k * lda], d22 = A[k + 1 + (k + 1) * lda];
                double det = d11 * d22 - d21 * d21;
                x[k] = (d22 * b1 - d21 * b2) / det;
                x[k + 1] = (d11 * b2 - d21 * b1) / det;
            }
        }
        __syncthreads();
        k += kstep;
    }
    // backward: L' x = y, then undo the interchanges
    k = n - 1;
    while (k >= 0) {
        // the block ending at k: a 2x2 block when the entry of column k is the "second column" mark
        const int second = (ipiv[3 * k] == 0);
        const int64_t k0 = second ? k - 1 : k;
        double a1 = 0.0, a2 = 0.0;
        for (int64_t i = k + 1 + tid; i < n; i += blockDim.x) {
            double xi = x[i];
            a1 += A[i + k0 * lda] * xi;
            if (second) a2 += A[i + k * lda] * xi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if ((tid & 31) == 0) {
            sred[0][tid >> 5] = a1;
            sred[1][tid >> 5] = a2;
        }
        __syncthreads();
        if (tid == 0) {
            double s1 = 0.0, s2 = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
                s1 += sred[0][w];
                s2 += sred[1][w];
            }
            int p1 = ipiv[3 * k0 + 1], p2 = ipiv[3 * k0 + 2];
            if (!second) {
                x[k] -= s1;
                if (p1 != k) { double t = x[k]; x[k] = x[p1]; x[p1] = t; }
            } else {
                x[k0] -= s1;
                x[k] -= s2;
                // dsytrs_rook: interchange K with -IPIV(K), then K-1 with -IPIV(K-1)
                if (p2 != k) { double t = x[k]; x[k] = x[p2]; x[p2] = t; }
                if (p1 != k0) { double t = x[k0]; x[k0] = x[p1]; x[p1] = t; }
            }
        }
        __syncthreads();
        k = k0 - 1;
    }
}

}  // namespace hypdev

// Batched symmetric eigendecomposition for the spectral cones (K9 of SURVEY.md section 2.3).
//
// replaces: update_eigen! = LAPACK.syev!('V', 'U', X), src/linearalgebra/dense.jl:69, called by
// update_feas / is_dual_feas of EpiPerSepSpectral{MatrixCSqr}, matrixcsqr.jl:91-138.
//
// One CTA per matrix, the whole problem resident in shared memory (A and V: 2 d (d|1) doubles, up to
// d = 119 inside the 227 KB of an sm_100 CTA; larger matrices run the same code on a global-memory
// scratch).  Algorithm: cyclic two-sided Jacobi with the round-robin ("chess tournament") parallel
// ordering - every step applies n/2 disjoint plane rotations at once:
//   phase A  thread t < n/2 picks its pair (p, q) and the rotation that annihilates a_pq;
//   phase B  A <- A J, V <- V J   (columns p, q of every pair; unit-stride, conflict-free);
//   phase C  A <- J' A            (rows p, q; leading dimension d|1 keeps the stride odd).
// Disjoint rotations commute, so a step equals n/2 sequential Jacobi rotations.  Sweeps stop when a
// whole sweep applied no rotation (|a_pq| <= eps * max(sqrt|a_pp a_qq|, 1e-4 |A|_F): at least the
// absolute accuracy of syev, relative accuracy on positive definite input).  Eigenvalues are
// returned ascending like syev (the oracles are invariant to the order; the sorted order only feeds
// the "sorted triple" rule of the second divided differences, matrixcsqr.jl:449-502).
// Latency-bound: ~3 (n-1) barriers per sweep, 8-20 sweeps.
#pragma once
#include "devdefs.cuh"

#define HYP_SYEVJ_THREADS 512
#define HYP_SYEVJ_MAX_SWEEPS 40

namespace hypdev {

// Ain / Vout: d x d column-major blocks with leading dimension lde = d rounded up to even at
// in_off[c] (both triangles of Ain filled).  lam: d eigenvalues at lam_off[c].  divv (optional):
// matrix c is divided by divv[div_off[c] + div_idx] before the decomposition (W / v; a divisor
// that is not > eps is replaced by 1 - the caller flags that cone infeasible).
template <bool SMEM, bool WANTV>
__global__ void __launch_bounds__(HYP_SYEVJ_THREADS, 1)
syevj_batched_kernel(int nmat, const int* __restrict__ sides, const int64_t* __restrict__ in_off,
                     const double* __restrict__ Ain, double* __restrict__ Vout,
                     const int64_t* __restrict__ lam_off, double* __restrict__ lam,
                     const double* __restrict__ divv, const int64_t* __restrict__ div_off, int div_idx,
                     double* __restrict__ gwork, int64_t gwork_stride) {
    HYP_DYN_SMEM(double, dyn);
    __shared__ double red[HYP_SYEVJ_THREADS / 32];
    __shared__ int s_rot;
    const int c = blockIdx.x;
    if (c >= nmat) return;
    const int d = sides[c];
    if (d <= 0) return;
    const int lde = (d + 1) & ~1;
    const int lda = d | 1;
    const int n = d + (d & 1);           // even number of players (one dummy when d is odd)
    const int half = n >> 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* A = SMEM ? dyn : gwork + (int64_t)c * gwork_stride;
    double* V = A + (int64_t)d * lda;
    // rotation table of the current step
    double* rc = SMEM ? dyn + (WANTV ? 2 : 1) * (int64_t)d * lda : gwork + (int64_t)c * gwork_stride + 2 * (int64_t)d * lda;
    double* rs = rc + half;
    int* rp = (int*)(rs + half);
    int* rq = rp + half;
    int* perm = rq + half;

    const double* Ac = Ain + in_off[c];
    double div = 1.0;
    if (divv) {
        div = divv[div_off[c] + div_idx];
        if (!(div > HYP_EPS)) div = 1.0;
    }
    const double idiv = 1.0 / div;
    double nrm = 0.0;
    for (int idx = tid; idx < d * d; idx += nt) {
        const int i = idx % d, j = idx / d;
        const double x = Ac[i + (int64_t)j * lde] * idiv;
        A[i + j * lda] = x;
        if (WANTV) V[i + j * lda] = (i == j) ? 1.0 : 0.0;
        nrm += x * x;
    }
    nrm = sqrt(hypdev::block_sum(nrm, red));
    const double floor_abs = 1e-4 * nrm;

    if (d > 1) {
        for (int sweep = 0; sweep < HYP_SYEVJ_MAX_SWEEPS; sweep++) {
            if (tid == 0) s_rot = 0;
            __syncthreads();
            for (int r = 0; r < n - 1; r++) {
                // ---- phase A: pairs and rotations ----
                if (tid < half) {
                    int p, q;
                    if (tid == 0) {
                        p = n - 1;
                        q = r;
                    } else {
                        p = (r + tid) % (n - 1);
                        q = (r - tid + n - 1) % (n - 1);
                    }
                    if (p > q) {
                        const int t = p;
                        p = q;
                        q = t;
                    }
                    double cs = 1.0, sn = 0.0;
                    int pp = -1;
                    if (q < d) {
                        const double apq = A[p + q * lda], app = A[p + p * lda], aqq = A[q + q * lda];
                        if (fabs(apq) > HYP_EPS * fmax(sqrt(fabs(app * aqq)), floor_abs)) {
                            const double th = (aqq - app) / (2.0 * apq);
                            const double t = copysign(1.0, th) / (fabs(th) + sqrt(th * th + 1.0));
                            cs = 1.0 / sqrt(t * t + 1.0);
                            sn = t * cs;
                            pp = p;
                            HYP_RAISE_FLAG(s_rot);
                        }
                    }
                    rc[tid] = cs;
                    rs[tid] = sn;
                    rp[tid] = pp;
                    rq[tid] = q;
                }
                __syncthreads();
                // ---- phase B: columns of A and V ----
                for (int idx = tid; idx < half * d; idx += nt) {
                    const int t = idx / d, i = idx - t * d;
                    const int p = rp[t];
                    if (p < 0) continue;
                    const int q = rq[t];
                    const double cs = rc[t], sn = rs[t];
                    const double ap = A[i + p * lda], aq = A[i + q * lda];
                    A[i + p * lda] = cs * ap - sn * aq;
                    A[i + q * lda] = sn * ap + cs * aq;
                    if (WANTV) {
                        const double vp = V[i + p * lda], vq = V[i + q * lda];
                        V[i + p * lda] = cs * vp - sn * vq;
                        V[i + q * lda] = sn * vp + cs * vq;
                    }
                }
                __syncthreads();
                // ---- phase C: rows of A ----
                for (int idx = tid; idx < half * d; idx += nt) {
                    const int t = idx / d, j = idx - t * d;
                    const int p = rp[t];
                    if (p < 0) continue;
                    const int q = rq[t];
                    const double cs = rc[t], sn = rs[t];
                    const double ap = A[p + j * lda], aq = A[q + j * lda];
                    A[p + j * lda] = cs * ap - sn * aq;
                    A[q + j * lda] = sn * ap + cs * aq;
                }
                __syncthreads();
            }
            const int any = s_rot;
            __syncthreads();
            if (!any) break;
        }
    }
    // ---- ascending order (rank sort), eigenvectors out ----
    double* lc = lam + lam_off[c];
    for (int i = tid; i < d; i += nt) {
        const double li = A[i + i * lda];
        int rank = 0;
        for (int j = 0; j < d; j++) {
            const double lj = A[j + j * lda];
            rank += (lj < li || (lj == li && j < i)) ? 1 : 0;
        }
        lc[rank] = li;
        perm[rank] = i;
    }
    __syncthreads();
    if (WANTV) {
        double* Vc = Vout + in_off[c];
        for (int idx = tid; idx < d * d; idx += nt) {
            const int i = idx % d, j = idx / d;
            Vc[i + (int64_t)j * lde] = V[i + perm[j] * lda];
        }
    }
}


// doubles of working storage per matrix: A (+ V), the rotation table (2 * n/2 doubles) and three
// int arrays (n/2, n/2, d)
__host__ __device__ inline int64_t syevj_work_doubles(int d, bool wantv) {
    const int lda = d | 1, n = d + (d & 1);
    return (int64_t)(wantv ? 2 : 1) * d * lda + n + (n + d + 2) / 2 + 2;
}

}  // namespace hypdev

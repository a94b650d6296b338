// One pass over G for BOTH products of apply_lhs / calc_convergence_params:  w = G x  and  y = G' z.
//
// reference call sites: mul!(res.x, G', dir.z) and mul!(res.z, G, dir.x) in apply_lhs (common.jl:91-103),
// mul!(x_residual, G', z) and mul!(z_residual, G, x) in calc_convergence_params (Solvers.jl:431-449): the
// reference reads G twice; both products only need each entry of G once.
// Thread = 2 consecutive rows, CTA = 128 threads = 256 rows, grid.y = column chunks (as gemv_n_kernel).
// G x: per-thread accumulators over the chunk -> partialN[chunk][row].  G' z: per column a sum over rows:
// eight columns at a time are reduced over the warp with a halving butterfly (9 double shuffles per 8 columns,
// fixed order => deterministic) -> partialT[row_block * 4 + warp][column]; two small reduce kernels finish.
// HBM-bound: 8 * rows * ncols algorithmic bytes for both products together.
#pragma once
#include "devdefs.cuh"

namespace hypdev {

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

static __global__ void __launch_bounds__(128)
gemv_nt_kernel(int64_t rows, int64_t ncols, const double* __restrict__ M, int64_t ld,
               const double* __restrict__ x, const double* __restrict__ z, int64_t cols_per_chunk,
               double* __restrict__ partialN, double* __restrict__ partialT) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 2;
    const int64_t j0 = (int64_t)blockIdx.y * cols_per_chunk;
    const int64_t j1 = j0 + cols_per_chunk < ncols ? j0 + cols_per_chunk : ncols;
    const bool ok0 = r < rows, ok1 = r + 1 < rows;
    const double z0 = ok0 ? z[r] : 0.0, z1 = ok1 ? z[r + 1] : 0.0;
    const double* base = M + r;
    double* pT = partialT + ((int64_t)blockIdx.x * 4 + warp) * ncols;
    double ax = 0.0, ay = 0.0;
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    int64_t j = j0;
    for (; j + 7 < j1; j += 8) {
        double t[8];
        double2 mm[8];
        if (ok1) {
            // all eight 16-byte loads are issued before the first use
#pragma unroll
            for (int u = 0; u < 8; u++) mm[u] = __ldg(reinterpret_cast<const double2*>(base + (j + u) * ld));
        } else {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                mm[u].x = ok0 ? base[(j + u) * ld] : 0.0;
                mm[u].y = 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double xj = x[j + u];
            ax += mm[u].x * xj;
            ay += mm[u].y * xj;
            t[u] = mm[u].x * z0 + mm[u].y * z1;
        }
        // halving butterfly: after the three steps lane L holds the partial sum of column
        // c = 4 * bit4(L) + 2 * bit3(L) + bit2(L) over the 8 lanes that share those bits
        double u4[4], u2[2], u1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double send = b4 ? t[i] : t[i + 4];
            const double keep = b4 ? t[i + 4] : t[i];
            u4[i] = keep + shfl_xor_d(send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double send = b3 ? u4[i] : u4[i + 2];
            const double keep = b3 ? u4[i + 2] : u4[i];
            u2[i] = keep + shfl_xor_d(send, 8);
        }
        {
            const double send = b2 ? u2[0] : u2[1];
            const double keep = b2 ? u2[1] : u2[0];
            u1 = keep + shfl_xor_d(send, 4);
        }
        u1 += shfl_xor_d(u1, 2);
        u1 += shfl_xor_d(u1, 1);
        if ((lane & 3) == 0) pT[j + (lane >> 2)] = u1;
    }
    for (; j < j1; j++) {
        double m0 = 0.0, m1 = 0.0;
        if (ok0) m0 = base[j * ld];
        if (ok1) m1 = base[j * ld + 1];
        const double xj = x[j];
        ax += m0 * xj;
        ay += m1 * xj;
        double t = m0 * z0 + m1 * z1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += shfl_xor_d(t, o);
        if (lane == 0) pT[j] = t;
    }
    double* outN = partialN + (int64_t)blockIdx.y * rows + r;
    if (ok0) outN[0] = ax;
    if (ok1) outN[1] = ay;
}

// y[j] = alpha * sum_p partial[p][j] + beta * y[j]
static __global__ void gemv_t_reduce_kernel(int64_t ncols, int nparts, const double* __restrict__ partial,
                                            double alpha, double beta, double* __restrict__ y) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < ncols;
         j += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < nparts; p++) acc += partial[(int64_t)p * ncols + j];
        y[j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[j]);
    }
}

// y[r] = alpha * sum_c partial[c][r] + beta * y[r]   (same as gemv_n_reduce_kernel of gemv.cu)
static __global__ void gemv_n_reduce2_kernel(int64_t rows, int nchunks, const double* __restrict__ partial,
                                             double alpha, double beta, double* __restrict__ y) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < nchunks; c++) acc += partial[(int64_t)c * rows + r];
        y[r] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[r]);
    }
}

// ---- the single-product kernels (K6): mul!(.., G', z) qrchol.jl:52, mul!(Gx, G, x) qrchol.jl:73, common.jl:91,94,144 ----
// y[j] = alpha * dot(M[:, j], x) + beta * y[j];  one CTA per column.
template <bool VEC>
__global__ void __launch_bounds__(256) gemv_t_cta_kernel(int64_t rows, int64_t ncols,
                                                         const double* __restrict__ M, int64_t ld,
                                                         const double* __restrict__ x, double alpha,
                                                         double beta, double* __restrict__ y) {
    __shared__ double sm[8];
    for (int64_t j = blockIdx.x; j < ncols; j += gridDim.x) {
        const double* col = M + j * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (VEC) {
            const double2* c2 = reinterpret_cast<const double2*>(col);
            const double2* x2 = reinterpret_cast<const double2*>(x);
            int64_t n2 = rows >> 1;
            int64_t i = threadIdx.x;
            for (; i + 3 * 256 < n2; i += 4 * 256) {
                double2 m0 = __ldg(c2 + i), m1 = __ldg(c2 + i + 256), m2 = __ldg(c2 + i + 512),
                        m3 = __ldg(c2 + i + 768);
                double2 v0 = x2[i], v1 = x2[i + 256], v2 = x2[i + 512], v3 = x2[i + 768];
                a0 += m0.x * v0.x + m0.y * v0.y;
                a1 += m1.x * v1.x + m1.y * v1.y;
                a2 += m2.x * v2.x + m2.y * v2.y;
                a3 += m3.x * v3.x + m3.y * v3.y;
            }
            for (; i < n2; i += 256) {
                double2 m0 = __ldg(c2 + i);
                double2 v0 = x2[i];
                a0 += m0.x * v0.x + m0.y * v0.y;
            }
            if ((rows & 1) && threadIdx.x == 0) a1 += col[rows - 1] * x[rows - 1];
        } else {
            for (int64_t i = threadIdx.x; i < rows; i += 256) a0 += col[i] * x[i];
        }
        double acc = warp_sum((a0 + a1) + (a2 + a3));
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) t += sm[w];
            y[j] = alpha * t + (beta == 0.0 ? 0.0 : beta * y[j]);
        }
        __syncthreads();
    }
}

// one warp per column (short columns)
static __global__ void __launch_bounds__(256) gemv_t_warp_kernel(int64_t rows, int64_t ncols,
                                                          const double* __restrict__ M, int64_t ld,
                                                          const double* __restrict__ x, double alpha,
                                                          double beta, double* __restrict__ y) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t j = warp; j < ncols; j += nwarps) {
        const double* col = M + j * ld;
        double acc = 0;
        for (int64_t i = lane; i < rows; i += 32) acc += col[i] * x[i];
        acc = warp_sum(acc);
        if (lane == 0) y[j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[j]);
    }
}

// partial[chunk][r] = sum_{j in chunk} M[r, j] x[j]; thread owns 2 consecutive rows.
template <bool VEC>
__global__ void __launch_bounds__(128) gemv_n_kernel(int64_t rows, int64_t ncols,
                                                     const double* __restrict__ M, int64_t ld,
                                                     const double* __restrict__ x, int64_t cols_per_chunk,
                                                     double* __restrict__ partial) {
    int64_t r = (blockIdx.x * 128 + threadIdx.x) * 2;
    int64_t j0 = blockIdx.y * cols_per_chunk;
    int64_t j1 = j0 + cols_per_chunk < ncols ? j0 + cols_per_chunk : ncols;
    if (r >= rows) return;
    double ax = 0, ay = 0, bx = 0, by = 0;
    if (VEC && r + 1 < rows) {
        const double* base = M + r;
        int64_t j = j0;
        for (; j + 7 < j1; j += 8) {
            double2 m[8];
#pragma unroll
            for (int u = 0; u < 8; u++) m[u] = __ldg(reinterpret_cast<const double2*>(base + (j + u) * ld));
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                double x0 = x[j + u], x1 = x[j + u + 1];
                ax += m[u].x * x0;
                ay += m[u].y * x0;
                bx += m[u + 1].x * x1;
                by += m[u + 1].y * x1;
            }
        }
        for (; j < j1; j++) {
            double2 m0 = __ldg(reinterpret_cast<const double2*>(base + j * ld));
            double x0 = x[j];
            ax += m0.x * x0;
            ay += m0.y * x0;
        }
        double* out = partial + blockIdx.y * rows + r;
        out[0] = ax + bx;
        out[1] = ay + by;
    } else {
        for (int64_t j = j0; j < j1; j++) {
            double x0 = x[j];
            ax += M[r + j * ld] * x0;
            if (r + 1 < rows) ay += M[r + 1 + j * ld] * x0;
        }
        double* out = partial + blockIdx.y * rows + r;
        out[0] = ax;
        if (r + 1 < rows) out[1] = ay;
    }
}

static __global__ void gemv_n_reduce_kernel(int64_t rows, int nchunks, const double* __restrict__ partial,
                                     double alpha, double beta, double* __restrict__ y) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0;
        for (int c = 0; c < nchunks; c++) acc += partial[(int64_t)c * rows + r];
        y[r] = alpha * acc + (beta == 0.0 ? 0.0 : beta * y[r]);
    }
}


// ---- two right-hand sides per pass over G (hyp_solve_system_multi / hyp_apply_lhs_multi) -------------------------------
// The stepper's data flow (steppers/combined.jl:67-79) allows {cent, pred} and {centadj, predadj} to be solved
// together; every entry of G then serves two products.  Same thread mappings and reduction orders as the
// single-vector kernels above, so column v of a two-column call is bit-identical to a single-column call.

// y_v[j] = alpha * dot(M[:, j], x_v) + beta * y_v[j], v = 0, 1;  one CTA per column of M.
static __global__ void __launch_bounds__(256)
gemv_t2_cta_kernel(int64_t rows, int64_t ncols, const double* __restrict__ M, int64_t ld, const double* __restrict__ x0,
                   const double* __restrict__ x1, double alpha, double beta, double* __restrict__ y0,
                   double* __restrict__ y1) {
    __shared__ double sm[2][8];
    for (int64_t j = blockIdx.x; j < ncols; j += gridDim.x) {
        const double* col = M + j * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        const double2* c2 = reinterpret_cast<const double2*>(col);
        const double2* p0 = reinterpret_cast<const double2*>(x0);
        const double2* p1 = reinterpret_cast<const double2*>(x1);
        const int64_t n2 = rows >> 1;
        int64_t i = threadIdx.x;
        for (; i + 3 * 256 < n2; i += 4 * 256) {
            const double2 m0 = __ldg(c2 + i), m1 = __ldg(c2 + i + 256), m2 = __ldg(c2 + i + 512), m3 = __ldg(c2 + i + 768);
            const double2 v0 = p0[i], v1 = p0[i + 256], v2 = p0[i + 512], v3 = p0[i + 768];
            const double2 w0 = p1[i], w1 = p1[i + 256], w2 = p1[i + 512], w3 = p1[i + 768];
            a0 += m0.x * v0.x + m0.y * v0.y;
            a1 += m1.x * v1.x + m1.y * v1.y;
            a2 += m2.x * v2.x + m2.y * v2.y;
            a3 += m3.x * v3.x + m3.y * v3.y;
            b0 += m0.x * w0.x + m0.y * w0.y;
            b1 += m1.x * w1.x + m1.y * w1.y;
            b2 += m2.x * w2.x + m2.y * w2.y;
            b3 += m3.x * w3.x + m3.y * w3.y;
        }
        for (; i < n2; i += 256) {
            const double2 m0 = __ldg(c2 + i);
            const double2 v0 = p0[i], w0 = p1[i];
            a0 += m0.x * v0.x + m0.y * v0.y;
            b0 += m0.x * w0.x + m0.y * w0.y;
        }
        if ((rows & 1) && threadIdx.x == 0) {
            a1 += col[rows - 1] * x0[rows - 1];
            b1 += col[rows - 1] * x1[rows - 1];
        }
        const double sa = warp_sum((a0 + a1) + (a2 + a3)), sb = warp_sum((b0 + b1) + (b2 + b3));
        if ((threadIdx.x & 31) == 0) {
            sm[0][threadIdx.x >> 5] = sa;
            sm[1][threadIdx.x >> 5] = sb;
        }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) t += sm[threadIdx.x][w];
            double* y = threadIdx.x ? y1 : y0;
            y[j] = alpha * t + (beta == 0.0 ? 0.0 : beta * y[j]);
        }
        __syncthreads();
    }
}

// partial_v[chunk][r] = sum_{j in chunk} M[r, j] x_v[j], v = 0, 1 (partial_1 = partial_0 + pstride)
static __global__ void __launch_bounds__(128)
gemv_n2_kernel(int64_t rows, int64_t ncols, const double* __restrict__ M, int64_t ld, const double* __restrict__ x0,
               const double* __restrict__ x1, int64_t cols_per_chunk, double* __restrict__ partial, int64_t pstride) {
    const int64_t r = (blockIdx.x * 128 + threadIdx.x) * 2;
    const int64_t j0 = blockIdx.y * cols_per_chunk;
    const int64_t j1 = j0 + cols_per_chunk < ncols ? j0 + cols_per_chunk : ncols;
    if (r >= rows) return;
    const bool ok1 = r + 1 < rows;
    double ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0, dx = 0, dy = 0;
    const double* base = M + r;
    int64_t j = j0;
    if (ok1) {
        for (; j + 7 < j1; j += 8) {
            double2 m[8];
#pragma unroll
            for (int u = 0; u < 8; u++) m[u] = __ldg(reinterpret_cast<const double2*>(base + (j + u) * ld));
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const double p0 = x0[j + u], p1 = x0[j + u + 1], q0 = x1[j + u], q1 = x1[j + u + 1];
                ax += m[u].x * p0;
                ay += m[u].y * p0;
                bx += m[u + 1].x * p1;
                by += m[u + 1].y * p1;
                cx += m[u].x * q0;
                cy += m[u].y * q0;
                dx += m[u + 1].x * q1;
                dy += m[u + 1].y * q1;
            }
        }
    }
    for (; j < j1; j++) {
        const double m0 = base[j * ld], m1 = ok1 ? base[j * ld + 1] : 0.0;
        const double p0 = x0[j], q0 = x1[j];
        ax += m0 * p0;
        ay += m1 * p0;
        cx += m0 * q0;
        cy += m1 * q0;
    }
    double* out = partial + blockIdx.y * rows + r;
    out[0] = ax + bx;
    out[pstride] = cx + dx;
    if (ok1) {
        out[1] = ay + by;
        out[pstride + 1] = cy + dy;
    }
}

// two (x, z) pairs per pass:  w_v = M x_v  and  y_v = M' z_v  (see gemv_nt_kernel); partial buffers of pair 1 follow
// those of pair 0 at pstrideN / pstrideT
static __global__ void __launch_bounds__(128)
gemv_nt2_kernel(int64_t rows, int64_t ncols, const double* __restrict__ M, int64_t ld, const double* __restrict__ xa,
                const double* __restrict__ xb, const double* __restrict__ za, const double* __restrict__ zb,
                int64_t cols_per_chunk, double* __restrict__ partialN, int64_t pstrideN, double* __restrict__ partialT,
                int64_t pstrideT) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 2;
    const int64_t j0 = (int64_t)blockIdx.y * cols_per_chunk;
    const int64_t j1 = j0 + cols_per_chunk < ncols ? j0 + cols_per_chunk : ncols;
    const bool ok0 = r < rows, ok1 = r + 1 < rows;
    const double za0 = ok0 ? za[r] : 0.0, za1 = ok1 ? za[r + 1] : 0.0;
    const double zb0 = ok0 ? zb[r] : 0.0, zb1 = ok1 ? zb[r + 1] : 0.0;
    const double* base = M + r;
    double* pTa = partialT + ((int64_t)blockIdx.x * 4 + warp) * ncols;
    double* pTb = pTa + pstrideT;
    double axa = 0.0, aya = 0.0, axb = 0.0, ayb = 0.0;
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
    int64_t j = j0;
    for (; j + 7 < j1; j += 8) {
        double ta[8], tb[8];
        double2 mm[8];
        if (ok1) {
#pragma unroll
            for (int u = 0; u < 8; u++) mm[u] = __ldg(reinterpret_cast<const double2*>(base + (j + u) * ld));
        } else {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                mm[u].x = ok0 ? base[(j + u) * ld] : 0.0;
                mm[u].y = 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double xja = xa[j + u], xjb = xb[j + u];
            axa += mm[u].x * xja;
            aya += mm[u].y * xja;
            axb += mm[u].x * xjb;
            ayb += mm[u].y * xjb;
            ta[u] = mm[u].x * za0 + mm[u].y * za1;
            tb[u] = mm[u].x * zb0 + mm[u].y * zb1;
        }
#pragma unroll
        for (int v = 0; v < 2; v++) {
            double* t = v ? tb : ta;
            double u4[4], u2[2], u1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double send = b4 ? t[i] : t[i + 4];
                const double keep = b4 ? t[i + 4] : t[i];
                u4[i] = keep + shfl_xor_d(send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double send = b3 ? u4[i] : u4[i + 2];
                const double keep = b3 ? u4[i + 2] : u4[i];
                u2[i] = keep + shfl_xor_d(send, 8);
            }
            {
                const double send = b2 ? u2[0] : u2[1];
                const double keep = b2 ? u2[1] : u2[0];
                u1 = keep + shfl_xor_d(send, 4);
            }
            u1 += shfl_xor_d(u1, 2);
            u1 += shfl_xor_d(u1, 1);
            if ((lane & 3) == 0) (v ? pTb : pTa)[j + (lane >> 2)] = u1;
        }
    }
    for (; j < j1; j++) {
        double m0 = 0.0, m1 = 0.0;
        if (ok0) m0 = base[j * ld];
        if (ok1) m1 = base[j * ld + 1];
        const double xja = xa[j], xjb = xb[j];
        axa += m0 * xja;
        aya += m1 * xja;
        axb += m0 * xjb;
        ayb += m1 * xjb;
        double t0 = m0 * za0 + m1 * za1, t1 = m0 * zb0 + m1 * zb1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t0 += shfl_xor_d(t0, o);
            t1 += shfl_xor_d(t1, o);
        }
        if (lane == 0) {
            pTa[j] = t0;
            pTb[j] = t1;
        }
    }
    double* outN = partialN + (int64_t)blockIdx.y * rows + r;
    if (ok0) {
        outN[0] = axa;
        outN[pstrideN] = axb;
    }
    if (ok1) {
        outN[1] = aya;
        outN[pstrideN + 1] = ayb;
    }
}

}  // namespace hypdev

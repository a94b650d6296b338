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

}  // namespace hypdev

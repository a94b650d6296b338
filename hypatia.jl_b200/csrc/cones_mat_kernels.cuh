// Pack / unpack kernels of the matrix-domain cones (svec <-> smat, arrayutilities.jl:163-236), shared by
// cones_mat.cu (PosSemidefTri, log-det, root-det) and cones_spec.cu (EpiPerSepSpectral).
#pragma once
#include "devdefs.cuh"

namespace hypdev {

#define HYP_RT2 1.4142135623730951
#define HYP_IRT2 0.7071067811865476

// smat of the matrix part of `vec` for every cone of the group -> A (and B if given), full symmetric
static __global__ void __launch_bounds__(256)
unpack_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                    const int64_t* __restrict__ moff, int lead, const double* __restrict__ vec,
                    double* __restrict__ A, double* __restrict__ B) {
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int d = sides[c], lde = (d + 1) & ~1;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    const double* v = vec + off[c] + lead;
    double* Ac = A + moff[c];
    double* Bc = B ? B + moff[c] : nullptr;
    for (int64_t idx = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; idx < len;
         idx += (int64_t)gridDim.y * blockDim.x) {
        int a, b;
        svec_rc(idx, a, b);
        double x = v[idx];
        if (a != b) x *= HYP_IRT2;
        Ac[a + (int64_t)b * lde] = x;
        Ac[b + (int64_t)a * lde] = x;
        if (Bc) {
            Bc[a + (int64_t)b * lde] = x;
            Bc[b + (int64_t)a * lde] = x;
        }
    }
}

// columns [j0, j0 + cc) of one cone block of `arr` -> Mall = [M_0 ... M_{cc-1}], each d x lde (ld lde)
static __global__ void __launch_bounds__(256)
unpack_cols_kernel(int d, int lde, int64_t len, const double* arr, int64_t ld_arr, int64_t cc,
                   double* __restrict__ Mall) {
    for (int64_t j = blockIdx.y; j < cc; j += gridDim.y) {
        const double* v = arr + j * ld_arr;
        double* Mj = Mall + j * (int64_t)lde * lde;
        for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < len;
             idx += (int64_t)gridDim.x * blockDim.x) {
            int a, b;
            svec_rc(idx, a, b);
            double x = v[idx];
            if (a != b) x *= HYP_IRT2;
            Mj[a + (int64_t)b * lde] = x;
            Mj[b + (int64_t)a * lde] = x;
        }
        // blocks are lde columns wide (TMA coordinates must be even): keep the pad column finite
        if (lde > d && blockIdx.x == 0)
            for (int a = threadIdx.x; a < lde; a += blockDim.x) Mj[a + (int64_t)d * lde] = 0.0;
    }
}

// prod[idx, j] = alpha_j * svec(Y_j)[idx] + beta_j * vecB[idx]
static __global__ void __launch_bounds__(256)
pack_cols_kernel(int d, int lde, int64_t len, const double* __restrict__ Yall, int64_t cc,
                 const double* __restrict__ alpha, const double* __restrict__ beta,
                 const double* __restrict__ vecB, double* prod, int64_t ld_prod) {
    for (int64_t j = blockIdx.y; j < cc; j += gridDim.y) {
        const double* Yj = Yall + j * (int64_t)lde * lde;
        double* pr = prod + j * ld_prod;
        const double al = alpha ? alpha[j] : 1.0;
        const double be = beta ? beta[j] : 0.0;
        for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < len;
             idx += (int64_t)gridDim.x * blockDim.x) {
            int a, b;
            svec_rc(idx, a, b);
            double x = Yj[a + (int64_t)b * lde];
            if (a != b) x *= HYP_RT2;
            x *= al;
            if (vecB) x += be * vecB[idx];
            pr[idx] = x;
        }
    }
}

}  // namespace hypdev

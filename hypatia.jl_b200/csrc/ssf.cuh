// Separable spectral functions of EpiPerSepSpectral (reference: src/Cones/epipersepspectral/sepspectralfun.jl:17-116):
// value, derivatives and convex conjugate of h at a point; shared by the matrix (cones_spec_kernels.cuh) and vector
// (cones_vec3_kernels.cuh) cones of squares.
#pragma once
#include "devdefs.cuh"

namespace hypdev {

// h, h', h'', h''' of a separable spectral function at x (sepspectralfun.jl:17-110)
__device__ __forceinline__ void ssf_eval(int kind, double p, double x, double& h, double& d1, double& d2,
                                         double& d3) {
    if (kind == SSF_INV) {
        const double xi = 1.0 / x;
        h = xi;
        d1 = -xi * xi;
        d2 = 2.0 * xi * xi * xi;
        d3 = -6.0 * xi * xi * xi * xi;
    } else if (kind == SSF_NEGLOG) {
        const double xi = 1.0 / x;
        h = -log(x);
        d1 = -xi;
        d2 = xi * xi;
        d3 = -2.0 * xi * xi * xi;
    } else if (kind == SSF_NEGENTROPY) {
        const double lx = log(x), xi = 1.0 / x;
        h = x * lx;
        d1 = 1.0 + lx;
        d2 = xi;
        d3 = -xi * xi;
    } else {
        h = pow(x, p);
        d1 = p * pow(x, p - 1.0);
        d2 = p * (p - 1.0) * pow(x, p - 2.0);
        d3 = p * (p - 1.0) * (p - 2.0) * pow(x, p - 3.0);
    }
}

// one term of the convex conjugate (sepspectralfun.jl:22, :42, :62, :85-89)
__device__ __forceinline__ double ssf_conj(int kind, double p, double x) {
    if (kind == SSF_INV) return -2.0 * sqrt(x);
    if (kind == SSF_NEGLOG) return -1.0 - log(x);
    if (kind == SSF_NEGENTROPY) return exp(-x - 1.0);
    const double qq = p / (p - 1.0);
    return x >= 0.0 ? 0.0 : (p - 1.0) * pow(fabs(x) / p, qq);
}

__device__ __forceinline__ bool ssf_conj_dom_pos(int kind) { return kind == SSF_INV || kind == SSF_NEGLOG; }

}  // namespace hypdev

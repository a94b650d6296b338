// Device oracles of GeneralizedPower and HypoPowerMean and the GENERIC inverse-Hessian product for cones without
// closed forms.
//
// reference: src/Cones/generalizedpower.jl:77-236 (update_feas, is_dual_feas, update_grad, update_hess, hess_prod!,
// dder3); the cone defines no inv_hess_prod!, so the reference falls back to the generic oracles of
// src/Cones/Cones.jl:113-118, 239-259: explicit Hessian -> posdef_fact_copy! -> hess_fact \ arr.  Here: the state
// kernel writes the explicit dim x dim Hessian of every cone, the batched Cholesky (chol.cu) factors it and
// returns U^-1, and gen_invhess_prod_kernel applies H^-1 = U^-1 U^-T with two triangular mat-vecs per column.
// Layout: point = (u in R^m, w in R^n), m = mu[c] powers alpha at aoff[c]; scal: 0 z, 1 |w|^2, 2 zw, 3 zwzwi.
// HypoPowerMean (hypopowermean.jl:74-203): point = (u, w in R^d), powers alpha (d entries); scal: 0 phi, 1 zeta.
// One warp per cone (state, dder3) / per (cone, column) (products); HBM-bound, 16 * dim B per column.
#pragma once
#include "devdefs.cuh"

namespace hypdev {

static __global__ void __launch_bounds__(256)
gpow_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ mu, const int64_t* __restrict__ aoff,
                  const double* __restrict__ alpha, const int* __restrict__ kidx,
                  const int64_t* __restrict__ moff, const double* __restrict__ point,
                  const double* __restrict__ dual, double* __restrict__ grad, double* __restrict__ scal,
                  double* __restrict__ H, uint8_t* feas, uint8_t* dual_feas) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], m = mu[c], lde = (d + 1) & ~1;
    const double* al = alpha + aoff[c];
    double nbad = 0.0, dbad = 0.0, sl = 0.0, dsl = 0.0, w2 = 0.0, dw2 = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double x = point[o + i], y = dual[o + i];
        if (i < m) {
            if (!(x > HYP_EPS)) nbad += 1.0;
            if (!(y > HYP_EPS)) dbad += 1.0;
            sl += al[i] * log(x);
            dsl += al[i] * log(y / al[i]);
        } else {
            w2 += x * x;
            dw2 += y * y;
        }
    }
    nbad = warp_sum(nbad);
    dbad = warp_sum(dbad);
    sl = warp_sum(sl);
    dsl = warp_sum(dsl);
    w2 = warp_sum(w2);
    dw2 = warp_sum(dw2);
    const double z = exp(2.0 * sl), zw = z - w2;
    const bool ok = nbad == 0.0 && zw > HYP_EPS;
    const bool dok = dbad == 0.0 && (exp(2.0 * dsl) - dw2) > HYP_EPS;
    const double zwzwi = (z + w2) / zw;
    for (int i = lane; i < d; i += 32) {
        const double x = point[o + i];
        grad[o + i] = i < m ? -(zwzwi * al[i] + 1.0) / x : 2.0 * x / zw;
    }
    if (lane == 0) {
        double* sc = scal + 8 * c;
        sc[0] = z;
        sc[1] = w2;
        sc[2] = zw;
        sc[3] = zwzwi;
        if (!ok) feas[kidx[c]] = 0;
        if (!dok) dual_feas[kidx[c]] = 0;
    }
    // explicit Hessian, both triangles (generalizedpower.jl:122-163)
    double* Hc = H + moff[c];
    const double zzwim1 = -w2 / zw, zwi = 2.0 / zw;
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx % d, j = idx / d;
        const double xi = point[o + i], xj = point[o + j];
        double v;
        if (i < m && j < m) {
            const double aui = 2.0 * al[i] / xi, auizzwj = -z * (2.0 * al[j] / xj) / zw;
            v = aui * auizzwj * zzwim1;
            if (i == j) v += (zwzwi * al[i] + 1.0) / (xi * xi);
        } else if (i >= m && j >= m) {
            v = (2.0 * xi / zw) * (2.0 * xj / zw);
            if (i == j) v += zwi;
        } else {
            const int iu = i < m ? i : j, iw = i < m ? j : i;
            const double xu = point[o + iu], xw = point[o + iw];
            v = (-z * (2.0 * al[iu] / xu) / zw) * (2.0 * xw / zw);
        }
        Hc[i + (int64_t)j * lde] = v;
    }
}

// hess_prod!, generalizedpower.jl:165-200.  want_dual: -1 = every cone, 0 / 1 = only cones with that dualf flag.
static __global__ void __launch_bounds__(256)
gpow_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ mu, const int64_t* __restrict__ aoff,
                 const double* __restrict__ alpha, const int* __restrict__ dualf,
                 const double* __restrict__ scal, const double* __restrict__ point, const double* arr,
                 int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    const int64_t o = off[c];
    const int d = dim[c], m = mu[c];
    const double* al = alpha + aoff[c];
    const double z = scal[8 * c], zw = scal[8 * c + 2], zwzwi = scal[8 * c + 3];
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        double s1 = 0.0, s2 = 0.0;
        for (int i = lane; i < d; i += 32) {
            const double x = point[o + i];
            if (i < m) s1 += al[i] * (a[i] / x);
            else s2 += x * (2.0 * a[i] / zw);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        const double dot1 = -4.0 * s1 * z / zw;
        const double dot2 = (dot1 + 2.0 * s2) / zw;
        const double dot3 = dot1 - dot2 * z;
        for (int i = lane; i < d; i += 32) {
            const double x = point[o + i], ai = a[i];
            pr[i] = i < m ? ((ai / x) * (1.0 + zwzwi * al[i]) + dot3 * al[i]) / x : 2.0 * ai / zw + dot2 * x;
        }
    }
}

// dder3, generalizedpower.jl:202-236
static __global__ void __launch_bounds__(256)
gpow_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ mu, const int64_t* __restrict__ aoff,
                  const double* __restrict__ alpha, const double* __restrict__ scal,
                  const double* __restrict__ point, const double* __restrict__ dir,
                  double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], m = mu[c];
    const double* al = alpha + aoff[c];
    const double z = scal[8 * c], w2 = scal[8 * c + 1], zw = scal[8 * c + 2], zwzwi = scal[8 * c + 3];
    double swd = 0.0, audu = 0.0, sau2 = 0.0, sdd = 0.0;
    for (int i = lane; i < d; i += 32) {
        const double x = point[o + i], di = dir[o + i];
        if (i < m) {
            const double udu = di / x;
            audu += al[i] * udu;
            sau2 += al[i] * udu * udu;
        } else {
            swd += x * di;
            sdd += di * di;
        }
    }
    swd = warp_sum(swd);
    audu = warp_sum(audu);
    sau2 = warp_sum(sau2);
    sdd = warp_sum(sdd);
    const double zzwi = 2.0 * z / zw, zwi = 2.0 / zw;
    const double wwd = 2.0 * swd, c15 = wwd / zw;
    const double c1 = 2.0 * zwzwi * audu * audu + sau2;
    const double c10 = sdd + wwd * c15;
    const double c13 = zzwi * (w2 * c1 - 2.0 * wwd * zwzwi * audu + c10) / zw;
    const double c14 = zzwi * (2.0 * audu * w2 - wwd) / zw;
    const double c6 = zwi * (z * (4.0 * audu * c15 - c1) - c10) / zw;
    const double c7 = zwi * (2.0 * z * audu - wwd) / zw;
    for (int i = lane; i < d; i += 32) {
        const double x = point[o + i], di = dir[o + i];
        if (i < m) {
            const double udu = di / x;
            out[o + i] = (c13 * al[i] + ((c14 + zwzwi * udu) * al[i] + udu) * udu) / x;
        } else {
            out[o + i] = c7 * di + c6 * x;
        }
    }
}

// ---- HypoPowerMean ----
static __global__ void __launch_bounds__(256)
hpm_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ aoff, const double* __restrict__ alpha,
                 const int* __restrict__ kidx, const int64_t* __restrict__ moff,
                 const double* __restrict__ point, const double* __restrict__ dual, double* __restrict__ grad,
                 double* __restrict__ scal, double* __restrict__ H, uint8_t* feas, uint8_t* dual_feas) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], lde = (d + 1) & ~1;
    const double* al = alpha + aoff[c] - 1;      // al[i] = power of entry i (i = 1 .. d - 1)
    const double u = point[o], du = dual[o];
    double nbad = 0.0, dbad = 0.0, sl = 0.0, dsl = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double w = point[o + i], y = dual[o + i];
        if (!(w > HYP_EPS)) nbad += 1.0;
        if (!(y > HYP_EPS)) dbad += 1.0;
        sl += al[i] * log(w);
        dsl += al[i] * log(y / al[i]);
    }
    nbad = warp_sum(nbad);
    dbad = warp_sum(dbad);
    sl = warp_sum(sl);
    dsl = warp_sum(dsl);
    const double phi = exp(sl), zeta = phi - u;
    const bool ok = nbad == 0.0 && zeta > HYP_EPS;
    const bool dok = du < -HYP_EPS && dbad == 0.0 && (exp(dsl) + du) > HYP_EPS;
    const double zip = phi / zeta;
    for (int i = 1 + lane; i < d; i += 32) grad[o + i] = (-zip * al[i] - 1.0) / point[o + i];
    if (lane == 0) {
        grad[o] = 1.0 / zeta;
        scal[8 * c] = phi;
        scal[8 * c + 1] = zeta;
        if (!ok) feas[kidx[c]] = 0;
        if (!dok) dual_feas[kidx[c]] = 0;
    }
    // explicit Hessian, both triangles (hypopowermean.jl:118-148)
    double* Hc = H + moff[c];
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx % d, j = idx / d;
        double v;
        if (i == 0 && j == 0) {
            v = 1.0 / (zeta * zeta);
        } else if (i == 0 || j == 0) {
            const int k = i + j;
            v = -(zip * al[k] / point[o + k]) / zeta;
        } else if (i == j) {
            const double w = point[o + i];
            v = (zip * (al[i] / w) * (1.0 + al[i] * (zip - 1.0)) + 1.0 / w) / w;
        } else {
            v = zip * (zip - 1.0) * (al[i] / point[o + i]) * (al[j] / point[o + j]);
        }
        Hc[i + (int64_t)j * lde] = v;
    }
}

// hess_prod!, hypopowermean.jl:150-175
static __global__ void __launch_bounds__(256)
hpm_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int64_t* __restrict__ aoff, const double* __restrict__ alpha,
                const int* __restrict__ dualf, const double* __restrict__ scal,
                const double* __restrict__ point, const double* arr, int64_t ld_arr, double* prod,
                int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    const int64_t o = off[c];
    const int d = dim[c];
    const double* al = alpha + aoff[c] - 1;
    const double phi = scal[8 * c], zeta = scal[8 * c + 1], zip = phi / zeta;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        const double p = a[0];
        double s = 0.0;
        for (int i = 1 + lane; i < d; i += 32) s += al[i] * (a[i] / point[o + i]);
        const double c0 = warp_sum(s);
        const double c1 = zip * c0 - p / zeta, c2 = c1 - c0;
        for (int i = 1 + lane; i < d; i += 32) {
            const double w = point[o + i], rwi = a[i] / w;
            pr[i] = (al[i] * zip * (c2 + rwi) + rwi) / w;
        }
        if (lane == 0) pr[0] = c1 / -zeta;
    }
}

// dder3, hypopowermean.jl:177-203
static __global__ void __launch_bounds__(256)
hpm_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ aoff, const double* __restrict__ alpha,
                 const double* __restrict__ scal, const double* __restrict__ point,
                 const double* __restrict__ dir, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    const double* al = alpha + aoff[c] - 1;
    const double phi = scal[8 * c], zeta = scal[8 * c + 1], zip = phi / zeta;
    const double p = dir[o];
    double s0 = 0.0, s6 = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double r = dir[o + i] / point[o + i];
        s0 += r * al[i];
        s6 += r * r * al[i];
    }
    const double c0 = warp_sum(s0), c6 = warp_sum(s6);
    const double zichi = (p - phi * c0) / zeta;
    const double c1 = zichi * zichi + zip * (c6 - c0 * c0) / 2;
    const double c7 = zip * (c1 - c6 / 2 + c0 * (zichi + c0 / 2));
    const double c8 = -zip * (zichi + c0);
    for (int i = 1 + lane; i < d; i += 32) {
        const double w = point[o + i], r = dir[o + i] / w;
        out[o + i] = (al[i] * (c7 + r * (c8 + zip * r)) + r * r) / w;
    }
    if (lane == 0) out[o] = -c1 / zeta;
}

// Generic inv_hess_prod! (Cones.jl:113-118): y = H^-1 x = U^-1 (U^-T x) with the inverse Cholesky factor
// Ui (upper triangular, dim x dim, leading dimension lde) of the explicit Hessian.  One warp per (cone, column);
// dims up to 128 (the batched Cholesky's limit).  want_dual as in gpow_prod_kernel.
static __global__ void __launch_bounds__(256)
gen_invhess_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                        const int64_t* __restrict__ moff, const int* __restrict__ dualf,
                        const double* __restrict__ Ui, const double* arr, int64_t ld_arr, double* prod,
                        int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    __shared__ double st[8][128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * (blockDim.x >> 5) + wid;
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    const int64_t o = off[c];
    const int d = dim[c], lde = (d + 1) & ~1;
    const double* U = Ui + moff[c];
    double* t = st[wid];
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        // t = Ui' x : t_k = sum_{i <= k} Ui[i, k] x_i  (column k of Ui is contiguous)
        for (int k = lane; k < d; k += 32) {
            const double* col = U + (int64_t)k * lde;
            double s = 0.0;
            for (int i = 0; i <= k; i++) s += col[i] * a[i];
            t[k] = s;
        }
        __syncwarp();
        // y = Ui t : y_i = sum_{k >= i} Ui[i, k] t_k
        for (int i = lane; i < d; i += 32) {
            double s = 0.0;
            for (int k = i; k < d; k++) s += U[i + (int64_t)k * lde] * t[k];
            pr[i] = s;
        }
        __syncwarp();
    }
}

}  // namespace hypdev

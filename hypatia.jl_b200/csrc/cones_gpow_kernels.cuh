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

// ---- EpiNormSpectral (real), epinormspectral.jl:107-294 ----
// point = (u, vec(W)) with W d1 x d2 column-major, d1 <= d2, dim = 1 + d1 d2 <= 128 (so d1 <= 11).  Per-cone
// workspace at vecs + voff[c]: tau = Z^-1 W (d1 d2), Zitau = Z^-1 tau (d1 d2), Zi = Z^-1 (d1^2), Uz = upper Cholesky
// factor of Z = u^2 I - W W' (d1^2), scratch (d1 d2).  scal: 0 Huu, 1 trZi2.  Every product with the d2 x d2 matrices
// of the reference (W' tau + I, W_dir' tau) is re-associated through d1 x d1 intermediates.

// x <- Z^-1 x for one column (Z = U'U, U upper, column-major with leading dimension d1); run by ONE thread
__device__ __forceinline__ void ens_zsolve_col(const double* U, int d1, double* x) {
    for (int i = 0; i < d1; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= U[k + i * d1] * x[k];
        x[i] = s / U[i + i * d1];
    }
    for (int i = d1 - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < d1; k++) s -= U[i + k * d1] * x[k];
        x[i] = s / U[i + i * d1];
    }
}
// warp-wide helpers (lane-strided); callers separate dependent steps with __syncwarp()
__device__ __forceinline__ void ens_zsolve(const double* U, double* X, int d1, int d2, int lane) {
    for (int k = lane; k < d2; k += 32) ens_zsolve_col(U, d1, X + k * d1);
}
// S (d1 x d1) = [S +] X Y'
__device__ __forceinline__ void ens_mm_nt(double* S, const double* X, const double* Y, int d1, int d2, int lane,
                                          bool acc) {
    for (int idx = lane; idx < d1 * d1; idx += 32) {
        const int a = idx % d1, b = idx / d1;
        double s = acc ? S[idx] : 0.0;
        for (int k = 0; k < d2; k++) s += X[a + k * d1] * Y[b + k * d1];
        S[idx] = s;
    }
}
// O (d1 x d2) = beta O + S Y
__device__ __forceinline__ void ens_mm_sy(double* O, const double* S, const double* Y, int d1, int d2, int lane,
                                          double beta) {
    for (int idx = lane; idx < d1 * d2; idx += 32) {
        const int r = idx % d1, k = idx / d1;
        double s = 0.0;
        for (int m = 0; m < d1; m++) s += S[r + m * d1] * Y[m + k * d1];
        O[idx] = beta == 0.0 ? s : beta * O[idx] + s;
    }
}

static __global__ void __launch_bounds__(256)
ens_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ d1s, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                 const int* __restrict__ kidx, const int64_t* __restrict__ moff,
                 const double* __restrict__ point, const double* __restrict__ dual, double* __restrict__ grad,
                 double* __restrict__ scal, double* __restrict__ H, uint8_t* feas, uint8_t* dual_feas) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], d1 = d1s[c], d2 = (d - 1) / d1, n12 = d1 * d2, lde = (d + 1) & ~1;
    double* tau = vecs + voff[c];
    double* Zitau = tau + n12;
    double* Zi = Zitau + n12;
    double* Uz = Zi + d1 * d1;
    double* scr = Uz + d1 * d1;
    const double u = point[o];
    const double* W = point + o + 1;
    // update_feas (:107-124): Z = u^2 I - W W', Cholesky
    for (int idx = lane; idx < d1 * d1; idx += 32) {
        const int a = idx % d1, b = idx / d1;
        double s = a == b ? u * u : 0.0;
        for (int k = 0; k < d2; k++) s -= W[a + k * d1] * W[b + k * d1];
        Uz[idx] = s;
    }
    __syncwarp();
    int ok = u > HYP_EPS ? 1 : 0;
    if (lane == 0) {
        for (int j = 0; j < d1; j++) {
            double s = Uz[j + j * d1];
            for (int k = 0; k < j; k++) s -= Uz[k + j * d1] * Uz[k + j * d1];
            if (!(s > 0.0)) {
                ok = 0;
                s = 1.0;
            }
            const double r = sqrt(s);
            Uz[j + j * d1] = r;
            for (int i = j + 1; i < d1; i++) {
                double t = Uz[j + i * d1];
                for (int k = 0; k < j; k++) t -= Uz[k + j * d1] * Uz[k + i * d1];
                Uz[j + i * d1] = t / r;
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    // update_grad (:135-151): tau = Z^-1 W, Zi = Z^-1; update_hess_aux (:153-172): Zitau, trZi2, Huu
    for (int b = lane; b < d1; b += 32) {
        double* x = Zi + b * d1;
        for (int i = 0; i < d1; i++) x[i] = i == b ? 1.0 : 0.0;
        ens_zsolve_col(Uz, d1, x);
    }
    for (int i = lane; i < n12; i += 32) tau[i] = W[i];
    __syncwarp();
    ens_zsolve(Uz, tau, d1, d2, lane);
    __syncwarp();
    for (int idx = lane; idx < d1 * d1; idx += 32) {     // copytri!(Zi, 'U')
        const int a = idx % d1, b = idx / d1;
        if (a > b) Zi[idx] = Zi[b + a * d1];
    }
    for (int i = lane; i < n12; i += 32) Zitau[i] = tau[i];
    __syncwarp();
    ens_zsolve(Uz, Zitau, d1, d2, lane);
    double tr = 0.0, tr2 = 0.0;
    for (int idx = lane; idx < d1 * d1; idx += 32) {
        const double z = Zi[idx];
        if (idx % d1 == idx / d1) tr += z;
        tr2 += z * z;
    }
    tr = warp_sum(tr);
    tr2 = warp_sum(tr2);
    __syncwarp();
    const double g0 = -2.0 * u * tr + (d1 - 1) / u;
    const double Huu = 4.0 * u * u * tr2 + (g0 - 2.0 * (d1 - 1) / u) / u;
    for (int i = lane; i < n12; i += 32) grad[o + 1 + i] = 2.0 * tau[i];
    // is_dual_feas (:126-133): u - sum of the singular values of the dual W; one-sided Jacobi on its d1 rows
    const double du = dual[o];
    for (int i = lane; i < n12; i += 32) scr[i] = dual[o + 1 + i];
    __syncwarp();
    for (int sweep = 0; sweep < 40; sweep++) {
        int rotated = 0;
        for (int p = 0; p < d1 - 1; p++)
            for (int q = p + 1; q < d1; q++) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int k = lane; k < d2; k += 32) {
                    const double x = scr[p + k * d1], y = scr[q + k * d1];
                    al += x * x;
                    be += y * y;
                    ga += x * y;
                }
                al = warp_sum(al);
                be = warp_sum(be);
                ga = warp_sum(ga);
                if (ga != 0.0 && fabs(ga) > HYP_EPS * sqrt(al * be)) {
                    rotated = 1;
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    for (int k = lane; k < d2; k += 32) {
                        const double x = scr[p + k * d1], y = scr[q + k * d1];
                        scr[p + k * d1] = cs * x - sn * y;
                        scr[q + k * d1] = sn * x + cs * y;
                    }
                    __syncwarp();
                }
            }
        if (!rotated) break;
    }
    double nuc = 0.0;
    for (int p = 0; p < d1; p++) {
        double al = 0.0;
        for (int k = lane; k < d2; k += 32) al += scr[p + k * d1] * scr[p + k * d1];
        nuc += sqrt(warp_sum(al));
    }
    const bool dok = du > HYP_EPS && (du - nuc) > HYP_EPS;
    if (lane == 0) {
        grad[o] = g0;
        scal[8 * c] = Huu;
        scal[8 * c + 1] = tr2;
        if (!ok) feas[kidx[c]] = 0;
        if (!dok) dual_feas[kidx[c]] = 0;
    }
    // explicit Hessian, both triangles (:174-214): H[(j,i),(l,k)] = 2 (Zi[l,j] (W'tau + I)[i,k] + tau[l,i] tau[j,k])
    double* Hc = H + moff[c];
    for (int idx = lane; idx < d * d; idx += 32) {
        const int r = idx % d, cc = idx / d;
        double v;
        if (r == 0 && cc == 0) {
            v = Huu;
        } else if (r == 0 || cc == 0) {
            v = -4.0 * u * Zitau[r + cc - 1];
        } else {
            const int j = (r - 1) % d1, i = (r - 1) / d1, l = (cc - 1) % d1, k = (cc - 1) / d1;
            double wt = i == k ? 1.0 : 0.0;
            for (int m = 0; m < d1; m++) wt += W[m + i * d1] * tau[m + k * d1];
            v = 2.0 * (Zi[l + j * d1] * wt + tau[l + i * d1] * tau[j + k * d1]);
        }
        Hc[r + (int64_t)cc * lde] = v;
    }
}

// hess_prod!, epinormspectral.jl:216-246
static __global__ void __launch_bounds__(256)
ens_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int* __restrict__ d1s, const int64_t* __restrict__ voff, const double* __restrict__ vecs,
                const int* __restrict__ dualf, const double* __restrict__ scal,
                const double* __restrict__ point, const double* arr, int64_t ld_arr, double* prod,
                int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    __shared__ double sh[8][256];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    double* T = sh[threadIdx.x >> 5];
    double* X = T + 128;
    const int64_t o = off[c];
    const int d = dim[c], d1 = d1s[c], d2 = (d - 1) / d1, n12 = d1 * d2;
    const double* tau = vecs + voff[c];
    const double* Zitau = tau + n12;
    const double* Uz = Zitau + n12 + d1 * d1;
    const double u = point[o], Huu = scal[8 * c];
    const double* W = point + o + 1;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        const double p = a[0];
        const double* R = a + 1;
        double s0 = 0.0;
        for (int i = lane; i < n12; i += 32) s0 += Zitau[i] * R[i];
        const double p0 = Huu * p - 4.0 * u * warp_sum(s0);
        for (int idx = lane; idx < d1 * d1; idx += 32) {
            const int r = idx % d1, b = idx / d1;
            double s = r == b ? -2.0 * u * p : 0.0;
            for (int k = 0; k < d2; k++) s += R[r + k * d1] * W[b + k * d1] + R[b + k * d1] * W[r + k * d1];
            T[idx] = s;
        }
        __syncwarp();
        for (int idx = lane; idx < n12; idx += 32) {
            const int r = idx % d1, k = idx / d1;
            double s = 0.0;
            for (int m = 0; m < d1; m++) s += T[r + m * d1] * tau[m + k * d1];
            X[idx] = 2.0 * s + 2.0 * R[idx];
        }
        __syncwarp();
        ens_zsolve(Uz, X, d1, d2, lane);
        __syncwarp();
        for (int idx = lane; idx < n12; idx += 32) pr[1 + idx] = X[idx];
        if (lane == 0) pr[0] = p0;
        __syncwarp();
    }
}

// dder3, epinormspectral.jl:248-294; one warp per cone, 4 warps per CTA
static __global__ void __launch_bounds__(128)
ens_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ d1s, const int64_t* __restrict__ voff, const double* __restrict__ vecs,
                 const double* __restrict__ scal, const double* __restrict__ point,
                 const double* __restrict__ dir, double* __restrict__ out) {
    __shared__ double sh[4][9 * 128];
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    double* M0 = sh[threadIdx.x >> 5];
    double *M1 = M0 + 128, *M2 = M0 + 256, *M3 = M0 + 384, *M4 = M0 + 512, *M5 = M0 + 640;
    double *S0 = M0 + 768, *S1 = M0 + 896, *S2 = M0 + 1024;
    const int64_t o = off[c];
    const int d = dim[c], d1 = d1s[c], d2 = (d - 1) / d1, n12 = d1 * d2;
    const double* tau = vecs + voff[c];
    const double* Zitau = tau + n12;
    const double* Zi = Zitau + n12;
    const double* Uz = Zi + d1 * d1;
    const double u = point[o], ud = dir[o], trZi2 = scal[8 * c + 1];
    const double* W = point + o + 1;
    const double* A = dir + o + 1;
    for (int i = lane; i < n12; i += 32) M0[i] = A[i];
    __syncwarp();
    ens_zsolve(Uz, M0, d1, d2, lane);                  // D = Z^-1 A
    __syncwarp();
    ens_mm_nt(S0, M0, W, d1, d2, lane, false);         // F = D W'
    ens_mm_nt(S2, M0, tau, d1, d2, lane, false);       // D tau'
    ens_mm_nt(S1, tau, A, d1, d2, lane, false);        // tau A'
    __syncwarp();
    for (int i = lane; i < n12; i += 32) M1[i] = M0[i];
    __syncwarp();
    ens_mm_sy(M1, S0, tau, d1, d2, lane, 1.0);         // E = D (W'tau + I)
    ens_mm_sy(M2, S2, A, d1, d2, lane, 0.0);           // C = D (A'tau)'
    ens_mm_sy(M3, S1, tau, d1, d2, lane, 0.0);         // P = tau (A'tau)
    __syncwarp();
    ens_mm_nt(S2, M3, A, d1, d2, lane, false);         // P A' + C W' + E A'
    __syncwarp();
    ens_mm_nt(S2, M2, W, d1, d2, lane, true);
    __syncwarp();
    ens_mm_nt(S2, M1, A, d1, d2, lane, true);
    __syncwarp();
    for (int i = lane; i < n12; i += 32) M4[i] = M2[i];
    __syncwarp();
    ens_mm_sy(M4, S2, tau, d1, d2, lane, 1.0);         // D2 = (...) tau + (tau A') E + C
    __syncwarp();
    ens_mm_sy(M4, S1, M1, d1, d2, lane, 1.0);
    for (int i = lane; i < n12; i += 32) M3[i] = M1[i];
    __syncwarp();
    ens_zsolve(Uz, M3, d1, d2, lane);                  // Z^-1 E
    ens_mm_nt(S2, Zitau, A, d1, d2, lane, false);      // Zitau A'
    __syncwarp();
    ens_mm_sy(M3, S2, tau, d1, d2, lane, 1.0);         // E2 = Z^-1 E + Zitau (A'tau)
    __syncwarp();
    for (int idx = lane; idx < d1 * d1; idx += 32) S2[idx] = S0[idx] + S1[idx];    // F2 = F + tau A'
    __syncwarp();
    ens_mm_sy(M3, S2, Zitau, d1, d2, lane, 1.0);
    const double const1 = 4.0 * u * ud * u;
    for (int i = lane; i < n12; i += 32) M5[i] = const1 * Zitau[i] - ud * tau[i];
    __syncwarp();
    ens_zsolve(Uz, M5, d1, d2, lane);                  // C2
    __syncwarp();
    double dot = 0.0;
    for (int i = lane; i < n12; i += 32) {
        const double e4 = -2.0 * u * M3[i] + M5[i];     // E4 = E3 + C2
        out[o + 1 + i] = -2.0 * ud * e4 - 2.0 * M4[i];
        dot += A[i] * (e4 + 3.0 * M5[i]);
    }
    dot = warp_sum(dot);
    // trZi3 = |U^-T Zi|_F^2
    double t3 = 0.0;
    for (int b = lane; b < d1; b += 32) {
        double* y = S0 + b * d1;
        for (int i = 0; i < d1; i++) {
            double s = Zi[i + b * d1];
            for (int k = 0; k < i; k++) s -= Uz[k + i * d1] * y[k];
            y[i] = s / Uz[i + i * d1];
            t3 += y[i] * y[i];
        }
    }
    t3 = warp_sum(t3);
    if (lane == 0)
        out[o] = -dot - u * ud * (6.0 * trZi2 - 8.0 * u * t3 * u) * ud - (d1 - 1) * (ud / u) * (ud / u) / u;
}


// ---- WSOSInterpNonnegative (real, U <= 128), wsosinterpnonnegative.jl:91-200 ----
// Region of cone c at vecs + voff[c]: [nP][L_1 .. L_nP][P_1 .. P_nP] (the data of hyp_set_cone_alpha: P_k is U x L_k,
// column-major), then per k the workspace F_k = L_k^-1 P_k' (L_k x U, "LFLP" of the reference) and the lower Cholesky
// factor of Lambda_k = P_k' Diagonal(point) P_k (L_k x L_k), then one L_max x L_max scratch for dder3.
// One CTA of 256 threads per cone; thread j < U owns entry j of the gradient / dder3.

static __global__ void __launch_bounds__(256)
wsos_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int64_t* __restrict__ voff, double* __restrict__ vecs, const int* __restrict__ kidx,
                  const int64_t* __restrict__ moff, const double* __restrict__ point, double* __restrict__ grad,
                  double* __restrict__ H, uint8_t* feas) {
    __shared__ int s_ok;
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int U = dim[c], lde = (U + 1) & ~1;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int sumL = 0;
    for (int k = 0; k < nP; k++) sumL += (int)reg[1 + k];
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* Hc = H + moff[c];
    if (tid == 0) s_ok = 1;
    double gacc = 0.0;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k];
        double* F = ws;
        double* Lc = F + (int64_t)L * U;
        ws = Lc + L * L;
        // update_feas (:91-121): Lambda = P' Diagonal(point) P and its Cholesky (lower, in place, right-looking)
        for (int idx = tid; idx < L * L; idx += 256) {
            const int a = idx % L, b = idx / L;
            double s = 0.0;
            for (int i = 0; i < U; i++) s += P[i + (int64_t)a * U] * point[o + i] * P[i + (int64_t)b * U];
            Lc[idx] = s;
        }
        __syncthreads();
        for (int j = 0; j < L; j++) {
            if (tid == 0) {
                double dg = Lc[j + j * L];
                if (!(dg > 0.0)) {
                    s_ok = 0;
                    dg = 1.0;
                }
                Lc[j + j * L] = sqrt(dg);
            }
            __syncthreads();
            const double dj = Lc[j + j * L];
            for (int i = j + 1 + tid; i < L; i += 256) Lc[i + j * L] /= dj;
            __syncthreads();
            const int r = L - j - 1;
            for (int idx = tid; idx < r * r; idx += 256) {
                const int ii = j + 1 + idx % r, kk = j + 1 + idx / r;
                if (kk <= ii) Lc[ii + kk * L] -= Lc[ii + j * L] * Lc[kk + j * L];
            }
            __syncthreads();
        }
        // update_grad (:123-138): F = L^-1 P', grad_j -= |F[:, j]|^2
        for (int jj = tid; jj < U; jj += 256) {
            double* f = F + (int64_t)jj * L;
            double acc = 0.0;
            for (int a = 0; a < L; a++) {
                double s = P[jj + (int64_t)a * U];
                for (int b = 0; b < a; b++) s -= Lc[a + b * L] * f[b];
                f[a] = s / Lc[a + a * L];
                acc += f[a] * f[a];
            }
            gacc -= acc;
        }
        __syncthreads();
        // update_hess (:140-156): H += (F'F).^2, both triangles
        for (int idx = tid; idx < U * U; idx += 256) {
            const int i = idx % U, j = idx / U;
            double s = 0.0;
            for (int a = 0; a < L; a++) s += F[a + (int64_t)i * L] * F[a + (int64_t)j * L];
            const double v = s * s;
            Hc[i + (int64_t)j * lde] = k == 0 ? v : Hc[i + (int64_t)j * lde] + v;
        }
        __syncthreads();
        P += (int64_t)U * L;
    }
    if (tid < U) grad[o + tid] = gacc;
    if (tid == 0 && !s_ok) feas[kidx[c]] = 0;
}

// dder3 (:180-200): per k, S = F Diagonal(dir) F', out_j += |S F[:, j]|^2
static __global__ void __launch_bounds__(256)
wsos_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int64_t* __restrict__ voff, double* __restrict__ vecs, const double* __restrict__ dir,
                  double* __restrict__ out) {
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int U = dim[c];
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k];
        sumL += L;
        wsz += (int64_t)L * U + (int64_t)L * L;
    }
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* S = ws + wsz;
    double acc = 0.0;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k];
        const double* F = ws;
        ws += (int64_t)L * U + (int64_t)L * L;
        for (int idx = tid; idx < L * L; idx += 256) {
            const int a = idx % L, b = idx / L;
            double s = 0.0;
            for (int j = 0; j < U; j++) s += F[a + (int64_t)j * L] * dir[o + j] * F[b + (int64_t)j * L];
            S[idx] = s;
        }
        __syncthreads();
        for (int jj = tid; jj < U; jj += 256) {
            const double* f = F + (int64_t)jj * L;
            for (int a = 0; a < L; a++) {
                double t = 0.0;
                for (int b = 0; b < L; b++) t += S[a + b * L] * f[b];
                acc += t * t;
            }
        }
        __syncthreads();
    }
    if (tid < U) out[o + tid] = acc;
}

// generic hess_prod! of the Cone API (Cones.jl:101-105): prod = H arr with the explicit Hessian; one warp per
// (cone, column), dim <= 128, in-place safe
static __global__ void __launch_bounds__(256)
gen_hess_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                     const int64_t* __restrict__ moff, const int* __restrict__ dualf, const double* __restrict__ H,
                     const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols,
                     int64_t row_shift) {
    __shared__ double xs[8][128];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    double* x = xs[threadIdx.x >> 5];
    const int64_t o = off[c];
    const int d = dim[c], lde = (d + 1) & ~1;
    const double* Hc = H + moff[c];
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        for (int i = lane; i < d; i += 32) x[i] = a[i];
        __syncwarp();
        for (int i = lane; i < d; i += 32) {
            double s = 0.0;
            for (int jj = 0; jj < d; jj++) s += Hc[i + (int64_t)jj * lde] * x[jj];
            pr[i] = s;
        }
        __syncwarp();
    }
}


// ---- LinMatrixIneq (real dense A_i, dim <= 128), linmatrixineq.jl:87-159 ----
// Region of cone c at vecs + voff[c]: [side][A_1 .. A_dim] (the data of hyp_set_cone_alpha, side x side column-major
// each), then the workspace: the lower Cholesky factor of S = sum_i w_i A_i (side^2), B_i = L^-1 A_i L^-T for every i
// (dim side^2, "sumAinvAs" of the reference) and two side^2 scratch matrices for dder3.
// One CTA of 256 threads per cone.

static __global__ void __launch_bounds__(256)
lmi_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ voff, double* __restrict__ vecs, const int* __restrict__ kidx,
                 const int64_t* __restrict__ moff, const double* __restrict__ point, double* __restrict__ grad,
                 double* __restrict__ H, uint8_t* feas) {
    __shared__ int s_ok;
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], lde = (d + 1) & ~1;
    double* reg = vecs + voff[c];
    const int sd = (int)reg[0], s2 = sd * sd;
    const double* A = reg + 1;
    double* Lc = reg + 1 + (int64_t)d * s2;
    double* B = Lc + s2;
    double* Hc = H + moff[c];
    if (tid == 0) s_ok = 1;
    // update_feas (:87-96): S = sum_i w_i A_i and its Cholesky (lower, in place, right-looking)
    for (int idx = tid; idx < s2; idx += 256) {
        double s = 0.0;
        for (int i = 0; i < d; i++) s += point[o + i] * A[(int64_t)i * s2 + idx];
        Lc[idx] = s;
    }
    __syncthreads();
    for (int j = 0; j < sd; j++) {
        if (tid == 0) {
            double dg = Lc[j + j * sd];
            if (!(dg > 0.0)) {
                s_ok = 0;
                dg = 1.0;
            }
            Lc[j + j * sd] = sqrt(dg);
        }
        __syncthreads();
        const double dj = Lc[j + j * sd];
        for (int i = j + 1 + tid; i < sd; i += 256) Lc[i + j * sd] /= dj;
        __syncthreads();
        const int r = sd - j - 1;
        for (int idx = tid; idx < r * r; idx += 256) {
            const int ii = j + 1 + idx % r, kk = j + 1 + idx / r;
            if (kk <= ii) Lc[ii + kk * sd] -= Lc[ii + j * sd] * Lc[kk + j * sd];
        }
        __syncthreads();
    }
    // update_grad (:98-109): B_i = L^-1 A_i L^-T.  Pass 1: every column of A_i through L^-1 (in place in B_i);
    // pass 2: every row of the result through L^-1 (row a of X L^-T is L^-1 applied to row a of X; rows are disjoint).
    for (int idx = tid; idx < d * sd; idx += 256) {
        const int i = idx / sd, col = idx % sd;
        const double* a = A + (int64_t)i * s2 + (int64_t)col * sd;
        double* x = B + (int64_t)i * s2 + (int64_t)col * sd;
        for (int r = 0; r < sd; r++) {
            double s = a[r];
            for (int b = 0; b < r; b++) s -= Lc[r + b * sd] * x[b];
            x[r] = s / Lc[r + r * sd];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < d * sd; idx += 256) {
        const int i = idx / sd, row = idx % sd;
        double* x = B + (int64_t)i * s2 + row;          // stride sd along the row
        for (int r = 0; r < sd; r++) {
            double s = x[(int64_t)r * sd];
            for (int b = 0; b < r; b++) s -= Lc[r + b * sd] * x[(int64_t)b * sd];
            x[(int64_t)r * sd] = s / Lc[r + r * sd];
        }
    }
    __syncthreads();
    for (int i = tid; i < d; i += 256) {
        double tr = 0.0;
        for (int r = 0; r < sd; r++) tr += B[(int64_t)i * s2 + r + r * sd];
        grad[o + i] = -tr;
    }
    // update_hess (:111-123): H_ij = <B_i, B_j>, both triangles
    for (int idx = tid; idx < d * d; idx += 256) {
        const int i = idx % d, j = idx / d;
        const double* bi = B + (int64_t)i * s2;
        const double* bj = B + (int64_t)j * s2;
        double s = 0.0;
        for (int e = 0; e < s2; e++) s += bi[e] * bj[e];
        Hc[i + (int64_t)j * lde] = s;
    }
    if (tid == 0 && !s_ok) feas[kidx[c]] = 0;
}

// dder3 (:147-159): D = sum_i dir_i B_i, Z = D D', out_i = <Z, B_i>
static __global__ void __launch_bounds__(256)
lmi_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ voff, double* __restrict__ vecs, const double* __restrict__ dir,
                 double* __restrict__ out) {
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    double* reg = vecs + voff[c];
    const int sd = (int)reg[0], s2 = sd * sd;
    const double* B = reg + 1 + (int64_t)d * s2 + s2;
    double* D = reg + 1 + (int64_t)d * s2 + s2 + (int64_t)d * s2;
    double* Z = D + s2;
    for (int idx = tid; idx < s2; idx += 256) {
        double s = 0.0;
        for (int i = 0; i < d; i++) s += dir[o + i] * B[(int64_t)i * s2 + idx];
        D[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < s2; idx += 256) {
        const int a = idx % sd, b = idx / sd;
        double s = 0.0;
        for (int k = 0; k < sd; k++) s += D[a + k * sd] * D[b + k * sd];
        Z[idx] = s;
    }
    __syncthreads();
    for (int i = tid; i < d; i += 256) {
        const double* bi = B + (int64_t)i * s2;
        double s = 0.0;
        for (int e = 0; e < s2; e++) s += Z[e] * bi[e];
        out[o + i] = s;
    }
}


// ---- DoublyNonnegativeTri (dim = svec_length(side) <= 128, so side <= 15), doublynonnegativetri.jl:128-205 ----
// Per-cone workspace at vecs + voff[c]: Zi = W^-1 (side^2, full symmetric) and the upper Cholesky factor of W (side^2).
// One warp per cone / per (cone, column); the congruences Zi M Zi run on side x side matrices in shared memory.

// svec index p -> (i, j), i <= j, column-major upper triangle: p = j (j + 1) / 2 + i
__device__ __forceinline__ void dnn_ij(int p, int& i, int& j) {
    j = 0;
    while ((j + 1) * (j + 2) / 2 <= p) j++;
    i = p - j * (j + 1) / 2;
}
// M (side x side, full) <- smat(v): off-diagonal entries of svec are scaled by 1/sqrt(2)
__device__ __forceinline__ void dnn_smat(double* Mx, const double* v, int side, int d, int lane) {
    for (int p = lane; p < d; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const double x = i == j ? v[p] : v[p] * 0.70710678118654752440;
        Mx[i + j * side] = x;
        Mx[j + i * side] = x;
    }
}
// O = X Y for side x side matrices
__device__ __forceinline__ void dnn_mm(double* O, const double* X, const double* Y, int side, int lane, bool y_transposed) {
    for (int idx = lane; idx < side * side; idx += 32) {
        const int a = idx % side, b = idx / side;
        double s = 0.0;
        for (int k = 0; k < side; k++) s += X[a + k * side] * (y_transposed ? Y[b + k * side] : Y[k + b * side]);
        O[idx] = s;
    }
}

static __global__ void __launch_bounds__(256)
dnn_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ sides, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                 const int* __restrict__ kidx, const int64_t* __restrict__ moff, const double* __restrict__ point,
                 double* __restrict__ grad, double* __restrict__ H, uint8_t* feas) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], side = sides[c], lde = (d + 1) & ~1;
    double* Zi = vecs + voff[c];
    double* Uz = Zi + side * side;
    const double* pt = point + o;
    // update_feas (:128-142): every svec entry positive, then Cholesky of smat(point)
    double nbad = 0.0;
    for (int p = lane; p < d; p += 32)
        if (!(pt[p] > HYP_EPS)) nbad += 1.0;
    nbad = warp_sum(nbad);
    dnn_smat(Uz, pt, side, d, lane);
    __syncwarp();
    int ok = nbad == 0.0 ? 1 : 0;
    if (lane == 0) {
        for (int j = 0; j < side; j++) {
            double s = Uz[j + j * side];
            for (int k = 0; k < j; k++) s -= Uz[k + j * side] * Uz[k + j * side];
            if (!(s > 0.0)) {
                ok = 0;
                s = 1.0;
            }
            const double r = sqrt(s);
            Uz[j + j * side] = r;
            for (int i = j + 1; i < side; i++) {
                double t = Uz[j + i * side];
                for (int k = 0; k < j; k++) t -= Uz[k + j * side] * Uz[k + i * side];
                Uz[j + i * side] = t / r;
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    // update_grad (:144-155): Zi = W^-1 (column by column), grad = -svec(Zi) - 1 / offdiag
    for (int b = lane; b < side; b += 32) {
        double* x = Zi + b * side;
        for (int i = 0; i < side; i++) x[i] = i == b ? 1.0 : 0.0;
        ens_zsolve_col(Uz, side, x);
    }
    __syncwarp();
    for (int idx = lane; idx < side * side; idx += 32) {     // copytri!(inv_mat, 'U')
        const int a = idx % side, b = idx / side;
        if (a > b) Zi[idx] = Zi[b + a * side];
    }
    __syncwarp();
    for (int p = lane; p < d; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        grad[o + p] = i == j ? -Zi[i + j * side] : -1.41421356237309504880 * Zi[i + j * side] - 1.0 / pt[p];
    }
    if (lane == 0 && !ok) feas[kidx[c]] = 0;
    // update_hess (:157-171): symm_kron(Zi) + Diagonal(1 / offdiag^2), both triangles
    double* Hc = H + moff[c];
    for (int idx = lane; idx < d * d; idx += 32) {
        const int p = idx % d, q = idx / d;
        int i, j, k, l;
        dnn_ij(p, i, j);
        dnn_ij(q, k, l);
        double v;
        if (i == j && k == l) v = Zi[i + k * side] * Zi[i + k * side];
        else if (i == j) v = 1.41421356237309504880 * Zi[i + k * side] * Zi[i + l * side];
        else if (k == l) v = 1.41421356237309504880 * Zi[i + k * side] * Zi[j + k * side];
        else v = Zi[i + k * side] * Zi[j + l * side] + Zi[i + l * side] * Zi[j + k * side];
        if (p == q && i != j) v += 1.0 / (pt[p] * pt[p]);
        Hc[p + (int64_t)q * lde] = v;
    }
}

// hess_prod!, doublynonnegativetri.jl:173-193
static __global__ void __launch_bounds__(256)
dnn_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int* __restrict__ sides, const int64_t* __restrict__ voff, const double* __restrict__ vecs,
                const int* __restrict__ dualf, const double* __restrict__ point, const double* arr, int64_t ld_arr,
                double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    __shared__ double sh[8][2 * 232];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    double* Mx = sh[threadIdx.x >> 5];
    double* T = Mx + 232;
    const int64_t o = off[c];
    const int d = dim[c], side = sides[c];
    const double* Zi = vecs + voff[c];
    const double* pt = point + o;
    for (int64_t jc = blockIdx.y; jc < ncols; jc += gridDim.y) {
        const double* a = arr + jc * ld_arr + (o - row_shift);
        double* pr = prod + jc * ld_prod + (o - row_shift);
        dnn_smat(Mx, a, side, d, lane);
        __syncwarp();
        dnn_mm(T, Zi, Mx, side, lane, false);          // T = Zi M
        __syncwarp();
        dnn_mm(Mx, T, Zi, side, lane, false);          // M <- Zi M Zi
        __syncwarp();
        for (int p = lane; p < d; p += 32) {
            int i, j;
            dnn_ij(p, i, j);
            const double ap = a[p];
            pr[p] = i == j ? Mx[i + j * side] : 1.41421356237309504880 * Mx[i + j * side] + ap / (pt[p] * pt[p]);
        }
        __syncwarp();
    }
}

// dder3, doublynonnegativetri.jl:195-205: svec(Zi D Zi D Zi) + (dir / s)^2 / s on the off-diagonal entries
static __global__ void __launch_bounds__(128)
dnn_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ sides, const int64_t* __restrict__ voff, const double* __restrict__ vecs,
                 const double* __restrict__ point, const double* __restrict__ dir, double* __restrict__ out) {
    __shared__ double sh[4][3 * 232];
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    double* D = sh[threadIdx.x >> 5];
    double* T1 = D + 232;
    double* T2 = T1 + 232;
    const int64_t o = off[c];
    const int d = dim[c], side = sides[c];
    const double* Zi = vecs + voff[c];
    const double* pt = point + o;
    dnn_smat(D, dir + o, side, d, lane);
    __syncwarp();
    dnn_mm(T1, Zi, D, side, lane, false);              // T1 = Zi D
    __syncwarp();
    dnn_mm(T2, T1, Zi, side, lane, false);             // T2 = Zi D Zi
    __syncwarp();
    dnn_mm(D, T2, T1, side, lane, true);               // D <- T2 T1' = Zi D Zi D Zi
    __syncwarp();
    for (int p = lane; p < d; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const double r = dir[o + p] / pt[p];
        out[o + p] = i == j ? D[i + j * side] : 1.41421356237309504880 * D[i + j * side] + r * r / pt[p];
    }
}


// ---- MatrixEpiPerSquare (real), matrixepipersquare.jl:118-397 ----
// point = (svec(U) [per = d1 (d1 + 1) / 2], v, vec(W) [d1 x d2 column-major]), d1 <= d2, dim <= 128 (so d1 <= 8).
// Per-cone state at vecs + voff[c]: Zi = Z^-1 (d1^2), Uz = upper Cholesky factor of Z = 2 v U - W W' (d1^2),
// Um = smat(U) (d1^2), ZiUZi (d1^2), ZiW (d1 d2), ZiUZiW (d1 d2), then the dder3 scratch (17 d1^2 + 11 d1 d2 + 3 d2^2).
// scal: 0 Hvv, 1 v.  One warp per cone / per (cone, column); every product below is a tiny dense GEMM done lane-strided.

// O (r x c) = beta O + alpha op(X) op(Y), op(X) r x k, op(Y) k x c, all column-major and dense; O must not alias X, Y
__device__ __forceinline__ void mep_mm(double* O, const double* X, const double* Y, int r, int k, int c, bool tx,
                                       bool ty, double alpha, double beta, int lane) {
    for (int idx = lane; idx < r * c; idx += 32) {
        const int i = idx % r, j = idx / r;
        double s = 0.0;
        for (int t = 0; t < k; t++) s += (tx ? X[t + i * k] : X[i + t * r]) * (ty ? Y[j + t * c] : Y[t + j * k]);
        O[idx] = (beta == 0.0 ? 0.0 : beta * O[idx]) + alpha * s;
    }
    __syncwarp();
}
// O = a X + b Y (elementwise, n entries; O may alias X or Y)
__device__ __forceinline__ void mep_axpby(double* O, double a, const double* X, double b, const double* Y, int n,
                                          int lane) {
    for (int i = lane; i < n; i += 32) O[i] = a * X[i] + (Y ? b * Y[i] : 0.0);
    __syncwarp();
}
// O = X + X' (d x d); O must not alias X
__device__ __forceinline__ void mep_symsum(double* O, const double* X, int d, int lane) {
    for (int idx = lane; idx < d * d; idx += 32) O[idx] = X[idx] + X[(idx / d) + (idx % d) * d];
    __syncwarp();
}
__device__ __forceinline__ double mep_dot(const double* X, const double* Y, int n, int lane) {
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += X[i] * Y[i];
    return warp_sum(s);
}
// X <- Z^-1 X for the d1 x cols matrix X (lane per column)
__device__ __forceinline__ void mep_zsolve(const double* Uz, double* X, int d1, int cols, int lane) {
    for (int k = lane; k < cols; k += 32) ens_zsolve_col(Uz, d1, X + k * d1);
    __syncwarp();
}

// hess_prod! for one column (matrixepipersquare.jl:281-325).  a: input column, pr: output column (may alias a),
// sc: 5 * 128 doubles of scratch private to the warp.
__device__ __forceinline__ void mep_hess_col(const double* a, double* pr, int d1, int d2, const double* st,
                                             const double* W, double v, double Hvv, double* sc, int lane) {
    const int per = d1 * (d1 + 1) / 2, n11 = d1 * d1, n12 = d1 * d2;
    const double* Zi = st;
    const double* Uz = st + n11;
    const double* Um = st + 2 * n11;
    const double* ZiUZi = st + 3 * n11;
    const double* ZiW = st + 4 * n11;
    double* tU = sc;
    double* U3 = sc + 128;
    double* T = sc + 256;
    double* U2 = sc + 384;
    double* ZiWd = sc + 512;
    const double va = a[per], v2 = 2.0 * v;
    dnn_smat(tU, a, d1, per, lane);
    for (int i = lane; i < n12; i += 32) ZiWd[i] = a[per + 1 + i];
    __syncwarp();
    mep_zsolve(Uz, ZiWd, d1, d2, lane);
    mep_mm(U3, ZiWd, ZiW, d1, d2, d1, false, true, 1.0, 0.0, lane);          // U3 = ZiWd ZiW'
    mep_mm(T, Zi, tU, d1, d1, d1, false, false, 1.0, 0.0, lane);
    mep_mm(U2, T, Zi, d1, d1, d1, false, false, 1.0, 0.0, lane);             // Zi tU Zi
    mep_symsum(T, U3, d1, lane);
    mep_axpby(U2, -v2, U2, 1.0, T, n11, lane);                               // U2 = U3 + U3' - 2 v Zi tU Zi
    const double dT2 = 2.0 * v2 * mep_dot(ZiUZi, tU, n11, lane) - 2.0 * mep_dot(Zi, tU, n11, lane);
    const double dUU3 = mep_dot(Um, U3, n11, lane);
    mep_axpby(T, 1.0, U2, -2.0 * va, ZiUZi, n11, lane);                      // T1 = U2 - 2 va ZiUZi
    mep_mm(U3, T, W, d1, d1, d2, false, false, 2.0, 0.0, lane);              // (U3 reused, d1 x d2) 2 T1 W
    for (int i = lane; i < n12; i += 32) pr[per + 1 + i] = U3[i] + 2.0 * ZiWd[i];
    // prod_U = svec(va (2 v2 ZiUZi - 2 Zi) - v2 U2)
    for (int p = lane; p < per; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const int e = i + j * d1;
        const double x = va * (2.0 * v2 * ZiUZi[e] - 2.0 * Zi[e]) - v2 * U2[e];
        pr[p] = i == j ? x : 1.41421356237309504880 * x;
    }
    if (lane == 0) pr[per] = dT2 - 4.0 * dUU3 + Hvv * va;
    __syncwarp();
}

static __global__ void __launch_bounds__(128)
mep_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ d1s, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                 const int* __restrict__ kidx, const int64_t* __restrict__ moff, const double* __restrict__ point,
                 const double* __restrict__ dual, double* __restrict__ grad, double* __restrict__ scal,
                 double* __restrict__ H, uint8_t* feas, uint8_t* dual_feas) {
    __shared__ double sh[4][6 * 128];
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    double* sc = sh[threadIdx.x >> 5];
    const int64_t o = off[c];
    const int d = dim[c], d1 = d1s[c], per = d1 * (d1 + 1) / 2, d2 = (d - per - 1) / d1;
    const int n11 = d1 * d1, n12 = d1 * d2, lde = (d + 1) & ~1;
    double* st = vecs + voff[c];
    double* Zi = st;
    double* Uz = st + n11;
    double* Um = st + 2 * n11;
    double* ZiUZi = st + 3 * n11;
    double* ZiW = st + 4 * n11;
    double* ZiUZiW = ZiW + n12;
    const double v = point[o + per];
    const double* W = point + o + per + 1;
    // update_feas (:118-135): Z = 2 v U - W W' and its Cholesky
    dnn_smat(Um, point + o, d1, per, lane);
    __syncwarp();
    mep_mm(Uz, W, W, d1, d2, d1, false, true, -1.0, 0.0, lane);
    for (int i = lane; i < n11; i += 32) Uz[i] += 2.0 * v * Um[i];
    __syncwarp();
    int ok = v > HYP_EPS ? 1 : 0;
    if (lane == 0) {
        for (int j = 0; j < d1; j++) {
            double s = Uz[j + j * d1];
            for (int k = 0; k < j; k++) s -= Uz[k + j * d1] * Uz[k + j * d1];
            if (!(s > 0.0)) {
                ok = 0;
                s = 1.0;
            }
            const double r = sqrt(s);
            Uz[j + j * d1] = r;
            for (int i = j + 1; i < d1; i++) {
                double t = Uz[j + i * d1];
                for (int k = 0; k < j; k++) t -= Uz[k + j * d1] * Uz[k + i * d1];
                Uz[j + i * d1] = t / r;
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __syncwarp();
    // update_grad (:152-170) and update_hess_aux (:172-185)
    for (int i = lane; i < n11; i += 32) Zi[i] = (i % d1 == i / d1) ? 1.0 : 0.0;
    for (int i = lane; i < n12; i += 32) ZiW[i] = W[i];
    __syncwarp();
    mep_zsolve(Uz, Zi, d1, d1, lane);
    mep_zsolve(Uz, ZiW, d1, d2, lane);
    for (int idx = lane; idx < n11; idx += 32) {          // symmetric part, as inv_fact! + Hermitian(:U) reads it
        const int a = idx % d1, b = idx / d1;
        if (a > b) Zi[idx] = Zi[b + a * d1];
    }
    __syncwarp();
    mep_mm(sc, Zi, Um, d1, d1, d1, false, false, 1.0, 0.0, lane);
    mep_mm(ZiUZi, sc, Zi, d1, d1, d1, false, false, 1.0, 0.0, lane);
    mep_symsum(sc, ZiUZi, d1, lane);
    mep_axpby(ZiUZi, 0.5, sc, 0.0, nullptr, n11, lane);
    mep_mm(ZiUZiW, ZiUZi, W, d1, d1, d2, false, false, 1.0, 0.0, lane);
    const double trZiU = mep_dot(Zi, Um, n11, lane);
    const double Hvv = 4.0 * mep_dot(ZiUZi, Um, n11, lane) - (d1 - 1) / v / v;
    for (int p = lane; p < per; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const double x = -2.0 * v * Zi[i + j * d1];
        grad[o + p] = i == j ? x : 1.41421356237309504880 * x;
    }
    for (int i = lane; i < n12; i += 32) grad[o + per + 1 + i] = 2.0 * ZiW[i];
    // is_dual_feas (:137-150): dual U positive definite and 2 v - |R^-T W|_F^2 > eps with U = R'R
    const double dv = dual[o + per];
    double* R = sc;                 // d1^2
    double* LW = sc + 128;          // d1 d2
    dnn_smat(R, dual + o, d1, per, lane);
    for (int i = lane; i < n12; i += 32) LW[i] = dual[o + per + 1 + i];
    __syncwarp();
    int dok = dv > HYP_EPS ? 1 : 0;
    if (lane == 0) {
        for (int j = 0; j < d1; j++) {
            double s = R[j + j * d1];
            for (int k = 0; k < j; k++) s -= R[k + j * d1] * R[k + j * d1];
            if (!(s > 0.0)) {
                dok = 0;
                s = 1.0;
            }
            const double r = sqrt(s);
            R[j + j * d1] = r;
            for (int i = j + 1; i < d1; i++) {
                double t = R[j + i * d1];
                for (int k = 0; k < j; k++) t -= R[k + j * d1] * R[k + i * d1];
                R[j + i * d1] = t / r;
            }
        }
    }
    dok = __shfl_sync(0xffffffffu, dok, 0);
    __syncwarp();
    double tr = 0.0;
    for (int k = lane; k < d2; k += 32) {
        double* x = LW + k * d1;
        for (int i = 0; i < d1; i++) {                      // forward substitution with R'
            double s = x[i];
            for (int b = 0; b < i; b++) s -= R[b + i * d1] * x[b];
            x[i] = s / R[i + i * d1];
            tr += x[i] * x[i];
        }
    }
    tr = warp_sum(tr);
    if (!(2.0 * dv - tr > HYP_EPS)) dok = 0;
    if (lane == 0) {
        grad[o + per] = -2.0 * trZiU + (d1 - 1) / v;
        scal[8 * c] = Hvv;
        scal[8 * c + 1] = v;
        if (!ok) feas[kidx[c]] = 0;
        if (!dok) dual_feas[kidx[c]] = 0;
    }
    __syncwarp();
    // explicit Hessian: hess_prod! applied to the unit vectors (equal to update_hess, :187-279), then symmetrised
    double* Hc = H + moff[c];
    double* unit = sc + 5 * 128;
    for (int j = 0; j < d; j++) {
        for (int i = lane; i < d; i += 32) unit[i] = i == j ? 1.0 : 0.0;
        __syncwarp();
        mep_hess_col(unit, Hc + (int64_t)j * lde, d1, d2, st, W, v, Hvv, sc, lane);
    }
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx % d, j = idx / d;
        if (i < j) {
            const double x = 0.5 * (Hc[i + (int64_t)j * lde] + Hc[j + (int64_t)i * lde]);
            Hc[i + (int64_t)j * lde] = x;
            Hc[j + (int64_t)i * lde] = x;
        }
    }
}

static __global__ void __launch_bounds__(256)
mep_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int* __restrict__ d1s, const int64_t* __restrict__ voff, const double* __restrict__ vecs,
                const int* __restrict__ dualf, const double* __restrict__ scal, const double* __restrict__ point,
                const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    __shared__ double sh[8][5 * 128];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    double* sc = sh[threadIdx.x >> 5];
    const int64_t o = off[c];
    const int d = dim[c], d1 = d1s[c], per = d1 * (d1 + 1) / 2, d2 = (d - per - 1) / d1;
    const double* st = vecs + voff[c];
    const double* W = point + o + per + 1;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y)
        mep_hess_col(arr + j * ld_arr + (o - row_shift), prod + j * ld_prod + (o - row_shift), d1, d2, st, W,
                     scal[8 * c + 1], scal[8 * c], sc, lane);
}

// dder3 (matrixepipersquare.jl:327-397), transcribed product by product; one warp per cone, scratch in global memory
static __global__ void __launch_bounds__(128)
mep_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int* __restrict__ d1s, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                 const double* __restrict__ scal, const double* __restrict__ point, const double* __restrict__ dir,
                 double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], d1 = d1s[c], per = d1 * (d1 + 1) / 2, d2 = (d - per - 1) / d1;
    const int n11 = d1 * d1, n12 = d1 * d2, n22 = d2 * d2;
    double* st = vecs + voff[c];
    const double* Zi = st;
    const double* Uz = st + n11;
    const double* U = st + 2 * n11;
    const double* ZiUZi = st + 3 * n11;
    const double* ZiW = st + 4 * n11;
    const double* ZiUZiW = ZiW + n12;
    double* ws = st + 4 * n11 + 2 * n12;
    double* S[17];
    double* Mx[11];
    for (int i = 0; i < 17; i++) S[i] = ws + i * n11;
    for (int i = 0; i < 11; i++) Mx[i] = ws + 17 * n11 + i * n12;
    double* B0 = ws + 17 * n11 + 11 * n12;
    double* B1 = B0 + n22;
    double* B2 = B1 + n22;
    const double v = point[o + per], vd = dir[o + per], v2 = 2.0 * v, vd2 = 2.0 * vd;
    const double* W = point + o + per + 1;
    const double* Wd = dir + o + per + 1;
    double *Ud = S[0], *ZiU = S[1], *ZiUd = S[2], *Z3 = S[3] /* ZiUZiUZi */, *Z2v = S[4] /* ZiUZi2v */,
           *WdWZi = S[5], *ZiWdWZi = S[6], *ZiWdWZi2 = S[7], *ZiUZiWdWZi = S[8], *ZiWdWZiUZi2 = S[9], *ZiUdZi = S[10],
           *ZiUZiUdZi2 = S[11], *ZiUdZiWdWZi = S[12], *vZ = S[13] /* vZiUZiUdZi2 */, *Ut = S[14], *T0 = S[15],
           *T1 = S[16];
    double *ZiWd = Mx[0], *UdZiW = Mx[1], *ZiUdZiW = Mx[2], *ZiUZiUdZiW = Mx[3], *ZiWdWtZiWI = Mx[4],
           *vdZW = Mx[5] /* vdZiUZiUZiW */, *WdWtZiWI = Mx[6], *ZWWI = Mx[7] /* ZiUZiWdWZiWI */, *Wt = Mx[8], *N0 = Mx[9],
           *N1 = Mx[10];
    double *WdZiW = B0, *WtZiWI = B1, *Bt = B2;

    dnn_smat(Ud, dir + o, d1, per, lane);
    for (int i = lane; i < n11; i += 32) {
        ZiU[i] = U[i];
        ZiUd[i] = 0.0;
    }
    for (int i = lane; i < n12; i += 32) ZiWd[i] = Wd[i];
    __syncwarp();
    for (int i = lane; i < n11; i += 32) ZiUd[i] = Ud[i];
    __syncwarp();
    mep_zsolve(Uz, ZiU, d1, d1, lane);                                             // ZiU = Z \ U
    mep_zsolve(Uz, ZiWd, d1, d2, lane);                                            // ZiWd = Z \ Wd
    mep_zsolve(Uz, ZiUd, d1, d1, lane);                                            // ZiUd = Z \ Ud
    mep_mm(T0, ZiUZi, ZiU, d1, d1, d1, false, true, 1.0, 0.0, lane);               // ZiUZi ZiU'
    mep_symsum(Z3, T0, d1, lane);
    mep_axpby(Z3, 0.5, Z3, 0.0, nullptr, n11, lane);                               // ZiUZiUZi (symmetric part)
    mep_axpby(Z2v, 1.0, ZiUZi, -v2, Z3, n11, lane);                                // ZiUZi2v
    mep_mm(WdWZi, Wd, ZiW, d1, d2, d1, false, true, 1.0, 0.0, lane);               // Wd ZiW'
    mep_mm(WdZiW, Wd, ZiW, d2, d1, d2, true, false, 1.0, 0.0, lane);               // Wd' ZiW
    mep_mm(UdZiW, Ud, ZiW, d1, d1, d2, false, false, 1.0, 0.0, lane);              // Ud ZiW
    for (int i = lane; i < n11; i += 32) ZiWdWZi[i] = WdWZi[i];
    for (int i = lane; i < n12; i += 32) ZiUdZiW[i] = UdZiW[i];
    __syncwarp();
    mep_zsolve(Uz, ZiWdWZi, d1, d1, lane);                                         // Z \ WdWZi
    mep_zsolve(Uz, ZiUdZiW, d1, d2, lane);                                         // Z \ UdZiW
    mep_symsum(ZiWdWZi2, ZiWdWZi, d1, lane);
    mep_mm(ZiUZiWdWZi, ZiU, ZiWdWZi, d1, d1, d1, false, false, 1.0, 0.0, lane);    // ZiU ZiWdWZi
    mep_mm(ZiUZiUdZiW, ZiU, ZiUdZiW, d1, d1, d2, false, false, 1.0, 0.0, lane);
    mep_mm(ZiUZiUdZiW, ZiUd, ZiUZiW, d1, d1, d2, false, false, 1.0, 1.0, lane);    // + ZiUd ZiUZiW
    mep_mm(T0, ZiWdWZi, ZiU, d1, d1, d1, false, true, 1.0, 0.0, lane);             // ZiWdWZiUZi = ZiWdWZi ZiU'
    mep_symsum(ZiWdWZiUZi2, T0, d1, lane);
    for (int idx = lane; idx < n11; idx += 32)                                     // + ZiUZiWdWZi'
        ZiWdWZiUZi2[idx] += ZiUZiWdWZi[(idx / d1) + (idx % d1) * d1];
    __syncwarp();
    mep_mm(T0, ZiUd, Zi, d1, d1, d1, false, false, 1.0, 0.0, lane);                // ZiUd / Z
    mep_symsum(ZiUdZi, T0, d1, lane);
    mep_axpby(ZiUdZi, 0.5, ZiUdZi, 0.0, nullptr, n11, lane);
    mep_mm(T0, ZiU, ZiUdZi, d1, d1, d1, false, false, 1.0, 0.0, lane);             // ZiUZiUdZi
    mep_symsum(ZiUZiUdZi2, T0, d1, lane);
    mep_mm(T0, ZiUd, ZiWdWZi, d1, d1, d1, false, false, 1.0, 0.0, lane);
    mep_mm(T0, ZiWdWZi, ZiUd, d1, d1, d1, false, true, 1.0, 1.0, lane);            // ZiUd ZiWdWZi + ZiWdWZi ZiUd'
    mep_axpby(ZiUdZiWdWZi, 1.0, T0, 0.0, nullptr, n11, lane);
    mep_mm(WtZiWI, W, ZiW, d2, d1, d2, true, false, 1.0, 0.0, lane);               // W' ZiW + I
    for (int i = lane; i < d2; i += 32) WtZiWI[i + i * d2] += 1.0;
    __syncwarp();
    mep_mm(ZiWdWtZiWI, ZiWd, WtZiWI, d1, d2, d2, false, false, 1.0, 0.0, lane);
    mep_mm(vdZW, Z3, W, d1, d1, d2, false, false, vd2, 0.0, lane);                 // vd2 ZiUZiUZi W
    mep_mm(WdWtZiWI, Wd, WtZiWI, d1, d2, d2, false, false, 1.0, 0.0, lane);
    mep_mm(ZWWI, ZiUZi, WdWtZiWI, d1, d1, d2, false, false, 1.0, 0.0, lane);
    mep_mm(ZWWI, ZiWdWZiUZi2, W, d1, d1, d2, false, false, 1.0, 1.0, lane);        // ZiUZiWdWZiWI
    mep_axpby(vZ, v, ZiUZiUdZi2, -1.0, ZiUdZi, n11, lane);                         // vZiUZiUdZi2

    // Utemp (:372-376)
    mep_mm(Ut, ZiWdWZi, WdWZi, d1, d1, d1, false, false, 1.0, 0.0, lane);
    mep_mm(Ut, WdWZi, ZiWdWZi2, d1, d1, d1, true, false, 1.0, 1.0, lane);
    mep_mm(Ut, ZiWdWtZiWI, ZiWd, d1, d2, d1, false, true, 1.0, 1.0, lane);
    mep_mm(T0, ZiUd, ZiUdZi, d1, d1, d1, false, false, 1.0, 0.0, lane);
    for (int idx = lane; idx < n11; idx += 32) {
        const int tr = (idx / d1) + (idx % d1) * d1;
        const double inner = v2 * T0[idx] - ZiUdZiWdWZi[idx] - ZiUdZiWdWZi[tr];
        const double a1 = -vd2 * Z2v[idx] + ZiWdWZi2[idx] - v2 * (ZiUZiWdWZi[idx] + ZiWdWZiUZi2[idx] - 2.0 * vZ[idx]);
        T1[idx] = vd2 * a1 + v2 * (Ut[idx] + v2 * inner);
    }
    __syncwarp();
    for (int p = lane; p < per; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const double x = 0.5 * (T1[i + j * d1] + T1[j + i * d1]);
        out[o + p] = i == j ? x : 1.41421356237309504880 * x;
    }
    // v_Wd_dot and the v entry (:379-383)
    for (int i = lane; i < n12; i += 32) N0[i] = -4.0 * (v * ZiUZiUdZiW[i] + vdZW[i]) + ZWWI[i] + 2.0 * ZiUdZiW[i];
    __syncwarp();
    const double dv1 = mep_dot(Z2v, Ud, n11, lane), dv2 = mep_dot(Z3, U, n11, lane), dv3 = mep_dot(vZ, Ud, n11, lane),
                 dv4 = mep_dot(N0, Wd, n12, lane);
    if (lane == 0)
        out[o + per] = vd * (-8.0 * dv1 + vd * (8.0 * dv2 - (d1 - 1) / v / v / v)) + 4.0 * v * dv3 + 2.0 * dv4;
    // Wtemp (:385-390)
    for (int i = lane; i < n12; i += 32)
        Wt[i] = 4.0 * vd * (ZiUdZiW[i] - v2 * ZiUZiUdZiW[i] + ZWWI[i] - vdZW[i]);
    __syncwarp();
    mep_mm(N0, ZiUdZiW, WdZiW, d1, d2, d2, false, false, 1.0, 0.0, lane);
    mep_mm(N0, ZiWdWZi, UdZiW, d1, d1, d2, false, false, 1.0, 1.0, lane);
    mep_mm(N0, WdWZi, ZiUdZiW, d1, d1, d2, true, false, 1.0, 1.0, lane);
    mep_mm(N0, ZiUdZi, WdWtZiWI, d1, d1, d2, false, false, 1.0, 1.0, lane);
    mep_mm(N0, ZiUd, ZiUdZiW, d1, d1, d2, false, false, -v2, 1.0, lane);
    mep_mm(Bt, WdZiW, WdZiW, d2, d2, d2, false, false, 1.0, 0.0, lane);
    mep_mm(N1, ZiW, Bt, d1, d2, d2, false, false, 1.0, 0.0, lane);                 // ZiW WdZiW WdZiW
    mep_mm(N1, WdWZi, ZiWdWtZiWI, d1, d1, d2, true, false, 1.0, 1.0, lane);
    mep_mm(N1, ZiWdWtZiWI, WdZiW, d1, d2, d2, false, false, 1.0, 1.0, lane);
    mep_mm(Bt, WdZiW, WtZiWI, d2, d2, d2, true, false, 1.0, 0.0, lane);            // WdZiW' WtZiWI
    mep_mm(N1, ZiWd, Bt, d1, d2, d2, false, false, 1.0, 1.0, lane);
    for (int i = lane; i < n12; i += 32) out[o + per + 1 + i] = Wt[i] + 4.0 * v * N0[i] - 2.0 * N1[i];
}


// ---- WSOSInterpPosSemidefTri (R x R matrix polynomials, dim = U svec_length(R) <= 128), wsosinterppossemideftri.jl:108-321 ----
// Region of cone c at vecs + voff[c]: [nP][L_1 .. L_nP][P_1 .. P_nP] (hyp_set_cone_alpha; P_k is U x L_k), then per k
// F_k = Lc_k^-1 (I_R kron P_k)' (R L_k x R U) and the lower Cholesky factor Lc_k of Lambda_k = (I kron P_k)' D(point)
// (I kron P_k) (R L_k x R L_k), then G = F'F (R U x R U) and the dder3 scratch S (R L_max)^2, T (R L_max x R U).
// D(s) has the diagonal blocks Diagonal(smat(s)_pq) (off-diagonal svec blocks scaled by 1 / sqrt 2).  Dense restatement
// of the reference's block-triangular algebra; one CTA of 256 threads per cone; thread i < dim owns entry i.

// entry (pL + a, qL + b) of (I kron P)' D(vec) (I kron P)
__device__ __forceinline__ double wpsd_lambda(const double* P, const double* vec, int U, int L, int row, int col) {
    const int p = row / L, a = row % L, q = col / L, b = col % L;
    const int hi = p > q ? p : q, lo = p > q ? q : p;
    const double* v = vec + (int64_t)(hi * (hi + 1) / 2 + lo) * U;
    double s = 0.0;
    for (int u = 0; u < U; u++) s += P[u + (int64_t)a * U] * v[u] * P[u + (int64_t)b * U];
    return p == q ? s : s * 0.70710678118654752440;
}

static __global__ void __launch_bounds__(256)
wpsd_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ Rs, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                  const int* __restrict__ kidx, const int64_t* __restrict__ moff, const double* __restrict__ point,
                  double* __restrict__ grad, double* __restrict__ H, uint8_t* feas) {
    __shared__ int s_ok;
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], R = Rs[c], nb = R * (R + 1) / 2, U = d / nb, RU = R * U, lde = (d + 1) & ~1;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0;
    for (int k = 0; k < nP; k++) {
        const int64_t L = (int64_t)reg[1 + k];
        sumL += L;
        wsz += R * L * RU + R * L * R * L;
    }
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* G = ws + wsz;
    double* Hc = H + moff[c];
    const double* pt = point + o;
    if (tid == 0) s_ok = 1;
    double gacc = 0.0;
    int gp = 0, gq = 0;                      // block (gp, gq), gq <= gp, and point index gu of entry tid
    if (tid < d) dnn_ij(tid / U, gq, gp);
    const int gu = tid % U;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k], RL = R * L;
        double* F = ws;
        double* Lc = F + (int64_t)RL * RU;
        ws = Lc + RL * RL;
        for (int idx = tid; idx < RL * RL; idx += 256) Lc[idx] = wpsd_lambda(P, pt, U, L, idx % RL, idx / RL);
        __syncthreads();
        for (int j = 0; j < RL; j++) {                     // Cholesky (lower, in place, right-looking)
            if (tid == 0) {
                double dg = Lc[j + j * RL];
                if (!(dg > 0.0)) {
                    s_ok = 0;
                    dg = 1.0;
                }
                Lc[j + j * RL] = sqrt(dg);
            }
            __syncthreads();
            const double dj = Lc[j + j * RL];
            for (int i = j + 1 + tid; i < RL; i += 256) Lc[i + j * RL] /= dj;
            __syncthreads();
            const int r = RL - j - 1;
            for (int idx = tid; idx < r * r; idx += 256) {
                const int ii = j + 1 + idx % r, kk = j + 1 + idx / r;
                if (kk <= ii) Lc[ii + kk * RL] -= Lc[ii + j * RL] * Lc[kk + j * RL];
            }
            __syncthreads();
        }
        // F = Lc^-1 (I kron P)': column (p, u) has P[u, :] in block p
        for (int jj = tid; jj < RU; jj += 256) {
            const int p = jj / U, u = jj % U;
            double* f = F + (int64_t)jj * RL;
            for (int a = 0; a < RL; a++) {
                double s = (a / L == p) ? P[u + (int64_t)(a % L) * U] : 0.0;
                for (int b = 0; b < a; b++) s -= Lc[a + b * RL] * f[b];
                f[a] = s / Lc[a + a * RL];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < RU * RU; idx += 256) {   // G = F'F ("PLambdaiP")
            const int i = idx % RU, j = idx / RU;
            double s = 0.0;
            for (int a = 0; a < RL; a++) s += F[a + (int64_t)i * RL] * F[a + (int64_t)j * RL];
            G[idx] = s;
        }
        __syncthreads();
        // gradient (:144-188) and Hessian (:190-239)
        if (tid < d)
            gacc -= G[(gq * U + gu) + (int64_t)(gp * U + gu) * RU] * (gp == gq ? 1.0 : 1.41421356237309504880);
        for (int idx = tid; idx < d * d; idx += 256) {
            const int e1 = idx % d, e2 = idx / d;
            int p1, q1, p2, q2;
            dnn_ij(e1 / U, q1, p1);
            dnn_ij(e2 / U, q2, p2);
            const int u1 = e1 % U, u2 = e2 % U;
#define WPSD_B(x, y) G[((x) * U + u1) + (int64_t)((y) * U + u2) * RU]
            double v = WPSD_B(p1, p2) * WPSD_B(q1, q2) * (((p1 == q1) != (p2 == q2)) ? 1.41421356237309504880 : 1.0);
            if (p1 != q1 && p2 != q2) v += WPSD_B(p1, q2) * WPSD_B(q1, p2);
#undef WPSD_B
            Hc[e1 + (int64_t)e2 * lde] = k == 0 ? v : Hc[e1 + (int64_t)e2 * lde] + v;
        }
        __syncthreads();
        P += (int64_t)U * L;
    }
    if (tid < d) grad[o + tid] = gacc;
    if (tid == 0 && !s_ok) feas[kidx[c]] = 0;
}

// dder3 = partial_prod! with use_symm_prod (:284-321): per k, S = Lc^-1 (I kron P)' D(dir) (I kron P) Lc^-T, T = S F,
// out[(p, q), u] += <T[:, (p, u)], T[:, (q, u)]> (sqrt 2 off the diagonal blocks)
static __global__ void __launch_bounds__(256)
wpsd_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ Rs, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                  const double* __restrict__ dir, double* __restrict__ out) {
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], R = Rs[c], nb = R * (R + 1) / 2, U = d / nb, RU = R * U;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0, Lmax = 0;
    for (int k = 0; k < nP; k++) {
        const int64_t L = (int64_t)reg[1 + k];
        sumL += L;
        wsz += R * L * RU + R * L * R * L;
        Lmax = L > Lmax ? L : Lmax;
    }
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* S = ws + wsz + (int64_t)RU * RU;
    double* T = S + (R * Lmax) * (R * Lmax);
    double acc = 0.0;
    int gp = 0, gq = 0;
    if (tid < d) dnn_ij(tid / U, gq, gp);
    const int gu = tid % U;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k], RL = R * L;
        const double* F = ws;
        const double* Lc = F + (int64_t)RL * RU;
        ws += (int64_t)RL * RU + RL * RL;
        for (int idx = tid; idx < RL * RL; idx += 256) S[idx] = wpsd_lambda(P, dir + o, U, L, idx % RL, idx / RL);
        __syncthreads();
        for (int col = tid; col < RL; col += 256) {        // S <- Lc^-1 S (columns)
            double* x = S + (int64_t)col * RL;
            for (int r = 0; r < RL; r++) {
                double s = x[r];
                for (int b = 0; b < r; b++) s -= Lc[r + b * RL] * x[b];
                x[r] = s / Lc[r + r * RL];
            }
        }
        __syncthreads();
        for (int row = tid; row < RL; row += 256) {        // S <- S Lc^-T (rows)
            double* x = S + row;
            for (int r = 0; r < RL; r++) {
                double s = x[(int64_t)r * RL];
                for (int b = 0; b < r; b++) s -= Lc[r + b * RL] * x[(int64_t)b * RL];
                x[(int64_t)r * RL] = s / Lc[r + r * RL];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < RL * RU; idx += 256) {   // T = S F
            const int a = idx % RL, j = idx / RL;
            double s = 0.0;
            for (int b = 0; b < RL; b++) s += S[a + (int64_t)b * RL] * F[b + (int64_t)j * RL];
            T[idx] = s;
        }
        __syncthreads();
        if (tid < d) {
            const double* t1 = T + (int64_t)(gp * U + gu) * RL;
            const double* t2 = T + (int64_t)(gq * U + gu) * RL;
            double s = 0.0;
            for (int a = 0; a < RL; a++) s += t1[a] * t2[a];
            acc += s * (gp == gq ? 1.0 : 1.41421356237309504880);
        }
        __syncthreads();
        P += (int64_t)U * L;
    }
    if (tid < d) out[o + tid] = acc;
}


// ---- WSOSInterpEpiNormEucl (R polynomials of U coefficients, dim = R U <= 128), wsosinterpepinormeucl.jl:119-382 ----
// Dense restatement: with A(s) the R L x R L block-arrow matrix (L11 = P' Diagonal(s_1) P on every diagonal block,
// L1r = P' Diagonal(s_r) P on the first block row / column) the reference's barrier -logdet L11 - logdet(Schur) equals
// -logdet A + (R - 2) logdet L11, two logdets of matrices that are linear in s.  The Cholesky factor of L11 is the leading
// L x L block of the factor of A, and L11^-1-type quantities are the leading rows of F = Lc^-1 (I kron P)'.
// Region of cone c: [nP][L_k ..][P_k ..] (hyp_set_cone_alpha), then per k F_k (R L x R U) and Lc_k (R L x R L), then
// G (R U)^2, G11 U^2 and the dder3 scratch S (R L_max)^2, T (R L_max x R U), S11 L_max^2, T11 (L_max x U).
// One CTA of 256 threads per cone; thread i < dim owns entry i = r U + u.

// entry (pL + a, qL + b) of the arrow matrix A(vec)
__device__ __forceinline__ double weuc_arrow(const double* P, const double* vec, int U, int L, int row, int col) {
    const int p = row / L, a = row % L, q = col / L, b = col % L;
    int blk;
    if (p == q) blk = 0;
    else if (p == 0 || q == 0) blk = p > q ? p : q;
    else return 0.0;
    const double* v = vec + (int64_t)blk * U;
    double s = 0.0;
    for (int u = 0; u < U; u++) s += P[u + (int64_t)a * U] * v[u] * P[u + (int64_t)b * U];
    return s;
}

static __global__ void __launch_bounds__(256)
weuc_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ Rs, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                  const int* __restrict__ kidx, const int64_t* __restrict__ moff, const double* __restrict__ point,
                  double* __restrict__ grad, double* __restrict__ H, uint8_t* feas) {
    __shared__ int s_ok;
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], R = Rs[c], U = d / R, RU = d, lde = (d + 1) & ~1;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0;
    for (int k = 0; k < nP; k++) {
        const int64_t L = (int64_t)reg[1 + k];
        sumL += L;
        wsz += R * L * RU + R * L * R * L;
    }
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* G = ws + wsz;
    double* G11 = G + (int64_t)RU * RU;
    double* Hc = H + moff[c];
    const double* pt = point + o;
    if (tid == 0) s_ok = 1;
    double gacc = 0.0;
    const int gr = tid / U, gu = tid % U;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k], RL = R * L;
        double* F = ws;
        double* Lc = F + (int64_t)RL * RU;
        ws = Lc + RL * RL;
        for (int idx = tid; idx < RL * RL; idx += 256) Lc[idx] = weuc_arrow(P, pt, U, L, idx % RL, idx / RL);
        __syncthreads();
        for (int j = 0; j < RL; j++) {                     // Cholesky (lower, in place, right-looking)
            if (tid == 0) {
                double dg = Lc[j + j * RL];
                if (!(dg > 0.0)) {
                    s_ok = 0;
                    dg = 1.0;
                }
                Lc[j + j * RL] = sqrt(dg);
            }
            __syncthreads();
            const double dj = Lc[j + j * RL];
            for (int i = j + 1 + tid; i < RL; i += 256) Lc[i + j * RL] /= dj;
            __syncthreads();
            const int r = RL - j - 1;
            for (int idx = tid; idx < r * r; idx += 256) {
                const int ii = j + 1 + idx % r, kk = j + 1 + idx / r;
                if (kk <= ii) Lc[ii + kk * RL] -= Lc[ii + j * RL] * Lc[kk + j * RL];
            }
            __syncthreads();
        }
        for (int jj = tid; jj < RU; jj += 256) {           // F = Lc^-1 (I kron P)'
            const int p = jj / U, u = jj % U;
            double* f = F + (int64_t)jj * RL;
            for (int a = 0; a < RL; a++) {
                double s = (a / L == p) ? P[u + (int64_t)(a % L) * U] : 0.0;
                for (int b = 0; b < a; b++) s -= Lc[a + b * RL] * f[b];
                f[a] = s / Lc[a + a * RL];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < RU * RU; idx += 256) {   // G = F'F; G11 from the leading L rows of the first block
            const int i = idx % RU, j = idx / RU;
            double s = 0.0;
            for (int a = 0; a < RL; a++) s += F[a + (int64_t)i * RL] * F[a + (int64_t)j * RL];
            G[idx] = s;
            if (i < U && j < U) {
                double s1 = 0.0;
                for (int a = 0; a < L; a++) s1 += F[a + (int64_t)i * RL] * F[a + (int64_t)j * RL];
                G11[i + (int64_t)j * U] = s1;
            }
        }
        __syncthreads();
#define WEUC_G(x, u1_, y, u2_) G[((x) * U + (u1_)) + (int64_t)((y) * U + (u2_)) * RU]
        if (tid < d) {                                      // gradient (:169-211)
            if (gr == 0) {
                double s = -(double)(R - 2) * G11[gu + (int64_t)gu * U];
                for (int rr = 0; rr < R; rr++) s += WEUC_G(rr, gu, rr, gu);
                gacc -= s;
            } else {
                gacc -= 2.0 * WEUC_G(0, gu, gr, gu);
            }
        }
        for (int idx = tid; idx < d * d; idx += 256) {      // Hessian (:213-290)
            const int e1 = idx % d, e2 = idx / d;
            const int r1 = e1 / U, u1 = e1 % U, r2 = e2 / U, u2 = e2 % U;
            double v;
            if (r1 == 0 && r2 == 0) {
                const double g11 = G11[u1 + (int64_t)u2 * U];
                v = -(double)(R - 2) * g11 * g11;
                for (int a = 0; a < R; a++)
                    for (int b = 0; b < R; b++) {
                        const double g = WEUC_G(a, u1, b, u2);
                        v += g * g;
                    }
            } else if (r1 == 0) {
                v = 0.0;
                for (int a = 0; a < R; a++) v += WEUC_G(a, u1, 0, u2) * WEUC_G(a, u1, r2, u2);
                v *= 2.0;
            } else if (r2 == 0) {
                v = 0.0;
                for (int a = 0; a < R; a++) v += WEUC_G(a, u2, 0, u1) * WEUC_G(a, u2, r1, u1);
                v *= 2.0;
            } else {
                v = 2.0 * (WEUC_G(0, u1, 0, u2) * WEUC_G(r1, u1, r2, u2) + WEUC_G(0, u1, r2, u2) * WEUC_G(r1, u1, 0, u2));
            }
            Hc[e1 + (int64_t)e2 * lde] = k == 0 ? v : Hc[e1 + (int64_t)e2 * lde] + v;
        }
#undef WEUC_G
        __syncthreads();
        P += (int64_t)U * L;
    }
    if (tid < d) grad[o + tid] = gacc;
    if (tid == 0 && !s_ok) feas[kidx[c]] = 0;
}

// dder3 (:292-382): Q = F' S^2 F with S = Lc^-1 A(dir) Lc^-T, Q11 likewise with the leading blocks and dir_1
static __global__ void __launch_bounds__(256)
weuc_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ Rs, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                  const double* __restrict__ dir, double* __restrict__ out) {
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], R = Rs[c], U = d / R, RU = d;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0, Lmax = 0;
    for (int k = 0; k < nP; k++) {
        const int64_t L = (int64_t)reg[1 + k];
        sumL += L;
        wsz += R * L * RU + R * L * R * L;
        Lmax = L > Lmax ? L : Lmax;
    }
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* S = ws + wsz + (int64_t)RU * RU + (int64_t)U * U;
    double* T = S + (R * Lmax) * (R * Lmax);
    double* S11 = T + (R * Lmax) * RU;
    double* T11 = S11 + Lmax * Lmax;
    double acc = 0.0;
    const int gr = tid / U, gu = tid % U;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k], RL = R * L;
        const double* F = ws;
        const double* Lc = F + (int64_t)RL * RU;
        ws += (int64_t)RL * RU + RL * RL;
        for (int idx = tid; idx < RL * RL; idx += 256) S[idx] = weuc_arrow(P, dir + o, U, L, idx % RL, idx / RL);
        for (int idx = tid; idx < L * L; idx += 256) S11[idx] = weuc_arrow(P, dir + o, U, L, idx % L, idx / L);
        __syncthreads();
        for (int col = tid; col < RL + L; col += 256) {    // Lc^-1 S and L11^-1 S11 (columns)
            const bool big = col < RL;
            const int n = big ? RL : L;
            double* x = big ? S + (int64_t)col * RL : S11 + (int64_t)(col - RL) * L;
            for (int r = 0; r < n; r++) {
                double s = x[r];
                for (int b = 0; b < r; b++) s -= Lc[r + b * RL] * x[b];
                x[r] = s / Lc[r + r * RL];
            }
        }
        __syncthreads();
        for (int row = tid; row < RL + L; row += 256) {    // ... Lc^-T and L11^-T (rows)
            const bool big = row < RL;
            const int n = big ? RL : L;
            double* x = big ? S + row : S11 + (row - RL);
            for (int r = 0; r < n; r++) {
                double s = x[(int64_t)r * n];
                for (int b = 0; b < r; b++) s -= Lc[r + b * RL] * x[(int64_t)b * n];
                x[(int64_t)r * n] = s / Lc[r + r * RL];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < RL * RU; idx += 256) {   // T = S F
            const int a = idx % RL, j = idx / RL;
            double s = 0.0;
            for (int b = 0; b < RL; b++) s += S[a + (int64_t)b * RL] * F[b + (int64_t)j * RL];
            T[idx] = s;
        }
        for (int idx = tid; idx < L * U; idx += 256) {     // T11 = S11 F11, F11 = leading L rows of the columns (0, u) of F
            const int a = idx % L, j = idx / L;
            double s = 0.0;
            for (int b = 0; b < L; b++) s += S11[a + (int64_t)b * L] * F[b + (int64_t)j * RL];
            T11[idx] = s;
        }
        __syncthreads();
        if (tid < d) {
            if (gr == 0) {
                double s = 0.0;
                for (int rr = 0; rr < R; rr++) {
                    const double* t = T + (int64_t)(rr * U + gu) * RL;
                    for (int a = 0; a < RL; a++) s += t[a] * t[a];
                }
                double s1 = 0.0;
                for (int a = 0; a < L; a++) s1 += T11[a + (int64_t)gu * L] * T11[a + (int64_t)gu * L];
                acc += s - (double)(R - 2) * s1;
            } else {
                const double* t1 = T + (int64_t)gu * RL;
                const double* t2 = T + (int64_t)(gr * U + gu) * RL;
                double s = 0.0;
                for (int a = 0; a < RL; a++) s += t1[a] * t2[a];
                acc += 2.0 * s;
            }
        }
        __syncthreads();
        P += (int64_t)U * L;
    }
    if (tid < d) out[o + tid] = acc;
}


// ---- WSOSInterpEpiNormOne (R polynomials of U coefficients, dim = R U <= 128), wsosinterpepinormone.jl:147-493 ----
// Dense restatement: the reference's barrier -logdet L11 - sum_{r >= 2} logdet(L11 - L1r L11^-1 L1r) equals
// sum_{r >= 2} -logdet A2(s_1, s_r) + (R - 2) logdet L11 with A2 the 2 L x 2 L arrow matrix [L11 L1r; L1r L11] of the
// R = 2 Euclidean-norm cone above: every oracle is a sum over the R - 1 pairs plus the L11 correction.
// Region of cone c: [nP][L_k ..][P_k ..], then per (k, r) F (2 L x 2 U) and Lc (2 L x 2 L), then G (2 U)^2, G11 U^2 and
// the dder3 scratch S (2 L_max)^2, T (2 L_max x 2 U), S11 L_max^2, T11 (L_max x U).  One CTA of 256 threads per cone.

// entry (pL + a, qL + b), p, q in {0, 1}, of the pair arrow matrix built from the coefficient vectors v0 (diagonal) and v1
__device__ __forceinline__ double wone_arrow(const double* P, const double* v0, const double* v1, int U, int L, int row,
                                             int col) {
    const int p = row / L, a = row % L, q = col / L, b = col % L;
    const double* v = p == q ? v0 : v1;
    double s = 0.0;
    for (int u = 0; u < U; u++) s += P[u + (int64_t)a * U] * v[u] * P[u + (int64_t)b * U];
    return s;
}

static __global__ void __launch_bounds__(256)
wone_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ Rs, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                  const int* __restrict__ kidx, const int64_t* __restrict__ moff, const double* __restrict__ point,
                  double* __restrict__ grad, double* __restrict__ H, uint8_t* feas) {
    __shared__ int s_ok;
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], R = Rs[c], U = d / R, U2 = 2 * U, lde = (d + 1) & ~1;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0;
    for (int k = 0; k < nP; k++) {
        const int64_t L = (int64_t)reg[1 + k];
        sumL += L;
        wsz += (R - 1) * (2 * L * U2 + 4 * L * L);
    }
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* G = ws + wsz;
    double* G11 = G + (int64_t)U2 * U2;
    double* Hc = H + moff[c];
    const double* pt = point + o;
    if (tid == 0) s_ok = 1;
    for (int idx = tid; idx < d * d; idx += 256) Hc[(idx % d) + (int64_t)(idx / d) * lde] = 0.0;
    __syncthreads();
    double gacc = 0.0;
    const int gr = tid / U, gu = tid % U;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k], L2 = 2 * L;
        for (int r = 1; r < R; r++) {
            double* F = ws;
            double* Lc = F + (int64_t)L2 * U2;
            ws = Lc + L2 * L2;
            for (int idx = tid; idx < L2 * L2; idx += 256)
                Lc[idx] = wone_arrow(P, pt, pt + (int64_t)r * U, U, L, idx % L2, idx / L2);
            __syncthreads();
            for (int j = 0; j < L2; j++) {                 // Cholesky (lower, in place, right-looking)
                if (tid == 0) {
                    double dg = Lc[j + j * L2];
                    if (!(dg > 0.0)) {
                        s_ok = 0;
                        dg = 1.0;
                    }
                    Lc[j + j * L2] = sqrt(dg);
                }
                __syncthreads();
                const double dj = Lc[j + j * L2];
                for (int i = j + 1 + tid; i < L2; i += 256) Lc[i + j * L2] /= dj;
                __syncthreads();
                const int rr = L2 - j - 1;
                for (int idx = tid; idx < rr * rr; idx += 256) {
                    const int ii = j + 1 + idx % rr, kk = j + 1 + idx / rr;
                    if (kk <= ii) Lc[ii + kk * L2] -= Lc[ii + j * L2] * Lc[kk + j * L2];
                }
                __syncthreads();
            }
            for (int jj = tid; jj < U2; jj += 256) {       // F = Lc^-1 (I_2 kron P)'
                const int p = jj / U, u = jj % U;
                double* f = F + (int64_t)jj * L2;
                for (int a = 0; a < L2; a++) {
                    double s = (a / L == p) ? P[u + (int64_t)(a % L) * U] : 0.0;
                    for (int b = 0; b < a; b++) s -= Lc[a + b * L2] * f[b];
                    f[a] = s / Lc[a + a * L2];
                }
            }
            __syncthreads();
            for (int idx = tid; idx < U2 * U2; idx += 256) {
                const int i = idx % U2, j = idx / U2;
                double s = 0.0;
                for (int a = 0; a < L2; a++) s += F[a + (int64_t)i * L2] * F[a + (int64_t)j * L2];
                G[idx] = s;
                if (r == 1 && i < U && j < U) {
                    double s1 = 0.0;
                    for (int a = 0; a < L; a++) s1 += F[a + (int64_t)i * L2] * F[a + (int64_t)j * L2];
                    G11[i + (int64_t)j * U] = s1;
                }
            }
            __syncthreads();
#define WONE_G(x, u1_, y, u2_) G[((x) * U + (u1_)) + (int64_t)((y) * U + (u2_)) * U2]
            if (tid < d) {
                if (gr == 0) gacc -= WONE_G(0, gu, 0, gu) + WONE_G(1, gu, 1, gu);
                else if (gr == r) gacc -= 2.0 * WONE_G(0, gu, 1, gu);
            }
            for (int idx = tid; idx < d * d; idx += 256) {
                const int e1 = idx % d, e2 = idx / d;
                const int r1 = e1 / U, u1 = e1 % U, r2 = e2 / U, u2 = e2 % U;
                double v;
                if (r1 == 0 && r2 == 0) {
                    v = 0.0;
                    for (int a = 0; a < 2; a++)
                        for (int b = 0; b < 2; b++) {
                            const double g = WONE_G(a, u1, b, u2);
                            v += g * g;
                        }
                } else if (r1 == 0 && r2 == r) {
                    v = 2.0 * (WONE_G(0, u1, 0, u2) * WONE_G(0, u1, 1, u2) + WONE_G(1, u1, 0, u2) * WONE_G(1, u1, 1, u2));
                } else if (r1 == r && r2 == 0) {
                    v = 2.0 * (WONE_G(0, u2, 0, u1) * WONE_G(0, u2, 1, u1) + WONE_G(1, u2, 0, u1) * WONE_G(1, u2, 1, u1));
                } else if (r1 == r && r2 == r) {
                    v = 2.0 * (WONE_G(0, u1, 0, u2) * WONE_G(1, u1, 1, u2) + WONE_G(0, u1, 1, u2) * WONE_G(1, u1, 0, u2));
                } else {
                    continue;
                }
                Hc[e1 + (int64_t)e2 * lde] += v;
            }
#undef WONE_G
            __syncthreads();
        }
        // the (R - 2) logdet L11 correction
        if (tid < d && gr == 0) gacc += (double)(R - 2) * G11[gu + (int64_t)gu * U];
        for (int idx = tid; idx < U * U; idx += 256) {
            const double g11 = G11[idx];
            Hc[(idx % U) + (int64_t)(idx / U) * lde] -= (double)(R - 2) * g11 * g11;
        }
        __syncthreads();
        P += (int64_t)U * L;
    }
    if (tid < d) grad[o + tid] = gacc;
    if (tid == 0 && !s_ok) feas[kidx[c]] = 0;
}

static __global__ void __launch_bounds__(256)
wone_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                  const int* __restrict__ Rs, const int64_t* __restrict__ voff, double* __restrict__ vecs,
                  const double* __restrict__ dir, double* __restrict__ out) {
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], R = Rs[c], U = d / R, U2 = 2 * U;
    double* reg = vecs + voff[c];
    const int nP = (int)reg[0];
    int64_t sumL = 0, wsz = 0, Lmax = 0;
    for (int k = 0; k < nP; k++) {
        const int64_t L = (int64_t)reg[1 + k];
        sumL += L;
        wsz += (R - 1) * (2 * L * U2 + 4 * L * L);
        Lmax = L > Lmax ? L : Lmax;
    }
    const double* P = reg + 1 + nP;
    double* ws = reg + 1 + nP + (int64_t)U * sumL;
    double* S = ws + wsz + (int64_t)U2 * U2 + (int64_t)U * U;
    double* T = S + 4 * Lmax * Lmax;
    double* S11 = T + 2 * Lmax * U2;
    double* T11 = S11 + Lmax * Lmax;
    const double* dr = dir + o;
    double acc = 0.0;
    const int gr = tid / U, gu = tid % U;
    for (int k = 0; k < nP; k++) {
        const int L = (int)reg[1 + k], L2 = 2 * L;
        for (int r = 1; r < R; r++) {
            const double* F = ws;
            const double* Lc = F + (int64_t)L2 * U2;
            ws += (int64_t)L2 * U2 + L2 * L2;
            for (int idx = tid; idx < L2 * L2; idx += 256)
                S[idx] = wone_arrow(P, dr, dr + (int64_t)r * U, U, L, idx % L2, idx / L2);
            if (r == 1)
                for (int idx = tid; idx < L * L; idx += 256) S11[idx] = wone_arrow(P, dr, dr, U, L, idx % L, idx / L);
            __syncthreads();
            const int ncol = L2 + (r == 1 ? L : 0);
            for (int col = tid; col < ncol; col += 256) {  // Lc^-1 S (and L11^-1 S11: leading block of Lc), columns
                const bool big = col < L2;
                const int n = big ? L2 : L;
                double* x = big ? S + (int64_t)col * L2 : S11 + (int64_t)(col - L2) * L;
                for (int a = 0; a < n; a++) {
                    double s = x[a];
                    for (int b = 0; b < a; b++) s -= Lc[a + b * L2] * x[b];
                    x[a] = s / Lc[a + a * L2];
                }
            }
            __syncthreads();
            for (int row = tid; row < ncol; row += 256) {  // ... Lc^-T, rows
                const bool big = row < L2;
                const int n = big ? L2 : L;
                double* x = big ? S + row : S11 + (row - L2);
                for (int a = 0; a < n; a++) {
                    double s = x[(int64_t)a * n];
                    for (int b = 0; b < a; b++) s -= Lc[a + b * L2] * x[(int64_t)b * n];
                    x[(int64_t)a * n] = s / Lc[a + a * L2];
                }
            }
            __syncthreads();
            for (int idx = tid; idx < L2 * U2; idx += 256) {
                const int a = idx % L2, j = idx / L2;
                double s = 0.0;
                for (int b = 0; b < L2; b++) s += S[a + (int64_t)b * L2] * F[b + (int64_t)j * L2];
                T[idx] = s;
            }
            if (r == 1)
                for (int idx = tid; idx < L * U; idx += 256) {
                    const int a = idx % L, j = idx / L;
                    double s = 0.0;
                    for (int b = 0; b < L; b++) s += S11[a + (int64_t)b * L] * F[b + (int64_t)j * L2];
                    T11[idx] = s;
                }
            __syncthreads();
            if (tid < d) {
                const double* t0 = T + (int64_t)gu * L2;
                const double* t1 = T + (int64_t)(U + gu) * L2;
                if (gr == 0) {
                    double s = 0.0;
                    for (int a = 0; a < L2; a++) s += t0[a] * t0[a] + t1[a] * t1[a];
                    acc += s;
                    if (r == 1) {
                        double s1 = 0.0;
                        for (int a = 0; a < L; a++) s1 += T11[a + (int64_t)gu * L] * T11[a + (int64_t)gu * L];
                        acc -= (double)(R - 2) * s1;
                    }
                } else if (gr == r) {
                    double s = 0.0;
                    for (int a = 0; a < L2; a++) s += t0[a] * t1[a];
                    acc += 2.0 * s;
                }
            }
            __syncthreads();
        }
        P += (int64_t)U * L;
    }
    if (tid < d) out[o + tid] = acc;
}


// ---- PosSemidefTriSparse (real, dense implementation as the reference's PSDSparseDense; dim = nnz <= 128),
// possemideftrisparse/denseimpl.jl:30-167 ----
// Region of cone c at vecs + voff[c]: [side][row_1 .. row_dim][col_1 .. col_dim] (hyp_set_cone_alpha; 0-based, col <= row,
// every diagonal entry present), then the workspace: lower Cholesky factor of the dense matrix (side^2), its inverse
// Li (side^2) and three side^2 scratch matrices for dder3.  One CTA of 256 threads per cone.

static __global__ void __launch_bounds__(256)
sps_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ voff, double* __restrict__ vecs, const int* __restrict__ kidx,
                 const int64_t* __restrict__ moff, const double* __restrict__ point, double* __restrict__ grad,
                 double* __restrict__ H, uint8_t* feas) {
    __shared__ int s_ok;
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c], lde = (d + 1) & ~1;
    double* reg = vecs + voff[c];
    const int sd = (int)reg[0], s2 = sd * sd;
    const double* rows = reg + 1;
    const double* cols = reg + 1 + d;
    double* Lc = reg + 1 + 2 * d;
    double* Li = Lc + s2;
    double* Hc = H + moff[c];
    if (tid == 0) s_ok = 1;
    for (int idx = tid; idx < s2; idx += 256) Lc[idx] = 0.0;
    __syncthreads();
    for (int e = tid; e < d; e += 256) {                  // svec_to_smat_sparse! (:207-222), both triangles
        const int i = (int)rows[e], j = (int)cols[e];
        const double x = i == j ? point[o + e] : point[o + e] * 0.70710678118654752440;
        Lc[i + j * sd] = x;
        Lc[j + i * sd] = x;
    }
    __syncthreads();
    for (int j = 0; j < sd; j++) {                        // Cholesky (lower, in place, right-looking)
        if (tid == 0) {
            double dg = Lc[j + j * sd];
            if (!(dg > 0.0)) {
                s_ok = 0;
                dg = 1.0;
            }
            Lc[j + j * sd] = sqrt(dg);
        }
        __syncthreads();
        const double dj = Lc[j + j * sd];
        for (int i = j + 1 + tid; i < sd; i += 256) Lc[i + j * sd] /= dj;
        __syncthreads();
        const int r = sd - j - 1;
        for (int idx = tid; idx < r * r; idx += 256) {
            const int ii = j + 1 + idx % r, kk = j + 1 + idx / r;
            if (kk <= ii) Lc[ii + kk * sd] -= Lc[ii + j * sd] * Lc[kk + j * sd];
        }
        __syncthreads();
    }
    for (int b = tid; b < sd; b += 256) {                 // Li = (L L')^-1, column by column
        double* x = Li + (int64_t)b * sd;
        for (int r = 0; r < sd; r++) {
            double s = r == b ? 1.0 : 0.0;
            for (int k = 0; k < r; k++) s -= Lc[r + k * sd] * x[k];
            x[r] = s / Lc[r + r * sd];
        }
        for (int r = sd - 1; r >= 0; r--) {
            double s = x[r];
            for (int k = r + 1; k < sd; k++) s -= Lc[k + r * sd] * x[k];
            x[r] = s / Lc[r + r * sd];
        }
    }
    __syncthreads();
    for (int e = tid; e < d; e += 256) {                  // update_grad (:43-55)
        const int i = (int)rows[e], j = (int)cols[e];
        grad[o + e] = -Li[i + j * sd] * (i == j ? 1.0 : 1.41421356237309504880);
    }
    for (int idx = tid; idx < d * d; idx += 256) {        // update_hess (:57-83), both triangles
        const int e1 = idx % d, e2 = idx / d;
        const int i1 = (int)rows[e1], j1 = (int)cols[e1], i2 = (int)rows[e2], j2 = (int)cols[e2];
        double v;
        if (i1 == j1 && i2 == j2) v = Li[i1 + i2 * sd] * Li[i1 + i2 * sd];
        else if ((i1 == j1) != (i2 == j2)) v = 1.41421356237309504880 * Li[i1 + i2 * sd] * Li[j1 + j2 * sd];
        else v = Li[i1 + i2 * sd] * Li[j1 + j2 * sd] + Li[i1 + j2 * sd] * Li[j1 + i2 * sd];
        Hc[e1 + (int64_t)e2 * lde] = v;
    }
    if (tid == 0 && !s_ok) feas[kidx[c]] = 0;
}

// dder3 (:153-167): entries of Li D Li D Li on the pattern, D = smat(dir)
static __global__ void __launch_bounds__(256)
sps_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ voff, double* __restrict__ vecs, const double* __restrict__ dir,
                 double* __restrict__ out) {
    const int c = blockIdx.x, tid = threadIdx.x;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    double* reg = vecs + voff[c];
    const int sd = (int)reg[0], s2 = sd * sd;
    const double* rows = reg + 1;
    const double* cols = reg + 1 + d;
    const double* Li = reg + 1 + 2 * d + s2;
    double* D = reg + 1 + 2 * d + 2 * s2;
    double* W1 = D + s2;
    double* W2 = W1 + s2;
    for (int idx = tid; idx < s2; idx += 256) D[idx] = 0.0;
    __syncthreads();
    for (int e = tid; e < d; e += 256) {
        const int i = (int)rows[e], j = (int)cols[e];
        const double x = i == j ? dir[o + e] : dir[o + e] * 0.70710678118654752440;
        D[i + j * sd] = x;
        D[j + i * sd] = x;
    }
    __syncthreads();
    for (int idx = tid; idx < s2; idx += 256) {           // W1 = Li D
        const int a = idx % sd, b = idx / sd;
        double s = 0.0;
        for (int k = 0; k < sd; k++) s += Li[a + k * sd] * D[k + b * sd];
        W1[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < s2; idx += 256) {           // W2 = W1 Li
        const int a = idx % sd, b = idx / sd;
        double s = 0.0;
        for (int k = 0; k < sd; k++) s += W1[a + k * sd] * Li[k + b * sd];
        W2[idx] = s;
    }
    __syncthreads();
    for (int e = tid; e < d; e += 256) {                  // (W2 W1')[i, j] = (Li D Li D Li)[i, j]
        const int i = (int)rows[e], j = (int)cols[e];
        double s = 0.0;
        for (int k = 0; k < sd; k++) s += W2[i + k * sd] * W1[j + k * sd];
        out[o + e] = s * (i == j ? 1.0 : 1.41421356237309504880);
    }
}


// ---- EpiTrRelEntropyTri (dim = 1 + 2 svec_length(d) <= 128, so d <= 10), epitrrelentropytri.jl:137-573 ----
// point = (u, svec(V), svec(W)); z = u - tr(W log W - W log V).  Restated through the Frechet derivatives of the matrix
// logarithm at V = Qv diag(lv) Qv' and W = Qw diag(lw) Qw' (Daleckii-Krein; D1 / D2 / D3 = first / second / third divided
// differences of log with confluent nodes, the "Delta" arrays of the reference).  Per-cone state at vecs + voff[c]:
// Qv, Qw, logV, logW, Vi, Wi, D1v, D1w, zV, zW (10 d^2), lv, lw (2 d), D2v, D2w (2 d^3), D3v (d^4), then 12 d^2 of scratch
// for dder3.  scal: 0 z.  One warp per cone / per (cone, column); eigen-decompositions by cyclic Jacobi inside the warp.

// log[x_0 .. x_k], k <= 3, with confluent nodes: sort, then the divided-difference table where a run of (nearly) equal
// nodes takes the derivative limit log^(m)(x) / m! = (-1)^(m-1) / (m x^m)
__device__ __forceinline__ double etr_logdd(double x0, double x1, double x2, double x3, int k) {
    double xs[4] = {x0, x1, x2, x3};
    for (int i = 1; i <= k; i++)
        for (int j = i; j > 0 && xs[j] < xs[j - 1]; j--) {
            const double t = xs[j];
            xs[j] = xs[j - 1];
            xs[j - 1] = t;
        }
    double tb[4];
    for (int i = 0; i <= k; i++) tb[i] = log(xs[i]);
    for (int m = 1; m <= k; m++)
        for (int i = 0; i + m <= k; i++) {
            const double a = xs[i], b = xs[i + m];
            if (fabs(b - a) <= 1e-9 * fmax(fabs(a), fabs(b))) {
                const double x = 0.5 * (a + b);
                double pw = x;
                for (int e = 1; e < m; e++) pw *= x;
                tb[i] = ((m & 1) ? 1.0 : -1.0) / (m * pw);
            } else {
                tb[i] = (tb[i + 1] - tb[i]) / (b - a);
            }
        }
    return tb[0];
}

// eigen-decomposition of the symmetric d x d matrix A (destroyed) by cyclic Jacobi, one warp: Q (columns), lam
__device__ __forceinline__ void etr_jacobi(double* A, double* Q, double* lam, int d, int lane) {
    for (int i = lane; i < d * d; i += 32) Q[i] = (i % d == i / d) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; sweep++) {
        double offn = 0.0, dn = 0.0;
        for (int i = lane; i < d * d; i += 32) {
            const double x = A[i] * A[i];
            if (i % d == i / d) dn += x;
            else offn += x;
        }
        offn = warp_sum(offn);
        dn = warp_sum(dn);
        if (offn <= 1e-30 * dn || offn == 0.0) break;
        for (int p = 0; p < d - 1; p++)
            for (int q = p + 1; q < d; q++) {
                const double apq = A[p + q * d];
                if (apq == 0.0) continue;
                const double app = A[p + p * d], aqq = A[q + q * d];
                const double th = (aqq - app) / (2.0 * apq);
                const double t = copysign(1.0, th) / (fabs(th) + sqrt(1.0 + th * th));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                __syncwarp();
                for (int k = lane; k < d; k += 32) {         // columns p, q of A and Q
                    const double akp = A[k + p * d], akq = A[k + q * d];
                    A[k + p * d] = cs * akp - sn * akq;
                    A[k + q * d] = sn * akp + cs * akq;
                    const double qkp = Q[k + p * d], qkq = Q[k + q * d];
                    Q[k + p * d] = cs * qkp - sn * qkq;
                    Q[k + q * d] = sn * qkp + cs * qkq;
                }
                __syncwarp();
                for (int k = lane; k < d; k += 32) {         // rows p, q of A
                    const double apk = A[p + k * d], aqk = A[q + k * d];
                    A[p + k * d] = cs * apk - sn * aqk;
                    A[q + k * d] = sn * apk + cs * aqk;
                }
                __syncwarp();
            }
    }
    for (int i = lane; i < d; i += 32) lam[i] = A[i + i * d];
    __syncwarp();
}

// O = Q diag(f) Q'
__device__ __forceinline__ void etr_spectral(double* O, const double* Q, const double* f, int d, int lane) {
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx % d, j = idx / d;
        double s = 0.0;
        for (int k = 0; k < d; k++) s += Q[i + k * d] * f[k] * Q[j + k * d];
        O[idx] = s;
    }
    __syncwarp();
}
// O = Q' H Q (t1: scratch)
__device__ __forceinline__ void etr_in(double* O, const double* Q, const double* H, double* t1, int d, int lane) {
    mep_mm(t1, Q, H, d, d, d, true, false, 1.0, 0.0, lane);
    mep_mm(O, t1, Q, d, d, d, false, false, 1.0, 0.0, lane);
}
// O = Q T Q'
__device__ __forceinline__ void etr_out(double* O, const double* Q, const double* T, double* t1, int d, int lane) {
    mep_mm(t1, Q, T, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(O, t1, Q, d, d, d, false, true, 1.0, 0.0, lane);
}
// O = Dlog(X)[H]; t1, t2 scratch
__device__ __forceinline__ void etr_d1(double* O, const double* Q, const double* D1, const double* H, double* t1,
                                       double* t2, int d, int lane) {
    etr_in(t2, Q, H, t1, d, lane);
    for (int i = lane; i < d * d; i += 32) t2[i] *= D1[i];
    __syncwarp();
    etr_out(O, Q, t2, t1, d, lane);
}
// O = D2log(X)[H, K] from the rotated Ht = Q'HQ, Kt = Q'KQ; t1, t2 scratch
__device__ __forceinline__ void etr_d2t(double* O, const double* Q, const double* D2, const double* Ht, const double* Kt,
                                        double* t1, double* t2, int d, int lane) {
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx % d, j = idx / d;
        double s = 0.0;
        for (int k = 0; k < d; k++)
            s += D2[i + d * (k + d * j)] * (Ht[i + k * d] * Kt[k + j * d] + Kt[i + k * d] * Ht[k + j * d]);
        t2[idx] = s;
    }
    __syncwarp();
    etr_out(O, Q, t2, t1, d, lane);
}

// hess_prod for one column (epitrrelentropytri.jl:210-267).  sc: 8 * 104 doubles of scratch private to the warp.
__device__ __forceinline__ void etr_hess_col(const double* a, double* pr, int d, const double* st, double z, double* sc,
                                             int lane) {
    const int n = d * d, vw = d * (d + 1) / 2;
    const double* Qv = st;
    const double* Qw = st + n;
    const double* Vi = st + 4 * n;
    const double* Wi = st + 5 * n;
    const double* D1v = st + 6 * n;
    const double* D1w = st + 7 * n;
    const double* zV = st + 8 * n;
    const double* zW = st + 9 * n;
    const double* D2v = st + 10 * n + 2 * d;
    const double* Wt = st + 10 * n + 2 * d + 2 * d * n + n * n;      // Qv' W Qv, stored by the state kernel
    double* dV = sc;
    double* dW = sc + 104;
    double* hV = sc + 2 * 104;
    double* hW = sc + 3 * 104;
    double* t1 = sc + 4 * 104;
    double* t2 = sc + 5 * 104;
    double* t3 = sc + 6 * 104;
    double* t4 = sc + 7 * 104;
    const double du = a[0];
    dnn_smat(dV, a + 1, d, vw, lane);
    dnn_smat(dW, a + 1 + vw, d, vw, lane);
    __syncwarp();
    const double dz = du + mep_dot(zV, dV, n, lane) + mep_dot(zW, dW, n, lane);
    // hV = D2log(V)[W, dV] + Dlog(V)[dW];  hW = Dlog(V)[dV] - Dlog(W)[dW]
    etr_in(t3, Qv, dV, t1, d, lane);
    etr_d2t(hV, Qv, D2v, Wt, t3, t1, t2, d, lane);
    etr_d1(t4, Qv, D1v, dW, t1, t2, d, lane);
    mep_axpby(hV, 1.0, hV, 1.0, t4, n, lane);
    etr_d1(hW, Qv, D1v, dV, t1, t2, d, lane);
    etr_d1(t4, Qw, D1w, dW, t1, t2, d, lane);
    mep_axpby(hW, 1.0, hW, -1.0, t4, n, lane);
    // + Vi dV Vi, + Wi dW Wi
    mep_mm(t1, Vi, dV, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(t3, t1, Vi, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(t1, Wi, dW, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(t4, t1, Wi, d, d, d, false, false, 1.0, 0.0, lane);
    const double c = dz / (z * z), zi = 1.0 / z;
    for (int p = lane; p < vw; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const int e = i + j * d, et = j + i * d;
        const double sV = c * zV[e] - zi * 0.5 * (hV[e] + hV[et]) + 0.5 * (t3[e] + t3[et]);
        const double sW = c * zW[e] - zi * 0.5 * (hW[e] + hW[et]) + 0.5 * (t4[e] + t4[et]);
        const double f = i == j ? 1.0 : 1.41421356237309504880;
        pr[1 + p] = f * sV;
        pr[1 + vw + p] = f * sW;
    }
    if (lane == 0) pr[0] = c;
    __syncwarp();
}

static __global__ void __launch_bounds__(128)
etr_state_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ voff, double* __restrict__ vecs, const int* __restrict__ kidx,
                 const int64_t* __restrict__ moff, const double* __restrict__ point, double* __restrict__ grad,
                 double* __restrict__ scal, double* __restrict__ H, uint8_t* feas) {
    __shared__ double sh[4][8 * 104 + 128];
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    double* sc = sh[threadIdx.x >> 5];
    const int64_t o = off[c];
    const int dm = dim[c], vw = (dm - 1) / 2, lde = (dm + 1) & ~1;
    int d = 1;
    while (d * (d + 1) / 2 < vw) d++;
    const int n = d * d;
    double* st = vecs + voff[c];
    double* Qv = st;
    double* Qw = st + n;
    double* logV = st + 2 * n;
    double* logW = st + 3 * n;
    double* Vi = st + 4 * n;
    double* Wi = st + 5 * n;
    double* D1v = st + 6 * n;
    double* D1w = st + 7 * n;
    double* zV = st + 8 * n;
    double* zW = st + 9 * n;
    double* lv = st + 10 * n;
    double* lw = lv + d;
    double* D2v = lw + d;
    double* D2w = D2v + d * n;
    double* D3v = D2w + d * n;
    double* Wt = D3v + n * n;
    double* Vm = sc;                 // smat(V), smat(W): kept (Jacobi works on copies)
    double* Wm = sc + 104;
    double* t1 = sc + 2 * 104;
    double* t2 = sc + 3 * 104;
    double* f = sc + 4 * 104;
    const double u = point[o];
    dnn_smat(Vm, point + o + 1, d, vw, lane);
    dnn_smat(Wm, point + o + 1 + vw, d, vw, lane);
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        t1[i] = Vm[i];
        t2[i] = Wm[i];
    }
    __syncwarp();
    etr_jacobi(t1, Qv, lv, d, lane);
    etr_jacobi(t2, Qw, lw, d, lane);
    // update_feas (:137-166): both matrices positive definite and z > 0
    double lmin = 1e300;
    for (int i = 0; i < d; i++) lmin = fmin(lmin, fmin(lv[i], lw[i]));
    const bool pd = lmin > 0.0;
    for (int i = lane; i < d; i += 32) f[i] = pd ? log(lv[i]) : 0.0;
    __syncwarp();
    etr_spectral(logV, Qv, f, d, lane);
    for (int i = lane; i < d; i += 32) f[i] = pd ? log(lw[i]) : 0.0;
    __syncwarp();
    etr_spectral(logW, Qw, f, d, lane);
    for (int i = lane; i < d; i += 32) f[i] = pd ? 1.0 / lv[i] : 1.0;
    __syncwarp();
    etr_spectral(Vi, Qv, f, d, lane);
    for (int i = lane; i < d; i += 32) f[i] = pd ? 1.0 / lw[i] : 1.0;
    __syncwarp();
    etr_spectral(Wi, Qw, f, d, lane);
    double tr = 0.0;
    for (int i = lane; i < n; i += 32) tr += Wm[i] * (logW[i] - logV[i]);
    const double z = u - warp_sum(tr);
    const bool ok = pd && z > 0.0;
    // divided differences of log at the eigenvalues (1.0 in place of a nonpositive eigenvalue keeps the arithmetic finite)
    for (int idx = lane; idx < n; idx += 32) {
        const int i = idx % d, j = idx / d;
        D1v[idx] = pd ? etr_logdd(lv[i], lv[j], 0, 0, 1) : 1.0;
        D1w[idx] = pd ? etr_logdd(lw[i], lw[j], 0, 0, 1) : 1.0;
    }
    for (int idx = lane; idx < d * n; idx += 32) {
        const int i = idx % d, k = (idx / d) % d, j = idx / n;
        D2v[idx] = pd ? etr_logdd(lv[i], lv[k], lv[j], 0, 2) : 0.0;
        D2w[idx] = pd ? etr_logdd(lw[i], lw[k], lw[j], 0, 2) : 0.0;
    }
    for (int idx = lane; idx < n * n; idx += 32) {
        const int i = idx % d, k = (idx / d) % d, l = (idx / n) % d, j = idx / (d * n);
        D3v[idx] = pd ? etr_logdd(lv[i], lv[k], lv[l], lv[j], 3) : 0.0;
    }
    __syncwarp();
    // dz/dV = Dlog(V)[W], dz/dW = -(log W + I - log V); Wt = Qv' W Qv is kept for the second derivatives
    etr_in(Wt, Qv, Wm, t1, d, lane);
    for (int i = lane; i < n; i += 32) t2[i] = Wt[i] * D1v[i];
    __syncwarp();
    etr_out(zV, Qv, t2, t1, d, lane);
    for (int i = lane; i < n; i += 32) zW[i] = -(logW[i] + ((i % d == i / d) ? 1.0 : 0.0) - logV[i]);
    __syncwarp();
    // update_grad (:168-208)
    const double zi = 1.0 / z;
    for (int p = lane; p < vw; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const int e = i + j * d;
        const double fs = i == j ? 1.0 : 1.41421356237309504880;
        grad[o + 1 + p] = fs * (-zi * zV[e] - Vi[e]);
        grad[o + 1 + vw + p] = fs * (-zi * zW[e] - Wi[e]);
    }
    if (lane == 0) {
        grad[o] = -zi;
        scal[8 * c] = z;
        if (!ok) feas[kidx[c]] = 0;
    }
    __syncwarp();
    // explicit Hessian: hess_prod applied to the unit vectors, then symmetrised
    double* Hc = H + moff[c];
    double* unit = sc + 8 * 104;
    for (int j = 0; j < dm; j++) {
        for (int i = lane; i < dm; i += 32) unit[i] = i == j ? 1.0 : 0.0;
        __syncwarp();
        etr_hess_col(unit, Hc + (int64_t)j * lde, d, st, z, sc, lane);
    }
    for (int idx = lane; idx < dm * dm; idx += 32) {
        const int i = idx % dm, j = idx / dm;
        if (i < j) {
            const double x = 0.5 * (Hc[i + (int64_t)j * lde] + Hc[j + (int64_t)i * lde]);
            Hc[i + (int64_t)j * lde] = x;
            Hc[j + (int64_t)i * lde] = x;
        }
    }
}

static __global__ void __launch_bounds__(128)
etr_prod_kernel(int ncones, int want_dual, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int64_t* __restrict__ voff, const double* __restrict__ vecs, const int* __restrict__ dualf,
                const double* __restrict__ scal, const double* arr, int64_t ld_arr, double* prod, int64_t ld_prod,
                int64_t ncols, int64_t row_shift) {
    __shared__ double sh[4][8 * 104];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    if (want_dual >= 0 && (dualf[c] != 0) != (want_dual != 0)) return;
    const int64_t o = off[c];
    const int vw = (dim[c] - 1) / 2;
    int d = 1;
    while (d * (d + 1) / 2 < vw) d++;
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y)
        etr_hess_col(arr + j * ld_arr + (o - row_shift), prod + j * ld_prod + (o - row_shift), d, vecs + voff[c],
                     scal[8 * c], sh[threadIdx.x >> 5], lane);
}

// dder3 (:269-383): -1/2 of the third directional derivative of the barrier; scratch in global memory
static __global__ void __launch_bounds__(128)
etr_dder3_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                 const int64_t* __restrict__ voff, double* __restrict__ vecs, const double* __restrict__ scal,
                 const double* __restrict__ dir, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int vw = (dim[c] - 1) / 2;
    int d = 1;
    while (d * (d + 1) / 2 < vw) d++;
    const int n = d * d;
    double* st = vecs + voff[c];
    const double* Qv = st;
    const double* Qw = st + n;
    const double* Vi = st + 4 * n;
    const double* Wi = st + 5 * n;
    const double* D1v = st + 6 * n;
    const double* D1w = st + 7 * n;
    const double* zV = st + 8 * n;
    const double* zW = st + 9 * n;
    const double* D2v = st + 10 * n + 2 * d;
    const double* D2w = D2v + d * n;
    const double* D3v = D2w + d * n;
    const double* Wt = D3v + n * n;
    double* ws = st + 10 * n + 2 * d + 2 * d * n + n * n + n;
    double *dV = ws, *dW = ws + n, *hV = ws + 2 * n, *hW = ws + 3 * n, *t1 = ws + 4 * n, *t2 = ws + 5 * n, *dVt = ws + 6 * n,
           *dWt = ws + 7 * n, *tV = ws + 8 * n, *tW = ws + 9 * n, *t3 = ws + 10 * n, *dWw = ws + 11 * n;
    const double z = scal[8 * c], du = dir[o];
    dnn_smat(dV, dir + o + 1, d, vw, lane);
    dnn_smat(dW, dir + o + 1 + vw, d, vw, lane);
    __syncwarp();
    const double dz = du + mep_dot(zV, dV, n, lane) + mep_dot(zW, dW, n, lane);
    etr_in(dVt, Qv, dV, t1, d, lane);          // Qv' dV Qv
    etr_in(dWt, Qv, dW, t1, d, lane);          // Qv' dW Qv
    etr_in(dWw, Qw, dW, t1, d, lane);          // Qw' dW Qw
    // second derivative of z along the direction: hV = D2log(V)[W, dV] + Dlog(V)[dW], hW = Dlog(V)[dV] - Dlog(W)[dW]
    etr_d2t(hV, Qv, D2v, Wt, dVt, t1, t2, d, lane);
    etr_d1(t3, Qv, D1v, dW, t1, t2, d, lane);
    mep_axpby(hV, 1.0, hV, 1.0, t3, n, lane);
    etr_d1(hW, Qv, D1v, dV, t1, t2, d, lane);
    etr_d1(t3, Qw, D1w, dW, t1, t2, d, lane);
    mep_axpby(hW, 1.0, hW, -1.0, t3, n, lane);
    const double dHd = mep_dot(hV, dV, n, lane) + mep_dot(hW, dW, n, lane);
    // third derivative of z: tV = D3log(V)[W, dV, dV] + 2 D2log(V)[dW, dV], tW = D2log(V)[dV, dV] - D2log(W)[dW, dW]
    for (int idx = lane; idx < n; idx += 32) {
        const int i = idx % d, j = idx / d;
        double s = 0.0;
        for (int k = 0; k < d; k++)
            for (int l = 0; l < d; l++) {
                const double w = D3v[i + d * (k + d * (l + d * j))];
                // the six orderings of (W, dV, dV): two of each distinct arrangement
                s += 2.0 * w * (Wt[i + k * d] * dVt[k + l * d] * dVt[l + j * d] + dVt[i + k * d] * Wt[k + l * d] * dVt[l + j * d] +
                                dVt[i + k * d] * dVt[k + l * d] * Wt[l + j * d]);
            }
        t2[idx] = s;
    }
    __syncwarp();
    etr_out(tV, Qv, t2, t1, d, lane);
    etr_d2t(t3, Qv, D2v, dWt, dVt, t1, t2, d, lane);
    mep_axpby(tV, 1.0, tV, 2.0, t3, n, lane);
    etr_d2t(tW, Qv, D2v, dVt, dVt, t1, t2, d, lane);
    etr_d2t(t3, Qw, D2w, dWw, dWw, t1, t2, d, lane);
    mep_axpby(tW, 1.0, tW, -1.0, t3, n, lane);
    // -2 Vi dV Vi dV Vi and -2 Wi dW Wi dW Wi
    mep_mm(t1, Vi, dV, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(t2, t1, Vi, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(dVt, t2, t1, d, d, d, false, true, 1.0, 0.0, lane);      // (Vi dV Vi)(Vi dV)' = Vi dV Vi dV Vi
    mep_mm(t1, Wi, dW, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(t2, t1, Wi, d, d, d, false, false, 1.0, 0.0, lane);
    mep_mm(dWt, t2, t1, d, d, d, false, true, 1.0, 0.0, lane);
    const double c1 = -2.0 * dz * dz / (z * z * z), c2 = 2.0 * dz / (z * z), c3 = dHd / (z * z), zi = 1.0 / z;
    for (int p = lane; p < vw; p += 32) {
        int i, j;
        dnn_ij(p, i, j);
        const int e = i + j * d, et = j + i * d;
        const double xV = (c1 + c3) * zV[e] + c2 * 0.5 * (hV[e] + hV[et]) - zi * 0.5 * (tV[e] + tV[et]) - (dVt[e] + dVt[et]);
        const double xW = (c1 + c3) * zW[e] + c2 * 0.5 * (hW[e] + hW[et]) - zi * 0.5 * (tW[e] + tW[et]) - (dWt[e] + dWt[et]);
        const double fs = i == j ? 1.0 : 1.41421356237309504880;
        out[o + 1 + p] = -0.5 * fs * xV;
        out[o + 1 + vw + p] = -0.5 * fs * xW;
    }
    if (lane == 0) out[o] = -0.5 * (c1 + c3);
}


}  // namespace hypdev

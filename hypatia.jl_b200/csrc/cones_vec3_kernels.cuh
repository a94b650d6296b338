// Device oracles of EpiPerSquare, HypoPerLog (two leading scalars (u, v) and a w block), EpiNormInf,
// HypoGeoMean (one leading scalar u and a w block), EpiPerSepSpectral{VectorCSqr} and EpiRelEntropy
// (u, v block, w block; src/Cones/epirelentropy.jl:91-364); one warp per cone for the state, one warp per (cone, column)
// for the products.
//
// reference: src/Cones/epipersquare.jl:59-274 (update_feas, is_dual_feas, update_grad, hess_prod!,
// inv_hess_prod!, sqrt_hess_prod!, inv_sqrt_hess_prod!, dder3), src/Cones/hypoperlog.jl:62-287
// (update_feas, is_dual_feas, update_grad, hess_prod!, inv_hess_prod!, dder3), src/Cones/epinorminf.jl:97-406
// (real case: update_feas, is_dual_feas, update_grad, update_hess_aux, hess_prod!, update_inv_hess_aux,
// inv_hess_prod!, dder3; the arrow-shaped Hessian is applied from u and w, nothing is stored per entry).
// scal (8 doubles per cone): EpiPerSquare 0 dist; HypoPerLog 1 phi, 2 zeta; EpiNormInf 0 Huu, 1 schur.
// HBM-bound streaming kernels: 16 * rows * ncols algorithmic bytes per product.
#pragma once
#include "devdefs.cuh"
#include "ssf.cuh"

#define V3_EPIPERSQUARE 6   // = HYP_CONE_EPIPERSQUARE
#define V3_HYPOPERLOG 7     // = HYP_CONE_HYPOPERLOG
#define V3_EPINORMINF 8     // = HYP_CONE_EPINORMINF
#define V3_HYPOGEOMEAN 10   // = HYP_CONE_HYPOGEOMEAN (hypogeomean.jl; one leading scalar; scal: 1 phi, 2 zeta, 5 phi / zeta / d)
#define V3_EPIRELENTROPY 13 // = HYP_CONE_EPIRELENTROPY (epirelentropy.jl; (u, v[d], w[d]); scal: 0 z, 1 Hiuu)
#define V3_SEPSPEC_VEC 9    // = HYP_CONE_EPIPERSEPSPECTRAL_VEC (vectorcsqr.jl; scal: 0 phi, 1 zeta, 2 sigma, 3 c0, 4 c4, 5 c5)
// product modes (= HYP_PROD_*)
#define V3_HESS 0
#define V3_INV_HESS 1
#define V3_SQRT_HESS 2
#define V3_INV_SQRT_HESS 3

namespace hypdev {

static __global__ void __launch_bounds__(256)
v3_state_kernel(int type, int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int* __restrict__ kidx, const int* __restrict__ hkind,
                const double* __restrict__ hparam, const double* __restrict__ point,
                const double* __restrict__ dual, double* __restrict__ grad, double* __restrict__ scal,
                uint8_t* feas, uint8_t* dual_feas) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;   // whole warps leave together
    const int64_t o = off[c];
    const int d = dim[c];
    const double u = point[o], v = point[o + 1], du = dual[o], dv = dual[o + 1];
    if (type == V3_EPIRELENTROPY) {
        // epirelentropy.jl:91-140 (feas, dual feas, grad), :188-222 (Hiuu of the inverse Hessian)
        const int n = (d - 1) / 2;
        double nbad = 0.0, ent = 0.0, dbad = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
            if (!(vi > HYP_EPS) || !(wi > HYP_EPS)) nbad += 1.0;
            ent += wi * log(wi / vi);
            const double dvi = dual[o + 1 + i], dwi = dual[o + 1 + n + i];
            if (!(dvi > HYP_EPS) || !(du * (1.0 + log(dvi / du)) + dwi > HYP_EPS)) dbad += 1.0;
        }
        nbad = warp_sum(nbad);
        ent = warp_sum(ent);
        dbad = warp_sum(dbad);
        const double z = u - ent;
        const bool ok = nbad == 0.0 && z > HYP_EPS;
        const bool dok = du > HYP_EPS && dbad == 0.0;
        double hh = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
            const double lwv = log(wi / vi);
            grad[o + 1 + i] = -wi / vi / z - 1.0 / vi;
            grad[o + 1 + n + i] = (lwv + 1.0) / z - 1.0 / wi;
            const double zw = z + wi, z2w = zw + wi, wz2w = wi / z2w;
            const double uvv = wi * (wi * lwv - z), uww = wi * (z + lwv * zw) * wz2w;
            hh += wz2w * uvv - uww * (lwv + 1.0);
        }
        hh = warp_sum(hh);
        if (lane == 0) {
            grad[o] = -1.0 / z;
            scal[8 * c] = z;
            scal[8 * c + 1] = z * z - hh;
            if (!ok) feas[kidx[c]] = 0;
            if (!dok) dual_feas[kidx[c]] = 0;
        }
    } else if (type == V3_HYPOGEOMEAN) {
        // hypogeomean.jl:69-110
        const int dw = d - 1;
        double nbad = 0.0, dbad = 0.0, sl = 0.0, dsl = 0.0;
        for (int i = 1 + lane; i < d; i += 32) {
            const double w = point[o + i], zd = dual[o + i];
            if (!(w > HYP_EPS)) nbad += 1.0;
            if (!(zd > HYP_EPS)) dbad += 1.0;
            sl += log(w);
            dsl += log(zd);
        }
        nbad = warp_sum(nbad);
        dbad = warp_sum(dbad);
        sl = warp_sum(sl);
        dsl = warp_sum(dsl);
        const double phi = exp(sl / dw), zeta = phi - u;
        const bool ok = nbad == 0.0 && zeta > HYP_EPS;
        const bool dok = du < -HYP_EPS && dbad == 0.0 && (dw * exp(dsl / dw) + du) > HYP_EPS;
        const double pzd = phi / zeta / dw;
        for (int i = 1 + lane; i < d; i += 32) grad[o + i] = (-pzd - 1.0) / point[o + i];
        if (lane == 0) {
            grad[o] = 1.0 / zeta;
            scal[8 * c + 1] = phi;
            scal[8 * c + 2] = zeta;
            scal[8 * c + 5] = pzd;
            if (!ok) feas[kidx[c]] = 0;
            if (!dok) dual_feas[kidx[c]] = 0;
        }
    } else if (type == V3_SEPSPEC_VEC) {
        // vectorcsqr.jl:61-114 (feas, dual feas, grad), :216-246 (inverse-Hessian scalars)
        const int kind = hkind[c];
        const double hp = hparam[c];
        double nbad = 0.0, phi = 0.0, s1 = 0.0, dbad = 0.0, conj = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], zd = dual[o + i];
            if (!(w > HYP_EPS)) nbad += 1.0;
            if (zd < HYP_EPS) dbad += 1.0;
            double h, a1, a2, a3;
            ssf_eval(kind, hp, w / v, h, a1, a2, a3);
            phi += h;
            s1 += (w / v) * a1;
            conj += ssf_conj(kind, hp, zd / du);
        }
        nbad = warp_sum(nbad);
        phi = warp_sum(phi);
        s1 = warp_sum(s1);
        dbad = warp_sum(dbad);
        conj = warp_sum(conj);
        const double zeta = u - v * phi, sigma = phi - s1, zetai = 1.0 / zeta, zetaivi = zetai / v;
        const bool ok = v > HYP_EPS && nbad == 0.0 && zeta > HYP_EPS;
        bool dok = !(du < HYP_EPS);
        if (ssf_conj_dom_pos(kind) && dbad > 0.0) dok = false;
        if (dok) dok = (dv - du * conj) > HYP_EPS;
        double r1 = 0.0, r2 = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], viw = w / v;
            double h, a1, a2, a3;
            ssf_eval(kind, hp, viw, h, a1, a2, a3);
            grad[o + i] = -1.0 / w + zetai * a1;
            const double w1 = zetaivi * a2, m = 1.0 / (w1 + 1.0 / (w * w));
            r1 += a1 * m * a1;               // dot(dh, alpha)
            r2 += a1 * m * w1 * viw;         // dot(dh, gamma)
        }
        r1 = warp_sum(r1);
        r2 = warp_sum(r2);
        const double zeta2beta = zeta * zeta + r1, c0 = sigma + r2, c1 = c0 / zeta2beta;
        double r3 = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], viw = w / v;
            double h, a1, a2, a3;
            ssf_eval(kind, hp, viw, h, a1, a2, a3);
            const double w1 = zetaivi * a2, m = 1.0 / (w1 + 1.0 / (w * w)), w1v = w1 * viw;
            r3 += (viw + c1 * m * a1 - m * w1v) * w1v;
        }
        r3 = warp_sum(r3);
        const double c3 = 1.0 / (v * v) + sigma * c1 + r3;
        if (lane == 0) {
            grad[o] = -zetai;
            grad[o + 1] = -1.0 / v + zetai * sigma;
            double* sc = scal + 8 * c;
            sc[0] = phi;
            sc[1] = zeta;
            sc[2] = sigma;
            sc[3] = c0;
            sc[4] = 1.0 / (c3 - c0 * c1);
            sc[5] = zeta2beta * c3;
            if (!ok) feas[kidx[c]] = 0;
            if (!dok) dual_feas[kidx[c]] = 0;
        }
    } else if (type == V3_EPINORMINF) {
        // epinorminf.jl:97-142, :144-168 (Huu), :275-300 (schur)
        const int n = d - 1;
        double wmax = 0.0, dsum = 0.0, sud = 0.0, sud2 = 0.0, sinv = 0.0;
        const double usqr = u * u;
        for (int i = 1 + lane; i < d; i += 32) {
            const double w = point[o + i];
            wmax = fmax(wmax, fabs(w));
            dsum += fabs(dual[o + i]);
            const double den = 0.5 * (usqr - w * w);
            const double uden = u / den;
            sud += uden;
            sud2 += uden * uden;
            sinv += 1.0 / (0.5 * (usqr + w * w));
            grad[o + i] = w / den;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, s));
        dsum = warp_sum(dsum);
        sud = warp_sum(sud);
        sud2 = warp_sum(sud2);
        sinv = warp_sum(sinv);
        if (lane == 0) {
            grad[o] = (n - 1) / u - sud;
            scal[8 * c] = sud2 - ((n - 1) / u + sud) / u;
            scal[8 * c + 1] = (1 - n) / usqr + sinv;
            if (!(u > HYP_EPS && u - wmax > HYP_EPS)) feas[kidx[c]] = 0;
            if (!(du > HYP_EPS && du - dsum > HYP_EPS)) dual_feas[kidx[c]] = 0;
        }
    } else if (type == V3_EPIPERSQUARE) {
        // epipersquare.jl:59-100
        double sw = 0.0, sdw = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], dw = dual[o + i];
            sw += w * w;
            sdw += dw * dw;
        }
        sw = warp_sum(sw);
        sdw = warp_sum(sdw);
        double dist = 0.0;
        bool ok = false;
        if (u > HYP_EPS && v > HYP_EPS) {
            dist = u * v - sw / 2;
            ok = dist > HYP_EPS;
        }
        const bool dok = (du > HYP_EPS && dv > HYP_EPS) && ((du * dv - sdw / 2) > HYP_EPS);
        for (int i = lane; i < d; i += 32) {
            double g;
            if (i == 0) g = -v / dist;
            else if (i == 1) g = -u / dist;
            else g = point[o + i] / dist;
            grad[o + i] = g;
        }
        if (lane == 0) {
            scal[8 * c] = dist;
            if (!ok) feas[kidx[c]] = 0;
            if (!dok) dual_feas[kidx[c]] = 0;
        }
    } else {
        // hypoperlog.jl:62-113
        const int dw = d - 2;
        double nbad = 0.0, dbad = 0.0, phi = 0.0, sumlog = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], z = dual[o + i];
            if (!(w > HYP_EPS)) nbad += 1.0;
            if (!(z > HYP_EPS)) dbad += 1.0;
            phi += log(w / v);
            sumlog += log(z / -du);
        }
        nbad = warp_sum(nbad);
        dbad = warp_sum(dbad);
        phi = warp_sum(phi);
        sumlog = warp_sum(sumlog);
        const double zeta = v * phi - u;
        const bool ok = v > HYP_EPS && nbad == 0.0 && zeta > HYP_EPS;
        const bool dok = dbad == 0.0 && du < -HYP_EPS && (dv - du * (sumlog + dw)) > HYP_EPS;
        const double vzi1 = -1.0 - v / zeta;
        for (int i = lane; i < d; i += 32) {
            double g;
            if (i == 0) g = 1.0 / zeta;
            else if (i == 1) g = -(phi - dw) / zeta - 1.0 / v;
            else g = vzi1 / point[o + i];
            grad[o + i] = g;
        }
        if (lane == 0) {
            scal[8 * c + 1] = phi;
            scal[8 * c + 2] = zeta;
            if (!ok) feas[kidx[c]] = 0;
            if (!dok) dual_feas[kidx[c]] = 0;
        }
    }
}

// prod[:, j] = oracle(arr[:, j]) on the rows of every cone of the group.  `dualf` (may be null):
// per-cone use_dual_barrier flag, only read when mode is one of the block modes (4: hess for primal
// / inv_hess for dual-barrier cones, 5: the other way round).
static __global__ void __launch_bounds__(256)
v3_prod_kernel(int type, int mode_in, int ncones, const int64_t* __restrict__ off,
               const int* __restrict__ dim, const int* __restrict__ dualf, const int* __restrict__ hkind,
               const double* __restrict__ hparam,
               const double* __restrict__ scal, const double* __restrict__ point, const double* arr,
               int64_t ld_arr, double* prod, int64_t ld_prod, int64_t ncols, int64_t row_shift) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    int mode = mode_in;
    if (mode == 4) mode = (dualf && dualf[c]) ? V3_INV_HESS : V3_HESS;
    if (mode == 5) mode = (dualf && dualf[c]) ? V3_HESS : V3_INV_HESS;
    const double u = point[o], v = point[o + 1];
    for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) {
        const double* a = arr + j * ld_arr + (o - row_shift);
        double* pr = prod + j * ld_prod + (o - row_shift);
        const double p = a[0], q = a[1];
        // in-place use (prod == arr): every lane holds its copy of the two leading entries before any
        // lane stores to them (found by the ThreadSanitizer run of the emulation, tools/emu_tsan.sh)
        __syncwarp();
        if (type == V3_EPIRELENTROPY) {
            const int n = (d - 1) / 2;
            const double z = scal[8 * c];
            const double* av = a + 1;
            const double* aw = a + 1 + n;
            if (mode == V3_HESS) {
                // epirelentropy.jl:260-293
                double sr = 0.0;
                for (int i = lane; i < n; i += 32) {
                    const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
                    sr += av[i] * (wi / vi / z) - aw[i] * ((log(wi / vi) + 1.0) / z);
                }
                const double up = warp_sum(sr) + p / z;
                for (int i = lane; i < n; i += 32) {
                    const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
                    const double sigma = wi / vi / z, tau = -(log(wi / vi) + 1.0) / z;
                    const double avi = av[i], awi = aw[i];
                    pr[1 + i] = sigma * up + (sigma * avi + avi / vi - awi / z) / vi;
                    pr[1 + n + i] = tau * up + (awi / z + awi / wi) / wi - avi / vi / z;
                }
                if (lane == 0) pr[0] = up / z;
            } else {
                // epirelentropy.jl:295-321 with the coefficients of :188-222 recomputed per entry
                const double Hiuu = scal[8 * c + 1];
                double sr = 0.0;
                for (int i = lane; i < n; i += 32) {
                    const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
                    const double lwv = log(wi / vi);
                    const double zw = z + wi, z2w = zw + wi, wz2w = wi / z2w, vz2w = vi / z2w;
                    const double uvv = wi * (wi * lwv - z);
                    const double Hiuv = vz2w * uvv, Hiuw = wi * (z + lwv * zw) * wz2w;
                    const double Hivw = wi * vi * wz2w, Hiww = wi * zw * wz2w, Hivv = vi * zw * vz2w;
                    const double avi = av[i], awi = aw[i];
                    sr += avi * Hiuv + awi * Hiuw;
                    pr[1 + i] = Hiuv * p + Hivv * avi + Hivw * awi;
                    pr[1 + n + i] = Hiuw * p + Hivw * avi + Hiww * awi;
                }
                sr = warp_sum(sr);
                if (lane == 0) pr[0] = Hiuu * p + sr;
            }
        } else if (type == V3_HYPOGEOMEAN) {
            const double phi = scal[8 * c + 1], zeta = scal[8 * c + 2], pzd = scal[8 * c + 5];
            const double di = 1.0 / (double)(d - 1);
            if (mode == V3_HESS) {
                // hypogeomean.jl:141-167
                double sr = 0.0;
                for (int i = 1 + lane; i < d; i += 32) sr += a[i] / point[o + i];
                const double c0 = pzd * warp_sum(sr);
                const double c1 = c0 - p / zeta;
                const double c2 = pzd * c1 - di * c0;
                for (int i = 1 + lane; i < d; i += 32) {
                    const double w = point[o + i];
                    pr[i] = (c2 + (pzd + 1.0) * (a[i] / w)) / w;
                }
                if (lane == 0) pr[0] = c1 / -zeta;
            } else {
                // hypogeomean.jl:202-230
                const double phidi = phi * di, c2 = 1.0 / (pzd + 1.0), c3 = c2 / zeta * di;
                const double c4 = zeta * zeta + phidi * phi;
                double sr = 0.0;
                for (int i = 1 + lane; i < d; i += 32) sr += a[i] * point[o + i];
                const double c5 = warp_sum(sr);
                const double c6 = phidi * (c3 * c5 + p);
                for (int i = 1 + lane; i < d; i += 32) {
                    const double w = point[o + i];
                    pr[i] = (c6 + c2 * (a[i] * w)) * w;
                }
                if (lane == 0) pr[0] = phidi * c5 + c4 * p;
            }
        } else if (type == V3_SEPSPEC_VEC) {
            const int kind = hkind[c];
            const double hp = hparam[c];
            const double* sc = scal + 8 * c;
            const double zeta = sc[1], sigma = sc[2], zetai = 1.0 / zeta, zetaivi = zetai / v;
            if (mode == V3_HESS) {
                // vectorcsqr.jl:172-204
                const double viq = q / v;
                double s1 = 0.0, s2 = 0.0;
                for (int i = 2 + lane; i < d; i += 32) {
                    const double w = point[o + i], ri = a[i];
                    double h, a1, a2, a3;
                    ssf_eval(kind, hp, w / v, h, a1, a2, a3);
                    s1 += a1 * ri;
                    s2 += (w / v) * zetaivi * a2 * (ri - viq * w);
                }
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                const double c1 = -zetai * (p - sigma * q - s1) * zetai;
                for (int i = 2 + lane; i < d; i += 32) {
                    const double w = point[o + i], ri = a[i];
                    double h, a1, a2, a3;
                    ssf_eval(kind, hp, w / v, h, a1, a2, a3);
                    pr[i] = c1 * a1 + zetaivi * a2 * (ri - viq * w) + ri / (w * w);
                }
                if (lane == 0) {
                    pr[0] = -c1;
                    pr[1] = c1 * sigma - s2 + viq / v;
                }
            } else {
                // vectorcsqr.jl:275-305
                const double c0 = sc[3], c4 = sc[4], c5 = sc[5];
                double s1 = 0.0, s2 = 0.0;
                for (int i = 2 + lane; i < d; i += 32) {
                    const double w = point[o + i], ri = a[i];
                    double h, a1, a2, a3;
                    ssf_eval(kind, hp, w / v, h, a1, a2, a3);
                    const double w1 = zetaivi * a2, m = 1.0 / (w1 + 1.0 / (w * w));
                    s1 += m * w1 * (w / v) * ri;     // dot(gamma, r)
                    s2 += m * a1 * ri;               // dot(alpha, r)
                }
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                const double qgr = q + s1;
                const double cu = c4 * (c5 * p + c0 * qgr), cv = c4 * (c0 * p + qgr);
                for (int i = 2 + lane; i < d; i += 32) {
                    const double w = point[o + i], ri = a[i];
                    double h, a1, a2, a3;
                    ssf_eval(kind, hp, w / v, h, a1, a2, a3);
                    const double w1 = zetaivi * a2, m = 1.0 / (w1 + 1.0 / (w * w));
                    pr[i] = p * m * a1 + cv * m * w1 * (w / v) + m * ri;
                }
                if (lane == 0) {
                    pr[0] = cu + s2;
                    pr[1] = cv;
                }
            }
        } else if (type == V3_EPINORMINF) {
            const double usqr = u * u, ua = p;
            if (mode == V3_HESS) {
                // epinorminf.jl:226-244: prod_u = Huu ua + Hure . wa; prod_w = Hure ua + Hrere wa
                double dot = 0.0;
                for (int i = 1 + lane; i < d; i += 32) {
                    const double w = point[o + i], den = 0.5 * (usqr - w * w);
                    const double hure = -(w / den) * (u / den);
                    const double ai = a[i];
                    dot += hure * ai;
                    pr[i] = hure * ua + ((w / den) * (w / den) + 1.0 / den) * ai;
                }
                dot = warp_sum(dot);
                if (lane == 0) pr[0] = scal[8 * c] * ua + dot;
            } else {
                // epinorminf.jl:347-364: prod_u = (ua + Hiure . wa) / schur; prod_w = Hiure prod_u + wa / Hrere
                double dot = 0.0;
                for (int i = 1 + lane; i < d; i += 32) {
                    const double w = point[o + i];
                    dot += u / (0.5 * (usqr + w * w)) * w * a[i];
                }
                const double pu = (ua + warp_sum(dot)) / scal[8 * c + 1];
                for (int i = 1 + lane; i < d; i += 32) {
                    const double w = point[o + i], den = 0.5 * (usqr - w * w);
                    const double hrere = (w / den) * (w / den) + 1.0 / den;
                    pr[i] = u / (0.5 * (usqr + w * w)) * w * pu + a[i] / hrere;
                }
                if (lane == 0) pr[0] = pu;
            }
        } else if (type == V3_EPIPERSQUARE) {
            // every oracle is  coef * vec + kap * J a  with J a = (-a_2, -a_1, a_w):
            //   vec = J point / point / sqrt-vectors of epipersquare.jl:156-191, coef from one dot product
            const double dist = scal[8 * c];
            const double rtdist = sqrt(dist), denom = 2 * rtdist + u + v;
            double v0, v1, sc, kap;   // vec = (v0, v1, sc * w)
            if (mode == V3_HESS) {
                v0 = -v; v1 = -u; sc = 1.0; kap = 1.0 / dist;
            } else if (mode == V3_INV_HESS) {
                v0 = u; v1 = v; sc = 1.0; kap = dist;
            } else if (mode == V3_SQRT_HESS) {
                v0 = -v / rtdist - 1.0; v1 = -u / rtdist - 1.0; sc = 1.0 / rtdist; kap = 1.0 / rtdist;
            } else {
                v0 = u + rtdist; v1 = v + rtdist; sc = 1.0; kap = rtdist;
            }
            double dot = 0.0;
            for (int i = 2 + lane; i < d; i += 32) dot += point[o + i] * a[i];
            dot = sc * warp_sum(dot) + v0 * p + v1 * q;
            double coef;
            if (mode == V3_HESS) coef = dot / (dist * dist);
            else if (mode == V3_INV_HESS) coef = dot;
            else coef = dot / denom;
            for (int i = lane; i < d; i += 32) {
                double x;
                if (i == 0) x = coef * v0 - kap * q;
                else if (i == 1) x = coef * v1 - kap * p;
                else x = coef * sc * point[o + i] + kap * a[i];
                pr[i] = x;
            }
        } else {
            const double phi = scal[8 * c + 1], zeta = scal[8 * c + 2];
            const double dd = (double)(d - 2);
            if (mode == V3_HESS) {
                // hypoperlog.jl:153-183
                const double sigma = phi - dd, vzi1 = v / zeta + 1.0;
                double s = 0.0;
                for (int i = 2 + lane; i < d; i += 32) s += a[i] / point[o + i];
                s = warp_sum(s);
                const double qzi = q / zeta, c0 = s / zeta;
                const double c1 = (v * c0 - p / zeta + sigma * qzi) / zeta;
                const double c3 = c1 * v - qzi;
                for (int i = lane; i < d; i += 32) {
                    double x;
                    if (i == 0) x = -c1;
                    else if (i == 1) x = c1 * sigma - c0 + (qzi * dd + q / v) / v;
                    else {
                        const double w = point[o + i];
                        x = (c3 + vzi1 * (a[i] / w)) / w;
                    }
                    pr[i] = x;
                }
            } else {
                // hypoperlog.jl:221-257
                const double zv = zeta + v, zzvi = zeta / zv;
                const double c3 = v / (zv + dd * v);
                const double c0 = phi - dd * zzvi;
                const double c4 = v * c3 * zv;
                const double t = zeta + v * phi;
                const double c6 = (v * phi) * (v * phi) + zeta * (zeta + dd * v) - dd * t * t * c3;
                const double c7 = c4 * c0, c8 = c7 + v * zeta;
                double s = 0.0;
                for (int i = 2 + lane; i < d; i += 32) s += a[i] * point[o + i];
                s = warp_sum(s);
                const double c1 = s / zv;
                const double c5 = c0 * p + q + c1;
                const double c2 = v * (zzvi * p + c3 * c5);
                for (int i = lane; i < d; i += 32) {
                    double x;
                    if (i == 0) x = c6 * p + c7 * q + c8 * c1;
                    else if (i == 1) x = c4 * c5;
                    else {
                        const double w = point[o + i];
                        x = (c2 + zzvi * (a[i] * w)) * w;
                    }
                    pr[i] = x;
                }
            }
        }
    }
}

// epipersquare.jl:246-274, hypoperlog.jl:259-287
static __global__ void __launch_bounds__(256)
v3_dder3_kernel(int type, int ncones, const int64_t* __restrict__ off, const int* __restrict__ dim,
                const int* __restrict__ hkind, const double* __restrict__ hparam,
                const double* __restrict__ scal, const double* __restrict__ point,
                const double* __restrict__ dir, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= ncones) return;
    const int64_t o = off[c];
    const int d = dim[c];
    const double u = point[o], v = point[o + 1], p = dir[o], q = dir[o + 1];
    if (type == V3_EPIRELENTROPY) {
        // epirelentropy.jl:323-364
        const int n = (d - 1) / 2;
        const double z = scal[8 * c], i2z = 1.0 / (2.0 * z);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
            const double dvi = dir[o + 1 + i], dwi = dir[o + 1 + n + i];
            const double vdv = dvi / vi, wdw = dwi / wi, tau = -(log(wi / vi) + 1.0) / z;
            s0 += wi * vdv;
            s1 += tau * dwi;
            s2 += wi * vdv * vdv + dwi * (wdw - 2.0 * vdv);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        const double const0 = (p + s0) / z + s1;
        const double const1 = const0 * const0 + s2 * i2z;
        for (int i = lane; i < n; i += 32) {
            const double vi = point[o + 1 + i], wi = point[o + 1 + n + i];
            const double dvi = dir[o + 1 + i], dwi = dir[o + 1 + n + i];
            const double vdv = dvi / vi, wdw = dwi / wi, tau = -(log(wi / vi) + 1.0) / z;
            double t = const1 + (const0 + vdv) * vdv - i2z * wdw * dwi;
            t = t * wi + (z * vdv - dwi) * vdv + (-const0 + i2z * dwi) * dwi;
            out[o + 1 + i] = t / vi / z;
            out[o + 1 + n + i] = const1 * tau + ((const0 - wi * vdv / z) / z + (1.0 / wi + i2z) * wdw) * wdw +
                                 (-const0 + dwi / z - vdv / 2.0) / z * vdv;
        }
        if (lane == 0) out[o] = const1 / z;
    } else if (type == V3_HYPOGEOMEAN) {
        // hypogeomean.jl:232-257
        const double phi = scal[8 * c + 1], zeta = scal[8 * c + 2], pzd = scal[8 * c + 5];
        const double di = 1.0 / (double)(d - 1);
        double s0 = 0.0, s6 = 0.0;
        for (int i = 1 + lane; i < d; i += 32) {
            const double r = dir[o + i] / point[o + i];
            s0 += r;
            s6 += r * r;
        }
        const double c0 = warp_sum(s0) * di, c6 = warp_sum(s6) * di;
        const double zichi = (p - phi * c0) / zeta;
        const double c1 = zichi * zichi + phi / zeta * (c6 - c0 * c0) / 2;
        const double c7 = pzd * (c1 - c6 / 2 + c0 * (zichi + c0 / 2));
        const double c8 = -pzd * (zichi + c0), c9 = pzd + 1.0;
        for (int i = 1 + lane; i < d; i += 32) {
            const double w = point[o + i], r = dir[o + i] / w;
            out[o + i] = (c7 + r * (c8 + c9 * r)) / w;
        }
        if (lane == 0) out[o] = c1 / -zeta;
    } else if (type == V3_SEPSPEC_VEC) {
        // vectorcsqr.jl:314-357
        const int kind = hkind[c];
        const double hp = hparam[c];
        const double* sc = scal + 8 * c;
        const double zeta = sc[1], sigma = sc[2], zetai = 1.0 / zeta, zetaivi = zetai / v;
        const double viq = q / v;
        double s1 = 0.0, s2 = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], ri = dir[o + i];
            double h, a1, a2, a3;
            ssf_eval(kind, hp, w / v, h, a1, a2, a3);
            const double xi = ri - viq * w;
            s1 += a1 * ri;
            s2 += zetaivi * a2 * xi * xi;
        }
        s1 = warp_sum(s1);
        const double xibxi = warp_sum(s2) / 2;
        const double zetaichi = zetai * (p - sigma * q - s1);
        const double c1 = -zetai * (zetaichi * zetaichi + xibxi), c2 = -zetai / 2;
        double s3 = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], ri = dir[o + i];
            double h, a1, a2, a3;
            ssf_eval(kind, hp, w / v, h, a1, a2, a3);
            const double xi = ri - viq * w, xiv = xi / v;
            const double waux = zetaivi * a2 * xi * (zetaichi + viq) + c2 * a3 * xiv * xiv;
            s3 += (w / v) * waux;
            const double rw = ri / w;
            out[o + i] = c1 * a1 + waux + rw * rw / w;
        }
        s3 = warp_sum(s3);
        if (lane == 0) {
            out[o] = -c1;
            out[o + 1] = c1 * sigma - s3 + (xibxi + viq * viq) / v;
        }
    } else if (type == V3_EPINORMINF) {
        // epinorminf.jl:366-406 (real case)
        const int n = d - 1;
        const double usqr = u * u, udir = p, u3 = 1.5 / u, udu = udir / u;
        double s0 = 0.0, s1 = 0.0;
        for (int i = 1 + lane; i < d; i += 32) {
            const double w = point[o + i], di = dir[o + i];
            const double den = 0.5 * (usqr - w * w), z = u / den;
            s0 += z * (u3 - z) * z;
            const double deni = -4.0 * den, udeni = 2.0 * z, wdeni = 2.0 * w / den;
            const double suuw = udir * (-1.0 + udeni * u);
            const double uuw = suuw * wdeni;
            const double uimim = 1.0 + wdeni * w;
            const double uimim2 = -udeni * uimim * di;
            s1 += di * (2.0 * uuw + uimim2) / deni;
            out[o + i] = (udir * (uuw + 2.0 * uimim2) + di * wdeni * (2.0 + uimim) * di) / deni;
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) out[o] = -udir * s0 * udir - udu * (n - 1) / u * udu + s1;
    } else if (type == V3_EPIPERSQUARE) {
        const double dist = scal[8 * c];
        double sww = 0.0, swd = 0.0, sdd = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double w = point[o + i], wd = dir[o + i];
            sww += w * w;
            swd += w * wd;
            sdd += wd * wd;
        }
        sww = warp_sum(sww);
        swd = warp_sum(swd);
        sdd = warp_sum(sdd);
        const double jdotpd = u * q + v * p - swd;
        const double ga = (swd - v * p - u * q) / dist;
        const double h0 = (-ga * v - q) / dist, h1 = (-ga * u - p) / dist;   // (H dir)[0], [1]
        // (H dir)[i] = (ga w_i + wd_i) / dist
        const double dHd = p * h0 + q * h1 + (ga * swd + sdd) / dist;
        const double pHd = u * h0 + v * h1 + (ga * sww + swd) / dist;
        const double dotdHd = -dHd, dotpHd = pHd;
        const double inv2d = 1.0 / (2 * dist);
        for (int i = lane; i < d; i += 32) {
            double r;
            if (i == 0) r = h0 * jdotpd - dotdHd * v - dotpHd * q;
            else if (i == 1) r = h1 * jdotpd - dotdHd * u - dotpHd * p;
            else {
                const double w = point[o + i], wd = dir[o + i];
                r = (ga * w + wd) / dist * jdotpd + dotdHd * w + dotpHd * wd;
            }
            out[o + i] = r * inv2d;
        }
    } else {
        const double phi = scal[8 * c + 1], zeta = scal[8 * c + 2];
        const double dd = (double)(d - 2);
        const double sigma = phi - dd, viq = q / v, viq2 = viq * viq, vzi = v / zeta, vzi1 = vzi + 1.0;
        double c0 = 0.0, c7 = 0.0;
        for (int i = 2 + lane; i < d; i += 32) {
            const double r = dir[o + i] / point[o + i];
            c0 += r;
            c7 += r * r;
        }
        c0 = warp_sum(c0);
        c7 = warp_sum(c7);
        const double zichi = (-p + sigma * q + c0 * v) / zeta;
        const double c4 = (viq * (-viq * dd + 2 * c0) - c7) / zeta / 2;
        const double c1 = (zichi * zichi - v * c4) / zeta;
        const double c3 = -(zichi + viq) / zeta;
        const double c5 = c3 * q + vzi * viq2;
        const double c6 = -2 * vzi * viq - c3 * v;
        const double c8 = c5 + c1 * v;
        for (int i = lane; i < d; i += 32) {
            double x;
            if (i == 0) x = -c1;
            else if (i == 1) x = c1 * sigma + (viq2 - (dd * c5 + c6 * c0 + vzi * c7)) / v - c4;
            else {
                const double w = point[o + i], r = dir[o + i] / w;
                x = (c8 + r * (c6 + vzi1 * r)) / w;
            }
            out[o + i] = x;
        }
    }
}

}  // namespace hypdev

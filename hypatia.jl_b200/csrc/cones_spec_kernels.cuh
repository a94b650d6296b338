// Device oracles of EpiPerSepSpectral{MatrixCSqr} (real symmetric case): points (u, v, svec W),
// barrier -log(u - v sum_i h(lambda_i(W / v))) - log v - logdet W, nu = 2 + d.
//
// reference: src/Cones/epipersepspectral/matrixcsqr.jl:91-564, sepspectralfun.jl:17-116.
// State per cone (after the batched eigensolver, eig_kernels.cuh): eigenvectors V of W / v and
//   vecs  (8 arrays of d doubles): 0 lambda, 1 h'(lambda), 2 h'', 3 h''', 4 1/(v lambda), 5 alpha,
//         6 gamma, 7 eigenvalues of dual W / dual u (dual feasibility only)
//   scal  (8 doubles): 0 phi, 1 zeta, 2 sigma, 3 u, 4 v, 5 c0, 6 c4, 7 c5   (matrixcsqr.jl:321-359)
//   Dh    first divided differences of h' (matrixcsqr.jl:167-217), theta = zeta^-1 v^-1 Dh + w w'
// Every Hessian-type product is  svec(M) -> R = V' M V (congruence, tensor GEMMs)
//   -> an elementwise (Hadamard) transform of R with theta plus rank-one scalar terms (this file)
//   -> V (.) V' (congruence) -> svec.
#pragma once
#include "devdefs.cuh"
#include "cones_mat_kernels.cuh"
#include "ssf.cuh"

namespace hypdev {

__device__ __forceinline__ double spec_lam_diff(double la, double lb) {
    const double t = la - lb;
    return fabs(t) < HYP_RTEPS ? 0.0 : t;   // matrixcsqr.jl:180-193
}

// update_feas (after the Cholesky gate and the eigendecomposition), update_grad, update_hess_aux,
// update_inv_hess_aux: matrixcsqr.jl:91-115, :140-217, :321-359.  One CTA per cone.
static __global__ void __launch_bounds__(256)
spec_post_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                 const int64_t* __restrict__ moff, const int64_t* __restrict__ voff,
                 const int* __restrict__ kidx, const int* __restrict__ hkind,
                 const double* __restrict__ hparam, const double* __restrict__ point,
                 const double* __restrict__ V, double* __restrict__ Vt, double* __restrict__ theta,
                 double* __restrict__ Dh, double* __restrict__ vecs, double* __restrict__ scal,
                 double* __restrict__ grad, uint8_t* __restrict__ feas) {
    __shared__ double sm[8];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int d = sides[c], lde = (d + 1) & ~1;
    const int64_t o = off[c], mo = moff[c];
    const int kind = hkind[c];
    const double hp = hparam[c];
    double* lam = vecs + voff[c];
    double* dh = lam + d;
    double* d2h = dh + d;
    double* d3h = d2h + d;
    double* wli = d3h + d;
    double* alpha = wli + d;
    double* gamma = alpha + d;
    const double u = point[o], v = point[o + 1];
    bool ok = v > HYP_EPS;
    double phi = 0.0, s1 = 0.0, nbad = 0.0;
    for (int i = tid; i < d; i += nt) {
        const double l = lam[i];
        if (!(l > HYP_EPS)) nbad += 1.0;
        double h, a1, a2, a3;
        ssf_eval(kind, hp, l, h, a1, a2, a3);
        dh[i] = a1;
        d2h[i] = a2;
        d3h[i] = a3;
        wli[i] = 1.0 / (v * l);
        phi += h;
        s1 += l * a1;
    }
    phi = block_sum(phi, sm);
    s1 = block_sum(s1, sm);
    nbad = block_sum(nbad, sm);
    const double zeta = u - v * phi;
    ok = ok && nbad == 0.0 && zeta > HYP_EPS;
    const double zetai = 1.0 / zeta, sigma = phi - s1, zetaivi = zetai / v;
    const double* Vc = V + mo;
    double* Vtc = Vt + mo;
    double* Dhc = Dh + mo;
    double* thc = theta + mo;
    for (int idx = tid; idx < d * d; idx += nt) {
        const int a = idx % d, b = idx / d;
        Vtc[a + (int64_t)b * lde] = Vc[b + (int64_t)a * lde];
        double dd;
        if (a == b) {
            dd = d2h[a];
        } else {
            const double t = spec_lam_diff(lam[a], lam[b]);
            dd = (t == 0.0) ? 0.5 * (d2h[a] + d2h[b]) : (dh[a] - dh[b]) / t;
        }
        Dhc[a + (int64_t)b * lde] = dd;
        thc[a + (int64_t)b * lde] = zetaivi * dd + wli[a] * wli[b];
    }
    __syncthreads();
    // inverse-Hessian scalars (matrixcsqr.jl:321-359)
    double r1 = 0.0, r2 = 0.0;
    for (int i = tid; i < d; i += nt) {
        const double th = thc[i + (int64_t)i * lde];
        const double wd = zetaivi * d2h[i] * lam[i];
        const double al = dh[i] / th, ga = wd / th;
        alpha[i] = al;
        gamma[i] = ga;
        r1 += dh[i] * al;
        r2 += dh[i] * ga;
    }
    r1 = block_sum(r1, sm);
    r2 = block_sum(r2, sm);
    const double zeta2beta = zeta * zeta + r1;
    const double c0 = sigma + r2;
    const double c1 = c0 / zeta2beta;
    double r3 = 0.0;
    for (int i = tid; i < d; i += nt) {
        const double wd = zetaivi * d2h[i] * lam[i];
        r3 += (lam[i] + c1 * alpha[i] - gamma[i]) * wd;
    }
    r3 = block_sum(r3, sm);
    const double c3 = 1.0 / (v * v) + sigma * c1 + r3;
    if (tid == 0) {
        double* sc = scal + 8 * c;
        sc[0] = phi;
        sc[1] = zeta;
        sc[2] = sigma;
        sc[3] = u;
        sc[4] = v;
        sc[5] = c0;
        sc[6] = 1.0 / (c3 - c0 * c1);
        sc[7] = zeta2beta * c3;
        grad[o] = -zetai;
        grad[o + 1] = -1.0 / v + zetai * sigma;
        if (!ok) feas[kidx[c]] = 0;
    }
    // gradient, W part: svec(V diag(zeta^-1 h' - 1/(v lambda)) V')   (matrixcsqr.jl:158-161)
    const int64_t len = (int64_t)d * (d + 1) / 2;
    for (int64_t idx = tid; idx < len; idx += nt) {
        int a, b;
        svec_rc(idx, a, b);
        const double* va = Vtc + (int64_t)a * lde;
        const double* vb = Vtc + (int64_t)b * lde;
        double s = 0.0;
        for (int k = 0; k < d; k++) s += va[k] * (zetai * dh[k] - wli[k]) * vb[k];
        if (a != b) s *= 1.4142135623730951;
        grad[o + 2 + idx] = s;
    }
}

// is_dual_feas, matrixcsqr.jl:119-138: u >= eps, (conjugate domain positive: Cholesky of the dual
// matrix succeeded), v - u * h_conj(eig(W / u)) > eps.  lamd: eigenvalues of dual W / dual u.
static __global__ void __launch_bounds__(128)
spec_dualfeas_kernel(int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                     const int64_t* __restrict__ lam_off, const int* __restrict__ kidx,
                     const int* __restrict__ hkind, const double* __restrict__ hparam,
                     const double* __restrict__ dual, const double* __restrict__ lamd,
                     const uint8_t* __restrict__ chol_ok, uint8_t* __restrict__ dual_feas) {
    __shared__ double sm[4];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int d = sides[c];
    const int kind = hkind[c];
    const double hp = hparam[c];
    const double* l = lamd + lam_off[c];
    double cj = 0.0;
    for (int i = threadIdx.x; i < d; i += blockDim.x) cj += ssf_conj(kind, hp, l[i]);
    cj = block_sum(cj, sm);
    if (threadIdx.x == 0) {
        const int64_t o = off[c];
        const double u = dual[o];
        bool ok = !(u < HYP_EPS);
        if (ssf_conj_dom_pos(kind) && !chol_ok[kidx[c]]) ok = false;
        if (ok) ok = (dual[o + 1] - u * cj) > HYP_EPS;
        if (!ok) dual_feas[kidx[c]] = 0;
    }
}

// Middle step of hess_prod! (inverse = 0, matrixcsqr.jl:273-319) and inv_hess_prod! (inverse = 1,
// matrixcsqr.jl:402-447) for ONE column: R = V' M V (d x d, leading dimension ldr, shared or global memory) is
// transformed in place into the matrix that is rotated back, and the two leading entries of the product are
// written to out.  theta / Dh have leading dimension lde.  Called by every thread of the CTA.
__device__ __forceinline__ void spec_mid_column(int inverse, int d, int lde, int ldr, const double* sc,
                                                const double* vecs, const double* theta, const double* Dh,
                                                double* R, double p, double q, double* out, double* sm) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* lam = vecs;
    const double* dh = vecs + d;
    const double* alpha = vecs + 5 * (int64_t)d;
    const double* gamma = vecs + 6 * (int64_t)d;
    const double zeta = sc[1], sigma = sc[2], v = sc[4], c0 = sc[5], c4 = sc[6], c5 = sc[7];
    const double zetai = 1.0 / zeta, zetaivi = zetai / v;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    if (!inverse) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = tid; i < d; i += nt) {
            const double rii = R[i + (int64_t)i * ldr];
            s1 += dh[i] * rii;
            s2 += lam[i] * zetaivi * Dh[i + (int64_t)i * lde] * (rii - q * lam[i]);
        }
        s1 = block_sum(s1, sm);
        s2 = block_sum(s2, sm);
        const double c1 = -zetai * (p - sigma * q - s1) * zetai;
        for (int64_t idx = tid; idx < len; idx += nt) {
            int a, b;
            svec_rc(idx, a, b);
            double x = theta[a + (int64_t)b * lde] * R[a + (int64_t)b * ldr];
            if (a == b) x += c1 * dh[a] - zetaivi * Dh[a + (int64_t)a * lde] * q * lam[a];
            R[a + (int64_t)b * ldr] = x;
            R[b + (int64_t)a * ldr] = x;
        }
        if (tid == 0) {
            out[0] = -c1;
            out[1] = c1 * sigma - s2 + q / v / v;
        }
    } else {
        double s1 = 0.0, s2 = 0.0;
        for (int i = tid; i < d; i += nt) {
            const double rii = R[i + (int64_t)i * ldr];
            s1 += gamma[i] * rii;
            s2 += alpha[i] * rii;
        }
        s1 = block_sum(s1, sm);
        s2 = block_sum(s2, sm);
        const double qgr = q + s1;
        const double cu = c4 * (c5 * p + c0 * qgr);
        const double cv = c4 * (c0 * p + qgr);
        for (int64_t idx = tid; idx < len; idx += nt) {
            int a, b;
            svec_rc(idx, a, b);
            double x = R[a + (int64_t)b * ldr] / theta[a + (int64_t)b * lde];
            if (a == b) x += p * alpha[a] + cv * gamma[a];
            R[a + (int64_t)b * ldr] = x;
            R[b + (int64_t)a * ldr] = x;
        }
        if (tid == 0) {
            out[0] = cu + s2;
            out[1] = cv;
        }
    }
    __syncthreads();
}

// The middle step for cc columns of one cone block: R_j sits in Mall (d x d, ld lde, stride lde * lde).
// One CTA per column.
static __global__ void __launch_bounds__(256)
spec_mid_kernel(int inverse, int d, int lde, const double* __restrict__ sc, const double* __restrict__ vecs,
                const double* __restrict__ theta, const double* __restrict__ Dh, double* __restrict__ Mall,
                const double* arr, int64_t ld_arr, double* pr, int64_t ld_prod, int64_t cc) {
    __shared__ double sm[8];
    for (int64_t j = blockIdx.x; j < cc; j += gridDim.x)
        spec_mid_column(inverse, d, lde, lde, sc, vecs, theta, Dh, Mall + j * (int64_t)lde * lde, arr[j * ld_arr],
                        arr[j * ld_arr + 1], pr + j * ld_prod, sm);
}

// Fused product for FEW columns and MANY spectral cones (cf. mat_small_prod_kernel): CTA (c, j) does
//     svec -> M -> R = V' M V -> middle step -> V R V' -> svec
// on chip (M / R and the intermediate T in shared memory, V and V' read through L1).
// mode: 0 hess, 1 inv_hess, 4 block (hess / inv_hess by dualf), 5 block_inv.
static __global__ void __launch_bounds__(256)
spec_small_prod_kernel(int mode_in, int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                       const int64_t* __restrict__ moff, const int64_t* __restrict__ voff,
                       const int* __restrict__ dualf, const double* __restrict__ V,
                       const double* __restrict__ Vt, const double* __restrict__ theta,
                       const double* __restrict__ Dh, const double* __restrict__ vecs,
                       const double* __restrict__ scal, const double* arr, int64_t ld_arr, double* prod,
                       int64_t ld_prod, int64_t row_shift) {
    HYP_DYN_SMEM(double, dyn);
    __shared__ double red[8];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int64_t j = blockIdx.y;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int d = sides[c], lde = (d + 1) & ~1;
    const int ldm = d | 1;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    int mode = mode_in;
    if (mode == 4) mode = (dualf && dualf[c]) ? 1 : 0;
    if (mode == 5) mode = (dualf && dualf[c]) ? 0 : 1;
    const int64_t o = off[c], mo = moff[c];
    const double* a = arr + j * ld_arr + (o - row_shift);
    double* pr = prod + j * ld_prod + (o - row_shift);
    double* sM = dyn;
    double* sT = dyn + (int64_t)d * ldm;
    const double p = a[0], q = a[1];
    for (int64_t idx = tid; idx < len; idx += nt) {
        int r, s;
        svec_rc(idx, r, s);
        double x = a[2 + idx];
        if (r != s) x *= HYP_IRT2;
        sM[r + s * ldm] = x;
        sM[s + r * ldm] = x;
    }
    __syncthreads();
    small_gemm_mx(d, ldm, sM, V + mo, lde, sT);                // T = M V
    __syncthreads();
    small_gemm_xtt(d, ldm, V + mo, lde, sT, [&](int r, int s2, double x) {   // R = V' T -> sM (both triangles)
        sM[r + s2 * ldm] = x;
        sM[s2 + r * ldm] = x;
    });
    __syncthreads();
    spec_mid_column(mode == 1, d, lde, ldm, scal + 8 * c, vecs + voff[c], theta + mo, Dh + mo, sM, p, q, pr, red);
    small_gemm_mx(d, ldm, sM, Vt + mo, lde, sT);               // T = R V'
    __syncthreads();
    small_gemm_xtt(d, ldm, Vt + mo, lde, sT, [&](int r, int s2, double x) {  // V T = V R V'
        if (r != s2) x *= HYP_RT2;
        pr[2 + (int64_t)s2 * (s2 + 1) / 2 + r] = x;
    });
}

// second divided difference of h' over the index triple (i, j, k), update_dder3_aux
// matrixcsqr.jl:449-502: the triple is taken in ascending index order a <= b <= c.
__device__ __forceinline__ double spec_d2h(int i, int j, int k, int lde, const double* lam, const double* d3h,
                                           const double* Dh) {
    int a = i, b = j, c = k;   // caller guarantees i <= j
    if (k < i) {
        a = k; b = i; c = j;
    } else if (k < j) {
        b = k; c = j;
    }
    const double dab = spec_lam_diff(lam[a], lam[b]);
    if (dab == 0.0) {
        const double dac = spec_lam_diff(lam[a], lam[c]);
        if (dac == 0.0) return (d3h[a] + d3h[b] + d3h[c]) / 6.0;
        return (Dh[a + (int64_t)b * lde] - Dh[b + (int64_t)c * lde]) / dac;
    }
    return (Dh[a + (int64_t)c * lde] - Dh[b + (int64_t)c * lde]) / dab;
}

// dder3, matrixcsqr.jl:504-564.  R = V' smat(dir_w) V (d x d, ld lde) in; X: scratch of the same
// size; OUT: the matrix rotated back by the caller.  One CTA.
static __global__ void __launch_bounds__(256)
spec_dder3_kernel(int d, int lde, const double* __restrict__ sc, const double* __restrict__ vecs,
                  const double* __restrict__ Dh, const double* __restrict__ R, double* __restrict__ X,
                  double* __restrict__ OUT, const double* __restrict__ dir, double* __restrict__ out) {
    __shared__ double sm[8];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* lam = vecs;
    const double* dh = vecs + d;
    const double* d3h = vecs + 3 * (int64_t)d;
    const double* wli = vecs + 4 * (int64_t)d;
    const double zeta = sc[1], sigma = sc[2], v = sc[4];
    const double zetai = 1.0 / zeta, vi = 1.0 / v;
    const double p = dir[0], q = dir[1];
    const double viq = vi * q;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    // xi = (R - q diag(lambda)) / v (symmetrised from the upper triangle), xi_b = zeta^-1 Dh .* xi
    double s1 = 0.0, s2 = 0.0;
    for (int64_t idx = tid; idx < len; idx += nt) {
        int a, b;
        svec_rc(idx, a, b);
        double x = vi * R[a + (int64_t)b * lde];
        if (a == b) {
            x -= viq * lam[a];
            s1 += dh[a] * R[a + (int64_t)a * lde];
        }
        X[a + (int64_t)b * lde] = x;
        X[b + (int64_t)a * lde] = x;
        s2 += (a == b ? 0.5 : 1.0) * zetai * Dh[a + (int64_t)b * lde] * x * x;
    }
    s1 = block_sum(s1, sm);
    const double xibxi = block_sum(s2, sm);
    const double zetaichi = zetai * (p - sigma * q - s1);
    const double c1 = -zetai * (zetaichi * zetaichi + v * xibxi);
    const double f = zetaichi + viq;
    for (int64_t idx = tid; idx < len; idx += nt) {
        int i, j;
        svec_rc(idx, i, j);
        const double* xi_i = X + (int64_t)i * lde;
        const double* xi_j = X + (int64_t)j * lde;
        double t = 0.0;
        for (int k = 0; k < d; k++) t += xi_i[k] * spec_d2h(i, j, k, lde, lam, d3h, Dh) * xi_j[k];
        OUT[i + (int64_t)j * lde] = zetai * Dh[i + (int64_t)j * lde] * X[i + (int64_t)j * lde] * f - zetai * t;
    }
    __syncthreads();
    double s3 = 0.0;
    for (int i = tid; i < d; i += nt) s3 += lam[i] * OUT[i + (int64_t)i * lde];
    const double c2 = block_sum(s3, sm);
    for (int64_t idx = tid; idx < len; idx += nt) {
        int i, j;
        svec_rc(idx, i, j);
        // (D R S)(D R S)' with D = diag(1/(v lambda)), S = D^(1/2): entry = w_i w_j sum_k w_k R_ik R_jk
        double t = 0.0;
        for (int k = 0; k < d; k++) {
            const double rik = (i <= k) ? R[i + (int64_t)k * lde] : R[k + (int64_t)i * lde];
            const double rjk = (j <= k) ? R[j + (int64_t)k * lde] : R[k + (int64_t)j * lde];
            t += wli[k] * rik * rjk;
        }
        double x = OUT[i + (int64_t)j * lde] + wli[i] * wli[j] * t;
        if (i == j) x += c1 * dh[i];
        OUT[i + (int64_t)j * lde] = x;
        OUT[j + (int64_t)i * lde] = x;
    }
    if (tid == 0) {
        out[0] = -c1;
        out[1] = c1 * sigma - c2 + xibxi + viq * viq / v;
    }
}

}  // namespace hypdev

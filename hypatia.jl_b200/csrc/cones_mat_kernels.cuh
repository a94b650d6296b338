// Pack / unpack kernels of the matrix-domain cones (svec <-> smat, arrayutilities.jl:163-236), shared by
// cones_mat.cu (PosSemidefTri, log-det, root-det) and cones_spec.cu (EpiPerSepSpectral).
#pragma once
#include "devdefs.cuh"

namespace hypdev {

#define HYP_RT2 1.4142135623730951
// cone type codes (= HYP_CONE_* of the ABI; cones_mat.cu static_asserts the equality)
#define MK_POSSEMIDEFTRI 2
#define MK_HYPOPERLOGDETTRI 3
#define MK_HYPOROOTDETTRI 4
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


// ---- on-chip d x d products for the fused small-matrix kernels: 4 x 4 register tiles per thread ----
// T = M X : M, T in shared memory (leading dimension ldm), X in global memory (leading dimension lde)
__device__ __forceinline__ void small_gemm_mx(int d, int ldm, const double* sM, const double* X, int lde,
                                              double* sT) {
    const int tpd = (d + 3) >> 2;
    for (int t = threadIdx.x; t < tpd * tpd; t += blockDim.x) {
        const int i0 = (t % tpd) * 4, j0 = (t / tpd) * 4;
        double acc[4][4] = {};
        for (int k = 0; k < d; k++) {
            double mv[4], xv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                mv[u] = (i0 + u < d) ? sM[i0 + u + k * ldm] : 0.0;
                xv[u] = (j0 + u < d) ? X[k + (int64_t)(j0 + u) * lde] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int w = 0; w < 4; w++) acc[u][w] += mv[u] * xv[w];
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int w = 0; w < 4; w++)
                if (i0 + u < d && j0 + w < d) sT[i0 + u + (j0 + w) * ldm] = acc[u][w];
    }
}

// Y = X' T for the upper triangle: epi(r, s, Y[r, s]) is called once for every r <= s
template <class Epi>
__device__ __forceinline__ void small_gemm_xtt(int d, int ldm, const double* X, int lde, const double* sT,
                                               Epi epi) {
    const int tpd = (d + 3) >> 2;
    for (int t = threadIdx.x; t < tpd * tpd; t += blockDim.x) {
        const int i0 = (t % tpd) * 4, j0 = (t / tpd) * 4;
        if (i0 > j0) continue;
        double acc[4][4] = {};
        for (int k = 0; k < d; k++) {
            double xv[4], tv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                xv[u] = (i0 + u < d) ? X[k + (int64_t)(i0 + u) * lde] : 0.0;
                tv[u] = (j0 + u < d) ? sT[k + (j0 + u) * ldm] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int w = 0; w < 4; w++) acc[u][w] += xv[u] * tv[w];
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const int r = i0 + u, s = j0 + w;
                if (r <= s && s < d) epi(r, s, acc[u][w]);
            }
    }
}

// Scalar parts of hess_prod! / inv_hess_prod! of the log-det / root-det cones for one column
// (hypoperlogdettri.jl:196-237, :274-319; hyporootdettri.jl:176-212, :246-283): given dot = <r, vecB> (vecB =
// svec(W^-1) for hess, svec(W) for inv_hess) and the leading entries (p, q) of the column, the leading entries of
// the product (out0, out1; out1 only for the log-det cone) and the coefficients of its matrix part
// alpha * svec(X' R X) + beta * vecB.  type: 3 = HypoPerLogdetTri, 4 = HypoRootdetTri.
__device__ __forceinline__ void mat_colscal(int type, int inverse, int d, const double* sc, double dot, double p,
                                            double q, double& out0, double& out1, double& alpha, double& beta) {
    const double phi = sc[1], zeta = sc[2], dd = (double)d;
    if (type == 3) {
        const double v = sc[4];
        if (!inverse) {
            const double sigma = phi - dd, qzi = q / zeta;
            const double c0 = dot / zeta;
            const double c1 = (v * c0 - p / zeta + sigma * qzi) / zeta;
            const double c3 = c1 * v - qzi;
            out0 = -c1;
            out1 = c1 * sigma - c0 + (qzi * dd + q / v) / v;
            alpha = v / zeta + 1.0;
            beta = c3;
        } else {
            const double zv = zeta + v, zzvi = zeta / zv;
            const double c3 = v / (zv + dd * v);
            const double c0 = phi - dd * zzvi;
            const double c4 = v * c3 * zv;
            const double t = zeta + v * phi;
            const double c6 = (v * phi) * (v * phi) + zeta * (zeta + dd * v) - dd * t * t * c3;
            const double c7 = c4 * c0, c8 = c7 + v * zeta;
            const double c1 = dot / zv;
            const double c5 = c0 * p + q + c1;
            const double c2 = v * (zzvi * p + c3 * c5);
            out0 = c6 * p + c7 * q + c8 * c1;
            out1 = c4 * c5;
            alpha = zzvi;
            beta = c2;
        }
    } else {
        const double pzd = sc[5], di = 1.0 / dd;
        out1 = 0.0;
        if (!inverse) {
            const double c0 = pzd * dot;
            const double c1 = c0 - p / zeta;
            const double c2 = pzd * c1 - di * c0;
            out0 = -c1 / zeta;
            alpha = pzd + 1.0;
            beta = c2;
        } else {
            const double phidi = phi * di;
            const double c2 = 1.0 / (pzd + 1.0);
            const double c3 = c2 / zeta * di;
            const double c4 = zeta * zeta + phidi * phi;
            const double c6 = phidi * (c3 * dot + p);
            out0 = phidi * dot + c4 * p;
            alpha = c2;
            beta = c6;
        }
    }
}

// Fused small-matrix product for FEW columns and MANY cones: CTA (c, j) computes column j of the product of
// cone c of a group of PosSemidefTri / HypoPerLogdetTri / HypoRootdetTri cones entirely on chip:
//     svec -> M (shared memory) -> T = M X (shared memory) -> Y = X' T -> alpha * svec(Y) + beta * vecB,
// X = W^-1 (hess), W (inv_hess), U^-1 (sqrt_hess), U' (inv_sqrt_hess) read through L1 from the per-cone state.
// Replaces the per-cone launch sequence unpack / colscal / two TMA GEMMs / pack (5 launches and 6 tensor maps
// per cone and product) that dominates models with many small matrix cones when ncols is 1..4 (apply_lhs,
// setup_rhs3, the z update of solve_subsystem3, the proximity oracles of the line search).
// mode: 0 hess, 1 inv_hess, 2 sqrt, 3 inv_sqrt, 4 block (hess / inv_hess by dualf), 5 block_inv.
// Each thread owns 4 x 4 register tiles; CUDA-core FP64 (4 d^3 flop per (cone, column)).
static __global__ void __launch_bounds__(256)
mat_small_prod_kernel(int type, int mode_in, int ncones, const int64_t* __restrict__ off,
                      const int* __restrict__ sides, const int64_t* __restrict__ moff,
                      const int* __restrict__ dualf, const double* __restrict__ W,
                      const double* __restrict__ Wi, const double* __restrict__ Ui,
                      const double* __restrict__ Ut, const double* __restrict__ scal,
                      const double* __restrict__ point, const double* __restrict__ wivec, const double* arr,
                      int64_t ld_arr, double* prod, int64_t ld_prod, int64_t row_shift) {
    HYP_DYN_SMEM(double, dyn);
    __shared__ double red[8];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int64_t j = blockIdx.y;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int d = sides[c], lde = (d + 1) & ~1;
    const int ldm = d | 1;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    const int lead = type == 2 ? 0 : type == 3 ? 2 : 1;
    int mode = mode_in;
    if (mode == 4) mode = (dualf && dualf[c]) ? 1 : 0;
    if (mode == 5) mode = (dualf && dualf[c]) ? 0 : 1;
    const int inverse = mode == 1;
    const double* X = (mode == 0 ? Wi : mode == 1 ? W : mode == 2 ? Ui : Ut) + moff[c];
    const int64_t o = off[c];
    const double* a = arr + j * ld_arr + (o - row_shift);
    double* pr = prod + j * ld_prod + (o - row_shift);
    double* sM = dyn;
    double* sT = dyn + (int64_t)d * ldm;
    // ---- svec -> M, and the dot product with vecB for the scalar parts ----
    const double* vecB = type == 2 ? nullptr : (inverse ? point : wivec) + o + lead;
    double dot = 0.0;
    for (int64_t idx = tid; idx < len; idx += nt) {
        int r, s;
        svec_rc(idx, r, s);
        double x = a[lead + idx];
        if (vecB) dot += x * vecB[idx];
        if (r != s) x *= HYP_IRT2;
        sM[r + s * ldm] = x;
        sM[s + r * ldm] = x;
    }
    double alpha = 1.0, beta = 0.0, o0 = 0.0, o1 = 0.0;
    if (vecB) {
        dot = block_sum(dot, red);
        mat_colscal(type, inverse, d, scal + 8 * c, dot, a[0], lead == 2 ? a[1] : 0.0, o0, o1, alpha, beta);
    } else {
        __syncthreads();
    }
    small_gemm_mx(d, ldm, sM, X, lde, sT);
    __syncthreads();
    // ---- Y = X' T (upper triangle), packed straight into the product column ----
    small_gemm_xtt(d, ldm, X, lde, sT, [&](int r, int s2, double x) {
        const int64_t idx = (int64_t)s2 * (s2 + 1) / 2 + r;
        if (r != s2) x *= HYP_RT2;
        x *= alpha;
        if (vecB) x += beta * vecB[idx];
        pr[lead + idx] = x;
    });
    if (tid == 0 && lead > 0) {
        pr[0] = o0;
        if (lead == 2) pr[1] = o1;
    }
}

// dder3 of the matrix cones (possemideftri.jl:197-207, hypoperlogdettri.jl:321-368, hyporootdettri.jl:285-324): with
// E = U^-T R U^-1, tr E and tr E^2, the result is U^-1 (k6 E + k1 E^2 + k8 I) U^-T plus the leading entries (out0, out1).
__device__ __forceinline__ void mat_dder3_coefs(int type, int d, const double* sc, double trE, double trE2, double p,
                                                double q, double& out0, double& out1, double& k6, double& k1,
                                                double& k8) {
    const double dd = (double)d;
    k6 = 0.0;
    k1 = 1.0;
    k8 = 0.0;
    out0 = out1 = 0.0;
    if (type == 3) {
        const double phi = sc[1], zeta = sc[2], v = sc[4];
        const double sigma = phi - dd, viq = q / v, viq2 = viq * viq, vzi = v / zeta, vzi1 = vzi + 1.0;
        const double c0 = trE, c7 = trE2;
        const double zichi = (-p + sigma * q + c0 * v) / zeta;
        const double c4 = (viq * (-viq * dd + 2 * c0) - c7) / zeta / 2;
        const double c1 = (zichi * zichi - v * c4) / zeta;
        const double c3 = -(zichi + viq) / zeta;
        const double c5 = c3 * q + vzi * viq2;
        const double c6 = -2 * vzi * viq - c3 * v;
        const double c8 = c5 + c1 * v;
        out0 = -c1;
        out1 = c1 * sigma + (viq2 - (dd * c5 + c6 * c0 + vzi * c7)) / v - c4;
        k6 = c6;
        k1 = vzi1;
        k8 = c8;
    } else if (type == 4) {
        const double phi = sc[1], zeta = sc[2], pzd = sc[5], di = 1.0 / dd;
        const double c0 = trE * di, c6 = trE2 * di;
        const double zichi = (p - phi * c0) / zeta;
        const double c1 = zichi * zichi + phi / zeta * (c6 - c0 * c0) / 2;
        const double c7 = pzd * (c1 - c6 / 2 + c0 * (zichi + c0 / 2));
        const double c8 = -pzd * (zichi + c0);
        out0 = -c1 / zeta;
        k6 = c8;
        k1 = pzd + 1.0;
        k8 = c7;
    }
}

// Fused dder3 for a group of small matrix cones: CTA c does  R -> E = U^-T R U^-1 -> E^2 -> M = k6 E + k1 E^2 + k8 I
// -> U^-1 M U^-T -> svec  on chip (two d x (d|1) buffers in shared memory); replaces 8 launches per cone.
static __global__ void __launch_bounds__(256)
mat_small_dder3_kernel(int type, int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                       const int64_t* __restrict__ moff, const double* __restrict__ Ui,
                       const double* __restrict__ Uit, const double* __restrict__ scal,
                       const double* __restrict__ dir, double* __restrict__ out) {
    HYP_DYN_SMEM(double, dyn);
    __shared__ double red[8];
    __shared__ double coef[3];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int d = sides[c], lde = (d + 1) & ~1;
    const int ldm = d | 1;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    const int lead = type == 2 ? 0 : type == 3 ? 2 : 1;
    const int64_t o = off[c], mo = moff[c];
    const double* a = dir + o;
    double* pr = out + o;
    double* sA = dyn;
    double* sB = dyn + (int64_t)d * ldm;
    const double p = a[0], q = lead == 2 ? a[1] : 0.0;
    for (int64_t idx = tid; idx < len; idx += nt) {
        int r, s;
        svec_rc(idx, r, s);
        double x = a[lead + idx];
        if (r != s) x *= HYP_IRT2;
        sA[r + s * ldm] = x;
        sA[s + r * ldm] = x;
    }
    __syncthreads();
    small_gemm_mx(d, ldm, sA, Ui + mo, lde, sB);                       // T = R U^-1
    __syncthreads();
    small_gemm_xtt(d, ldm, Ui + mo, lde, sB, [&](int r, int s2, double x) {   // E = U^-T T -> sA, both triangles
        sA[r + s2 * ldm] = x;
        sA[s2 + r * ldm] = x;
    });
    __syncthreads();
    small_gemm_mx(d, ldm, sA, sA, ldm, sB);                            // E^2 -> sB
    __syncthreads();
    double t0 = 0.0, t7 = 0.0;
    for (int k = tid; k < d; k += nt) {
        t0 += sA[k + k * ldm];
        t7 += sB[k + k * ldm];
    }
    const double trE = block_sum(t0, red);
    const double trE2 = block_sum(t7, red);
    if (tid == 0) {
        double o0, o1, k6, k1, k8;
        mat_dder3_coefs(type, d, scal + 8 * c, trE, trE2, p, q, o0, o1, k6, k1, k8);
        coef[0] = k6;
        coef[1] = k1;
        coef[2] = k8;
        if (lead >= 1) pr[0] = o0;
        if (lead == 2) pr[1] = o1;
    }
    __syncthreads();
    const double k6 = coef[0], k1 = coef[1], k8 = coef[2];
    for (int idx = tid; idx < d * d; idx += nt) {
        const int r = idx % d, s = idx / d;
        double x = k6 * sA[r + s * ldm] + k1 * sB[r + s * ldm];
        if (r == s) x += k8;
        sB[r + s * ldm] = x;                                           // M
    }
    __syncthreads();
    small_gemm_mx(d, ldm, sB, Uit + mo, lde, sA);                      // T = M U^-T
    __syncthreads();
    small_gemm_xtt(d, ldm, Uit + mo, lde, sA, [&](int r, int s2, double x) {  // U^-1 T
        if (r != s2) x *= HYP_RT2;
        pr[lead + (int64_t)s2 * (s2 + 1) / 2 + r] = x;
    });
}

// ---- state of a group of PosSemidefTri / HypoPerLogdetTri / HypoRootdetTri cones after the batched Cholesky ----
// scal layout (8 doubles per cone): 0 logdet W, 1 phi, 2 zeta, 3 u, 4 v, 5 pzd (rootdet) / sigma (logdet)
// One CTA per cone: U', U^-T, W^-1 (side <= 128; larger cones get W^-1 from a GEMM beforehand),
// log det, the cone's scalars, feasibility, gradient and svec(W^-1).
static __global__ void __launch_bounds__(256)
mat_post_kernel(int type, int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                const int64_t* __restrict__ moff, const int* __restrict__ kidx,
                const double* __restrict__ point, const double* __restrict__ U,
                const double* __restrict__ Ui, double* __restrict__ Ut, double* __restrict__ Uit,
                double* __restrict__ Wi, double* __restrict__ scal, double* __restrict__ grad,
                double* __restrict__ wivec, uint8_t* __restrict__ feas) {
    __shared__ double sm[8];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int d = sides[c], lde = (d + 1) & ~1;
    const int64_t mo = moff[c], o = off[c];
    const double* Uc = U + mo;
    const double* Uic = Ui + mo;
    double* Wic = Wi + mo;
    const int lead = type == MK_POSSEMIDEFTRI ? 0 : type == MK_HYPOPERLOGDETTRI ? 2 : 1;
    for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
        int a = idx % d, b = idx / d;
        Ut[mo + a + (int64_t)b * lde] = Uc[b + (int64_t)a * lde];
        Uit[mo + a + (int64_t)b * lde] = Uic[b + (int64_t)a * lde];
        if (d <= 128 && a <= b) {
            // W^-1 = U^-1 U^-T : entry (a, b) = sum_{k >= b} Ui[a, k] Ui[b, k]
            double s = 0.0;
            for (int k = b; k < d; k++) s += Uic[a + (int64_t)k * lde] * Uic[b + (int64_t)k * lde];
            Wic[a + (int64_t)b * lde] = s;
            Wic[b + (int64_t)a * lde] = s;
        }
    }
    double ld = 0.0;
    for (int k = threadIdx.x; k < d; k += blockDim.x) ld += log(Uc[k + (int64_t)k * lde]);
    ld = 2.0 * block_sum(ld, sm);
    __syncthreads();
    double gscale = -1.0;   // grad matrix part = gscale * svec(W^-1)
    bool ok = true;
    double* sc = scal + 8 * c;
    if (type == MK_HYPOPERLOGDETTRI) {
        // hypoperlogdettri.jl:96-151
        const double u = point[o], v = point[o + 1];
        double phi = 0, zeta = 0;
        if (v > HYP_EPS) {
            phi = ld - d * log(v);
            zeta = v * phi - u;
            ok = zeta > HYP_EPS;
        } else {
            ok = false;
        }
        gscale = -1.0 - v / zeta;
        if (threadIdx.x == 0) {
            sc[0] = ld; sc[1] = phi; sc[2] = zeta; sc[3] = u; sc[4] = v; sc[5] = phi - d;
            grad[o] = 1.0 / zeta;
            grad[o + 1] = -1.0 / v - (phi - d) / zeta;
            wivec[o] = 0.0;
            wivec[o + 1] = 0.0;
        }
    } else if (type == MK_HYPOROOTDETTRI) {
        // hyporootdettri.jl:100-145
        const double u = point[o];
        const double phi = exp(ld / d), zeta = phi - u;
        ok = zeta > HYP_EPS;
        const double pzd = phi / zeta / d;
        gscale = -pzd - 1.0;
        if (threadIdx.x == 0) {
            sc[0] = ld; sc[1] = phi; sc[2] = zeta; sc[3] = u; sc[4] = 0.0; sc[5] = pzd;
            grad[o] = 1.0 / zeta;
            wivec[o] = 0.0;
        }
    } else if (threadIdx.x == 0) {
        sc[0] = ld;
    }
    if (!ok && threadIdx.x == 0) feas[kidx[c]] = 0;
    const int64_t len = (int64_t)d * (d + 1) / 2;
    for (int64_t idx = threadIdx.x; idx < len; idx += blockDim.x) {
        int a, b;
        svec_rc(idx, a, b);
        double x = Wic[a + (int64_t)b * lde];
        if (a != b) x *= HYP_RT2;
        wivec[o + lead + idx] = x;
        grad[o + lead + idx] = gscale * x;
    }
}

// dual feasibility beyond "Cholesky of smat(dual) succeeded" (hypoperlogdettri.jl:119-132,
// hyporootdettri.jl:117-129); U2 holds the factor of the dual matrix
static __global__ void __launch_bounds__(128)
mat_dualfeas_kernel(int type, int ncones, const int64_t* __restrict__ off, const int* __restrict__ sides,
                    const int64_t* __restrict__ moff, const int* __restrict__ kidx,
                    const double* __restrict__ dual, const double* __restrict__ U2,
                    uint8_t* __restrict__ dual_feas) {
    __shared__ double sm[4];
    const int c = blockIdx.x;
    if (c >= ncones) return;
    const int d = sides[c], lde = (d + 1) & ~1;
    const double* Uc = U2 + moff[c];
    double ld = 0.0;
    for (int k = threadIdx.x; k < d; k += blockDim.x) ld += log(Uc[k + (int64_t)k * lde]);
    ld = 2.0 * block_sum(ld, sm);
    if (threadIdx.x == 0) {
        const int64_t o = off[c];
        const double u = dual[o];
        bool ok = false;
        if (u < -HYP_EPS) {
            if (type == MK_HYPOPERLOGDETTRI) {
                const double v = dual[o + 1];
                ok = (v - u * (ld + d * (1.0 - log(-u)))) > HYP_EPS;
            } else {
                ok = (ld - d * log(-u / d)) > HYP_EPS;
            }
        }
        if (!ok) dual_feas[kidx[c]] = 0;
    }
}

}  // namespace hypdev

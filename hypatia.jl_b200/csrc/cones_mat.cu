// Matrix-domain cone oracles: PosSemidefTri, HypoPerLogdetTri, HypoRootdetTri (K7-K10 of
// SURVEY.md section 2.3).
//
// reference: src/Cones/possemideftri.jl:80-207, hypoperlogdettri.jl:96-368,
// hyporootdettri.jl:100-324, arrayutilities.jl:163-236 (svec <-> smat).
//
// State per cone (side d, point matrix W = smat(w)): the Cholesky factor W = U'U, U^-1, U', U^-T
// and W^-1, all d x d column-major, produced by one batched Cholesky + triangular-inverse launch
// (chol.cu) and one batched post kernel.  Every Hessian-type product of these cones is a congruence
//     svec(M) -> svec(X' M X),   X = W^-1 (hess), W (inv_hess), U^-1 (sqrt_hess), U' (inv_sqrt_hess)
// plus, for the log-det / root-det cones, a rank-one correction in the scalars (u, v) and the
// vectors svec(W^-1) / svec(W).  The reference applies X column by column with triangular solves
// (possemideftri.jl:126-195); here the columns of a cone block are unpacked side by side and the
// congruence is two TMA + DMMA GEMMs over the whole column chunk:
//     T  = [M_1 ... M_c]' X          (one (d c) x d product, row block j = M_j X)
//     Y_j = X' T_j                   (c grouped d x d products in one launch)
// Bound: tensor (FP64 DMMA), 4 d^3 flops per column; the unpack / pack passes are HBM-bound.
#include "cones_mat.cuh"
#include <cstdlib>
#include "cones_mat_kernels.cuh"

static_assert(MK_POSSEMIDEFTRI == HYP_CONE_POSSEMIDEFTRI && MK_HYPOPERLOGDETTRI == HYP_CONE_HYPOPERLOGDETTRI &&
                  MK_HYPOROOTDETTRI == HYP_CONE_HYPOROOTDETTRI,
              "cones_mat_kernels.cuh type codes must match the ABI");

using hypdev::block_sum;
using hypdev::mat_dualfeas_kernel;
using hypdev::mat_post_kernel;
using hypdev::pack_cols_kernel;
using hypdev::svec_rc;
using hypdev::unpack_cols_kernel;
using hypdev::unpack_state_kernel;
using hypdev::warp_sum;

namespace {

constexpr double RT2 = 1.4142135623730951;

__global__ void info_to_flag_kernel(const int* __restrict__ info, uint8_t* __restrict__ flag, int k) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && info[0] != 0) flag[k] = 0;
}

__global__ void zero_lower_kernel(double* __restrict__ A, int d, int lde) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d * d; idx += gridDim.x * blockDim.x) {
        int a = idx % d, b = idx / d;
        if (a > b) A[a + (int64_t)b * lde] = 0.0;
    }
}

__global__ void copy_block_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src,
                                  int64_t lds, int rows, int cols, double scale) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < rows * cols; idx += gridDim.x * blockDim.x) {
        int a = idx % rows, b = idx / rows;
        dst[a + (int64_t)b * ldd] = scale * src[a + (int64_t)b * lds];
    }
}

// Per column of a log-det / root-det cone block: the scalar parts of hess_prod! / inv_hess_prod!
// (hypoperlogdettri.jl:196-237, :274-319; hyporootdettri.jl:176-212, :246-283) and the coefficients
// (alpha_j, beta_j) of the matrix part  alpha_j * svec(X' R_j X) + beta_j * vecB.
// One warp per column.  `a` / `pr` point at the first row of the cone block.
__global__ void __launch_bounds__(256)
colscal_kernel(int type, int inverse, int d, int64_t len, const double* __restrict__ sc,
               const double* __restrict__ vecB, const double* a, int64_t ld_arr, double* pr,
               int64_t ld_prod, int64_t cc, double* __restrict__ alpha, double* __restrict__ beta) {
    const int lane = threadIdx.x & 31;
    const int64_t j = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= cc) return;
    const int lead = type == HYP_CONE_HYPOPERLOGDETTRI ? 2 : 1;
    const double* col = a + j * ld_arr;
    double dot = 0.0;
    for (int64_t i = lane; i < len; i += 32) dot += col[lead + i] * vecB[i];
    dot = warp_sum(dot);
    if (lane != 0) return;
    double o0 = 0.0, o1 = 0.0, al = 1.0, be = 0.0;
    hypdev::mat_colscal(type, inverse, d, sc, dot, col[0], col[1], o0, o1, al, be);
    double* out = pr + j * ld_prod;
    out[0] = o0;
    if (type == HYP_CONE_HYPOPERLOGDETTRI) out[1] = o1;
    alpha[j] = al;
    beta[j] = be;
}

// dder3 combination step (hypoperlogdettri.jl:321-368, hyporootdettri.jl:285-324,
// possemideftri.jl:197-207): given E = U^-T R U^-1 and E2 = E E, with tr E = <r, svec W^-1> and
// tr E2, form the matrix  k6 E + k1 E2 + k8 I  (in E2's storage) and the leading entries of dder3.
__global__ void __launch_bounds__(256)
dder3_combine_kernel(int type, int d, int lde, const double* __restrict__ sc, const double* __restrict__ E,
                     double* __restrict__ E2, const double* __restrict__ dir, double* __restrict__ out,
                     const double* __restrict__ wivec) {
    __shared__ double sm[8];
    __shared__ double coef[3];
    const int64_t len = (int64_t)d * (d + 1) / 2;
    const int lead = type == HYP_CONE_POSSEMIDEFTRI ? 0 : type == HYP_CONE_HYPOPERLOGDETTRI ? 2 : 1;
    double t0 = 0.0, t7 = 0.0;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        t0 += E[k + (int64_t)k * lde];
        t7 += E2[k + (int64_t)k * lde];
    }
    const double trE = block_sum(t0, sm);
    const double trE2 = block_sum(t7, sm);
    (void)wivec;
    (void)len;
    if (threadIdx.x == 0) {
        double k6 = 0.0, k1 = 1.0, k8 = 0.0, o0 = 0.0, o1 = 0.0;
        hypdev::mat_dder3_coefs(type, d, sc, trE, trE2, dir[0], type == HYP_CONE_HYPOPERLOGDETTRI ? dir[1] : 0.0, o0, o1,
                                k6, k1, k8);
        if (type != HYP_CONE_POSSEMIDEFTRI) out[0] = o0;
        if (type == HYP_CONE_HYPOPERLOGDETTRI) out[1] = o1;
        coef[0] = k6; coef[1] = k1; coef[2] = k8;
    }
    __syncthreads();
    const double k6 = coef[0], k1 = coef[1], k8 = coef[2];
    (void)lead;
    for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
        int a = idx % d, b = idx / d;
        double x = k6 * E[a + (int64_t)b * lde] + k1 * E2[a + (int64_t)b * lde];
        if (a == b) x += k8;
        E2[a + (int64_t)b * lde] = x;
    }
}

void ensure_matwork(hyp_ctx* ctx, int64_t doubles) {
    if (doubles <= ctx->matwork_doubles) return;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_matwork) cudaFree(ctx->d_matwork);
    ctx->d_matwork = nullptr;
    CUDA_TRY(cudaMalloc(&ctx->d_matwork, (size_t)doubles * sizeof(double)));
    ctx->matwork_doubles = doubles;
}

inline int lead_of(int type) {
    return type == HYP_CONE_POSSEMIDEFTRI ? 0 : type == HYP_CONE_HYPOPERLOGDETTRI ? 2 : 1;
}

// Cholesky + triangular inverse of one large (side > 128) matrix with the blocked kernels
void big_chol_inverse(hyp_ctx* ctx, double* U, double* Ui, int d, int lde, uint8_t* d_flag, int kidx) {
    int nblk = ceil_div(d, 128);
    int64_t need = (int64_t)nblk * 128 * 128 + (int64_t)lde * 128 + 16;
    ensure_matwork(ctx, need);
    double* dinv = ctx->d_matwork;
    double* T = dinv + (int64_t)nblk * 128 * 128;
    hyp_potrf_upper(ctx, U, lde, d, dinv, ctx->d_info + 12);
    info_to_flag_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_info + 12, d_flag, kidx);
    zero_lower_kernel<<<ceil_div((int64_t)d * d, 256), 256, 0, ctx->stream>>>(U, d, lde);
    CUDA_TRY(cudaMemsetAsync(Ui, 0, (size_t)lde * d * 8, ctx->stream));
    ctx->launches += 2;
    // blocked upper-triangular inverse: Ui[jj] = Dinv_j ; Ui[0:j0, jj] = -Ui[0:j0, 0:j0] U[0:j0, jj] Dinv_j
    for (int jb = 0; jb < nblk; jb++) {
        int j0 = jb * 128, nbj = std::min(128, d - j0);
        const double* Dj = dinv + (int64_t)jb * 128 * 128;
        copy_block_kernel<<<ceil_div(nbj * nbj, 256), 256, 0, ctx->stream>>>(Ui + j0 + (int64_t)j0 * lde, lde, Dj,
                                                                            128, nbj, nbj, 1.0);
        ctx->launches++;
        if (j0 > 0) {
            hyp_gemm_simple(ctx, false, false, j0, nbj, nbj, U + (int64_t)j0 * lde, lde, Dj, 128, T, lde);
            hyp_gemm_simple(ctx, false, false, j0, nbj, j0, Ui, lde, T, lde, Ui + (int64_t)j0 * lde, lde);
            copy_block_kernel<<<ceil_div(j0 * nbj, 256), 256, 0, ctx->stream>>>(
                Ui + (int64_t)j0 * lde, lde, Ui + (int64_t)j0 * lde, lde, j0, nbj, -1.0);
            ctx->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
}

// Y_j = X' M_j X for cc matrices stored side by side in Mall (in place); C1 is (d*cc) x d scratch
void congruence(hyp_ctx* ctx, const double* X, int d, int lde, double* Mall, int64_t cc, double* C1,
                int64_t ldc1) {
    // Large cones (side >= 512, e.g. the side-1000 log-det cone of the natvsext-shaped config 5): both products on the
    // int8 tensor pipe by digit slicing (ozaki.cu, hyp_ozaki_gemm_tn) - 2.5 x the FP64 DMMA rate at this depth; below
    // that the contraction is too short for the two-pass tcgen05 kernel and the DMMA products stay.
    static const int i8_min = getenv("HYP_CONG_I8_MIN") ? atoi(getenv("HYP_CONG_I8_MIN")) : 512;
    const bool i8 = d >= i8_min && ctx->syrk_mode == 1;
    // T = [M_1 ... M_cc]' X : row block j (lde rows, the last one padding when d is odd) = M_j X
    if (!(i8 && hyp_ozaki_gemm_tn(ctx, Mall, lde, X, lde, d, (int64_t)lde * cc, d, C1, ldc1, 1.0, 0.0)))
        hyp_gemm_tn(ctx, Mall, lde, X, lde, d, (int64_t)lde * cc, d, C1, ldc1, 1.0, 0.0);
    // Y_j = X' T_j
    if (!(i8 && hyp_ozaki_gemm_tn(ctx, X, lde, C1, ldc1, d, d, d, Mall, lde, 1.0, 0.0, (int)cc, lde, (int64_t)lde * lde)))
        hyp_gemm_tn_grouped(ctx, X, lde, C1, ldc1, d, d, d, (int)cc, lde, Mall, lde, (int64_t)lde * lde, 1.0, 0.0);
}

}  // namespace

void hyp_mat_ensure_work(hyp_ctx* ctx, int64_t doubles) { ensure_matwork(ctx, doubles); }

void hyp_mat_congruence(hyp_ctx* ctx, const double* X, int d, int lde, double* Mall, int64_t cc, double* C1,
                        int64_t ldc1) {
    congruence(ctx, X, d, lde, Mall, cc, C1, ldc1);
}

void hyp_mat_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    (void)ctx;
    // per-cone matrices use an even leading dimension (TMA strides are multiples of 16 bytes)
    g.mat_total = 0;
    for (int i = 0; i < g.count; i++) {
        int d = g.h_side[i], lde = (d + 1) & ~1;
        g.h_moff[i] = g.mat_total;
        g.mat_total += (int64_t)lde * d;
    }
    cudaFree(g.d_moff);
    g.d_moff = nullptr;
    CUDA_TRY(cudaMalloc(&g.d_moff, g.count * sizeof(int64_t)));
    CUDA_TRY(cudaMemcpy(g.d_moff, g.h_moff.data(), g.count * sizeof(int64_t), cudaMemcpyHostToDevice));
    double** mats[] = {&g.d_W, &g.d_U, &g.d_Ut, &g.d_Ui, &g.d_Uit, &g.d_Wi};
    for (double** m : mats) {
        CUDA_TRY(cudaMalloc(m, (size_t)std::max<int64_t>(g.mat_total, 1) * sizeof(double)));
        CUDA_TRY(cudaMemset(*m, 0, (size_t)std::max<int64_t>(g.mat_total, 1) * sizeof(double)));
    }
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
}

void hyp_mat_update_state(hyp_ctx* ctx, ConeGroup& g) {
    const int lead = lead_of(g.type);
    int64_t maxlen = (int64_t)g.max_side * (g.max_side + 1) / 2;
    dim3 ugrid(g.count, (unsigned)std::max<int64_t>(1, std::min<int64_t>((maxlen + 255) / 256, 64)));
    // ---- primal: W, Cholesky, inverse (possemideftri.jl:80-107) ----
    unpack_state_kernel<<<ugrid, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_moff, lead,
                                                        ctx->d_point, g.d_W, g.d_U);
    ctx->launches++;
    hyp_chol_batched(ctx, g.count, g.d_side, g.d_moff, g.d_kidx, g.d_U, g.d_Ui, ctx->d_feas);
    for (int i = 0; i < g.count; i++) {
        int d = g.h_side[i];
        if (d <= 128) continue;
        int lde = (d + 1) & ~1;
        big_chol_inverse(ctx, g.d_U + g.h_moff[i], g.d_Ui + g.h_moff[i], d, lde, ctx->d_feas, g.h_kidx[i]);
        // W^-1 = U^-1 U^-T
        hyp_gemm_simple(ctx, false, true, d, d, d, g.d_Ui + g.h_moff[i], lde, g.d_Ui + g.h_moff[i], lde,
                        g.d_Wi + g.h_moff[i], lde);
    }
    mat_post_kernel<<<g.count, 256, 0, ctx->stream>>>(g.type, g.count, g.d_off, g.d_side, g.d_moff, g.d_kidx,
                                                     ctx->d_point, g.d_U, g.d_Ui, g.d_Ut, g.d_Uit, g.d_Wi,
                                                     g.d_scal, ctx->d_grad, ctx->d_wivec, ctx->d_feas);
    ctx->launches++;
    // ---- dual feasibility (possemideftri.jl:92-95): Cholesky of smat(dual) on scratch ----
    ensure_matwork(ctx, 2 * g.mat_total + 16);
    double* U2 = ctx->d_matwork;
    double* Ui2 = ctx->d_matwork + g.mat_total;
    bool any_big = g.max_side > 128;
    if (!any_big) {
        unpack_state_kernel<<<ugrid, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_moff, lead,
                                                            ctx->d_dual, U2, nullptr);
        ctx->launches++;
        hyp_chol_batched(ctx, g.count, g.d_side, g.d_moff, g.d_kidx, U2, Ui2, ctx->d_dual_feas);
        if (g.type != HYP_CONE_POSSEMIDEFTRI) {
            mat_dualfeas_kernel<<<g.count, 128, 0, ctx->stream>>>(g.type, g.count, g.d_off, g.d_side, g.d_moff,
                                                                 g.d_kidx, ctx->d_dual, U2, ctx->d_dual_feas);
            ctx->launches++;
        }
    } else {
        // large cones: factor the dual matrices one by one in a scratch copy that lives behind the Dinv blocks of the
        // blocked Cholesky in the context's matrix workspace (no cudaMalloc / cudaFree - and no implicit device
        // synchronisation - inside the per-iteration state update)
        int64_t dinv_len = 0;
        for (int i = 0; i < g.count; i++)
            dinv_len = std::max<int64_t>(dinv_len, (int64_t)ceil_div(g.h_side[i], 128) * 128 * 128 + 16);
        ensure_matwork(ctx, dinv_len + g.mat_total + 16);
        double* scratch = ctx->d_matwork + dinv_len;
        unpack_state_kernel<<<ugrid, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_side, g.d_moff, lead,
                                                            ctx->d_dual, scratch, nullptr);
        ctx->launches++;
        for (int i = 0; i < g.count; i++) {
            int d = g.h_side[i], lde = (d + 1) & ~1;
            hyp_potrf_upper(ctx, scratch + g.h_moff[i], lde, d, ctx->d_matwork, ctx->d_info + 12);
            info_to_flag_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_info + 12, ctx->d_dual_feas, g.h_kidx[i]);
            ctx->launches++;
        }
        if (g.type != HYP_CONE_POSSEMIDEFTRI) {
            mat_dualfeas_kernel<<<g.count, 128, 0, ctx->stream>>>(g.type, g.count, g.d_off, g.d_side, g.d_moff,
                                                                 g.d_kidx, ctx->d_dual, scratch, ctx->d_dual_feas);
            ctx->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
}

void hyp_mat_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols,
                  int64_t ld_prod, int64_t ld_arr, int mode, int64_t row_shift) {
    const int lead = lead_of(g.type);
    if ((mode == HYP_PROD_SQRT_HESS || mode == HYP_PROD_INV_SQRT_HESS) && g.type != HYP_CONE_POSSEMIDEFTRI)
        throw HypError{"sqrt_hess_prod is not defined for the log-det / root-det cones"};
    // few columns, sides that fit in shared memory: one fused launch for the whole group (CUDA-core FP64,
    // M and T on chip) instead of the per-cone tensor-GEMM sequence below
    {
        static int max_small_cols = -1;
        if (max_small_cols < 0) {
            const char* e = getenv("HYP_MAT_SMALL_MAXCOLS");
            max_small_cols = e ? atoi(e) : 8;
        }
        const int64_t smem = (int64_t)2 * g.max_side * (g.max_side | 1) * sizeof(double);
        if (ncols <= max_small_cols && g.count > 0 && smem <= 226 * 1024) {
            static bool attr = false;
            if (!attr) {
                CUDA_TRY(cudaFuncSetAttribute(hypdev::mat_small_prod_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                attr = true;
            }
            dim3 grid(g.count, (unsigned)ncols);
            hypdev::mat_small_prod_kernel<<<grid, 256, smem, ctx->stream>>>(
                g.type, mode, g.count, g.d_off, g.d_side, g.d_moff, g.d_dual, g.d_W, g.d_Wi, g.d_Ui, g.d_Ut, g.d_scal,
                ctx->d_point, ctx->d_wivec, arr, ld_arr, prod, ld_prod, row_shift);
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            return;
        }
    }
    const int64_t budget = (int64_t)48 << 20;   // doubles per workspace matrix (384 MB)
    for (int i = 0; i < g.count; i++) {
        const int d = g.h_side[i], lde = (d + 1) & ~1;
        const int64_t len = (int64_t)d * (d + 1) / 2;
        int m = mode;
        if (m == HYP_PROD_BLOCK) m = g.h_dual[i] ? HYP_PROD_INV_HESS : HYP_PROD_HESS;
        if (m == HYP_PROD_BLOCK_INV) m = g.h_dual[i] ? HYP_PROD_HESS : HYP_PROD_INV_HESS;
        const double* X = (m == HYP_PROD_HESS ? g.d_Wi : m == HYP_PROD_INV_HESS ? g.d_W
                           : m == HYP_PROD_SQRT_HESS ? g.d_Ui : g.d_Ut) + g.h_moff[i];
        const int inverse = (m == HYP_PROD_INV_HESS) ? 1 : 0;
        const int64_t per_col = (int64_t)lde * lde;
        int64_t cmax = std::max<int64_t>(1, std::min<int64_t>(ncols, budget / per_col));
        const int64_t ldc1 = (int64_t)lde * cmax;
        ensure_matwork(ctx, per_col * cmax + ldc1 * d + 2 * cmax + 16);
        double* Mall = ctx->d_matwork;
        double* C1 = Mall + per_col * cmax;
        double* alpha = C1 + ldc1 * d;
        double* beta = alpha + cmax;
        const int64_t row0 = g.h_off[i] - row_shift;
        const double* vecB = nullptr;
        if (g.type != HYP_CONE_POSSEMIDEFTRI)
            vecB = (inverse ? ctx->d_point : ctx->d_wivec) + g.h_off[i] + lead;
        for (int64_t j0 = 0; j0 < ncols; j0 += cmax) {
            const int64_t cc = std::min(cmax, ncols - j0);
            const double* a0 = arr + row0 + j0 * ld_arr;
            double* p0 = prod + row0 + j0 * ld_prod;
            dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((len + 255) / 256, cc > 64 ? 8 : 64)),
                      (unsigned)std::min<int64_t>(cc, 65535));
            unpack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, a0 + lead, ld_arr, cc, Mall);
            ctx->launches++;
            if (g.type != HYP_CONE_POSSEMIDEFTRI) {
                colscal_kernel<<<ceil_div(cc, 8), 256, 0, ctx->stream>>>(g.type, inverse, d, len, g.d_scal + 8 * i,
                                                                       vecB, a0, ld_arr, p0, ld_prod, cc, alpha,
                                                                       beta);
                ctx->launches++;
            }
            congruence(ctx, X, d, lde, Mall, cc, C1, ldc1);
            pack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, Mall, cc,
                                                            vecB ? alpha : nullptr, vecB ? beta : nullptr, vecB,
                                                            p0 + lead, ld_prod);
            ctx->launches++;
        }
    }
    CUDA_TRY(cudaGetLastError());
}

// dder3 for the matrix cones: E = U^-T R U^-1, E2 = E E, M = k6 E + k1 E2 + k8 I, result U^-1 M U^-T
void hyp_mat_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir) {
    const int lead = lead_of(g.type);
    {
        // sides that fit in shared memory: one fused launch for the whole group
        static int use_small = -1;
        if (use_small < 0) {
            const char* e = getenv("HYP_MAT_SMALL_MAXCOLS");
            use_small = (e && atoi(e) == 0) ? 0 : 1;
        }
        const int64_t smem = (int64_t)2 * g.max_side * (g.max_side | 1) * sizeof(double);
        if (use_small && g.count > 0 && smem <= 226 * 1024) {
            static bool attr = false;
            if (!attr) {
                CUDA_TRY(cudaFuncSetAttribute(hypdev::mat_small_dder3_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
                attr = true;
            }
            hypdev::mat_small_dder3_kernel<<<g.count, 256, smem, ctx->stream>>>(
                g.type, g.count, g.d_off, g.d_side, g.d_moff, g.d_Ui, g.d_Uit, g.d_scal, dir, out);
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            return;
        }
    }
    for (int i = 0; i < g.count; i++) {
        const int d = g.h_side[i], lde = (d + 1) & ~1;
        const int64_t len = (int64_t)d * (d + 1) / 2;
        const int64_t per = (int64_t)lde * lde;
        const int64_t ldc1 = lde;
        ensure_matwork(ctx, 2 * per + ldc1 * d + 16);
        double* E = ctx->d_matwork;
        double* E2 = E + per;
        double* C1 = E2 + per;
        const int64_t o = g.h_off[i];
        dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((len + 255) / 256, 64)), 1);
        unpack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, dir + o + lead, ctx->q, 1, E);
        ctx->launches++;
        congruence(ctx, g.d_Ui + g.h_moff[i], d, lde, E, 1, C1, ldc1);           // E = U^-T R U^-1
        hyp_gemm_tn(ctx, E, lde, E, lde, d, d, d, E2, lde, 1.0, 0.0);            // E2 = E' E = E E
        dder3_combine_kernel<<<1, 256, 0, ctx->stream>>>(g.type, d, lde, g.d_scal + 8 * i, E, E2, dir + o,
                                                        out + o, ctx->d_wivec + o + lead);
        ctx->launches++;
        congruence(ctx, g.d_Uit + g.h_moff[i], d, lde, E2, 1, C1, ldc1);         // U^-1 M U^-T
        pack_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d, lde, len, E2, 1, nullptr, nullptr, nullptr,
                                                        out + o + lead, ctx->q);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
}

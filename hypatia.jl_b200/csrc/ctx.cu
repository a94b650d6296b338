// C ABI of libhypatia_b200 (include/hypatia_b200.h): context, model load, and the
// SystemSolver entry points - device restatement of the reference's QRCholDenseSystemSolver.
//
// reference: src/Solvers/systemsolvers/qrchol.jl:16-37 (setup_rhs3), :39-85 (solve_subsystem3),
// :138-179 (load), :181-257 (update_lhs, update_lhs_fact); common.jl:79-121 (apply_lhs),
// :129-151 (solve_system), :154-182 (solve_subsystem4), :184-211 (setup_point_sub, dot_obj);
// src/linearalgebra/dense.jl:194-215 (posdef_fact_copy! fallback chain).
#include "common.cuh"

#include <cmath>
#include <cstring>

namespace {

const char* kTimingNames[T_NUM] = {"cone_state", "schur_prepass", "schur_syrk", "allreduce", "potrf",
                                   "ldlt",       "trsv",          "gemv",       "cone_prod", "vec"};

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

template <typename T>
void dalloc(T** p, int64_t count) {
    *p = nullptr;
    if (count <= 0) count = 1;
    CUDA_TRY(cudaMalloc((void**)p, (size_t)count * sizeof(T)));
    CUDA_TRY(cudaMemset(*p, 0, (size_t)count * sizeof(T)));
    // the memset runs on the legacy stream, which does not order with the context's non-blocking
    // stream: finish it before anything is enqueued on the new buffer
    CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));
}

template <typename T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// device view of an input array: the pointer itself when it already lives on the device, else a
// copy in `buf` (which must hold `len` doubles)
const double* stage_in(hyp_ctx* ctx, const double* p, int64_t len, double* buf) {
    if (len <= 0) return buf;
    if (!p) throw HypError{"null input pointer"};
    if (is_device_ptr(p)) return p;
    CUDA_TRY(cudaMemcpyAsync(buf, p, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    return buf;
}

void stage_out(hyp_ctx* ctx, double* p, int64_t len, const double* dev) {
    if (len <= 0) return;
    if (!p) throw HypError{"null output pointer"};
    if (is_device_ptr(p)) {
        if (p != dev)
            CUDA_TRY(cudaMemcpyAsync(p, dev, (size_t)len * sizeof(double), cudaMemcpyDeviceToDevice,
                                     ctx->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(p, dev, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
}

void stage_out_u8(hyp_ctx* ctx, uint8_t* p, int64_t len, const uint8_t* dev) {
    if (len <= 0 || !p) return;
    if (is_device_ptr(p)) {
        CUDA_TRY(cudaMemcpyAsync(p, dev, (size_t)len, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(p, dev, (size_t)len, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
}

void ensure_stage(hyp_ctx* ctx, int64_t doubles) {
    if (doubles <= ctx->stage_doubles) return;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dfree(ctx->d_stage);
    dalloc(&ctx->d_stage, doubles);
    ctx->stage_doubles = doubles;
}

// copy a (rows x cols, ld) host-or-device matrix into a device matrix with leading dim dld
void upload_matrix(hyp_ctx* ctx, double* dst, int64_t dld, const double* src, int64_t sld,
                   int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return;
    if (!src) throw HypError{"null matrix pointer"};
    cudaMemcpyKind kind = is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dld * 8, src, (size_t)sld * 8, (size_t)rows * 8, (size_t)cols,
                               kind, ctx->stream));
}

void free_model(hyp_ctx* ctx) {
    hyp_cones_free_groups(ctx);
    if (ctx->d_GQ != ctx->d_Graw) dfree(ctx->d_GQ);
    ctx->d_GQ = nullptr;
    dfree(ctx->d_Graw);
    dfree(ctx->d_HG);
    dfree(ctx->d_PG);
    dfree(ctx->d_A);
    dfree(ctx->d_Q);
    dfree(ctx->d_R);
    dfree(ctx->d_Rdinv);
    dfree(ctx->d_cbh);
    dfree(ctx->d_cone_nu);
    dfree(ctx->d_cone_off);
    dfree(ctx->d_cone_dim);
    dfree(ctx->d_cone_type);
    dfree(ctx->d_row_dual);
    dfree(ctx->d_point);
    dfree(ctx->d_dual);
    dfree(ctx->d_grad);
    dfree(ctx->d_wivec);
    dfree(ctx->d_row_ns);
    dfree(ctx->d_feas);
    dfree(ctx->d_dual_feas);
    dfree(ctx->d_num_ok);
    dfree(ctx->d_tmpflag);
    dfree(ctx->d_proxsqr);
    dfree(ctx->d_matwork);
    ctx->matwork_doubles = 0;
    dfree(ctx->d_S);
    dfree(ctx->d_F);
    dfree(ctx->d_Dinv);
    dfree(ctx->d_info);
    dfree(ctx->d_ipiv);
    dfree(ctx->d_ldl_work);
    dfree(ctx->d_flags);
    dfree(ctx->d_digits);
    dfree(ctx->d_expo);
    dfree(ctx->d_dscale);
    dfree(ctx->d_L3);
    dfree(ctx->d_F3);
    dfree(ctx->d_row_cone);
    dfree(ctx->d_blk_arr);
    ctx->solver_kind = 0;
    dfree(ctx->d_rhs);
    dfree(ctx->d_sol);
    dfree(ctx->d_sub_sol);
    dfree(ctx->d_sub_rhs);
    dfree(ctx->d_const_sol);
    dfree(ctx->d_const_rhs);
    dfree(ctx->d_Gx_const);
    dfree(ctx->d_t);
    dfree(ctx->d_t2);
    dfree(ctx->d_Gx);
    dfree(ctx->d_HGx);
    dfree(ctx->d_vq1);
    dfree(ctx->d_vq2);
    dfree(ctx->d_vq3);
    dfree(ctx->d_vq4);
    dfree(ctx->d_vp1);
    dfree(ctx->d_vp2);
    dfree(ctx->d_partial);
    dfree(ctx->d_scalars);
    dfree(ctx->d_partial2);
    dfree(ctx->d_partial3);
    dfree(ctx->d_partial4);
    ctx->partial3_doubles = ctx->partial4_doubles = 0;
    dfree(ctx->d_multi);
    ctx->multi_doubles = 0;
    dfree(ctx->d_colbits);
    dfree(ctx->d_digitsP);
    dfree(ctx->d_expoP);
    dfree(ctx->d_dscaleP);
    ctx->partial2_doubles = 0;
    dfree(ctx->d_stage);
    ctx->stage_doubles = 0;
    ctx->model_loaded = ctx->lhs_ready = ctx->cones_loaded = false;
}

// ---- small device kernels of the Point-level glue ------------------------------------------
__global__ void negate_kernel(int64_t len, double* __restrict__ out, const double* __restrict__ in) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = -in[i];
}

// setup_rhs3 pre-step: v = rhs.z on primal-barrier rows, -rhs.z - rhs.s on dual-barrier rows
__global__ void rhs3_pre_kernel(int64_t q, const uint8_t* __restrict__ row_dual,
                                const double* __restrict__ rz, const double* __restrict__ rs,
                                double* __restrict__ v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q;
         i += (int64_t)gridDim.x * blockDim.x)
        v[i] = (row_dual && row_dual[i]) ? (-rz[i] - rs[i]) : rz[i];
}

// setup_rhs3 post-step: out = -Hv - rhs.s (primal rows) or Hinv v (dual rows)
__global__ void rhs3_post_kernel(int64_t q, const uint8_t* __restrict__ row_dual,
                                 const double* __restrict__ hv, const double* __restrict__ rs,
                                 double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (row_dual && row_dual[i]) ? hv[i] : (-hv[i] - rs[i]);
}

__global__ void symindef_rhs3_kernel(int64_t q, const uint8_t* __restrict__ row_dual,
                                     const double* __restrict__ hinv_rs, const double* __restrict__ rz,
                                     const double* __restrict__ rs, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (row_dual && row_dual[i]) ? (-rz[i] - rs[i]) : (-hinv_rs[i] - rz[i]);
}

// primal / dual views of a direction (point.jl:46-51): primal = s (z on dual-barrier rows)
__global__ void primal_dual_kernel(int64_t q, const uint8_t* __restrict__ row_dual,
                                   const double* __restrict__ z, const double* __restrict__ s,
                                   double* __restrict__ prim, double* __restrict__ dual) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q;
         i += (int64_t)gridDim.x * blockDim.x) {
        bool d = row_dual && row_dual[i];
        prim[i] = d ? z[i] : s[i];
        dual[i] = d ? s[i] : z[i];
    }
}

// tau lift of solve_subsystem4 / solve_system (common.jl:171-175,147-148); scal[0] = dot_obj(sol_sub),
// scal[1] = dot_obj(sol_const); writes scal[2] = tau, sol[tau_idx] = tau, sol[kap_idx] = kap
__global__ void tau_kernel(double* __restrict__ scal, const double* __restrict__ rhs,
                           double* __restrict__ sol, int64_t tau_idx, int64_t kap_idx, double mu_tt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double num = rhs[tau_idx] + rhs[kap_idx] + scal[0];
        double den = mu_tt - scal[1];
        double tau = num / den;
        scal[2] = tau;
        sol[tau_idx] = tau;
        sol[kap_idx] = -mu_tt * tau + rhs[kap_idx];
    }
}

// s = -(Gx_sub + tau * Gx_const) + h * tau - rhs.z   (common.jl:143-144 with G*x split)
__global__ void s_lift_kernel(int64_t q, const double* __restrict__ scal,
                              const double* __restrict__ gx, const double* __restrict__ gxc,
                              const double* __restrict__ h, const double* __restrict__ rz,
                              double* __restrict__ s) {
    const double tau = scal[2];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q;
         i += (int64_t)gridDim.x * blockDim.x)
        s[i] = -(gx[i] + tau * gxc[i]) + h[i] * tau - rz[i];
}

// apply_lhs tail (common.jl:97-120): res.tau and res.kap from the device dot products
__global__ void lhs_tail_kernel(const double* __restrict__ scal, const double* __restrict__ dir,
                                double* __restrict__ res, int64_t tau_idx, int64_t kap_idx,
                                double mu_tt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double tau_dir = dir[tau_idx], kap_dir = dir[kap_idx];
        res[tau_idx] = -scal[3] - kap_dir;
        res[kap_idx] = mu_tt * tau_dir + kap_dir;
    }
}

// out = a * x + b * y + cs * dir[idx] * z   (scalar taken from a Point entry on the device)
__global__ void axpbypcz_dev_kernel(int64_t len, double* __restrict__ out, double a,
                                    const double* __restrict__ x, double b, const double* __restrict__ y,
                                    double cs, const double* __restrict__ sptr,
                                    const double* __restrict__ z) {
    const double c = cs * sptr[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = c * z[i];
        if (x) v += a * x[i];
        if (y) v += b * y[i];
        out[i] = v;
    }
}

inline int vgrid(hyp_ctx* ctx, int64_t len) {
    int64_t blocks = (len + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

// ---- G passes (local row panel; results re-replicated across ranks) ------------------------
// y(n) = alpha * G' v(q) + beta * y       (qrchol.jl:52, common.jl:91)
void G_t(hyp_ctx* ctx, const double* Gp, const double* vq, double alpha, double beta, double* yn,
         int64_t ncols = -1) {
    if (ncols < 0) ncols = ctx->n;
    if (!hyp_row_sharded(ctx)) {
        hyp_gemv_t(ctx, ctx->qloc, ncols, Gp, ctx->ldg, vq + ctx->row_lo, alpha, beta, yn);
    } else {
        hyp_gemv_t(ctx, ctx->qloc, ncols, Gp, ctx->ldg, vq + ctx->row_lo, alpha, 0.0, ctx->d_t2);
        hyp_allreduce_sum(ctx, ctx->d_t2, ncols);
        hyp_axpby(ctx, ncols, 1.0, ctx->d_t2, beta, yn);
    }
}
// out(q)[local rows] = G x   (qrchol.jl:73); other rows untouched
void G_n(hyp_ctx* ctx, const double* Gp, const double* xn, double* outq, int64_t ncols = -1) {
    if (ncols < 0) ncols = ctx->n;
    hyp_gemv_n(ctx, ctx->qloc, ncols, Gp, ctx->ldg, xn, 1.0, 0.0, outq + ctx->row_lo);
}

void potrs(hyp_ctx* ctx, double* x) {
    if (ctx->fact_kind == 0) {
        hyp_trsv_upper(ctx, ctx->d_F, ctx->lds, ctx->nmp, ctx->d_Dinv, x, true);
        hyp_trsv_upper(ctx, ctx->d_F, ctx->lds, ctx->nmp, ctx->d_Dinv, x, false);
    } else {
        hyp_ldlt_solve(ctx, ctx->d_F, ctx->lds, ctx->nmp, ctx->d_ipiv, x);
    }
}

// solve_subsystem3 (qrchol.jl:39-85) on device sub Points; leaves G*sol.x in ctx->d_Gx and
// H*G*sol.x in ctx->d_HGx
void solve_subsystem3_dev(hyp_ctx* ctx, double* sol, const double* rhs) {
    const int64_t n = ctx->n, p = ctx->p, q = ctx->q, nmp = ctx->nmp;
    if (sol != rhs) hyp_copy(ctx, n + p + q, sol, rhs);
    if (ctx->solver_kind == 1) {
        // SymIndefDense: ldiv!(sol.vec, fact, rhs.vec) (symindef.jl:263-271); G x kept for the s lift
        hyp_ldlt_solve(ctx, ctx->d_F3, ctx->ld3, n + p + q, ctx->d_ipiv, sol);
        if (q > 0) G_n(ctx, ctx->d_Graw, sol, ctx->d_Gx);
        return;
    }
    double* x = sol;
    double* y = sol + n;
    double* z = sol + n + p;
    double* t = ctx->d_t;

    // t = Q'(x + G'z)
    hyp_copy(ctx, n, t, x);
    G_t(ctx, ctx->d_Graw, z, 1.0, 1.0, t);
    if (ctx->d_Q) {
        hyp_gemv_t(ctx, n, n, ctx->d_Q, ctx->ldqm, t, 1.0, 0.0, ctx->d_vq4 /*n <= ? see alloc*/);
        hyp_copy(ctx, n, t, ctx->d_vq4);
    }
    if (p > 0) {
        // y = R' \ y ; x_sub1 = y
        hyp_trsv_upper(ctx, ctx->d_R, ctx->ldr, p, ctx->d_Rdinv, y, true);
        hyp_copy(ctx, p, x, y);
        if (nmp > 0) {
            // Q2div -= GQ2' H GQ1 y
            G_n(ctx, ctx->d_GQ, y, ctx->d_vq1, p);
            hyp_cones_prod(ctx, ctx->d_vq2, ctx->d_vq1, 1, q, q, HYP_PROD_BLOCK, 0);
            G_t(ctx, ctx->d_GQ + p * ctx->ldg, ctx->d_vq2, -1.0, 1.0, t + p, nmp);
        }
    }
    if (nmp > 0) {
        potrs(ctx, t + p);
        hyp_copy(ctx, nmp, x + p, t + p);
    }
    if (ctx->d_Q) {
        // x = Q x
        hyp_gemv_n(ctx, n, n, ctx->d_Q, ctx->ldqm, x, 1.0, 0.0, ctx->d_vq4);
        hyp_copy(ctx, n, x, ctx->d_vq4);
    }
    // z = H G x - z
    G_n(ctx, ctx->d_Graw, x, ctx->d_Gx);
    {
        TimeScope ts(ctx, T_CONE_PROD);
        hyp_cones_prod(ctx, ctx->d_HGx, ctx->d_Gx, 1, q, q, HYP_PROD_BLOCK, 0);
    }
    if (hyp_row_sharded(ctx)) {
        hyp_replicate_q(ctx, ctx->d_Gx);
        hyp_replicate_q(ctx, ctx->d_HGx);
    }
    hyp_axpby(ctx, q, 1.0, ctx->d_HGx, -1.0, z);
    if (p > 0) {
        // y = R \ (Q1'(x + G'z) - GQ1' H G x)
        hyp_copy(ctx, p, y, t);
        G_t(ctx, ctx->d_GQ, ctx->d_HGx, -1.0, 1.0, y, p);
        hyp_trsv_upper(ctx, ctx->d_R, ctx->ldr, p, ctx->d_Rdinv, y, false);
    }
}

// solve_system (common.jl:129-151) on device full Points
void solve_system_dev(hyp_ctx* ctx, double* sol, const double* rhs) {
    const int64_t n = ctx->n, p = ctx->p, q = ctx->q;
    const int64_t dim3 = n + p + q, tau_idx = dim3, kap_idx = dim3 + q + 1;
    const double* rz = rhs + n + p;
    const double* rs = rhs + tau_idx + 1;
    double* sub_rhs = ctx->d_sub_rhs;
    double* sub_sol = ctx->d_sub_sol;
    // rhs_sub.x = rhs.x ; rhs_sub.y = -rhs.y
    hyp_copy(ctx, n, sub_rhs, rhs);
    if (p > 0) {
        negate_kernel<<<vgrid(ctx, p), 256, 0, ctx->stream>>>(p, sub_rhs + n, rhs + n);
        ctx->launches++;
    }
    if (q > 0 && ctx->solver_kind == 1) {
        // setup_rhs3 (symindef.jl:31-52): -Hinv rs - rz on primal-barrier rows, -rz - rs on dual rows
        TimeScope ts(ctx, T_CONE_PROD);
        hyp_cones_prod(ctx, ctx->d_vq2, rs, 1, q, q, HYP_PROD_BLOCK_INV, 0);
        symindef_rhs3_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, ctx->d_vq2, rz, rs,
                                                                   sub_rhs + n + p);
        ctx->launches++;
    } else
    // setup_rhs3 (qrchol.jl:16-37)
    if (q > 0) {
        TimeScope ts(ctx, T_CONE_PROD);
        const double* v = rz;
        if (ctx->any_dual) {
            rhs3_pre_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, rz, rs, ctx->d_vq1);
            ctx->launches++;
            v = ctx->d_vq1;
        }
        hyp_cones_prod(ctx, ctx->d_vq2, v, 1, q, q, HYP_PROD_BLOCK, 0);
        if (hyp_row_sharded(ctx)) hyp_replicate_q(ctx, ctx->d_vq2);
        rhs3_post_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, ctx->d_vq2, rs,
                                                              sub_rhs + n + p);
        ctx->launches++;
    }
    solve_subsystem3_dev(ctx, sub_sol, sub_rhs);
    // tau lift (common.jl:171-179)
    hyp_dot(ctx, dim3, ctx->d_cbh, sub_sol, ctx->d_scalars + 0, false);
    const double mu_tt = ctx->mu / ctx->tau_bar / ctx->tau_bar;
    tau_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_scalars, rhs, sol, tau_idx, kap_idx, mu_tt);
    ctx->launches++;
    hyp_axpy_dev(ctx, dim3, sol, 1.0, sub_sol, ctx->d_scalars + 2, 1.0, ctx->d_const_sol);
    // s = -G x + h tau - rhs.z  with  G x = G x_sub + tau G x_const
    if (q > 0) {
        s_lift_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_scalars, ctx->d_Gx, ctx->d_Gx_const,
                                                           ctx->d_cbh + n + p, rz, sol + tau_idx + 1);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
}

// apply_lhs (common.jl:79-121) on device full Points
void apply_lhs_dev(hyp_ctx* ctx, double* res, const double* dir) {
    const int64_t n = ctx->n, p = ctx->p, q = ctx->q;
    const int64_t dim3 = n + p + q, tau_idx = dim3, kap_idx = dim3 + q + 1;
    const double* dx = dir;
    const double* dy = dir + n;
    const double* dz = dir + n + p;
    const double* ds = dir + tau_idx + 1;
    const double* c = ctx->d_cbh;
    const double* b = ctx->d_cbh + n;
    const double* h = ctx->d_cbh + n + p;
    // res.x = G'z + A'y + c tau
    axpbypcz_dev_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, res, 0.0, nullptr, 0.0, nullptr, 1.0,
                                                             dir + tau_idx, c);
    ctx->launches++;
    // one pass over G for G'z (here) and G x (res.z below) when the panel is not sharded
    const bool fuse_g = !hyp_row_sharded(ctx) && q > 0 && n > 0;
    if (fuse_g)
        hyp_gemv_nt(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, dx, dz, 1.0, 0.0, ctx->d_vq1, 1.0, 1.0, res);
    else
        G_t(ctx, ctx->d_Graw, dz, 1.0, 1.0, res);
    if (p > 0) {
        hyp_gemv_t(ctx, p, n, ctx->d_A, ctx->lda, dy, 1.0, 1.0, res);
        // res.y = b tau - A x
        axpbypcz_dev_kernel<<<vgrid(ctx, p), 256, 0, ctx->stream>>>(p, res + n, 0.0, nullptr, 0.0, nullptr,
                                                                 1.0, dir + tau_idx, b);
        ctx->launches++;
        hyp_gemv_n(ctx, p, n, ctx->d_A, ctx->lda, dx, -1.0, 1.0, res + n);
    }
    if (q > 0) {
        // res.z = h tau - s - G x
        if (!fuse_g) G_n(ctx, ctx->d_Graw, dx, ctx->d_vq1);
        if (hyp_row_sharded(ctx)) hyp_replicate_q(ctx, ctx->d_vq1);
        axpbypcz_dev_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, res + n + p, -1.0, ds, -1.0, ctx->d_vq1,
                                                                 1.0, dir + tau_idx, h);
        ctx->launches++;
        // res.s = H prim + dual   (hess_prod_slow! = hess_prod! for these cones)
        TimeScope ts(ctx, T_CONE_PROD);
        const double* prim = ds;
        const double* dual = dz;
        if (ctx->any_dual) {
            primal_dual_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, dz, ds, ctx->d_vq2,
                                                                    ctx->d_vq3);
            ctx->launches++;
            prim = ctx->d_vq2;
            dual = ctx->d_vq3;
        }
        hyp_cones_prod(ctx, ctx->d_vq4, prim, 1, q, q, HYP_PROD_HESS, 0);
        if (hyp_row_sharded(ctx)) hyp_replicate_q(ctx, ctx->d_vq4);
        hyp_lincomb3(ctx, q, res + tau_idx + 1, 1.0, ctx->d_vq4, 1.0, dual, 0.0, nullptr);
    }
    // res.tau = -c'x - b'y - h'z - kap ; res.kap = mu/tau^2 * tau_dir + kap_dir
    hyp_dot(ctx, dim3, ctx->d_cbh, dir, ctx->d_scalars + 3, false);
    const double mu_tt = ctx->mu / ctx->tau_bar / ctx->tau_bar;
    lhs_tail_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_scalars, dir, res, tau_idx, kap_idx, mu_tt);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

// max |v_i| into out[0] (out must be zeroed): non-negative doubles order like their bit patterns
__global__ void absmax_kernel(int64_t len, const double* __restrict__ v, double* __restrict__ out) {
    double m = 0.0;
    bool nan = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double a = fabs(v[i]);
        if (a != a) nan = true;
        m = fmax(m, a);
    }
    if (nan) m = __longlong_as_double(0x7ff8000000000000LL);
    unsigned long long bits = (unsigned long long)__double_as_longlong(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if ((threadIdx.x & 31) == 0 && bits) atomicMax((unsigned long long*)out, bits);
}

void absmax(hyp_ctx* ctx, int64_t len, const double* v, double* out) {
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double), ctx->stream));
    if (len <= 0) return;
    absmax_kernel<<<vgrid(ctx, len), 256, 0, ctx->stream>>>(len, v, out);
    ctx->launches++;
}

// Residuals of calc_convergence_params (Solvers.jl:425-483) for the full Point `pt` on the device:
//   xres = -(G'z + A'y + c tau), yres = A x - b tau, zres = s + G x - h tau, and ten scalars in
//   st: |G'z + A'y|_inf, |.. + c tau|_inf, |A x|_inf, |A x - b tau|_inf, |s + G x|_inf,
//   |s + G x - h tau|_inf, c'x, b'y, h'z, z's.
void calc_residuals_dev(hyp_ctx* ctx, const double* pt, double* xres, double* yres, double* zres, double* st) {
    const int64_t n = ctx->n, p = ctx->p, q = ctx->q;
    const int64_t tau_idx = n + p + q;
    const double* px = pt;
    const double* py = pt + n;
    const double* pz = pt + n + p;
    const double* ps = pt + tau_idx + 1;
    const double* c = ctx->d_cbh;
    const double* b = ctx->d_cbh + n;
    const double* h = ctx->d_cbh + n + p;
    CUDA_TRY(cudaMemsetAsync(st, 0, 10 * sizeof(double), ctx->stream));
    {
        TimeScope ts(ctx, T_GEMV);
        const bool fuse_g = !hyp_row_sharded(ctx) && q > 0 && n > 0;
        if (fuse_g) hyp_gemv_nt(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, px, pz, 1.0, 0.0, ctx->d_vq1, 1.0, 0.0, ctx->d_t);
        if (n > 0) {
            if (fuse_g) {
            } else if (q > 0) G_t(ctx, ctx->d_Graw, pz, 1.0, 0.0, ctx->d_t);
            else hyp_fill(ctx, n, ctx->d_t, 0.0);
            if (p > 0) hyp_gemv_t(ctx, p, n, ctx->d_A, ctx->lda, py, 1.0, 1.0, ctx->d_t);
        }
        if (p > 0) hyp_gemv_n(ctx, p, n, ctx->d_A, ctx->lda, px, 1.0, 0.0, ctx->d_vp1);
        if (q > 0) {
            if (fuse_g) {
            } else if (n > 0) G_n(ctx, ctx->d_Graw, px, ctx->d_vq1);
            else hyp_fill(ctx, q, ctx->d_vq1, 0.0);
            if (hyp_row_sharded(ctx)) hyp_replicate_q(ctx, ctx->d_vq1);
        }
    }
    TimeScope tv(ctx, T_VEC);
    if (n > 0) {
        absmax(ctx, n, ctx->d_t, st + 0);
        axpbypcz_dev_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, ctx->d_t, 1.0, ctx->d_t, 0.0, nullptr, 1.0,
                                                                 pt + tau_idx, c);
        ctx->launches++;
        absmax(ctx, n, ctx->d_t, st + 1);
        hyp_lincomb3(ctx, n, xres, -1.0, ctx->d_t, 0.0, nullptr, 0.0, nullptr);
        hyp_dot(ctx, n, c, px, st + 6, false);
    }
    if (p > 0) {
        absmax(ctx, p, ctx->d_vp1, st + 2);
        axpbypcz_dev_kernel<<<vgrid(ctx, p), 256, 0, ctx->stream>>>(p, yres, 1.0, ctx->d_vp1, 0.0, nullptr, -1.0,
                                                                 pt + tau_idx, b);
        ctx->launches++;
        absmax(ctx, p, yres, st + 3);
        hyp_dot(ctx, p, b, py, st + 7, false);
    }
    if (q > 0) {
        hyp_lincomb3(ctx, q, ctx->d_vq1, 1.0, ctx->d_vq1, 1.0, ps, 0.0, nullptr);
        absmax(ctx, q, ctx->d_vq1, st + 4);
        axpbypcz_dev_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, zres, 1.0, ctx->d_vq1, 0.0, nullptr, -1.0,
                                                                 pt + tau_idx, h);
        ctx->launches++;
        absmax(ctx, q, zres, st + 5);
        hyp_dot(ctx, q, h, pz, st + 8, false);
        hyp_dot(ctx, q, pz, ps, st + 9, false);
    }
    CUDA_TRY(cudaGetLastError());
}

// ---- two-column solves (hyp_solve_system_multi / hyp_apply_lhs_multi) -----------------------------------------------
// The pair {cent, pred} and the pair {centadj, predadj} of one iteration are independent (combined.jl:67-79): solved
// together, every pass over G and every triangular sweep serves two right-hand sides.  Supported for the reduced model
// the default preprocessing produces (p = 0), the Cholesky factor, one rank (or column sharding); anything else
// falls back to one column at a time in the C entry points.  Column v of every product below is bit-identical to the
// single-column call (same kernels' thread mappings and reduction orders).
bool hyp_multi_supported(hyp_ctx* ctx, int ncols) {
    return ncols >= 2 && ctx->p == 0 && ctx->solver_kind == 0 && ctx->fact_kind == 0 && !ctx->d_Q && !hyp_row_sharded(ctx) &&
           ctx->q > 0 && ctx->n > 0 && !getenv("HYP_NO_MULTI");
}

struct Multi2 {
    double *sub_rhs[2], *sub_sol[2], *t[2], *Gx[2], *HGx[2], *vq1[2], *vq2[2], *vq3[2], *vq4[2];
};

Multi2 multi_buffers(hyp_ctx* ctx) {
    const int64_t n = ctx->n, q = ctx->q, dim3 = n + ctx->p + q;
    const int64_t e = [](int64_t v) { return (v + 1) & ~(int64_t)1; }(1);   // keep every buffer 16-byte aligned
    (void)e;
    auto ev = [](int64_t v) { return (v + 1) & ~(int64_t)1; };
    const int64_t per = 2 * ev(dim3) + ev(n) + 6 * ev(q);
    if (ctx->multi_doubles < 2 * per) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        dfree(ctx->d_multi);
        dalloc(&ctx->d_multi, 2 * per);
        ctx->multi_doubles = 2 * per;
    }
    Multi2 b;
    double* w = ctx->d_multi;
    // the q-vector pairs are laid out as two consecutive columns (ld = ev(q)) so that hyp_cones_prod can take both at once
    for (int v = 0; v < 2; v++) b.sub_rhs[v] = w + v * ev(dim3);
    w += 2 * ev(dim3);
    for (int v = 0; v < 2; v++) b.sub_sol[v] = w + v * ev(dim3);
    w += 2 * ev(dim3);
    for (int v = 0; v < 2; v++) b.t[v] = w + v * ev(n);
    w += 2 * ev(n);
    double** qs[6] = {b.Gx, b.HGx, b.vq1, b.vq2, b.vq3, b.vq4};
    for (auto arr : qs) {
        for (int v = 0; v < 2; v++) arr[v] = w + v * ev(q);
        w += 2 * ev(q);
    }
    return b;
}

void solve_system_pair_dev(hyp_ctx* ctx, double* sol, const double* rhs, int64_t ld) {
    const int64_t n = ctx->n, q = ctx->q;
    const int64_t dim3 = n + q, tau_idx = dim3, kap_idx = dim3 + q + 1;
    const int64_t ldq = (q + 1) & ~(int64_t)1;
    Multi2 b = multi_buffers(ctx);
    const double* rz[2] = {rhs + n, rhs + ld + n};
    const double* rs[2] = {rhs + tau_idx + 1, rhs + ld + tau_idx + 1};
    for (int v = 0; v < 2; v++) hyp_copy(ctx, n, b.sub_rhs[v], rhs + v * ld);
    {
        // setup_rhs3 (qrchol.jl:16-37) for both columns
        TimeScope ts(ctx, T_CONE_PROD);
        const double* arr = rz[0];
        int64_t ld_arr = ld;
        if (ctx->any_dual) {
            for (int v = 0; v < 2; v++) {
                rhs3_pre_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, rz[v], rs[v], b.vq1[v]);
                ctx->launches++;
            }
            arr = b.vq1[0];
            ld_arr = ldq;
        }
        hyp_cones_prod(ctx, b.vq2[0], arr, 2, ldq, ld_arr, HYP_PROD_BLOCK, 0);
        for (int v = 0; v < 2; v++) {
            rhs3_post_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, b.vq2[v], rs[v], b.sub_rhs[v] + n);
            ctx->launches++;
        }
    }
    // solve_subsystem3 (qrchol.jl:39-85 with p = 0, Ap_Q = I): t = x + G'z ; x = S^-1 t ; z = H G x - z
    for (int v = 0; v < 2; v++) {
        hyp_copy(ctx, dim3, b.sub_sol[v], b.sub_rhs[v]);
        hyp_copy(ctx, n, b.t[v], b.sub_sol[v]);
    }
    if (hyp_gemv2_ok(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, b.sub_sol[0] + n, b.sub_sol[1] + n))
        hyp_gemv_t2(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, b.sub_sol[0] + n, b.sub_sol[1] + n, 1.0, 1.0, b.t[0], b.t[1]);
    else
        for (int v = 0; v < 2; v++) hyp_gemv_t(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, b.sub_sol[v] + n, 1.0, 1.0, b.t[v]);
    hyp_trsv_upper2(ctx, ctx->d_F, ctx->lds, ctx->nmp, ctx->d_Dinv, b.t[0], b.t[1] - b.t[0], true);
    hyp_trsv_upper2(ctx, ctx->d_F, ctx->lds, ctx->nmp, ctx->d_Dinv, b.t[0], b.t[1] - b.t[0], false);
    for (int v = 0; v < 2; v++) hyp_copy(ctx, n, b.sub_sol[v], b.t[v]);
    if (hyp_gemv2_ok(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, b.sub_sol[0], b.sub_sol[1]))
        hyp_gemv_n2(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, b.sub_sol[0], b.sub_sol[1], 1.0, 0.0, b.Gx[0], b.Gx[1]);
    else
        for (int v = 0; v < 2; v++) hyp_gemv_n(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, b.sub_sol[v], 1.0, 0.0, b.Gx[v]);
    {
        TimeScope ts(ctx, T_CONE_PROD);
        hyp_cones_prod(ctx, b.HGx[0], b.Gx[0], 2, ldq, ldq, HYP_PROD_BLOCK, 0);
    }
    for (int v = 0; v < 2; v++) hyp_axpby(ctx, q, 1.0, b.HGx[v], -1.0, b.sub_sol[v] + n);
    // tau lift, s lift, kap (common.jl:129-182) column by column: O(q) work
    const double mu_tt = ctx->mu / ctx->tau_bar / ctx->tau_bar;
    for (int v = 0; v < 2; v++) {
        double* so = sol + v * ld;
        const double* rh = rhs + v * ld;
        hyp_dot(ctx, dim3, ctx->d_cbh, b.sub_sol[v], ctx->d_scalars + 0, false);
        tau_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_scalars, rh, so, tau_idx, kap_idx, mu_tt);
        ctx->launches++;
        hyp_axpy_dev(ctx, dim3, so, 1.0, b.sub_sol[v], ctx->d_scalars + 2, 1.0, ctx->d_const_sol);
        s_lift_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_scalars, b.Gx[v], ctx->d_Gx_const, ctx->d_cbh + n, rz[v],
                                                           so + tau_idx + 1);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
}

void apply_lhs_pair_dev(hyp_ctx* ctx, double* res, const double* dir, int64_t ld) {
    const int64_t n = ctx->n, q = ctx->q;
    const int64_t dim3 = n + q, tau_idx = dim3, kap_idx = dim3 + q + 1;
    const int64_t ldq = (q + 1) & ~(int64_t)1;
    Multi2 b = multi_buffers(ctx);
    const double* c = ctx->d_cbh;
    const double* h = ctx->d_cbh + n;
    const double *dx[2], *dz[2], *ds[2];
    double* r[2];
    for (int v = 0; v < 2; v++) {
        dx[v] = dir + v * ld;
        dz[v] = dx[v] + n;
        ds[v] = dx[v] + tau_idx + 1;
        r[v] = res + v * ld;
        // res.x = c tau (+ G'z below)
        axpbypcz_dev_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, r[v], 0.0, nullptr, 0.0, nullptr, 1.0, dx[v] + tau_idx, c);
        ctx->launches++;
    }
    if (hyp_gemv2_ok(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, dx[0], dx[1]) && ((uintptr_t)dz[0] % 16 == 0) &&
        ((uintptr_t)dz[1] % 16 == 0))
        hyp_gemv_nt2(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, dx[0], dx[1], dz[0], dz[1], 1.0, 0.0, b.vq1[0], b.vq1[1], 1.0, 1.0,
                     r[0], r[1]);
    else
        for (int v = 0; v < 2; v++)
            hyp_gemv_nt(ctx, ctx->qloc, n, ctx->d_Graw, ctx->ldg, dx[v], dz[v], 1.0, 0.0, b.vq1[v], 1.0, 1.0, r[v]);
    for (int v = 0; v < 2; v++) {
        // res.z = h tau - s - G x
        axpbypcz_dev_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, r[v] + n, -1.0, ds[v], -1.0, b.vq1[v], 1.0, dx[v] + tau_idx, h);
        ctx->launches++;
    }
    {
        // res.s = H prim + dual
        TimeScope ts(ctx, T_CONE_PROD);
        const double* prim = ds[0];
        int64_t ld_prim = ld;
        if (ctx->any_dual) {
            for (int v = 0; v < 2; v++) {
                primal_dual_kernel<<<vgrid(ctx, q), 256, 0, ctx->stream>>>(q, ctx->d_row_dual, dz[v], ds[v], b.vq2[v], b.vq3[v]);
                ctx->launches++;
            }
            prim = b.vq2[0];
            ld_prim = ldq;
        }
        hyp_cones_prod(ctx, b.vq4[0], prim, 2, ldq, ld_prim, HYP_PROD_HESS, 0);
        for (int v = 0; v < 2; v++)
            hyp_lincomb3(ctx, q, r[v] + tau_idx + 1, 1.0, b.vq4[v], 1.0, ctx->any_dual ? b.vq3[v] : dz[v], 0.0, nullptr);
    }
    const double mu_tt = ctx->mu / ctx->tau_bar / ctx->tau_bar;
    for (int v = 0; v < 2; v++) {
        hyp_dot(ctx, dim3, ctx->d_cbh, dx[v], ctx->d_scalars + 3, false);
        lhs_tail_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_scalars, dx[v], r[v], tau_idx, kap_idx, mu_tt);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
}

void solve_system_multi_dev(hyp_ctx* ctx, double* sol, const double* rhs, int ncols, int64_t ld) {
    int j = 0;
    for (; j + 1 < ncols; j += 2) solve_system_pair_dev(ctx, sol + j * ld, rhs + j * ld, ld);
    for (; j < ncols; j++) solve_system_dev(ctx, sol + j * ld, rhs + j * ld);
}
void apply_lhs_multi_dev(hyp_ctx* ctx, double* res, const double* dir, int ncols, int64_t ld) {
    int j = 0;
    for (; j + 1 < ncols; j += 2) apply_lhs_pair_dev(ctx, res + j * ld, dir + j * ld, ld);
    for (; j < ncols; j++) apply_lhs_dev(ctx, res + j * ld, dir + j * ld);
}

// ---- packed upper triangle for the Schur reduction ------------------------------------------
// Only the upper triangle of the partial Schur matrices is meaningful (outer_prod! = syrk 'U', dense.jl:80-86), so the
// allreduce moves the upper-triangular 128-column blocks only: block column b (columns [128 b, 128 b + 128)) keeps its
// rows [0, 128 (b + 1)), stored contiguously one block column after the other - 0.41 GB instead of the 0.8 GB square at
// m = 10000.  pack: S -> buf, unpack: buf -> S.
__global__ void tri_pack_kernel(int64_t m, const double* __restrict__ S, int64_t lds, double* __restrict__ buf, int unpack) {
    const int64_t b = blockIdx.y;                                   // block column
    const int64_t c0 = b * 128, nc = min((int64_t)128, m - c0);
    const int64_t nr = min(m, c0 + 128);                            // rows kept
    const int64_t off = 128 * 128 * (b * (b + 1) / 2);              // doubles in front of this block column
    const int64_t total = nr * nc;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx % nr, c = idx / nr;
        double* sp = const_cast<double*>(S) + r + (c0 + c) * lds;
        if (unpack) *sp = buf[off + idx];
        else buf[off + idx] = *sp;
    }
}

int64_t tri_packed_len(int64_t m) {
    const int64_t nb = (m + 127) / 128;
    return 128 * 128 * (nb * (nb + 1) / 2);
}

void allreduce_upper(hyp_ctx* ctx, double* S, int64_t lds, int64_t m, double* buf) {
    if (m <= 0) return;
    const int64_t nb = (m + 127) / 128;
    dim3 grid(std::max(1, std::min(64, ctx->sm_count)), (unsigned)nb);
    {
        TimeScope ts(ctx, T_VEC);
        tri_pack_kernel<<<grid, 256, 0, ctx->stream>>>(m, S, lds, buf, 0);
        ctx->launches++;
    }
    hyp_allreduce_sum(ctx, buf, tri_packed_len(m));
    TimeScope ts(ctx, T_VEC);
    tri_pack_kernel<<<grid, 256, 0, ctx->stream>>>(m, S, lds, buf, 1);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

// Schur assembly + factorisation (update_lhs_fact, qrchol.jl:201-257)
int update_lhs_fact(hyp_ctx* ctx) {
    const int64_t nmp = ctx->nmp, p = ctx->p;
    if (ctx->col_shard && ctx->nranks > 1) {
        // single giant cone: hess_prod! is independent per column of G_k (hypoperlogdettri.jl:196-237,
        // possemideftri.jl:126-142), so rank r takes the columns J_r = [c_lo, c_hi) of GQ2, forms (H GQ2)[:, J_r] for
        // every cone (the hess_prod! + mul! branch qrchol.jl:240-246 applied to all cones) and the upper part of the
        // column panel S[0:c_hi, J_r] = GQ2[:, 0:c_hi]' (H GQ2)[:, J_r]; the panels are contiguous column ranges of the
        // column-major S, so one in-place ncclAllGather completes S on every rank - no reduction.
        const int64_t cw = ctx->col_shard_width;
        const int64_t c_lo = std::min<int64_t>(nmp, (int64_t)ctx->rank * cw), c_hi = std::min<int64_t>(nmp, c_lo + cw);
        const double* GQ2 = ctx->d_GQ + p * ctx->ldg;
        if (c_hi > c_lo) {
            {
                TimeScope ts(ctx, T_SQRT_PREPASS);
                hyp_cones_prod(ctx, ctx->d_HG + c_lo * ctx->ldg, GQ2 + c_lo * ctx->ldg, c_hi - c_lo, ctx->ldg, ctx->ldg,
                               HYP_PROD_BLOCK, 0);
            }
            TimeScope ts(ctx, T_SYRK);
            hyp_gemm_tn(ctx, GQ2, ctx->ldg, ctx->d_HG + c_lo * ctx->ldg, ctx->ldg, ctx->qloc, c_hi, c_hi - c_lo,
                        ctx->d_S + c_lo * ctx->lds, ctx->lds, 1.0, 0.0);
        }
        hyp_allgather_inplace(ctx, ctx->d_S, cw * ctx->lds);
    } else {
    // digit-sliced SYRK on second-order-cone models: pre-pass and slicing fused, H^{1/2} G is never written
    int sliced = 0;       // 1: the digit slices are written, 2: H^{1/2} G is stored and the column exponents are ready
    const bool want_i8 = ctx->qloc > 0 && ctx->syrk_mode == 1 && !ctx->d_PG;
    if (want_i8) {
        if (!ctx->d_digits) {
            ctx->ldd = round_up(std::max<int64_t>(ctx->qloc, 16), 16);
            dalloc(&ctx->d_digits, 8 * ctx->ldd * nmp);
            dalloc(&ctx->d_expo, nmp);
            dalloc(&ctx->d_dscale, nmp);
        }
        sliced = hyp_cones_prepass_sliced(ctx, ctx->d_digits, ctx->ldd, ctx->ldd * nmp, ctx->d_expo, ctx->d_dscale);
    }
    if (!sliced) hyp_cones_schur_prepass(ctx);
    const bool have_expo = sliced == 2;
    if (sliced == 2) sliced = 0;
    {
        TimeScope ts(ctx, T_SYRK);
        const double* P = ctx->d_HG;
        if (ctx->d_PG) {
            // mixed model: P rows = H^{1/2} G (sqrt cones) / G (log-det family), R rows = HG
            hyp_build_pg(ctx);
            P = ctx->d_PG;
        }
        if (ctx->qloc > 0 && ctx->syrk_mode == 1 && !ctx->d_PG) {
            // FP64-accurate SYRK on the int8 tensor pipe (tcgen05): slice HG, multiply digit pairs
            if (!ctx->d_digits) {
                ctx->ldd = round_up(std::max<int64_t>(ctx->qloc, 16), 16);
                dalloc(&ctx->d_digits, 8 * ctx->ldd * nmp);
                dalloc(&ctx->d_expo, nmp);
                dalloc(&ctx->d_dscale, nmp);
            }
            if (!sliced)
                hyp_ozaki_slice(ctx, ctx->d_HG, ctx->ldg, ctx->qloc, nmp, ctx->d_digits, ctx->ldd, ctx->ldd * nmp,
                                ctx->d_expo, ctx->d_dscale, have_expo);
            hyp_ozaki_syrk(ctx, ctx->d_digits, ctx->ldd, ctx->ldd * nmp, ctx->d_expo, ctx->d_dscale, ctx->qloc, nmp, ctx->d_S,
                           ctx->lds, 1.0, 0.0);
        } else if (ctx->qloc > 0 && ctx->syrk_mode == 1 && ctx->d_PG && nmp >= 16 && hyp_ozaki_pair64_ready(ctx) &&
                   !getenv("HYP_K2_DMMA")) {
            // mixed / log-det models (the hess_prod! + mul! branch, qrchol.jl:240-246): S = P' (HG) with two DIFFERENT
            // operands, also on the int8 tensor pipe - both are cut into digit slices (their own column scales) and the
            // CTA-pair kernel takes the A-side tiles from P and the B-side tiles from HG
            if (!ctx->d_digits) {
                ctx->ldd = round_up(std::max<int64_t>(ctx->qloc, 16), 16);
                dalloc(&ctx->d_digits, 8 * ctx->ldd * nmp);
                dalloc(&ctx->d_expo, nmp);
                dalloc(&ctx->d_dscale, nmp);
            }
            if (!ctx->d_digitsP) {
                dalloc(&ctx->d_digitsP, 8 * ctx->ldd * nmp);
                dalloc(&ctx->d_expoP, nmp);
                dalloc(&ctx->d_dscaleP, nmp);
            }
            hyp_ozaki_slice(ctx, P, ctx->ldg, ctx->qloc, nmp, ctx->d_digitsP, ctx->ldd, ctx->ldd * nmp, ctx->d_expoP, ctx->d_dscaleP);
            hyp_ozaki_slice(ctx, ctx->d_HG, ctx->ldg, ctx->qloc, nmp, ctx->d_digits, ctx->ldd, ctx->ldd * nmp, ctx->d_expo,
                            ctx->d_dscale);
            hyp_ozaki_syrk(ctx, ctx->d_digitsP, ctx->ldd, ctx->ldd * nmp, ctx->d_expoP, ctx->d_dscaleP, ctx->qloc, nmp, ctx->d_S,
                           ctx->lds, 1.0, 0.0, ctx->d_digits, ctx->d_dscale);
        } else if (ctx->qloc > 0)
            hyp_atb_upper(ctx, P, ctx->ldg, ctx->d_HG, ctx->ldg, ctx->qloc, nmp, ctx->d_S, ctx->lds, 1.0, 0.0);
        else
            CUDA_TRY(cudaMemsetAsync(ctx->d_S, 0, (size_t)ctx->lds * nmp * 8, ctx->stream));
    }
    if (ctx->nranks > 1) {
        // the factor buffer is idle until the copy below: it holds the packed upper triangle during the reduction
        // (the ragged last block column is padded to 128 x 128-row blocks, still within lds * nmp doubles for m >= 256)
        static int packed = -1;
        if (packed < 0) packed = getenv("HYP_ALLREDUCE_FULL") ? 0 : 1;
        if (packed && tri_packed_len(nmp) <= ctx->lds * nmp) allreduce_upper(ctx, ctx->d_S, ctx->lds, nmp, ctx->d_F);
        else hyp_allreduce_sum(ctx, ctx->d_S, ctx->lds * nmp);
    }
    }
    (void)p;
    // posdef_fact_copy! (dense.jl:194-215): Cholesky -> Bunch-Kaufman -> shifted Bunch-Kaufman
    int info = 0;
    CUDA_TRY(cudaMemcpyAsync(ctx->d_F, ctx->d_S, (size_t)ctx->lds * nmp * 8, cudaMemcpyDeviceToDevice,
                             ctx->stream));
    hyp_potrf_upper(ctx, ctx->d_F, ctx->lds, nmp, ctx->d_Dinv, ctx->d_info);
    CUDA_TRY(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->fact_kind = 0;
    if (info == 0) return 0;
    for (int attempt = 1; attempt <= 2; attempt++) {
        TimeScope ts(ctx, T_LDLT);
        CUDA_TRY(cudaMemcpyAsync(ctx->d_F, ctx->d_S, (size_t)ctx->lds * nmp * 8, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
        if (attempt == 2) hyp_increase_diag(ctx, ctx->d_F, ctx->lds, nmp);
        if (!ctx->d_ipiv) {
            dalloc(&ctx->d_ipiv, 3 * nmp + 8);
            dalloc(&ctx->d_ldl_work, 4 * nmp + 64);
        }
        hyp_ldlt_factor(ctx, ctx->d_F, ctx->lds, nmp, ctx->d_ipiv, ctx->d_info);
        CUDA_TRY(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->fact_kind = attempt;
        if (info == 0) return 1;
    }
    return 2;
}


// ---- SymIndefDense (symindef.jl:203-271) ---------------------------------------------------
// out[i + (col0 + r) * ldo] = M[r + i * ldm] : M' into the columns col0.. of the big matrix
__global__ void transpose_into_kernel(double* __restrict__ out, int64_t ldo, int64_t col0,
                                      const double* __restrict__ M, int64_t ldm, int64_t rows, int64_t n) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += 8) {
        int64_t r = r0 + threadIdx.x, i = i0 + dy;
        tile[dy][threadIdx.x] = (r < rows && i < n) ? M[r + i * ldm] : 0.0;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += 8) {
        int64_t i = i0 + threadIdx.x, r = r0 + dy;
        if (r < rows && i < n) out[i + (col0 + r) * ldo] = tile[threadIdx.x][dy];
    }
}

// arr[r + j * q] = 1 where j is the index of row r inside its cone
__global__ void block_identity_kernel(int64_t q, int64_t maxdim, const int* __restrict__ row_cone,
                                      const int64_t* __restrict__ coff, double* __restrict__ arr) {
    for (int64_t j = blockIdx.y; j < maxdim; j += gridDim.y)
        for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < q;
             r += (int64_t)gridDim.x * blockDim.x)
            arr[r + j * q] = (r - coff[row_cone[r]] == j) ? 1.0 : 0.0;
}

// dst[(z0 + r) + (z0 + off_k + j) * ld] = sign * prod[r + j * q] for j < dim_k (k = cone of row r)
__global__ void scatter_blocks_kernel(int64_t q, int64_t maxdim, const int* __restrict__ row_cone,
                                      const int64_t* __restrict__ coff, const int64_t* __restrict__ cdim,
                                      const double* __restrict__ prod, double sign, double* __restrict__ dst,
                                      int64_t ld, int64_t z0) {
    for (int64_t j = blockIdx.y; j < maxdim; j += gridDim.y)
        for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < q;
             r += (int64_t)gridDim.x * blockDim.x) {
            const int k = row_cone[r];
            if (j < cdim[k]) dst[(z0 + r) + (z0 + coff[k] + j) * ld] = sign * prod[r + j * q];
        }
}

// packed explicit blocks: out[boff_k + i + j * dim_k] = prod[off_k + i + j * q]
__global__ void gather_blocks_kernel(int64_t q, int64_t maxdim, const int* __restrict__ row_cone,
                                     const int64_t* __restrict__ coff, const int64_t* __restrict__ cdim,
                                     const int64_t* __restrict__ boff, const double* __restrict__ prod,
                                     double* __restrict__ out) {
    for (int64_t j = blockIdx.y; j < maxdim; j += gridDim.y)
        for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < q;
             r += (int64_t)gridDim.x * blockDim.x) {
            const int k = row_cone[r];
            if (j < cdim[k]) out[boff[k] + (r - coff[k]) + j * cdim[k]] = prod[r + j * q];
        }
}

void ensure_block_buffers(hyp_ctx* ctx) {
    if (ctx->d_blk_arr) return;
    const int64_t q = ctx->q;
    int64_t maxdim = 1;
    std::vector<int> rc((size_t)std::max<int64_t>(q, 1), 0);
    for (int k = 0; k < ctx->K; k++) {
        maxdim = std::max(maxdim, ctx->h_cone_dim[k]);
        for (int64_t r = ctx->h_cone_off[k]; r < ctx->h_cone_off[k + 1]; r++) rc[r] = k;
    }
    ctx->blk_maxdim = maxdim;
    dalloc(&ctx->d_row_cone, q);
    if (q) CUDA_TRY(cudaMemcpyAsync(ctx->d_row_cone, rc.data(), q * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dalloc(&ctx->d_blk_arr, 2 * q * maxdim);
}

// explicit Hessian-type blocks of every cone (update_hess / update_inv_hess, K10): the oracle
// applied to the identity pattern; result (q x maxdim, ld q) in d_blk_arr + q * maxdim
double* cone_blocks_dev(hyp_ctx* ctx, int mode) {
    ensure_block_buffers(ctx);
    const int64_t q = ctx->q, md = ctx->blk_maxdim;
    if (ctx->nranks > 1) throw HypError{"explicit cone Hessian blocks are single-rank only"};
    dim3 grid(std::max(1, std::min(ceil_div(q, 256), ctx->sm_count * 2)), (unsigned)std::min<int64_t>(md, 65535));
    block_identity_kernel<<<grid, 256, 0, ctx->stream>>>(q, md, ctx->d_row_cone, ctx->d_cone_off, ctx->d_blk_arr);
    ctx->launches++;
    double* prod = ctx->d_blk_arr + q * md;
    hyp_cones_prod(ctx, prod, ctx->d_blk_arr, md, q, q, mode, 0);
    return prod;
}

void symindef_setup(hyp_ctx* ctx) {
    if (ctx->nranks > 1) throw HypError{"SymIndefDense is single-rank only"};
    const int64_t n = ctx->n, p = ctx->p, q = ctx->q, N3 = n + p + q;
    ctx->ld3 = round_up(std::max<int64_t>(N3, 2), 2);
    dfree(ctx->d_digits);
    dfree(ctx->d_expo);
    dfree(ctx->d_dscale);
    dfree(ctx->d_L3);
    dfree(ctx->d_F3);
    dalloc(&ctx->d_L3, ctx->ld3 * std::max<int64_t>(N3, 1));
    dalloc(&ctx->d_F3, ctx->ld3 * std::max<int64_t>(N3, 1));
    dim3 blk(32, 8);
    if (p > 0 && n > 0) {
        dim3 grid(ceil_div(p, 32), ceil_div(n, 32));
        transpose_into_kernel<<<grid, blk, 0, ctx->stream>>>(ctx->d_L3, ctx->ld3, n, ctx->d_A, ctx->lda, p, n);
        ctx->launches++;
    }
    if (q > 0 && n > 0) {
        dim3 grid(ceil_div(q, 32), ceil_div(n, 32));
        transpose_into_kernel<<<grid, blk, 0, ctx->stream>>>(ctx->d_L3, ctx->ld3, n + p, ctx->d_Graw, ctx->ldg, q, n);
        ctx->launches++;
    }
    dfree(ctx->d_ipiv);
    dfree(ctx->d_ldl_work);
    dalloc(&ctx->d_ipiv, 3 * N3 + 8);
    dalloc(&ctx->d_ldl_work, 4 * N3 + 64);
    ensure_block_buffers(ctx);
    // rhs_const.z = h for this solver (common.jl:203-205; no H h product)
    if (q) hyp_copy(ctx, q, ctx->d_const_rhs + n + p, ctx->d_cbh + n + p);
    CUDA_TRY(cudaGetLastError());
}

// update_lhs of SymIndefDense: z/z block = -inv_hess (primal) / -hess (dual), symm_fact_copy! chain
int symindef_update(hyp_ctx* ctx) {
    const int64_t n = ctx->n, p = ctx->p, q = ctx->q, N3 = n + p + q;
    const size_t bytes = (size_t)ctx->ld3 * N3 * 8;
    int info = 0, rc = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        CUDA_TRY(cudaMemcpyAsync(ctx->d_F3, ctx->d_L3, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (q > 0) {
            double* prod = cone_blocks_dev(ctx, HYP_PROD_BLOCK_INV);
            dim3 grid(std::max(1, std::min(ceil_div(q, 256), ctx->sm_count * 2)),
                      (unsigned)std::min<int64_t>(ctx->blk_maxdim, 65535));
            scatter_blocks_kernel<<<grid, 256, 0, ctx->stream>>>(q, ctx->blk_maxdim, ctx->d_row_cone, ctx->d_cone_off,
                                                               ctx->d_cone_dim, prod, -1.0, ctx->d_F3, ctx->ld3, n + p);
            ctx->launches++;
        }
        if (attempt == 1) hyp_increase_diag(ctx, ctx->d_F3, ctx->ld3, N3);   // dense.jl:170-184
        {
            TimeScope ts(ctx, T_LDLT);
            hyp_ldlt_factor(ctx, ctx->d_F3, ctx->ld3, N3, ctx->d_ipiv, ctx->d_info);
        }
        CUDA_TRY(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->fact_kind = 1 + attempt;
        if (info == 0) return rc;
        rc = (attempt == 0) ? 1 : 2;
    }
    return rc;
}

template <typename F>
int guarded(hyp_ctx* ctx, F&& f) {
    if (!ctx) return -1;
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) throw HypError{std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
        return f();
    } catch (HypError& e) {
        ctx->last_error = e.msg;
        cudaGetLastError();
        return -1;
    } catch (std::exception& e) {
        ctx->last_error = e.what();
        return -1;
    } catch (...) {
        ctx->last_error = "unknown exception";
        return -1;
    }
}

void need_model(hyp_ctx* ctx) {
    if (!ctx->model_loaded) throw HypError{"no model loaded (call hyp_load_model first)"};
}

}  // namespace

// P operand of the mixed Schur product: copy of HG on sqrt-cone rows, GQ2 on the other rows
__global__ void build_pg_kernel(int64_t qloc, int64_t nmp, int64_t ldg, const uint8_t* __restrict__ row_ns,
                                const double* __restrict__ GQ2, const double* __restrict__ HG,
                                double* __restrict__ PG) {
    for (int64_t j = blockIdx.y; j < nmp; j += gridDim.y)
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < qloc;
             i += (int64_t)gridDim.x * blockDim.x)
            PG[i + j * ldg] = row_ns[i] ? GQ2[i + j * ldg] : HG[i + j * ldg];
}

void hyp_build_pg(hyp_ctx* ctx) {
    dim3 grid(std::max(1, std::min(ceil_div(ctx->qloc, 256), ctx->sm_count * 4)),
              (unsigned)std::min<int64_t>(ctx->nmp, 65535));
    build_pg_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->qloc, ctx->nmp, ctx->ldg, ctx->d_row_ns,
                                                   ctx->d_GQ + ctx->p * ctx->ldg, ctx->d_HG, ctx->d_PG);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

extern "C" {

int hyp_version(void) { return 100; }

hyp_ctx* hyp_create(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
    if (prop.major != 10) {
        fprintf(stderr, "libhypatia_b200: device %d is sm_%d%d; this library is built for sm_100a only\n",
                device, prop.major, prop.minor);
        return nullptr;
    }
    hyp_ctx* ctx = new hyp_ctx();
    ctx->device = device;
    ctx->syrk_mode = 1;   // default: FP64-accurate digit slicing on tcgen05 (ozaki.cu)
    if (const char* e = getenv("HYP_SCHUR_SYRK")) ctx->syrk_mode = (strcmp(e, "dmma") == 0) ? 0 : 1;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return nullptr;
    }
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&ctx->ev_chain[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_bulk[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_near[i], cudaEventDisableTiming);
    }
    return ctx;
}

void hyp_destroy(hyp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_model(ctx);
    hyp_comm_destroy(ctx);
    for (int i = 0; i < T_NUM; i++) {
        if (ctx->timing[i].e0) cudaEventDestroy(ctx->timing[i].e0);
        if (ctx->timing[i].e1) cudaEventDestroy(ctx->timing[i].e1);
    }
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_chain[i]) cudaEventDestroy(ctx->ev_chain[i]);
        if (ctx->ev_bulk[i]) cudaEventDestroy(ctx->ev_bulk[i]);
        if (ctx->ev_near[i]) cudaEventDestroy(ctx->ev_near[i]);
    }
    if (ctx->d_chol_digits) cudaFree(ctx->d_chol_digits);
    if (ctx->d_chol_dscale) cudaFree(ctx->d_chol_dscale);
    if (ctx->d_trsv_part) cudaFree(ctx->d_trsv_part);
    if (ctx->d_trsv_pkt) cudaFree(ctx->d_trsv_pkt);
    if (ctx->d_gemm_digA) cudaFree(ctx->d_gemm_digA);
    if (ctx->d_gemm_digB) cudaFree(ctx->d_gemm_digB);
    if (ctx->d_gemm_scal) cudaFree(ctx->d_gemm_scal);
    if (ctx->d_dag_ver) cudaFree(ctx->d_dag_ver);
    if (ctx->d_dag_dbg) cudaFree(ctx->d_dag_dbg);
    cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* hyp_last_error(hyp_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
void* hyp_stream(hyp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int hyp_sync(hyp_ctx* ctx) {
    return guarded(ctx, [&] {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return 0;
    });
}

// nu of a WSOSInterpNonnegative cone = sum of the L_k in its packed data (hyp_set_cone_alpha)
static double wsos_nu(hyp_ctx* ctx, int k) {
    if ((int)ctx->h_cone_aoff.size() != ctx->K + 1) return 0.0;
    const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
    if (a1 - a0 < 2) return 0.0;
    const int nP = (int)ctx->h_cone_alpha[a0];
    double nu = 0.0;
    for (int j = 0; j < nP && a0 + 1 + j < a1; j++) nu += ctx->h_cone_alpha[a0 + 1 + j];
    return nu;
}

// nu of a LinMatrixIneq cone = the side of its matrices (first entry of its packed data)
static double lmi_nu(hyp_ctx* ctx, int k) {
    if ((int)ctx->h_cone_aoff.size() != ctx->K + 1) return 0.0;
    const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
    return a1 - a0 >= 1 ? ctx->h_cone_alpha[a0] : 0.0;
}

int hyp_load_model(hyp_ctx* ctx, int64_t n, int64_t p, int64_t q, const double* G_local, int64_t ldG,
                   const double* A, int64_t ldA, const double* c, const double* b, const double* h, int K,
                   const int* cone_type, const int64_t* cone_dim, const int* cone_dual, int cone_lo,
                   int cone_hi, const double* Ap_Q, const double* Ap_R) {
    return guarded(ctx, [&] {
        if (n < 0 || p < 0 || q < 0 || K < 0 || p > n) throw HypError{"hyp_load_model: bad dimensions"};
        if (cone_lo < 0 || cone_hi > K || cone_lo > cone_hi) throw HypError{"hyp_load_model: bad cone range"};
        if (p > 0 && !A) throw HypError{"hyp_load_model: p > 0 needs A"};
        if ((Ap_Q == nullptr) != (Ap_R == nullptr)) throw HypError{"hyp_load_model: pass both Ap_Q and Ap_R or neither"};
        // per-cone parameters staged by hyp_set_cone_params / hyp_set_cone_alpha survive the free below
        std::vector<int> hkind_in = ctx->h_cone_hkind;
        std::vector<double> hparam_in = ctx->h_cone_hparam;
        // only parameters staged SINCE the previous load count: a stale table of a previous model with the same cone
        // count must not be reused when the caller skipped hyp_set_cone_params
        const bool have_params = ctx->params_staged && (int)hkind_in.size() == K && K > 0;
        ctx->params_staged = false;
        free_model(ctx);
        ctx->h_cone_hkind = have_params ? hkind_in : std::vector<int>((size_t)K, 0);
        ctx->h_cone_hparam = have_params ? hparam_in : std::vector<double>((size_t)K, 0.0);
        ctx->n = n;
        ctx->p = p;
        ctx->q = q;
        ctx->nmp = n - p;
        ctx->K = K;
        ctx->cone_lo = cone_lo;
        ctx->cone_hi = cone_hi;
        ctx->h_cone_type.assign(cone_type, cone_type + K);
        ctx->h_cone_dim.assign(cone_dim, cone_dim + K);
        ctx->h_cone_dual.assign(K, 0);
        if (cone_dual) ctx->h_cone_dual.assign(cone_dual, cone_dual + K);
        ctx->h_cone_off.assign(K + 1, 0);
        ctx->h_cone_nu.assign(K, 0.0);
        ctx->any_dual = false;
        bool any_ns = false, any_sqrt = false;
        for (int k = 0; k < K; k++) {
            int t = cone_type[k];
            int64_t d = cone_dim[k];
            if (t < 0 || t >= HYP_NUM_CONE_TYPES || d < 1) throw HypError{"hyp_load_model: bad cone entry"};
            ctx->h_cone_off[k + 1] = ctx->h_cone_off[k] + d;
            if (ctx->h_cone_dual[k]) {
                if (!cone_allows_dual(t))
                    throw HypError{"hyp_load_model: use_dual_barrier is not defined for this cone type"};
                ctx->any_dual = true;
            }
            double side = 0;
            if (cone_is_matrix(t)) side = std::floor((std::sqrt(1.0 + 8.0 * (d - cone_mat_lead(t))) - 1) / 2 + 0.5);
            // get_nu: nonnegative.jl:40, epinormeucl.jl:42, possemideftri.jl:67, hypoperlogdettri.jl:80,
            // hyporootdettri.jl:80, epipersepspectral.jl:79, epipersquare.jl:50, hypoperlog.jl:54
            ctx->h_cone_nu[k] = t == HYP_CONE_NONNEGATIVE ? (double)d
                                : t == HYP_CONE_EPINORMEUCL ? 2.0
                                : t == HYP_CONE_POSSEMIDEFTRI ? side
                                : t == HYP_CONE_HYPOPERLOGDETTRI ? 2.0 + side
                                : t == HYP_CONE_HYPOROOTDETTRI ? 1.0 + side
                                : t == HYP_CONE_EPIPERSEPSPECTRAL_MAT ? 2.0 + side
                                : t == HYP_CONE_EPIPERSQUARE ? 2.0
                                : (t == HYP_CONE_EPINORMSPECTRAL || t == HYP_CONE_MATRIXEPIPERSQUARE)
                                    ? (double)ctx->h_cone_hkind[k] + 1.0   // epinormspectral.jl:95, matrixepipersquare.jl:101
                                : t == HYP_CONE_EPITRRELENTROPYTRI                                    // epitrrelentropytri.jl:119: 2 d + 1
                                    ? 2.0 * std::floor((std::sqrt(1.0 + 4.0 * (d - 1)) - 1) / 2 + 0.5) + 1.0
                                : t == HYP_CONE_WSOSINTERPNONNEGATIVE ? wsos_nu(ctx, k)               // wsosinterpnonnegative.jl:62
                                : t == HYP_CONE_WSOSINTERPEPINORMEUCL ? 2.0 * wsos_nu(ctx, k)         // wsosinterpepinormeucl.jl:68
                                : (t == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI || t == HYP_CONE_WSOSINTERPEPINORMONE)   // wsosinterppossemideftri.jl:66, wsosinterpepinormone.jl:88
                                    ? (double)ctx->h_cone_hkind[k] * wsos_nu(ctx, k)
                                : (t == HYP_CONE_LINMATRIXINEQ || t == HYP_CONE_POSSEMIDEFTRISPARSE)
                                    ? lmi_nu(ctx, k)   // linmatrixineq.jl:72, possemideftrisparse.jl:101: first entry of the packed data = side
                                : t == HYP_CONE_GENERALIZEDPOWER
                                    ? ((int)ctx->h_cone_aoff.size() == K + 1
                                           ? (double)(ctx->h_cone_aoff[k + 1] - ctx->h_cone_aoff[k]) + 1.0
                                           : 0.0)
                                                             : (double)d;   // HypoPerLog, EpiNormInf, EpiPerSepSpectral{VectorCSqr}, HypoGeoMean, HypoPowerMean, EpiRelEntropy: nu = dim
            if ((t == HYP_CONE_EPINORMINF || t == HYP_CONE_HYPOGEOMEAN) && d < 2)
                throw HypError{"hyp_load_model: EpiNormInf / HypoGeoMean need dimension >= 2"};
            if ((t == HYP_CONE_EPIPERSQUARE || t == HYP_CONE_HYPOPERLOG || t == HYP_CONE_EPIPERSEPSPECTRAL_MAT ||
                 t == HYP_CONE_EPIPERSEPSPECTRAL_VEC) && d < 3)
                throw HypError{"hyp_load_model: this cone type needs dimension >= 3"};
            if ((t == HYP_CONE_EPINORMSPECTRAL || t == HYP_CONE_MATRIXEPIPERSQUARE) &&
                (!have_params || ctx->h_cone_hkind[k] < 1))
                throw HypError{"hyp_load_model: EpiNormSpectral / MatrixEpiPerSquare cones need hyp_set_cone_params (number of rows d1) first"};
            if (t == HYP_CONE_EPIRELENTROPY && (d < 3 || d % 2 == 0))
                throw HypError{"hyp_load_model: EpiRelEntropy needs an odd dimension >= 3"};   // epirelentropy.jl:51-52
            if (t == HYP_CONE_EPIPERSEPSPECTRAL_MAT || t == HYP_CONE_EPIPERSEPSPECTRAL_VEC) {
                if (!have_params) throw HypError{"hyp_load_model: EpiPerSepSpectral cones need hyp_set_cone_params first"};
                const int hk = ctx->h_cone_hkind[k];
                const double hp = ctx->h_cone_hparam[k];
                if (hk < HYP_SSF_INV || hk > HYP_SSF_POWER12 || (hk == HYP_SSF_POWER12 && !(hp > 1.0 && hp <= 2.0)))
                    throw HypError{"hyp_load_model: bad separable spectral function"};
            }
            if (k >= cone_lo && k < cone_hi) {
                if (!cone_has_sqrt(t)) any_ns = true; else any_sqrt = true;
            }
        }
        if (ctx->h_cone_off[K] != q) throw HypError{"hyp_load_model: cone dimensions do not sum to q"};
        ctx->row_lo = ctx->h_cone_off[cone_lo];
        ctx->row_hi = ctx->h_cone_off[cone_hi];
        ctx->qloc = ctx->row_hi - ctx->row_lo;
        ctx->ldg = round_up(std::max<int64_t>(ctx->qloc, 2), 2);
        const int64_t nmp = ctx->nmp;

        // G panel (+ GQ = G * Ap_Q when p > 0, qrchol.jl:154)
        dalloc(&ctx->d_Graw, ctx->ldg * std::max<int64_t>(n, 1));
        upload_matrix(ctx, ctx->d_Graw, ctx->ldg, G_local, ldG, ctx->qloc, n);
        ctx->d_GQ = ctx->d_Graw;
        if (p > 0) {
            ctx->lda = round_up(p, 2);
            ctx->ldqm = round_up(n, 2);
            ctx->ldr = round_up(p, 2);
            dalloc(&ctx->d_A, ctx->lda * n);
            upload_matrix(ctx, ctx->d_A, ctx->lda, A, ldA, p, n);
        }
        if (p > 0 && Ap_Q) {
            dalloc(&ctx->d_Q, ctx->ldqm * n);
            dalloc(&ctx->d_R, ctx->ldr * p);
            upload_matrix(ctx, ctx->d_Q, ctx->ldqm, Ap_Q, n, n, n);
            upload_matrix(ctx, ctx->d_R, ctx->ldr, Ap_R, p, p, p);
            dalloc(&ctx->d_Rdinv, (int64_t)ceil_div(p, 128) * 128 * 128);
            hyp_trtri_diag(ctx, ctx->d_R, ctx->ldr, p, ctx->d_Rdinv);
            dalloc(&ctx->d_GQ, ctx->ldg * n);
            hyp_gemm_simple(ctx, false, false, ctx->qloc, n, n, ctx->d_Graw, ctx->ldg, ctx->d_Q, ctx->ldqm,
                            ctx->d_GQ, ctx->ldg);
        }
        dalloc(&ctx->d_HG, ctx->ldg * std::max<int64_t>(nmp, 1));
        if (any_ns && nmp > 0) {
            dalloc(&ctx->d_PG, ctx->ldg * nmp);
            std::vector<uint8_t> row_ns((size_t)std::max<int64_t>(ctx->qloc, 1), 0);
            for (int k = cone_lo; k < cone_hi; k++)
                if (!cone_has_sqrt(cone_type[k]))
                    for (int64_t r = ctx->h_cone_off[k]; r < ctx->h_cone_off[k + 1]; r++) row_ns[r - ctx->row_lo] = 1;
            dalloc(&ctx->d_row_ns, (int64_t)row_ns.size());
            CUDA_TRY(cudaMemcpyAsync(ctx->d_row_ns, row_ns.data(), row_ns.size(), cudaMemcpyHostToDevice, ctx->stream));
        }
        (void)any_sqrt;

        // (c, b, h) and the constant right-hand side (-c, b, h)  (common.jl:203-205)
        const int64_t dim3 = n + p + q, dim6 = n + p + 2 * q + 2;
        dalloc(&ctx->d_cbh, dim3);
        dalloc(&ctx->d_const_rhs, dim3);
        dalloc(&ctx->d_const_sol, dim3);
        dalloc(&ctx->d_sub_rhs, dim3);
        dalloc(&ctx->d_sub_sol, dim3);
        dalloc(&ctx->d_rhs, dim6);
        dalloc(&ctx->d_sol, dim6);
        if (n) CUDA_TRY(cudaMemcpyAsync(ctx->d_cbh, c, n * 8, cudaMemcpyDefault, ctx->stream));
        if (p) CUDA_TRY(cudaMemcpyAsync(ctx->d_cbh + n, b, p * 8, cudaMemcpyDefault, ctx->stream));
        if (q) CUDA_TRY(cudaMemcpyAsync(ctx->d_cbh + n + p, h, q * 8, cudaMemcpyDefault, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_const_rhs, ctx->d_cbh, dim3 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (n) {
            negate_kernel<<<vgrid(ctx, n), 256, 0, ctx->stream>>>(n, ctx->d_const_rhs, ctx->d_cbh);
            ctx->launches++;
        }

        // cone table
        dalloc(&ctx->d_cone_nu, K);
        dalloc(&ctx->d_cone_off, K + 1);
        dalloc(&ctx->d_cone_dim, K);
        dalloc(&ctx->d_cone_type, K);
        if (K) {
            CUDA_TRY(cudaMemcpyAsync(ctx->d_cone_nu, ctx->h_cone_nu.data(), K * 8, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(ctx->d_cone_off, ctx->h_cone_off.data(), (K + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(ctx->d_cone_dim, ctx->h_cone_dim.data(), K * 8, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(ctx->d_cone_type, ctx->h_cone_type.data(), K * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (ctx->any_dual) {
            std::vector<uint8_t> rd((size_t)q, 0);
            for (int k = 0; k < K; k++)
                if (ctx->h_cone_dual[k])
                    for (int64_t r = ctx->h_cone_off[k]; r < ctx->h_cone_off[k + 1]; r++) rd[r] = 1;
            dalloc(&ctx->d_row_dual, q);
            CUDA_TRY(cudaMemcpyAsync(ctx->d_row_dual, rd.data(), q, cudaMemcpyHostToDevice, ctx->stream));
        }
        hyp_cones_build_groups(ctx);

        // state / work vectors
        const int64_t qn = std::max<int64_t>(std::max(q, n), 1);
        dalloc(&ctx->d_point, q);
        dalloc(&ctx->d_dual, q);
        dalloc(&ctx->d_grad, q);
        dalloc(&ctx->d_wivec, q);
        dalloc(&ctx->d_feas, K);
        dalloc(&ctx->d_dual_feas, K);
        dalloc(&ctx->d_num_ok, K);
        dalloc(&ctx->d_tmpflag, K);
        dalloc(&ctx->d_proxsqr, K);
        dalloc(&ctx->d_Gx, q);
        dalloc(&ctx->d_HGx, q);
        dalloc(&ctx->d_Gx_const, q);
        dalloc(&ctx->d_vq1, qn);
        dalloc(&ctx->d_vq2, qn);
        dalloc(&ctx->d_vq3, qn);
        dalloc(&ctx->d_vq4, qn);
        dalloc(&ctx->d_t, n);
        dalloc(&ctx->d_t2, n);
        dalloc(&ctx->d_vp1, p);
        dalloc(&ctx->d_vp2, p);
        dalloc(&ctx->d_scalars, 64);
        ctx->partial_doubles = std::max<int64_t>(4096, 32 * std::max<int64_t>(std::max(ctx->qloc, n), p));
        dalloc(&ctx->d_partial, ctx->partial_doubles);
        dalloc(&ctx->d_info, 16);
        dalloc(&ctx->d_flags, 2 * ceil_div(std::max<int64_t>(std::max(nmp, p), 1), 128) + 8);
        ctx->trsv_epoch = 0;

        // Schur matrix, factor, inverted diagonal blocks
        ctx->lds = round_up(std::max<int64_t>(nmp, 2), 2);
        // column sharding: equal panels of col_shard_width columns per rank (ncclAllGather needs equal counts)
        ctx->col_shard_width = (ctx->col_shard && ctx->nranks > 1) ? round_up(ceil_div(std::max<int64_t>(nmp, 1), ctx->nranks), 2) : 0;
        dalloc(&ctx->d_S, ctx->lds * std::max<int64_t>(std::max<int64_t>(nmp, ctx->col_shard_width * ctx->nranks), 1));
        dalloc(&ctx->d_F, ctx->lds * std::max<int64_t>(nmp, 1));
        dalloc(&ctx->d_Dinv, (int64_t)ceil_div(std::max<int64_t>(nmp, 1), 128) * 128 * 128);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaDeviceSynchronize());
        ctx->model_loaded = true;
        return 0;
    });
}

int hyp_set_cone_params(hyp_ctx* ctx, int K, const int* ssf_kind, const double* ssf_param) {
    return guarded(ctx, [&] {
        if (K < 0 || (K > 0 && (!ssf_kind || !ssf_param))) throw HypError{"hyp_set_cone_params: bad arguments"};
        ctx->h_cone_hkind.assign(ssf_kind, ssf_kind + K);
        ctx->h_cone_hparam.assign(ssf_param, ssf_param + K);
        ctx->params_staged = true;
        return 0;
    });
}

int hyp_set_cone_alpha(hyp_ctx* ctx, int K, const int64_t* alpha_off, const double* alpha) {
    return guarded(ctx, [&] {
        if (K < 0 || (K > 0 && !alpha_off)) throw HypError{"hyp_set_cone_alpha: bad arguments"};
        ctx->h_cone_aoff.assign(alpha_off, alpha_off + K + 1);
        const int64_t tot = K > 0 ? alpha_off[K] : 0;
        if (tot < 0 || (tot > 0 && !alpha)) throw HypError{"hyp_set_cone_alpha: bad offsets"};
        ctx->h_cone_alpha.assign(alpha, alpha + tot);
        return 0;
    });
}

int hyp_set_column_sharding(hyp_ctx* ctx, int on) {
    return guarded(ctx, [&] {
        if (ctx->model_loaded) throw HypError{"hyp_set_column_sharding: call before hyp_load_model"};
        ctx->col_shard = on != 0;
        return 0;
    });
}
int hyp_set_syrk_mode(hyp_ctx* ctx, int mode) {
    return guarded(ctx, [&] {
        if (mode != 0 && mode != 1) throw HypError{"hyp_set_syrk_mode: 0 = FP64 DMMA, 1 = sliced int8 tcgen05"};
        ctx->syrk_mode = mode;
        ctx->lhs_ready = false;
        return 0;
    });
}

int hyp_set_syssolver(hyp_ctx* ctx, int kind) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (kind != 0 && kind != 1) throw HypError{"hyp_set_syssolver: kind must be 0 (QRCholDense) or 1 (SymIndefDense)"};
        ctx->solver_kind = kind;
        ctx->lhs_ready = false;
        if (kind == 1) symindef_setup(ctx);
        else if (ctx->q) hyp_copy(ctx, ctx->q, ctx->d_const_rhs + ctx->n + ctx->p, ctx->d_cbh + ctx->n + ctx->p);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return 0;
    });
}

int hyp_cones_hess_blocks(hyp_ctx* ctx, double* blocks, int inverse) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_cones_hess_blocks: no point loaded"};
        const int64_t q = ctx->q;
        if (q == 0) return 0;
        double* prod = cone_blocks_dev(ctx, inverse ? HYP_PROD_INV_HESS : HYP_PROD_HESS);
        std::vector<int64_t> boff(ctx->K + 1, 0);
        for (int k = 0; k < ctx->K; k++) boff[k + 1] = boff[k] + ctx->h_cone_dim[k] * ctx->h_cone_dim[k];
        int64_t* d_boff = nullptr;
        double* d_out = nullptr;
        dalloc(&d_boff, ctx->K + 1);
        CUDA_TRY(cudaMemcpyAsync(d_boff, boff.data(), (ctx->K + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        bool dev_out = is_device_ptr(blocks);
        if (dev_out) d_out = blocks; else dalloc(&d_out, boff[ctx->K]);
        dim3 grid(std::max(1, std::min(ceil_div(q, 256), ctx->sm_count * 2)),
                  (unsigned)std::min<int64_t>(ctx->blk_maxdim, 65535));
        gather_blocks_kernel<<<grid, 256, 0, ctx->stream>>>(q, ctx->blk_maxdim, ctx->d_row_cone, ctx->d_cone_off,
                                                          ctx->d_cone_dim, d_boff, prod, d_out);
        ctx->launches++;
        if (!dev_out) stage_out(ctx, blocks, boff[ctx->K], d_out);
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_boff);
        if (!dev_out) cudaFree(d_out);
        return 0;
    });
}

int hyp_set_mu_tau(hyp_ctx* ctx, double mu, double tau_bar) {
    return guarded(ctx, [&] {
        ctx->mu = mu;
        ctx->tau_bar = tau_bar;
        return 0;
    });
}

int hyp_cones_load_point(hyp_ctx* ctx, const double* primal, const double* dual, double scal) {
    return guarded(ctx, [&] {
        need_model(ctx);
        const int64_t q = ctx->q;
        ensure_stage(ctx, 2 * q + 2);
        const double* dp = stage_in(ctx, primal, q, ctx->d_stage);
        const double* dd = stage_in(ctx, dual, q, ctx->d_stage + q);
        // load_point(cone, point, scal): cone.point = scal * point (Cones.jl:157-161)
        hyp_lincomb3(ctx, q, ctx->d_point, scal, dp, 0.0, nullptr, 0.0, nullptr);
        hyp_copy(ctx, q, ctx->d_dual, dd);
        hyp_cones_update_state(ctx);
        ctx->cones_loaded = true;
        ctx->lhs_ready = false;
        return 0;
    });
}

int hyp_cones_feas(hyp_ctx* ctx, uint8_t* is_feas, uint8_t* is_dual_feas) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_cones_feas: no point loaded"};
        stage_out_u8(ctx, is_feas, ctx->K, ctx->d_feas);
        stage_out_u8(ctx, is_dual_feas, ctx->K, ctx->d_dual_feas);
        return 0;
    });
}

int hyp_cones_grad(hyp_ctx* ctx, double* grad) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_cones_grad: no point loaded"};
        stage_out(ctx, grad, ctx->q, ctx->d_grad);
        return 0;
    });
}

int hyp_cones_hess_prod(hyp_ctx* ctx, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                        int64_t ld_arr, int mode) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_cones_hess_prod: no point loaded"};
        if (mode < 0 || mode > HYP_PROD_BLOCK) throw HypError{"hyp_cones_hess_prod: bad mode"};
        const int64_t q = ctx->q;
        if (ncols <= 0 || q == 0) return 0;
        if (ld_arr < q || ld_prod < q) throw HypError{"hyp_cones_hess_prod: leading dimension < q"};
        bool dev_in = is_device_ptr(arr), dev_out = is_device_ptr(prod);
        const double* darr = arr;
        double* dprod = prod;
        int64_t lda_d = ld_arr, ldp_d = ld_prod;
        int64_t need = (dev_in ? 0 : q * ncols) + (dev_out ? 0 : q * ncols);
        ensure_stage(ctx, need + 2);
        double* buf = ctx->d_stage;
        if (!dev_in) {
            CUDA_TRY(cudaMemcpy2DAsync(buf, q * 8, arr, ld_arr * 8, q * 8, ncols, cudaMemcpyHostToDevice,
                                       ctx->stream));
            darr = buf;
            lda_d = q;
            buf += q * ncols;
        }
        if (!dev_out) {
            dprod = buf;
            ldp_d = q;
        }
        {
            TimeScope ts(ctx, T_CONE_PROD);
            hyp_cones_prod(ctx, dprod, darr, ncols, ldp_d, lda_d, mode, 0);
        }
        if (ctx->nranks > 1)
            for (int64_t j = 0; j < ncols; j++) hyp_replicate_q(ctx, dprod + j * ldp_d);
        if (!dev_out) {
            CUDA_TRY(cudaMemcpy2DAsync(prod, ld_prod * 8, dprod, q * 8, q * 8, ncols, cudaMemcpyDeviceToHost,
                                       ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        return 0;
    });
}

int hyp_cones_dder3(hyp_ctx* ctx, double* out, const double* dir) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_cones_dder3: no point loaded"};
        const int64_t q = ctx->q;
        ensure_stage(ctx, q + 2);
        const double* dd = stage_in(ctx, dir, q, ctx->d_stage);
        hyp_cones_dder3_dev(ctx, ctx->d_vq4, dd);
        if (hyp_row_sharded(ctx)) hyp_replicate_q(ctx, ctx->d_vq4);
        stage_out(ctx, out, q, ctx->d_vq4);
        return 0;
    });
}

int hyp_cones_proxsqr(hyp_ctx* ctx, double irtmu, int use_max, double* proxsqr, uint8_t* numerics_ok) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_cones_proxsqr: no point loaded"};
        hyp_cones_prox_dev(ctx, irtmu, use_max);
        stage_out(ctx, proxsqr, ctx->K, ctx->d_proxsqr);
        stage_out_u8(ctx, numerics_ok, ctx->K, ctx->d_num_ok);
        return 0;
    });
}

int hyp_update_lhs(hyp_ctx* ctx, int* fact_kind) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_update_lhs: no cone point loaded"};
        int rc = 0;
        ctx->fact_kind = 0;
        const int64_t n = ctx->n, p = ctx->p, q = ctx->q;
        if (ctx->solver_kind == 1) {
            rc = symindef_update(ctx);
        } else if (ctx->p > 0 && !ctx->d_R) {
            throw HypError{"QRCholDense with p > 0 needs the QR factors Ap_Q, Ap_R (pass them to hyp_load_model)"};
        } else if (ctx->nmp > 0) {
            rc = update_lhs_fact(ctx);
        }
        if (fact_kind) *fact_kind = ctx->fact_kind;
        // rhs_const.z = H h (block_hess_prod!, qrchol.jl:191-195), then the constant column
        if (q > 0 && ctx->solver_kind == 0) {
            TimeScope ts(ctx, T_CONE_PROD);
            hyp_cones_prod(ctx, ctx->d_const_rhs + n + p, ctx->d_cbh + n + p, 1, q, q, HYP_PROD_BLOCK, 0);
            if (hyp_row_sharded(ctx)) hyp_replicate_q(ctx, ctx->d_const_rhs + n + p);
        }
        solve_subsystem3_dev(ctx, ctx->d_const_sol, ctx->d_const_rhs);
        hyp_copy(ctx, q, ctx->d_Gx_const, ctx->d_Gx);
        hyp_dot(ctx, n + p + q, ctx->d_cbh, ctx->d_const_sol, ctx->d_scalars + 1, false);
        ctx->lhs_ready = true;
        return rc;
    });
}

int hyp_solve_subsystem3(hyp_ctx* ctx, double* sol, const double* rhs) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->lhs_ready && ctx->nmp > 0) throw HypError{"hyp_solve_subsystem3: call hyp_update_lhs first"};
        const int64_t dim3 = ctx->n + ctx->p + ctx->q;
        const double* drhs = stage_in(ctx, rhs, dim3, ctx->d_sub_rhs);
        solve_subsystem3_dev(ctx, ctx->d_sub_sol, drhs);
        stage_out(ctx, sol, dim3, ctx->d_sub_sol);
        return 0;
    });
}

int hyp_solve_system(hyp_ctx* ctx, double* sol, const double* rhs) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->lhs_ready) throw HypError{"hyp_solve_system: call hyp_update_lhs first"};
        const int64_t dim6 = ctx->n + ctx->p + 2 * ctx->q + 2;
        const double* drhs = stage_in(ctx, rhs, dim6, ctx->d_rhs);
        double* dsol = is_device_ptr(sol) ? sol : ctx->d_sol;
        if (dsol == drhs) dsol = ctx->d_sol;
        solve_system_dev(ctx, dsol, drhs);
        stage_out(ctx, sol, dim6, dsol);
        return 0;
    });
}

int hyp_apply_lhs(hyp_ctx* ctx, double* res, const double* dir) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_apply_lhs: no cone point loaded"};
        const int64_t dim6 = ctx->n + ctx->p + 2 * ctx->q + 2;
        const double* ddir = stage_in(ctx, dir, dim6, ctx->d_rhs);
        double* dres = is_device_ptr(res) ? res : ctx->d_sol;
        if (dres == ddir) dres = ctx->d_sol;
        apply_lhs_dev(ctx, dres, ddir);
        stage_out(ctx, res, dim6, dres);
        return 0;
    });
}

// Multi-column variants (the data flow of combined.jl:67-79 lets {cent, pred} and {centadj, predadj}
// share one sweep each).  Column j of `sol` / `rhs` / `res` / `dir` is the full Point at offset j * ld (ld >= n+p+2q+2).
int hyp_solve_system_multi(hyp_ctx* ctx, double* sol, const double* rhs, int ncols, int64_t ld) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->lhs_ready) throw HypError{"hyp_solve_system_multi: call hyp_update_lhs first"};
        const int64_t dim6 = ctx->n + ctx->p + 2 * ctx->q + 2;
        if (ncols < 0 || (ncols > 0 && ld < dim6)) throw HypError{"hyp_solve_system_multi: bad ncols / ld"};
        if (ncols == 0) return 0;
        const bool dev = is_device_ptr(sol) && is_device_ptr(rhs) && sol != rhs;
        if (dev && ncols > 1 && hyp_multi_supported(ctx, ncols)) {
            solve_system_multi_dev(ctx, sol, rhs, ncols, ld);
            return 0;
        }
        for (int j = 0; j < ncols; j++) {
            const double* drhs = stage_in(ctx, rhs + j * ld, dim6, ctx->d_rhs);
            double* dsol = is_device_ptr(sol) ? sol + j * ld : ctx->d_sol;
            if (dsol == drhs) dsol = ctx->d_sol;
            solve_system_dev(ctx, dsol, drhs);
            stage_out(ctx, sol + j * ld, dim6, dsol);
        }
        return 0;
    });
}
int hyp_apply_lhs_multi(hyp_ctx* ctx, double* res, const double* dir, int ncols, int64_t ld) {
    return guarded(ctx, [&] {
        need_model(ctx);
        if (!ctx->cones_loaded) throw HypError{"hyp_apply_lhs_multi: no cone point loaded"};
        const int64_t dim6 = ctx->n + ctx->p + 2 * ctx->q + 2;
        if (ncols < 0 || (ncols > 0 && ld < dim6)) throw HypError{"hyp_apply_lhs_multi: bad ncols / ld"};
        if (ncols == 0) return 0;
        const bool dev = is_device_ptr(res) && is_device_ptr(dir) && res != dir;
        if (dev && ncols > 1 && hyp_multi_supported(ctx, ncols)) {
            apply_lhs_multi_dev(ctx, res, dir, ncols, ld);
            return 0;
        }
        for (int j = 0; j < ncols; j++) {
            const double* ddir = stage_in(ctx, dir + j * ld, dim6, ctx->d_rhs);
            double* dres = is_device_ptr(res) ? res + j * ld : ctx->d_sol;
            if (dres == ddir) dres = ctx->d_sol;
            apply_lhs_dev(ctx, dres, ddir);
            stage_out(ctx, res + j * ld, dim6, dres);
        }
        return 0;
    });
}
int hyp_calc_residuals(hyp_ctx* ctx, const double* point, double* x_residual, double* y_residual,
                       double* z_residual, double* stats) {
    return guarded(ctx, [&] {
        need_model(ctx);
        const int64_t n = ctx->n, p = ctx->p, q = ctx->q, dim6 = n + p + 2 * q + 2;
        const double* dpt = stage_in(ctx, point, dim6, ctx->d_rhs);
        // results are staged in d_sol: [xres (n) | yres (p) | zres (q) | 10 scalars]
        double* dx = is_device_ptr(x_residual) ? x_residual : ctx->d_sol;
        double* dy = is_device_ptr(y_residual) ? y_residual : ctx->d_sol + n;
        double* dz = is_device_ptr(z_residual) ? z_residual : ctx->d_sol + n + p;
        double* dst = is_device_ptr(stats) ? stats : ctx->d_scalars + 16;
        calc_residuals_dev(ctx, dpt, dx, dy, dz, dst);
        if (n) stage_out(ctx, x_residual, n, dx);
        if (p) stage_out(ctx, y_residual, p, dy);
        if (q) stage_out(ctx, z_residual, q, dz);
        stage_out(ctx, stats, 10, dst);
        return 0;
    });
}

int hyp_get_schur(hyp_ctx* ctx, double* S, int64_t ld) {
    return guarded(ctx, [&] {
        need_model(ctx);
        const int64_t nmp = ctx->nmp;
        if (nmp == 0) return 0;
        cudaMemcpyKind kind = is_device_ptr(S) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        CUDA_TRY(cudaMemcpy2DAsync(S, ld * 8, ctx->d_S, ctx->lds * 8, nmp * 8, nmp, kind, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return 0;
    });
}

int64_t hyp_launch_count(hyp_ctx* ctx) { return ctx ? ctx->launches : -1; }

int hyp_timing_enable(hyp_ctx* ctx, int on) {
    if (!ctx) return -1;
    ctx->timing_enabled = on != 0;
    return 0;
}
int hyp_timing_get(hyp_ctx* ctx, int slot, double* total_ms, int64_t* calls) {
    if (!ctx || slot < 0 || slot >= T_NUM) return -1;
    if (total_ms) *total_ms = ctx->timing[slot].total_ms;
    if (calls) *calls = ctx->timing[slot].count;
    return 0;
}
int hyp_timing_reset(hyp_ctx* ctx) {
    if (!ctx) return -1;
    for (int i = 0; i < T_NUM; i++) {
        ctx->timing[i].total_ms = 0;
        ctx->timing[i].count = 0;
    }
    return 0;
}
int hyp_timing_slots(void) { return T_NUM; }
const char* hyp_timing_name(int slot) { return (slot >= 0 && slot < T_NUM) ? kTimingNames[slot] : ""; }

// ---- unit-test entry points --------------------------------------------------------------
namespace {
struct TmpDev {
    hyp_ctx* ctx;
    std::vector<void*> ptrs;
    ~TmpDev() {
        cudaStreamSynchronize(ctx->stream);
        for (void* p : ptrs) cudaFree(p);
    }
    // device copy (or the pointer itself) of a rows x cols matrix; returns new leading dim
    double* in(const double* src, int64_t ld, int64_t rows, int64_t cols, int64_t* ld_out) {
        if (is_device_ptr(src)) {
            *ld_out = ld;
            return const_cast<double*>(src);
        }
        int64_t dld = round_up(std::max<int64_t>(rows, 2), 2);
        double* d = nullptr;
        dalloc(&d, dld * std::max<int64_t>(cols, 1));
        ptrs.push_back(d);
        upload_matrix(ctx, d, dld, src, ld, rows, cols);
        *ld_out = dld;
        return d;
    }
    void out(double* dst, int64_t ld, const double* d, int64_t dld, int64_t rows, int64_t cols) {
        if (dst == d) return;
        CUDA_TRY(cudaMemcpy2DAsync(dst, ld * 8, d, dld * 8, rows * 8, cols, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
};
}  // namespace

int hyp_test_atb_upper(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr, int64_t klen,
                       int64_t ncols, double* C, int64_t ldc, double alpha, double beta) {
    return guarded(ctx, [&] {
        TmpDev tmp{ctx};
        int64_t lp, lr, lc;
        double* dP = tmp.in(P, ldp, klen, ncols, &lp);
        double* dR = (R == P) ? dP : tmp.in(R, ldr, klen, ncols, &lr);
        if (R == P) lr = lp;
        double* dC = tmp.in(C, ldc, ncols, ncols, &lc);
        hyp_atb_upper(ctx, dP, lp, dR, lr, klen, ncols, dC, lc, alpha, beta);
        tmp.out(C, ldc, dC, lc, ncols, ncols);
        return 0;
    });
}

int hyp_test_gemm_tn(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr, int64_t klen,
                     int64_t mrows, int64_t ncols, double* C, int64_t ldc, double alpha, double beta) {
    return guarded(ctx, [&] {
        TmpDev tmp{ctx};
        int64_t lp, lr, lc;
        double* dP = tmp.in(P, ldp, klen, mrows, &lp);
        double* dR = tmp.in(R, ldr, klen, ncols, &lr);
        double* dC = tmp.in(C, ldc, mrows, ncols, &lc);
        hyp_gemm_tn(ctx, dP, lp, dR, lr, klen, mrows, ncols, dC, lc, alpha, beta);
        tmp.out(C, ldc, dC, lc, mrows, ncols);
        return 0;
    });
}

namespace {
struct TestFactor {
    double* dinv = nullptr;
    int* info = nullptr;
    int* flags = nullptr;
    int64_t m = 0;
} g_tf;
}  // namespace

int hyp_test_potrf(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, int* info) {
    return guarded(ctx, [&] {
        TmpDev tmp{ctx};
        int64_t la;
        double* dA = tmp.in(A, lda, m, m, &la);
        if (g_tf.dinv) cudaFree(g_tf.dinv);
        if (g_tf.info) cudaFree(g_tf.info);
        dalloc(&g_tf.dinv, (int64_t)ceil_div(std::max<int64_t>(m, 1), 128) * 128 * 128);
        dalloc(&g_tf.info, 8);
        g_tf.m = m;
        hyp_potrf_upper(ctx, dA, la, m, g_tf.dinv, g_tf.info);
        int hinfo = 0;
        CUDA_TRY(cudaMemcpyAsync(&hinfo, g_tf.info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (info) *info = hinfo;
        tmp.out(A, lda, dA, la, m, m);
        return 0;
    });
}

int hyp_test_set_trsv_pkt(int on) {
    hyp_trsv_set_pkt(on);
    return 0;
}

int hyp_test_potrs(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, double* x) {
    return guarded(ctx, [&] {
        if (!g_tf.dinv || g_tf.m != m) throw HypError{"hyp_test_potrs: call hyp_test_potrf with the same m first"};
        TmpDev tmp{ctx};
        int64_t lf, lx;
        double* dF = tmp.in(F, ldf, m, m, &lf);
        double* dx = tmp.in(x, m, m, 1, &lx);
        int* saved = ctx->d_flags;
        int* flags = nullptr;
        dalloc(&flags, 2 * ceil_div(m, 128) + 8);
        tmp.ptrs.push_back(flags);
        ctx->d_flags = flags;
        hyp_trsv_upper(ctx, dF, lf, m, g_tf.dinv, dx, true);
        hyp_trsv_upper(ctx, dF, lf, m, g_tf.dinv, dx, false);
        ctx->d_flags = saved;
        tmp.out(x, m, dx, lx, m, 1);
        return 0;
    });
}

int hyp_test_gemv(hyp_ctx* ctx, int trans, int64_t rows, int64_t cols, const double* M, int64_t ld,
                  const double* x, double alpha, double beta, double* y) {
    return guarded(ctx, [&] {
        TmpDev tmp{ctx};
        int64_t lm, l1, l2;
        double* dM = tmp.in(M, ld, rows, cols, &lm);
        int64_t xl = trans ? rows : cols, yl = trans ? cols : rows;
        double* dx = tmp.in(x, xl, xl, 1, &l1);
        double* dy = tmp.in(y, yl, yl, 1, &l2);
        double* saved = ctx->d_partial;
        int64_t saved_n = ctx->partial_doubles;
        double* part = nullptr;
        int64_t pn = std::max<int64_t>(4096, 32 * std::max(rows, cols));
        dalloc(&part, pn);
        tmp.ptrs.push_back(part);
        ctx->d_partial = part;
        ctx->partial_doubles = pn;
        if (trans) hyp_gemv_t(ctx, rows, cols, dM, lm, dx, alpha, beta, dy);
        else hyp_gemv_n(ctx, rows, cols, dM, lm, dx, alpha, beta, dy);
        ctx->d_partial = saved;
        ctx->partial_doubles = saved_n;
        tmp.out(y, yl, dy, l2, yl, 1);
        return 0;
    });
}

int hyp_test_ldlt_solve(hyp_ctx* ctx, const double* A, int64_t lda, int64_t m, double* x, int* info) {
    return guarded(ctx, [&] {
        TmpDev tmp{ctx};
        int64_t la, lx;
        double* dA0 = tmp.in(A, lda, m, m, &la);
        double* dA = nullptr;
        dalloc(&dA, la * m);
        tmp.ptrs.push_back(dA);
        CUDA_TRY(cudaMemcpyAsync(dA, dA0, (size_t)la * m * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        double* dx = tmp.in(x, m, m, 1, &lx);
        int *ipiv = nullptr, *dinfo = nullptr;
        dalloc(&ipiv, 3 * m + 8);
        dalloc(&dinfo, 16);
        tmp.ptrs.push_back(ipiv);
        tmp.ptrs.push_back(dinfo);
        double* saved_w = ctx->d_ldl_work;
        double* work = nullptr;
        dalloc(&work, 4 * m + 64);
        tmp.ptrs.push_back(work);
        ctx->d_ldl_work = work;
        hyp_ldlt_factor(ctx, dA, la, m, ipiv, dinfo);
        int hinfo = 0;
        CUDA_TRY(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (info) *info = hinfo;
        hyp_ldlt_solve(ctx, dA, la, m, ipiv, dx);
        ctx->d_ldl_work = saved_w;
        tmp.out(x, m, dx, lx, m, 1);
        return 0;
    });
}

}  // extern "C"

// GeneralizedPower, HypoPowerMean and EpiNormSpectral (real; src/Cones/epinormspectral.jl:107-294) on the device, and the generic inverse-Hessian fallback of the Cone API (kernels:
// cones_gpow_kernels.cuh).
//
// reference: src/Cones/generalizedpower.jl:77-236, src/Cones/hypopowermean.jl:74-203; generic oracles src/Cones/Cones.jl:113-118 (inv_hess_prod! =
// hess_fact \\ arr), :239-259 (update_hess_fact, update_inv_hess).  ConeGroup fields reused: d_side = dim of the cone
// (side of its explicit Hessian), d_hkind = number of powers m, d_vecs / d_voff = the powers, d_W = explicit
// Hessian, d_U = its Cholesky factor (scratch), d_Ui = U^-1.  A Hessian whose Cholesky fails marks the cone
// infeasible (the reference would go on to Bunch-Kaufman, dense.jl:194-215; the line search backtracks instead).
#include "cones_mat.cuh"
#include "cones_gpow_kernels.cuh"

namespace {

template <typename T>
T* upload_vec(const std::vector<T>& v) {
    T* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace

// EpiNormSpectral: d_hkind = d1 (rows of W), d_vecs / d_voff = per-cone workspace (tau, Zitau, Zi, Cholesky factor of Z,
// scratch: 3 d1 d2 + 2 d1^2 doubles)
static void ens_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    g.h_voff.assign(g.count, 0);
    int64_t tot = 0;
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        const int d = g.h_dim[i], d1 = ctx->h_cone_hkind[k];
        if (d1 < 1 || d < 2 || (d - 1) % d1 != 0 || d1 > (d - 1) / d1)
            throw HypError{"EpiNormSpectral: hyp_set_cone_params must give the number of rows d1 with dim = 1 + d1 * d2, d1 <= d2"};
        if (d > 128) throw HypError{"EpiNormSpectral: dim above 128 is not supported (batched Cholesky limit)"};
        g.h_hkind.push_back(d1);
        g.h_side[i] = d;
        g.h_voff[i] = tot;
        tot += 3 * (int64_t)(d - 1) + 2 * (int64_t)d1 * d1;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    CUDA_TRY(cudaMemset(g.d_vecs, 0, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    hyp_mat_alloc_group(ctx, g);
}

// WSOSInterpNonnegative: d_vecs / d_voff = per-cone region [nP][L_1 .. L_nP][P_1 .. P_nP] (as passed to
// hyp_set_cone_alpha) followed by the workspace of wsos_state_kernel / wsos_dder3_kernel
static void wsos_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    if ((int)ctx->h_cone_aoff.size() != ctx->K + 1)
        throw HypError{"WSOSInterpNonnegative cones need hyp_set_cone_alpha (packed Ps) before hyp_load_model"};
    g.h_voff.assign(g.count, 0);
    std::vector<double> buf;
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
        const int U = g.h_dim[i];
        if (U > 128) throw HypError{"WSOSInterpNonnegative: dimension above 128 is not supported (batched Cholesky limit)"};
        if (a1 - a0 < 2) throw HypError{"WSOSInterpNonnegative: missing Ps data"};
        const int nP = (int)ctx->h_cone_alpha[a0];
        if (nP < 1 || a1 - a0 < 1 + nP) throw HypError{"WSOSInterpNonnegative: bad number of Ps matrices"};
        int64_t sumL = 0, wsz = 0, Lmax = 0;
        for (int j = 0; j < nP; j++) {
            const int64_t L = (int64_t)ctx->h_cone_alpha[a0 + 1 + j];
            if (L < 1 || L > U) throw HypError{"WSOSInterpNonnegative: need 1 <= L_k <= U"};
            sumL += L;
            wsz += L * U + L * L;
            Lmax = std::max(Lmax, L);
        }
        if (a1 - a0 != 1 + nP + (int64_t)U * sumL) throw HypError{"WSOSInterpNonnegative: Ps data has the wrong length"};
        g.h_voff[i] = (int64_t)buf.size();
        buf.insert(buf.end(), ctx->h_cone_alpha.begin() + a0, ctx->h_cone_alpha.begin() + a1);
        buf.resize(buf.size() + (size_t)(wsz + Lmax * Lmax), 0.0);
        g.h_hkind.push_back(nP);
        g.h_side[i] = U;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, std::max<size_t>(buf.size(), 1) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(g.d_vecs, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    hyp_mat_alloc_group(ctx, g);
}

// LinMatrixIneq: d_vecs / d_voff = per-cone region [side][A_1 .. A_dim] (as passed to hyp_set_cone_alpha) followed by
// the workspace of lmi_state_kernel / lmi_dder3_kernel (Cholesky factor, B_i, two scratch matrices)
static void lmi_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    if ((int)ctx->h_cone_aoff.size() != ctx->K + 1)
        throw HypError{"LinMatrixIneq cones need hyp_set_cone_alpha (packed As) before hyp_load_model"};
    g.h_voff.assign(g.count, 0);
    std::vector<double> buf;
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
        const int d = g.h_dim[i];
        if (d > 128) throw HypError{"LinMatrixIneq: dimension above 128 is not supported (batched Cholesky limit)"};
        if (d < 2 || a1 - a0 < 2) throw HypError{"LinMatrixIneq: need at least two matrices"};
        const int64_t sd = (int64_t)ctx->h_cone_alpha[a0];
        if (sd < 1 || sd > 256 || sd * (sd + 1) / 2 < d || a1 - a0 != 1 + (int64_t)d * sd * sd)
            throw HypError{"LinMatrixIneq: As data has the wrong length (need svec_length(side) >= dim, side <= 256)"};
        g.h_voff[i] = (int64_t)buf.size();
        buf.insert(buf.end(), ctx->h_cone_alpha.begin() + a0, ctx->h_cone_alpha.begin() + a1);
        buf.resize(buf.size() + (size_t)((d + 3) * sd * sd), 0.0);
        g.h_hkind.push_back((int)sd);
        g.h_side[i] = d;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, std::max<size_t>(buf.size(), 1) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(g.d_vecs, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    hyp_mat_alloc_group(ctx, g);
}

// DoublyNonnegativeTri: d_hkind = side of the matrix, d_vecs / d_voff = per-cone workspace (W^-1 and the Cholesky factor)
static void dnn_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    g.h_voff.assign(g.count, 0);
    int64_t tot = 0;
    for (int i = 0; i < g.count; i++) {
        const int d = g.h_dim[i];
        if (d > 128) throw HypError{"DoublyNonnegativeTri: dim above 128 is not supported (batched Cholesky limit)"};
        const int side = (int)std::floor((std::sqrt(1.0 + 8.0 * d) - 1) / 2 + 0.5);
        if (side * (side + 1) / 2 != d) throw HypError{"DoublyNonnegativeTri: dim is not a triangular number"};
        g.h_hkind.push_back(side);
        g.h_side[i] = d;
        g.h_voff[i] = tot;
        tot += 2 * (int64_t)side * side;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    CUDA_TRY(cudaMemset(g.d_vecs, 0, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    hyp_mat_alloc_group(ctx, g);
}

// MatrixEpiPerSquare: d_hkind = d1, d_vecs / d_voff = per-cone state (Zi, Cholesky factor of Z, smat(U), ZiUZi, ZiW,
// ZiUZiW) followed by the scratch of mep_dder3_kernel
static void mep_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    g.h_voff.assign(g.count, 0);
    int64_t tot = 0;
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        const int d = g.h_dim[i], d1 = ctx->h_cone_hkind[k];
        const int rest = d - d1 * (d1 + 1) / 2 - 1;
        if (d1 < 1 || rest < d1 * d1 || rest % d1 != 0)
            throw HypError{"MatrixEpiPerSquare: hyp_set_cone_params must give d1 with dim = svec_length(d1) + 1 + d1 * d2, d1 <= d2"};
        if (d > 128) throw HypError{"MatrixEpiPerSquare: dim above 128 is not supported (batched Cholesky limit)"};
        const int64_t d2 = rest / d1, n11 = (int64_t)d1 * d1, n12 = d1 * d2, n22 = d2 * d2;
        g.h_hkind.push_back(d1);
        g.h_side[i] = d;
        g.h_voff[i] = tot;
        tot += (4 + 17) * n11 + (2 + 11) * n12 + 3 * n22;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    CUDA_TRY(cudaMemset(g.d_vecs, 0, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    hyp_mat_alloc_group(ctx, g);
}

// WSOSInterpPosSemidefTri: d_hkind = R, d_vecs / d_voff = per-cone region [nP][L_1 .. L_nP][P_1 .. P_nP] (as passed to
// hyp_set_cone_alpha) followed by the workspace of wpsd_state_kernel / wpsd_dder3_kernel
static void wpsd_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    if ((int)ctx->h_cone_aoff.size() != ctx->K + 1)
        throw HypError{"WSOSInterpPosSemidefTri cones need hyp_set_cone_alpha (packed Ps) before hyp_load_model"};
    g.h_voff.assign(g.count, 0);
    std::vector<double> buf;
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
        const int d = g.h_dim[i], R = ctx->h_cone_hkind[k];
        const bool one = g.type == HYP_CONE_WSOSINTERPEPINORMONE;       // R - 1 pair (2 L x 2 L) factorisations per P_k
        const bool eucl = g.type == HYP_CONE_WSOSINTERPEPINORMEUCL || one;   // dim = U R (R >= 2) instead of U svec_length(R)
        if (d > 128) throw HypError{"WSOSInterpPosSemidefTri / EpiNormEucl: dimension above 128 is not supported (batched Cholesky limit)"};
        const int nblk = eucl ? R : R * (R + 1) / 2;
        if (R < (eucl ? 2 : 1) || d % nblk != 0)
            throw HypError{"WSOSInterpPosSemidefTri / EpiNormEucl: hyp_set_cone_params must give R with dim = U * svec_length(R) / U * R"};
        const int64_t U = d / nblk;
        if (a1 - a0 < 2) throw HypError{"WSOSInterpPosSemidefTri: missing Ps data"};
        const int nP = (int)ctx->h_cone_alpha[a0];
        if (nP < 1 || a1 - a0 < 1 + nP) throw HypError{"WSOSInterpPosSemidefTri: bad number of Ps matrices"};
        int64_t sumL = 0, wsz = 0, Lmax = 0;
        for (int j = 0; j < nP; j++) {
            const int64_t L = (int64_t)ctx->h_cone_alpha[a0 + 1 + j];
            if (L < 1 || L > U) throw HypError{"WSOSInterpPosSemidefTri: need 1 <= L_k <= U"};
            sumL += L;
            wsz += one ? (R - 1) * (4 * L * U + 4 * L * L) : R * L * R * U + R * L * R * L;
            Lmax = std::max(Lmax, L);
        }
        if (a1 - a0 != 1 + nP + U * sumL) throw HypError{"WSOSInterpPosSemidefTri: Ps data has the wrong length"};
        g.h_voff[i] = (int64_t)buf.size();
        buf.insert(buf.end(), ctx->h_cone_alpha.begin() + a0, ctx->h_cone_alpha.begin() + a1);
        const int64_t Rb = one ? 2 : R;      // block count of the factorised matrices
        buf.resize(buf.size() + (size_t)(wsz + Rb * U * Rb * U + Rb * Lmax * Rb * Lmax + Rb * Lmax * Rb * U +
                                         (eucl ? U * U + Lmax * Lmax + Lmax * U : 0)),
                   0.0);
        g.h_hkind.push_back(R);
        g.h_side[i] = d;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, std::max<size_t>(buf.size(), 1) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(g.d_vecs, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    hyp_mat_alloc_group(ctx, g);
}

// PosSemidefTriSparse: d_vecs / d_voff = per-cone region [side][rows][cols] (as passed to hyp_set_cone_alpha) followed by
// the workspace of sps_state_kernel / sps_dder3_kernel (5 side^2 doubles)
static void sps_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    if ((int)ctx->h_cone_aoff.size() != ctx->K + 1)
        throw HypError{"PosSemidefTriSparse cones need hyp_set_cone_alpha (side and the sparsity pattern) before hyp_load_model"};
    g.h_voff.assign(g.count, 0);
    std::vector<double> buf;
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
        const int d = g.h_dim[i];
        if (d > 128) throw HypError{"PosSemidefTriSparse: more than 128 nonzeros are not supported (batched Cholesky limit)"};
        if (a1 - a0 != 1 + 2 * (int64_t)d) throw HypError{"PosSemidefTriSparse: pattern data has the wrong length"};
        const int64_t sd = (int64_t)ctx->h_cone_alpha[a0];
        if (sd < 1 || sd > d) throw HypError{"PosSemidefTriSparse: bad side"};
        std::vector<int> seen((size_t)sd, 0);
        for (int e = 0; e < d; e++) {
            const int64_t r = (int64_t)ctx->h_cone_alpha[a0 + 1 + e], c = (int64_t)ctx->h_cone_alpha[a0 + 1 + d + e];
            if (c < 0 || c > r || r >= sd) throw HypError{"PosSemidefTriSparse: need 0 <= col <= row < side"};
            if (r == c) seen[(size_t)r]++;
        }
        for (int64_t r = 0; r < sd; r++)
            if (seen[(size_t)r] != 1) throw HypError{"PosSemidefTriSparse: every diagonal entry must appear exactly once"};
        g.h_voff[i] = (int64_t)buf.size();
        buf.insert(buf.end(), ctx->h_cone_alpha.begin() + a0, ctx->h_cone_alpha.begin() + a1);
        buf.resize(buf.size() + (size_t)(5 * sd * sd), 0.0);
        g.h_hkind.push_back((int)sd);
        g.h_side[i] = d;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, std::max<size_t>(buf.size(), 1) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(g.d_vecs, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    hyp_mat_alloc_group(ctx, g);
}

// EpiTrRelEntropyTri: d_vecs / d_voff = per-cone state (eigen-decompositions, matrix logs, inverses, divided differences
// of log up to third order) followed by the scratch of etr_dder3_kernel
static void etr_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    g.h_voff.assign(g.count, 0);
    int64_t tot = 0;
    for (int i = 0; i < g.count; i++) {
        const int dm = g.h_dim[i];
        if (dm > 128) throw HypError{"EpiTrRelEntropyTri: dim above 128 is not supported (batched Cholesky limit)"};
        const int vw = (dm - 1) / 2;
        int64_t d = 1;
        while (d * (d + 1) / 2 < vw) d++;
        if (dm < 3 || dm % 2 == 0 || d * (d + 1) / 2 != vw)
            throw HypError{"EpiTrRelEntropyTri: need dim = 1 + 2 * svec_length(d)"};
        const int64_t n = d * d;
        g.h_hkind.push_back((int)d);
        g.h_side[i] = dm;
        g.h_voff[i] = tot;
        tot += 23 * n + 2 * d + 2 * d * n + n * n;
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    CUDA_TRY(cudaMemset(g.d_vecs, 0, (size_t)std::max<int64_t>(tot, 1) * sizeof(double)));
    hyp_mat_alloc_group(ctx, g);
}

void hyp_gpow_alloc_group(hyp_ctx* ctx, ConeGroup& g) {
    if (g.type == HYP_CONE_EPITRRELENTROPYTRI) {
        etr_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_POSSEMIDEFTRISPARSE) {
        sps_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI || g.type == HYP_CONE_WSOSINTERPEPINORMEUCL ||
        g.type == HYP_CONE_WSOSINTERPEPINORMONE) {
        wpsd_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_MATRIXEPIPERSQUARE) {
        mep_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_DOUBLYNONNEGATIVETRI) {
        dnn_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_LINMATRIXINEQ) {
        lmi_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_EPINORMSPECTRAL) {
        ens_alloc_group(ctx, g);
        return;
    }
    if (g.type == HYP_CONE_WSOSINTERPNONNEGATIVE) {
        wsos_alloc_group(ctx, g);
        return;
    }
    std::vector<double> alpha;
    g.h_voff.assign(g.count, 0);
    for (int i = 0; i < g.count; i++) {
        const int k = g.h_kidx[i];
        if ((int)ctx->h_cone_aoff.size() != ctx->K + 1)
            throw HypError{"GeneralizedPower / HypoPowerMean cones need hyp_set_cone_alpha before hyp_load_model"};
        const int64_t a0 = ctx->h_cone_aoff[k], a1 = ctx->h_cone_aoff[k + 1];
        const int m = (int)(a1 - a0);
        if (g.type == HYP_CONE_GENERALIZEDPOWER && (m < 1 || m >= g.h_dim[i]))
            throw HypError{"GeneralizedPower: need 1 <= number of powers < dim"};
        if (g.type == HYP_CONE_HYPOPOWERMEAN && m != g.h_dim[i] - 1)
            throw HypError{"HypoPowerMean: need dim - 1 powers"};
        if (g.h_dim[i] > 128)
            throw HypError{"GeneralizedPower / HypoPowerMean: dim above 128 is not supported (batched Cholesky limit)"};
        double sum = 0.0;
        for (int64_t a = a0; a < a1; a++) {
            if (!(ctx->h_cone_alpha[a] > 0.0)) throw HypError{"cone powers must be positive"};
            sum += ctx->h_cone_alpha[a];
        }
        if (std::abs(sum - 1.0) > 1e-10) throw HypError{"cone powers must sum to one"};
        g.h_voff[i] = (int64_t)alpha.size();
        alpha.insert(alpha.end(), ctx->h_cone_alpha.begin() + a0, ctx->h_cone_alpha.begin() + a1);
        g.h_hkind.push_back(m);
        g.h_side[i] = g.h_dim[i];
    }
    g.max_side = g.max_dim;
    cudaFree(g.d_side);
    g.d_side = upload_vec(g.h_side);
    g.d_hkind = upload_vec(g.h_hkind);
    g.d_voff = upload_vec(g.h_voff);
    CUDA_TRY(cudaMalloc(&g.d_vecs, std::max<size_t>(alpha.size(), 1) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(g.d_vecs, alpha.data(), alpha.size() * sizeof(double), cudaMemcpyHostToDevice));
    hyp_mat_alloc_group(ctx, g);   // per-cone dim x dim matrices (even leading dimension)
}

void hyp_gpow_update_state(hyp_ctx* ctx, ConeGroup& g) {
    if (g.type == HYP_CONE_EPITRRELENTROPYTRI)
        hypdev::etr_state_kernel<<<ceil_div(g.count, 4), 128, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_kidx, g.d_moff, ctx->d_point, ctx->d_grad, g.d_scal, g.d_W,
            ctx->d_feas);
    else if (g.type == HYP_CONE_POSSEMIDEFTRISPARSE)
        hypdev::sps_state_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_kidx,
                                                                  g.d_moff, ctx->d_point, ctx->d_grad, g.d_W, ctx->d_feas);
    else if (g.type == HYP_CONE_WSOSINTERPEPINORMONE)
        hypdev::wone_state_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs,
                                                                   g.d_kidx, g.d_moff, ctx->d_point, ctx->d_grad, g.d_W,
                                                                   ctx->d_feas);
    else if (g.type == HYP_CONE_WSOSINTERPEPINORMEUCL)
        hypdev::weuc_state_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs,
                                                                   g.d_kidx, g.d_moff, ctx->d_point, ctx->d_grad, g.d_W,
                                                                   ctx->d_feas);
    else if (g.type == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI)
        hypdev::wpsd_state_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs,
                                                                   g.d_kidx, g.d_moff, ctx->d_point, ctx->d_grad, g.d_W,
                                                                   ctx->d_feas);
    else if (g.type == HYP_CONE_MATRIXEPIPERSQUARE)
        hypdev::mep_state_kernel<<<ceil_div(g.count, 4), 128, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_kidx, g.d_moff, ctx->d_point, ctx->d_dual,
            ctx->d_grad, g.d_scal, g.d_W, ctx->d_feas, ctx->d_dual_feas);
    else if (g.type == HYP_CONE_DOUBLYNONNEGATIVETRI)
        hypdev::dnn_state_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_kidx, g.d_moff, ctx->d_point, ctx->d_grad,
            g.d_W, ctx->d_feas);
    else if (g.type == HYP_CONE_LINMATRIXINEQ)
        hypdev::lmi_state_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_kidx,
                                                                  g.d_moff, ctx->d_point, ctx->d_grad, g.d_W, ctx->d_feas);
    else if (g.type == HYP_CONE_WSOSINTERPNONNEGATIVE)
        hypdev::wsos_state_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_kidx,
                                                                   g.d_moff, ctx->d_point, ctx->d_grad, g.d_W, ctx->d_feas);
    else if (g.type == HYP_CONE_EPINORMSPECTRAL)
        hypdev::ens_state_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_kidx, g.d_moff, ctx->d_point, ctx->d_dual,
            ctx->d_grad, g.d_scal, g.d_W, ctx->d_feas, ctx->d_dual_feas);
    else if (g.type == HYP_CONE_HYPOPOWERMEAN)
        hypdev::hpm_state_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_kidx, g.d_moff, ctx->d_point, ctx->d_dual, ctx->d_grad,
            g.d_scal, g.d_W, ctx->d_feas, ctx->d_dual_feas);
    else
        hypdev::gpow_state_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_kidx, g.d_moff, ctx->d_point, ctx->d_dual,
            ctx->d_grad, g.d_scal, g.d_W, ctx->d_feas, ctx->d_dual_feas);
    ctx->launches++;
    // hess_fact: Cholesky of the explicit Hessian (copy) and its inverse factor
    CUDA_TRY(cudaMemcpyAsync(g.d_U, g.d_W, (size_t)std::max<int64_t>(g.mat_total, 1) * sizeof(double),
                             cudaMemcpyDeviceToDevice, ctx->stream));
    hyp_chol_batched(ctx, g.count, g.d_side, g.d_moff, g.d_kidx, g.d_U, g.d_Ui, ctx->d_feas);
    CUDA_TRY(cudaGetLastError());
}

void hyp_gpow_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                   int64_t ld_arr, int mode, int64_t row_shift) {
    if (mode == HYP_PROD_SQRT_HESS || mode == HYP_PROD_INV_SQRT_HESS)
        throw HypError{"sqrt_hess_prod of GeneralizedPower is not exported (the Schur assembly uses hess_prod)"};
    dim3 grid(ceil_div(g.count, 8), (unsigned)std::min<int64_t>(ncols, 65535));
    // which cones take hess_prod / inv_hess_prod: everyone, or split by use_dual_barrier for the block modes
    int hess_dual = -2, inv_dual = -2;      // -2: kernel not launched, -1: every cone, 0 / 1: cones with that flag
    if (mode == HYP_PROD_HESS) hess_dual = -1;
    else if (mode == HYP_PROD_INV_HESS) inv_dual = -1;
    else if (mode == HYP_PROD_BLOCK) { hess_dual = 0; inv_dual = 1; }
    else if (mode == HYP_PROD_BLOCK_INV) { hess_dual = 1; inv_dual = 0; }
    else throw HypError{"hyp_gpow_prod: bad mode"};
    if (hess_dual > -2) {
        if (g.type == HYP_CONE_EPITRRELENTROPYTRI) {
            dim3 grid4(ceil_div(g.count, 4), (unsigned)std::min<int64_t>(ncols, 65535));
            hypdev::etr_prod_kernel<<<grid4, 128, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_voff, g.d_vecs,
                                                                   g.d_dual, g.d_scal, arr, ld_arr, prod, ld_prod, ncols,
                                                                   row_shift);
        } else if (g.type == HYP_CONE_MATRIXEPIPERSQUARE)
            hypdev::mep_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_hkind,
                                                                  g.d_voff, g.d_vecs, g.d_dual, g.d_scal, ctx->d_point,
                                                                  arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (g.type == HYP_CONE_DOUBLYNONNEGATIVETRI)
            hypdev::dnn_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_hkind,
                                                                  g.d_voff, g.d_vecs, g.d_dual, ctx->d_point, arr, ld_arr,
                                                                  prod, ld_prod, ncols, row_shift);
        else if (g.type == HYP_CONE_WSOSINTERPNONNEGATIVE || g.type == HYP_CONE_LINMATRIXINEQ ||
                 g.type == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI || g.type == HYP_CONE_WSOSINTERPEPINORMEUCL ||
                 g.type == HYP_CONE_WSOSINTERPEPINORMONE || g.type == HYP_CONE_POSSEMIDEFTRISPARSE)
            hypdev::gen_hess_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_moff,
                                                                       g.d_dual, g.d_W, arr, ld_arr, prod, ld_prod, ncols,
                                                                       row_shift);
        else if (g.type == HYP_CONE_EPINORMSPECTRAL)
            hypdev::ens_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_hkind,
                                                                  g.d_voff, g.d_vecs, g.d_dual, g.d_scal, ctx->d_point,
                                                                  arr, ld_arr, prod, ld_prod, ncols, row_shift);
        else if (g.type == HYP_CONE_HYPOPOWERMEAN)
            hypdev::hpm_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_voff,
                                                                  g.d_vecs, g.d_dual, g.d_scal, ctx->d_point, arr,
                                                                  ld_arr, prod, ld_prod, ncols, row_shift);
        else
            hypdev::gpow_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, hess_dual, g.d_off, g.d_dim, g.d_hkind,
                                                                   g.d_voff, g.d_vecs, g.d_dual, g.d_scal,
                                                                   ctx->d_point, arr, ld_arr, prod, ld_prod, ncols,
                                                                   row_shift);
        ctx->launches++;
    }
    if (inv_dual > -2) {
        hypdev::gen_invhess_prod_kernel<<<grid, 256, 0, ctx->stream>>>(g.count, inv_dual, g.d_off, g.d_dim, g.d_moff,
                                                                      g.d_dual, g.d_Ui, arr, ld_arr, prod, ld_prod,
                                                                      ncols, row_shift);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
}

void hyp_gpow_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir) {
    if (g.type == HYP_CONE_EPITRRELENTROPYTRI)
        hypdev::etr_dder3_kernel<<<ceil_div(g.count, 4), 128, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_scal, dir, out);
    else if (g.type == HYP_CONE_POSSEMIDEFTRISPARSE)
        hypdev::sps_dder3_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, dir, out);
    else if (g.type == HYP_CONE_WSOSINTERPEPINORMONE)
        hypdev::wone_dder3_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs,
                                                                   dir, out);
    else if (g.type == HYP_CONE_WSOSINTERPEPINORMEUCL)
        hypdev::weuc_dder3_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs,
                                                                   dir, out);
    else if (g.type == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI)
        hypdev::wpsd_dder3_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs,
                                                                   dir, out);
    else if (g.type == HYP_CONE_MATRIXEPIPERSQUARE)
        hypdev::mep_dder3_kernel<<<ceil_div(g.count, 4), 128, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_scal, ctx->d_point, dir, out);
    else if (g.type == HYP_CONE_DOUBLYNONNEGATIVETRI)
        hypdev::dnn_dder3_kernel<<<ceil_div(g.count, 4), 128, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, ctx->d_point, dir, out);
    else if (g.type == HYP_CONE_LINMATRIXINEQ)
        hypdev::lmi_dder3_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, dir, out);
    else if (g.type == HYP_CONE_WSOSINTERPNONNEGATIVE)
        hypdev::wsos_dder3_kernel<<<g.count, 256, 0, ctx->stream>>>(g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, dir, out);
    else if (g.type == HYP_CONE_EPINORMSPECTRAL)
        hypdev::ens_dder3_kernel<<<ceil_div(g.count, 4), 128, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_scal, ctx->d_point, dir, out);
    else if (g.type == HYP_CONE_HYPOPOWERMEAN)
        hypdev::hpm_dder3_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_voff, g.d_vecs, g.d_scal, ctx->d_point, dir, out);
    else
        hypdev::gpow_dder3_kernel<<<ceil_div(g.count, 8), 256, 0, ctx->stream>>>(
            g.count, g.d_off, g.d_dim, g.d_hkind, g.d_voff, g.d_vecs, g.d_scal, ctx->d_point, dir, out);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
}

// Shared declarations of libhypatia_b200 (sm_100a only).
//
// Device data layout.  Everything is Float64, column-major.
//   G        local row panel of model.G (rows of the cones this rank owns), qloc x n, ld = ldg
//   HG       same shape as the panel restricted to the n-p reduced columns: per cone either
//            H_k^{1/2} G_k (cones with a closed-form square root) or H_k G_k (others)
//   S, F     (n-p) x (n-p) Schur matrix / its factor, upper triangle, ld = lds
//   q-vectors (point, dual, grad, z, s, h ...) have GLOBAL length q on every rank, in model
//   order; a rank reads/writes only its rows [row_lo, row_hi) in the cone kernels and G
//   passes, results are re-replicated with an allreduce (no-op for one rank).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/hypatia_b200.h"

#define HYP_NUM_CONE_TYPES 24
// internal product mode: inv_hess for primal-barrier cones, hess for dual-barrier cones
#define HYP_PROD_BLOCK_INV 5
#define HYP_EPS 2.220446049250313e-16

struct HypError {
    std::string msg;
};

#define CUDA_TRY(call)                                                                     \
    do {                                                                                   \
        cudaError_t err__ = (call);                                                        \
        if (err__ != cudaSuccess) {                                                        \
            char buf__[512];                                                               \
            snprintf(buf__, sizeof(buf__), "%s:%d: %s failed: %s", __FILE__, __LINE__,     \
                     #call, cudaGetErrorString(err__));                                    \
            throw HypError{buf__};                                                         \
        }                                                                                  \
    } while (0)

// ---- cone type classes ----
// matrix-domain cones keep `lead` scalars in front of an svec block
static inline bool cone_is_matrix(int t) {
    return t == HYP_CONE_POSSEMIDEFTRI || t == HYP_CONE_HYPOPERLOGDETTRI || t == HYP_CONE_HYPOROOTDETTRI ||
           t == HYP_CONE_EPIPERSEPSPECTRAL_MAT;
}
static inline int cone_mat_lead(int t) {
    return t == HYP_CONE_POSSEMIDEFTRI ? 0 : t == HYP_CONE_HYPOROOTDETTRI ? 1 : 2;
}
// cones whose Schur contribution uses the closed-form sqrt_hess_prod! (use_sqrt_hess_oracles = true:
// nonnegative.jl:38, epinormeucl.jl:40, possemideftri.jl:54, epipersquare.jl:48)
static inline bool cone_has_sqrt(int t) {
    return t == HYP_CONE_NONNEGATIVE || t == HYP_CONE_EPINORMEUCL || t == HYP_CONE_POSSEMIDEFTRI ||
           t == HYP_CONE_EPIPERSQUARE;
}
// cones that accept use_dual_barrier = true
static inline bool cone_allows_dual(int t) {
    return t == HYP_CONE_HYPOPERLOGDETTRI || t == HYP_CONE_HYPOROOTDETTRI ||
           t == HYP_CONE_EPIPERSEPSPECTRAL_MAT || t == HYP_CONE_HYPOPERLOG || t == HYP_CONE_EPINORMINF ||
           t == HYP_CONE_EPIPERSEPSPECTRAL_VEC || t == HYP_CONE_HYPOGEOMEAN || t == HYP_CONE_GENERALIZEDPOWER ||
           t == HYP_CONE_HYPOPOWERMEAN || t == HYP_CONE_EPIRELENTROPY || t == HYP_CONE_EPINORMSPECTRAL ||
           t == HYP_CONE_WSOSINTERPNONNEGATIVE || t == HYP_CONE_LINMATRIXINEQ || t == HYP_CONE_DOUBLYNONNEGATIVETRI ||
           t == HYP_CONE_MATRIXEPIPERSQUARE || t == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI ||
           t == HYP_CONE_WSOSINTERPEPINORMEUCL || t == HYP_CONE_WSOSINTERPEPINORMONE ||
           t == HYP_CONE_POSSEMIDEFTRISPARSE || t == HYP_CONE_EPITRRELENTROPYTRI;
}
// cones whose inverse-Hessian oracles are the generic Hessian-factorisation fallback (cones_gpow.cu)
static inline bool cone_is_genfact(int t) {
    return t == HYP_CONE_GENERALIZEDPOWER || t == HYP_CONE_HYPOPOWERMEAN || t == HYP_CONE_EPINORMSPECTRAL ||
           t == HYP_CONE_WSOSINTERPNONNEGATIVE || t == HYP_CONE_LINMATRIXINEQ || t == HYP_CONE_DOUBLYNONNEGATIVETRI ||
           t == HYP_CONE_MATRIXEPIPERSQUARE || t == HYP_CONE_WSOSINTERPPOSSEMIDEFTRI ||
           t == HYP_CONE_WSOSINTERPEPINORMEUCL || t == HYP_CONE_WSOSINTERPEPINORMONE ||
           t == HYP_CONE_POSSEMIDEFTRISPARSE || t == HYP_CONE_EPITRRELENTROPYTRI;
}
// vector cones served by cones_vec3_kernels.cuh
static inline bool cone_is_vec3(int t) {
    return t == HYP_CONE_EPIPERSQUARE || t == HYP_CONE_HYPOPERLOG || t == HYP_CONE_EPINORMINF ||
           t == HYP_CONE_EPIPERSEPSPECTRAL_VEC || t == HYP_CONE_HYPOGEOMEAN || t == HYP_CONE_EPIRELENTROPY;
}

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// One batched group of locally-owned cones of the same type (device arrays indexed by the
// position of the cone inside the group).
struct ConeGroup {
    int type = 0;
    int count = 0;
    int max_dim = 0;
    int max_side = 0;
    int64_t rows = 0;             // total rows of the group
    int64_t* d_off = nullptr;     // GLOBAL row offset of each cone
    int* d_dim = nullptr;         // dimension q_k
    int* d_kidx = nullptr;        // GLOBAL cone index
    int* d_dual = nullptr;        // use_dual_barrier flag
    int* d_side = nullptr;        // matrix side (matrix cones)
    int64_t* d_moff = nullptr;    // offset (doubles) of the cone's side x side state matrices
    std::vector<int64_t> h_off;
    std::vector<int> h_dim, h_kidx, h_dual, h_side;
    std::vector<int64_t> h_moff;
    int64_t mat_total = 0;        // doubles of one set of per-cone matrices (sum side^2, padded)
    // matrix-cone state: per cone side x side col-major at d_moff
    double* d_W = nullptr;        // point matrix (full symmetric)
    double* d_U = nullptr;        // upper Cholesky factor, W = U'U (lower part zero)
    double* d_Ut = nullptr;       // U' (lower triangular, stored explicitly)
    double* d_Ui = nullptr;       // U^-1 (upper)
    double* d_Uit = nullptr;      // U^-T (lower)
    double* d_Wi = nullptr;       // W^-1 (full symmetric)
    double* d_scal = nullptr;     // per-cone scalars (8 per cone)
    // EpiPerSepSpectral (cones_spec.cu): spectral function of each cone, per-cone vectors
    // (8 arrays of side doubles at d_voff; d_voff7 = offset of the 8th array)
    std::vector<int> h_hkind;
    std::vector<double> h_hparam;
    std::vector<int64_t> h_voff;
    int* d_hkind = nullptr;
    double* d_hparam = nullptr;
    int64_t *d_voff = nullptr, *d_voff7 = nullptr;
    double* d_vecs = nullptr;
    // row list for elementwise cones (Nonnegative): global row of every element
    int* d_rows = nullptr;
    int* d_rowcone = nullptr;     // GLOBAL cone index of every element of d_rows
    // chunk table of the many-column EpiNormEucl kernel (cones.cu)
    int n_chunks = 0, chunk_smem = 0;
    bool chunks_cover_all = false;
    int chunk_align8 = -1;        // fused pre-pass: every chunk starts at a multiple of 8 local rows (-1 = not checked yet)
    int64_t* d_crow0 = nullptr;
    int *d_crows = nullptr, *d_ccone0 = nullptr, *d_ccount = nullptr;
};

struct TimingSlot {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double total_ms = 0;
    long count = 0;
    bool open = false;
};

enum {
    T_CONE_STATE = 0,  // load_point: feas / grad / per-cone factorisations (K9)
    T_SQRT_PREPASS,    // H^{1/2} G or H G (K8)
    T_SYRK,            // Schur SYRK / GEMM (K1, K2)
    T_ALLREDUCE,       // NCCL
    T_POTRF,           // Cholesky (K3)
    T_LDLT,            // Bunch-Kaufman fallback (K4)
    T_TRSV,            // triangular solves (K5)
    T_GEMV,            // passes over G / A (K6)
    T_CONE_PROD,       // 1..2-column cone products (K7)
    T_VEC,             // O(q) vector glue
    T_NUM
};

struct hyp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;    // side stream of the Cholesky look-ahead chain
    cudaStream_t launch_stream = nullptr;   // stream the GEMM / panel launch helpers use (stream or stream2)
    int grid_cap = 0;                  // > 0: persistent GEMM grids leave SMs free for the side stream
    bool small_tiles = false;          // GEMM launch helpers use the 128 x 32 latency tile shape (Cholesky chain stream)
    cudaEvent_t ev_chain[2] = {nullptr, nullptr}, ev_bulk[2] = {nullptr, nullptr}, ev_near[2] = {nullptr, nullptr};
    int8_t* d_chol_digits = nullptr;   // digit slices of one 512-row block row of the Cholesky (chol.cu, potrf_upper_i8)
    double* d_chol_dscale = nullptr;
    int64_t chol_digits_cols = 0;
    std::string last_error;

    // ---- model ----
    int64_t n = 0, p = 0, q = 0;       // global dims
    int64_t nmp = 0;                   // n - p
    int K = 0;                         // global number of cones
    int cone_lo = 0, cone_hi = 0;      // locally owned cones [lo, hi)
    // Column sharding (single giant cone): every rank holds ALL rows of G and all cone state; only the
    // Schur assembly is split - rank r builds the column panel S[:, J_r] = GQ2' (H GQ2)[:, J_r] and the panels are
    // all-gathered.  Everything else (solves, oracles) runs replicated without communication.
    bool col_shard = false;
    int64_t col_shard_width = 0;       // columns of S per rank (0 when not column-sharded)
    int64_t row_lo = 0, row_hi = 0;    // locally owned rows
    int64_t qloc = 0;                  // row_hi - row_lo
    int64_t ldg = 0;                   // leading dim of G / HG panels (even)
    std::vector<int> h_cone_type, h_cone_dual;
    std::vector<int64_t> h_cone_dim, h_cone_off;
    std::vector<double> h_cone_nu;
    bool params_staged = false;        // hyp_set_cone_params was called since the last hyp_load_model (which consumes it)
    std::vector<int64_t> h_cone_aoff;  // per global cone: offsets into h_cone_alpha (hyp_set_cone_alpha), K + 1 entries
    std::vector<double> h_cone_alpha;  // powers of the GeneralizedPower cones
    std::vector<int> h_cone_hkind;     // per global cone: HYP_SSF_* (EpiPerSepSpectral), set by hyp_set_cone_params
    std::vector<double> h_cone_hparam;
    std::vector<int> h_cone_sqrt;      // per global cone: 1 = sqrt-form in the Schur assembly
    double* d_Graw = nullptr;          // qloc x n (model.G panel)
    double* d_GQ = nullptr;            // qloc x n, G*Ap_Q when p > 0 (else alias of d_Graw)
    double* d_HG = nullptr;            // qloc x nmp
    double* d_PG = nullptr;            // qloc x nmp, only for models mixing sqrt and non-sqrt cones
    uint8_t* d_row_ns = nullptr;       // qloc: 1 on rows of cones without a square-root oracle
    double* d_A = nullptr;             // p x n, ld = lda
    int64_t lda = 0;
    double* d_Q = nullptr;             // n x n (Ap_Q), ld = ldqm, or null
    int64_t ldqm = 0;
    double* d_R = nullptr;             // p x p upper (Ap_R), ld = ldr, or null
    int64_t ldr = 0;
    double* d_Rdinv = nullptr;         // inverted 128-diagonal blocks of R
    double* d_cbh = nullptr;           // (c, b, h)  length n+p+q
    std::vector<ConeGroup> groups;     // local cones grouped by type
    double* d_cone_nu = nullptr;       // K (global)
    int64_t* d_cone_off = nullptr;     // K (global)
    int64_t* d_cone_dim = nullptr;     // K
    int* d_cone_type = nullptr;        // K
    uint8_t* d_row_dual = nullptr;     // q: 1 on rows of dual-barrier cones (null if none)
    bool any_dual = false;

    // ---- cone state (global-length q-vectors) ----
    double *d_point = nullptr, *d_dual = nullptr, *d_grad = nullptr;
    double* d_wivec = nullptr;         // svec(W^-1) on the rows of matrix cones (q)
    uint8_t *d_feas = nullptr, *d_dual_feas = nullptr, *d_num_ok = nullptr;  // K
    uint8_t* d_tmpflag = nullptr;      // K scratch flags (Cholesky gate of the spectral cones' dual check)
    double* d_proxsqr = nullptr;       // K
    double* d_matwork = nullptr;       // scratch for matrix-cone products
    int64_t matwork_doubles = 0;
    bool cones_loaded = false;

    // ---- Schur complement / factorisation ----
    int64_t lds = 0;
    double* d_S = nullptr;             // nmp x nmp Schur matrix (upper)
    double* d_F = nullptr;             // factor
    double* d_Dinv = nullptr;          // inverted 128 x 128 diagonal blocks of the Cholesky factor
    int* d_info = nullptr;             // device info word(s)
    int* d_ipiv = nullptr;             // LDL' pivots (fallback)
    double* d_ldl_work = nullptr;      // LDL' workspace
    int fact_kind = 0;
    double mu = 1.0, tau_bar = 1.0;
    // ---- Schur SYRK on tcgen05 (ozaki.cu) ----
    int syrk_mode = 1;                 // 0: FP64 DMMA (syrk.cu), 1: sliced int8 on tcgen05 (ozaki.cu)
    int8_t* d_digits = nullptr;        // 8 x ldd x nmp digit slices of HG
    int* d_expo = nullptr;             // nmp column exponents
    double* d_dscale = nullptr;        // 2^expo
    int64_t ldd = 0;
    // ---- SymIndefDense variant (symindef.jl:203-271) ----
    int solver_kind = 0;               // 0 QRCholDense, 1 SymIndefDense
    int64_t ld3 = 0;                   // leading dim of the (n+p+q)^2 matrices
    double* d_L3 = nullptr;            // static part [0 A' G'; . 0 0; . . 0] (upper triangle)
    double* d_F3 = nullptr;            // factored copy
    int* d_row_cone = nullptr;         // q: global cone index of every row
    double* d_blk_arr = nullptr;       // q x maxdim identity pattern / products (2 buffers)
    int64_t blk_maxdim = 0;
    int* d_flags = nullptr;            // TRSV ticket + block-ready flags + segment counters (1 + 2 nblk ints)
    double* d_trsv_part = nullptr;     // partial sums of the segmented triangular solve
    unsigned long long* d_trsv_pkt = nullptr;   // {value half, epoch} packets of the triangular solves (trsv_pkt_kernel)
    int64_t trsv_pkt_words = 0;
    int64_t trsv_part_len = 0;
    // two-column solves (hyp_solve_system_multi): partial buffers of the two-vector GEMV kernels and a second set of
    // the work vectors of solve_system_dev / apply_lhs_dev
    double *d_partial3 = nullptr, *d_partial4 = nullptr;
    int64_t partial3_doubles = 0, partial4_doubles = 0;
    unsigned long long* d_colbits = nullptr;   // column maxima (bit patterns) of the fused pre-pass
    int8_t *d_gemm_digA = nullptr, *d_gemm_digB = nullptr;   // digit workspaces of hyp_ozaki_gemm_tn (matrix-cone congruences)
    double* d_gemm_scal = nullptr;
    int64_t gemm_digA_bytes = 0, gemm_digB_bytes = 0, gemm_scal_bytes = 0;
    int8_t* d_digitsP = nullptr;               // digit slices / scales of the P operand of the two-operand Schur product (mixed models)
    int* d_expoP = nullptr;
    double* d_dscaleP = nullptr;
    double* d_multi = nullptr;
    int64_t multi_doubles = 0;
    int* d_dag_ver = nullptr;          // task-graph Cholesky (chol_dag.cu): tile version counters + tickets
    int64_t dag_ver_len = 0;
    unsigned long long* d_dag_dbg = nullptr;   // HYP_POTRF_DEBUG: per-CTA wait / busy times
    int trsv_epoch = 0;

    // ---- vectors (device) ----
    double* d_rhs = nullptr;           // n+p+2q+2 staged full Point
    double* d_sol = nullptr;
    double* d_sub_sol = nullptr;       // n+p+q
    double* d_sub_rhs = nullptr;
    double* d_const_sol = nullptr;
    double* d_const_rhs = nullptr;
    double* d_Gx_const = nullptr;      // G * x_const (q), cached by update_lhs
    double *d_t = nullptr, *d_t2 = nullptr;            // n-vectors
    double *d_Gx = nullptr, *d_HGx = nullptr;          // q-vectors
    double *d_vq1 = nullptr, *d_vq2 = nullptr, *d_vq3 = nullptr, *d_vq4 = nullptr;
    double *d_vp1 = nullptr, *d_vp2 = nullptr;         // p-vectors
    double* d_partial = nullptr;       // reduction workspace
    int64_t partial_doubles = 0;
    double* d_partial2 = nullptr;      // per-warp partials of the fused G x / G' z pass (gemv.cu)
    int64_t partial2_doubles = 0;
    double* d_scalars = nullptr;       // small device scalar block (64 doubles)
    double* d_stage = nullptr;         // device staging for host-pointer arguments
    int64_t stage_doubles = 0;
    double* h_pinned = nullptr;        // pinned staging buffer
    int64_t pinned_doubles = 0;

    // ---- multi-GPU ----
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;

    // ---- timing / counters ----
    TimingSlot timing[T_NUM];
    bool timing_enabled = false;
    long launches = 0;
    bool model_loaded = false;
    bool lhs_ready = false;
};

// ---- timing helpers (vec_kernels.cu) ----
void hyp_time_begin(hyp_ctx* ctx, int slot);
void hyp_time_end(hyp_ctx* ctx, int slot);
struct TimeScope {
    hyp_ctx* c;
    int s;
    TimeScope(hyp_ctx* ctx, int slot) : c(ctx), s(slot) { hyp_time_begin(c, s); }
    ~TimeScope() { hyp_time_end(c, s); }
};

// ---- vec_kernels.cu ----
void hyp_axpby(hyp_ctx* ctx, int64_t len, double a, const double* x, double b, double* y);
void hyp_lincomb3(hyp_ctx* ctx, int64_t len, double* out, double a, const double* x, double b,
                  const double* y, double c, const double* z);
// y = a * x + bscal[0] * z   (device scalar)
void hyp_axpy_dev(hyp_ctx* ctx, int64_t len, double* out, double a, const double* x,
                  const double* dscal, double sign, const double* z);
void hyp_copy(hyp_ctx* ctx, int64_t len, double* dst, const double* src);
void hyp_fill(hyp_ctx* ctx, int64_t len, double* dst, double val);
void hyp_dot(hyp_ctx* ctx, int64_t len, const double* x, const double* y, double* d_out,
             bool accumulate);
void hyp_zero_outside(hyp_ctx* ctx, double* v);   // zero rows outside [row_lo,row_hi) of a q-vector

// ---- nccl_shim.cu ----
void hyp_allreduce_sum(hyp_ctx* ctx, double* buf, int64_t count);
// in-place all-gather: rank r's `count` doubles sit at buf + r * count
void hyp_allgather_inplace(hyp_ctx* ctx, double* buf, int64_t count);
// rows of G (and cones) are split over ranks: partial results need the exchange steps
static inline bool hyp_row_sharded(const hyp_ctx* ctx) { return ctx->nranks > 1 && !ctx->col_shard; }
void hyp_allreduce_min_u8(hyp_ctx* ctx, uint8_t* buf, int64_t count);
// make the local rows of a global q-vector visible on every rank
void hyp_replicate_q(hyp_ctx* ctx, double* v);
void hyp_comm_destroy(hyp_ctx* ctx);

// ---- ctx.cu ----
void hyp_build_pg(hyp_ctx* ctx);

// ---- gemv.cu ----
// y[0:ncols] = alpha * M' x + beta * y   (M: rows x ncols col-major)
void hyp_gemv_t(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld,
                const double* x, double alpha, double beta, double* y);
// y[0:rows] = alpha * M x + beta * y
void hyp_gemv_n(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld,
                const double* x, double alpha, double beta, double* y);
// both products in one pass over M: w = alphaN * M x + betaN * w, y = alphaT * M' z + betaT * y
bool hyp_gemv2_ok(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* a, const double* b);
void hyp_gemv_t2(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x0,
                 const double* x1, double alpha, double beta, double* y0, double* y1);
void hyp_gemv_n2(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x0,
                 const double* x1, double alpha, double beta, double* y0, double* y1);
void hyp_gemv_nt2(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* xa,
                  const double* xb, const double* za, const double* zb, double alphaN, double betaN, double* wa,
                  double* wb, double alphaT, double betaT, double* ya, double* yb);
void hyp_trsv_upper2(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, const double* d_dinv, double* x,
                     int64_t xstride, bool trans);
// 1: the triangular solves publish {value half, epoch} packets (trsv_pkt_kernel) instead of flags; 0 (default): flags
void hyp_trsv_set_pkt(int on);
void hyp_gemv_nt(hyp_ctx* ctx, int64_t rows, int64_t ncols, const double* M, int64_t ld, const double* x,
                 const double* z, double alphaN, double betaN, double* w, double alphaT, double betaT, double* y);

// ---- cones.cu ----
void hyp_cones_build_groups(hyp_ctx* ctx);
void hyp_cones_free_groups(hyp_ctx* ctx);
void hyp_cones_update_state(hyp_ctx* ctx);  // after points were loaded: feas, grad, factors
// prod = oracle(arr) on the local rows of q x ncols arrays (global row indexing, base pointers
// point at global row 0).  `row_shift`: subtracted from the global row to index arr/prod (used
// for the G panel, whose row 0 is global row row_lo).
void hyp_cones_prod(hyp_ctx* ctx, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                    int64_t ld_arr, int mode, int64_t row_shift);
// Schur pre-pass: HG[:, j] = H_k^{1/2} G_k[:, j] (sqrt cones) or H_k G_k[:, j] (others)
void hyp_cones_schur_prepass(hyp_ctx* ctx);
void hyp_cones_dder3_dev(hyp_ctx* ctx, double* out, const double* dir);
void hyp_cones_prox_dev(hyp_ctx* ctx, double irtmu, int use_max);

// ---- cones_gpow.cu: GeneralizedPower + the generic inverse-Hessian fallback ----
void hyp_gpow_alloc_group(hyp_ctx* ctx, ConeGroup& g);
void hyp_gpow_update_state(hyp_ctx* ctx, ConeGroup& g);
void hyp_gpow_prod(hyp_ctx* ctx, ConeGroup& g, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                   int64_t ld_arr, int mode, int64_t row_shift);
void hyp_gpow_dder3(hyp_ctx* ctx, ConeGroup& g, double* out, const double* dir);

// ---- syrk.cu ----
// C(upper 128-tiles) = alpha * P' R + beta * C over k in [0, klen); P, R: klen x ncols col-major.
void hyp_atb_upper(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                   int64_t klen, int64_t ncols, double* C, int64_t ldc, double alpha, double beta);
// C (mrows x ncols, full) = alpha * P' R + beta * C; P: klen x mrows, R: klen x ncols.
void hyp_gemm_tn(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                 int64_t klen, int64_t mrows, int64_t ncols, double* C, int64_t ldc, double alpha,
                 double beta);
// `ngroups` independent products C_g = alpha * P' R_g + beta * C_g, where R_g is the row block
// [g * r_kstride, g * r_kstride + klen) of R and C_g = C + g * c_group_stride
void hyp_gemm_tn_grouped(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                         int64_t klen, int64_t mrows, int64_t ncols, int ngroups, int64_t r_kstride,
                         double* C, int64_t ldc, int64_t c_group_stride, double alpha, double beta);
// simple CUDA-core GEMM for one-off products at load time: C = op(A) op(B)
void hyp_gemm_simple(hyp_ctx* ctx, bool transA, bool transB, int64_t M, int64_t N, int64_t Kd,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                     int64_t ldc);

// ---- ozaki.cu ----
int hyp_ozaki_radix();
// fused Schur pre-pass + digit slicing for second-order-cone models (cones.cu); false = not applicable
int hyp_cones_prepass_sliced(hyp_ctx* ctx, int8_t* digits, int64_t ldd, int64_t slice_stride, int* expo, double* dscale);
bool hyp_ozaki_pair64_ready(hyp_ctx* ctx);
bool hyp_ozaki_gemm_tn(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr, int64_t klen, int64_t mrows,
                       int64_t ncols, double* C, int64_t ldc, double alpha, double beta, int ngroups = 1,
                       int64_t r_kstride = 0, int64_t c_group_stride = 0);
void hyp_ozaki_slice_short(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols, int8_t* digits, int64_t ldd,
                           int64_t slice_stride, double* dscale);
void hyp_ozaki_syrk_rows(hyp_ctx* ctx, const int8_t* digits, int64_t ldd, int64_t slice_stride, const double* dscale,
                         int64_t K, int64_t ncols, double* C, int64_t ldc, double alpha, double beta, int p_lo, int p_hi,
                         int skip_diag);
// have_expo: expo / dscale are already filled (the pre-pass collected the column maxima): only the digit slicing runs
void hyp_ozaki_slice(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols, int8_t* digits,
                     int64_t ldd, int64_t slice_stride, int* expo, double* dscale, bool have_expo = false);
void hyp_ozaki_syrk(hyp_ctx* ctx, const int8_t* digits, int64_t ldd, int64_t slice_stride, const int* expo,
                    const double* dscale, int64_t K, int64_t ncols, double* C, int64_t ldc, double alpha, double beta,
                    const int8_t* digitsB = nullptr, const double* dscaleB = nullptr);   // B side of C = P' R (nullptr: SYRK)

// ---- chol.cu ----
// in-place blocked upper Cholesky; d_dinv receives the inverted 128 x 128 diagonal blocks
// (ceil(m/128) blocks of 128*128 doubles); d_info[0] = 0 or 1-based index of the first bad pivot
void hyp_potrf_upper(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info);
bool hyp_potrf_upper_dag(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* d_dinv, int* d_info);
// batched Cholesky + triangular inverse of the (side <= 128) matrices of a cone group
void hyp_chol_batched(hyp_ctx* ctx, int ncones, const int* d_sides, const int64_t* d_moff,
                      const int* d_kidx, double* U, double* Ui, uint8_t* d_flag);
// inverted diagonal blocks of an already-triangular upper matrix (Ap_R)
void hyp_trtri_diag(hyp_ctx* ctx, const double* U, int64_t ldu, int64_t m, double* d_dinv);
// x <- U^-1 x (trans = false) or U^-T x (trans = true)
void hyp_trsv_upper(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, const double* d_dinv,
                    double* x, bool trans);

// ---- ldlt.cu ----
// rook-pivoted LDL' (upper storage) of the m x m matrix in A (in place); returns nothing,
// d_info[0] = 0 ok / > 0 zero pivot.  Solve: x <- A^-1 x.
void hyp_ldlt_factor(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, int* d_ipiv, int* d_info);
void hyp_ldlt_solve(hyp_ctx* ctx, const double* A, int64_t lda, int64_t m, const int* d_ipiv,
                    double* x);
void hyp_increase_diag(hyp_ctx* ctx, double* A, int64_t lda, int64_t m);

/*
 * libhypatia_b200 - C ABI of the B200-native (sm_100a) Hypatia KKT / cone-oracle hot path.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b): the entry points a Julia
 * `Solvers.SystemSolver{Float64}` subtype and a batched `Cones.Cone{Float64}` container bind
 * with `ccall` (julia/HypatiaB200.jl), and that the Python test driver binds with ctypes
 * (hypatia.jl_b200/capi.py).  Plain pointers and sizes only; every array is Float64, Julia
 * (column-major) layout.  All file:line citations are relative to the reference tree
 * (chriscoey/Hypatia.jl v0.5.1).
 *
 * Conventions
 *   - every function returns int: 0 ok; > 0 numerical condition (documented per call);
 *     < 0 CUDA / NCCL / argument error, message in hyp_last_error(ctx).  Nothing throws or
 *     aborts across the ABI (the reference never throws on numerical failure either:
 *     qrchol.jl:252-254, dense.jl:194-215).
 *   - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer of
 *     the context's GPU; the library detects which (cudaPointerGetAttributes) and stages host
 *     buffers through its own pinned buffer.  Calls are synchronous on return for host
 *     pointers and stream-ordered on hyp_stream(ctx) for device pointers.
 *   - one host thread per context; a context owns one GPU (one process per GPU; multi-GPU
 *     runs create one context per rank and join them with hyp_comm_init).
 *   - vectors crossing the ABI use the reference's Point layout (point.jl:24-54):
 *     full Point  = [x(n); y(p); z(q); tau; s(q); kap]   length n+p+2q+2
 *     sub Point   = [x(n); y(p); z(q)]                   length n+p+q
 *     q-vectors are cone after cone in model order (Models.jl:56-66), GLOBAL length q on
 *     every rank.
 */
#ifndef HYPATIA_B200_H
#define HYPATIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hyp_ctx hyp_ctx;

/* cone type codes (reference types in src/Cones/) */
#define HYP_CONE_NONNEGATIVE 0      /* nonnegative.jl      */
#define HYP_CONE_EPINORMEUCL 1      /* epinormeucl.jl      */
#define HYP_CONE_POSSEMIDEFTRI 2    /* possemideftri.jl (real) */
#define HYP_CONE_HYPOPERLOGDETTRI 3 /* hypoperlogdettri.jl */
#define HYP_CONE_HYPOROOTDETTRI 4   /* hyporootdettri.jl   */
#define HYP_CONE_EPIPERSEPSPECTRAL_MAT 5 /* epipersepspectral/{epipersepspectral,matrixcsqr}.jl (real) */
#define HYP_CONE_EPIPERSQUARE 6     /* epipersquare.jl     */
#define HYP_CONE_HYPOPERLOG 7       /* hypoperlog.jl       */
#define HYP_CONE_EPINORMINF 8       /* epinorminf.jl (real); use_dual = 1: l1-norm epigraph */
#define HYP_CONE_EPIPERSEPSPECTRAL_VEC 9 /* epipersepspectral/{epipersepspectral,vectorcsqr}.jl */
#define HYP_CONE_HYPOGEOMEAN 10     /* hypogeomean.jl      */
#define HYP_CONE_GENERALIZEDPOWER 11 /* generalizedpower.jl (powers via hyp_set_cone_alpha; dim <= 128) */
#define HYP_CONE_HYPOPOWERMEAN 12   /* hypopowermean.jl (dim - 1 powers via hyp_set_cone_alpha; dim <= 128) */
#define HYP_CONE_EPIRELENTROPY 13   /* epirelentropy.jl (u, v[d], w[d]); dim = 1 + 2 d */
#define HYP_CONE_WSOSINTERPNONNEGATIVE 15 /* wsosinterpnonnegative.jl (real): dim = U <= 128; the interpolation matrices travel
                                       in the per-cone array of hyp_set_cone_alpha as [nP, L_1 .. L_nP, vec(P_1) .. vec(P_nP)]
                                       (P_k is U x L_k, column-major); use_dual_barrier = 1 is the reference's default */
#define HYP_CONE_LINMATRIXINEQ 16   /* linmatrixineq.jl (real dense A_i): dim = number of matrices <= 128; the matrices travel in the
                                       per-cone array of hyp_set_cone_alpha as [side, vec(A_1) .. vec(A_dim)] */
#define HYP_CONE_DOUBLYNONNEGATIVETRI 17 /* doublynonnegativetri.jl: svec of a psd AND entrywise nonnegative matrix; dim <= 128 */
#define HYP_CONE_MATRIXEPIPERSQUARE 18 /* matrixepipersquare.jl (real): (svec(U), v, vec(W)), U d1 x d1, W d1 x d2, d1 <= d2;
                                       d1 is given as the integer parameter of hyp_set_cone_params; dim <= 128 */
#define HYP_CONE_WSOSINTERPPOSSEMIDEFTRI 19 /* wsosinterppossemideftri.jl: R x R matrix polynomials, dim = U svec_length(R) <= 128;
                                       R is the integer parameter of hyp_set_cone_params, the Ps travel like code 15 */
#define HYP_CONE_WSOSINTERPEPINORMEUCL 20 /* wsosinterpepinormeucl.jl: R >= 2 polynomials, dim = U R <= 128; R and the Ps as for 19 */
#define HYP_CONE_WSOSINTERPEPINORMONE 21 /* wsosinterpepinormone.jl: R >= 2 polynomials, dim = U R <= 128; R and the Ps as for 19 */
#define HYP_CONE_POSSEMIDEFTRISPARSE 22 /* possemideftrisparse/ (real; dense implementation like the reference's PSDSparseDense):
                                       dim = number of pattern entries <= 128; hyp_set_cone_alpha carries
                                       [side, row_1 .. row_dim, col_1 .. col_dim] (0-based, col <= row, every diagonal once) */
#define HYP_CONE_EPITRRELENTROPYTRI 23 /* epitrrelentropytri.jl: (u, svec(V), svec(W)), dim = 1 + 2 svec_length(d) <= 128 */
#define HYP_CONE_EPINORMSPECTRAL 14 /* epinormspectral.jl (real): (u, vec(W)), W d1 x d2 column-major, d1 <= d2; d1 is given
                                       as the integer parameter of hyp_set_cone_params; use_dual = 1: nuclear norm; dim <= 128 */

/* separable spectral functions of EpiPerSepSpectral (epipersepspectral/sepspectralfun.jl:17-116) */
#define HYP_SSF_INV 0        /* InvSSF        x -> 1/x      */
#define HYP_SSF_NEGLOG 1     /* NegLogSSF     x -> -log x   */
#define HYP_SSF_NEGENTROPY 2 /* NegEntropySSF x -> x log x  */
#define HYP_SSF_POWER12 3    /* Power12SSF(p) x -> x^p, 1 < p <= 2 */

/* modes of hyp_cones_hess_prod (Cones.jl oracle names) */
#define HYP_PROD_HESS 0          /* hess_prod!          */
#define HYP_PROD_INV_HESS 1      /* inv_hess_prod!      */
#define HYP_PROD_SQRT_HESS 2     /* sqrt_hess_prod!     (Nonnegative/EpiNormEucl/PosSemidefTri) */
#define HYP_PROD_INV_SQRT_HESS 3 /* inv_sqrt_hess_prod! (same cones) */
#define HYP_PROD_BLOCK 4         /* block_hess_prod!, qrchol.jl:87-98 */

/* ---- life cycle -------------------------------------------------------------------- */
int hyp_version(void);
/* one context on CUDA device `device`; NULL on failure (no CUDA device => no library) */
hyp_ctx* hyp_create(int device);
/* replaces free_memory(syssolver), Solvers.jl:407,582-584 */
void hyp_destroy(hyp_ctx* ctx);
const char* hyp_last_error(hyp_ctx* ctx);
/* cudaStream_t the context launches on (as void*) */
void* hyp_stream(hyp_ctx* ctx);
int hyp_sync(hyp_ctx* ctx);

/* ---- multi-GPU: cones / row panels of G sharded over ranks ----------- */
/* rank 0 creates a 128-byte NCCL unique id, the host broadcasts it, every rank joins. */
int hyp_comm_unique_id(char* id128);
int hyp_comm_init(hyp_ctx* ctx, int rank, int nranks, const char* id128);

/* ---- per-cone parameters beyond (type, dim, use_dual): the `h::SepSpectralFun` field of
 *      EpiPerSepSpectral (epipersepspectral.jl:28-32).  ssf_kind[k] = HYP_SSF_*, ssf_param[k] = the power
 *      of Power12SSF; entries of other cone types are ignored.  Call BEFORE hyp_load_model (the
 *      values are consumed by the next load); models without such cones need not call it. */
int hyp_set_cone_params(hyp_ctx* ctx, int K, const int* ssf_kind, const double* ssf_param);
/* the `alpha::Vector` field of GeneralizedPower (generalizedpower.jl:11): cone k owns alpha[alpha_off[k] ..
 * alpha_off[k + 1]) (empty for other cone types; n = dim_k - number of powers).  Call BEFORE hyp_load_model. */
int hyp_set_cone_alpha(hyp_ctx* ctx, int K, const int64_t* alpha_off, const double* alpha);

/* ---- load: replaces load(syssolver::QRCholDenseSystemSolver, solver), qrchol.jl:138-179,
 *      setup_point_sub common.jl:184-208, and setup_data!(cone) for every cone.
 * G_local: the rows of model.G owned by this rank, i.e. rows of cones [cone_lo, cone_hi)
 *          (all q rows when the context is not sharded), q_local x n, leading dim ldG.
 * A: p x n (NULL when p == 0); c (n), b (p), h (q, GLOBAL).
 * cone_type/cone_dim/cone_dual: K entries (GLOBAL cone list; offsets are the running sum).
 * Ap_Q (n x n) / Ap_R (p x p upper): solver.Ap_Q / solver.Ap_R when p > 0 (QR of A'), else NULL.
 */
int hyp_load_model(hyp_ctx* ctx, int64_t n, int64_t p, int64_t q, const double* G_local,
                   int64_t ldG, const double* A, int64_t ldA, const double* c, const double* b,
                   const double* h, int K, const int* cone_type, const int64_t* cone_dim,
                   const int* cone_dual, int cone_lo, int cone_hi, const double* Ap_Q,
                   const double* Ap_R);

/* ---- cone oracles (plugin slot 2; batched over all K cones) ---------------------------- */
/* load_point(cone, primal_k, scal) + load_dual_point + reset_data for every cone
 * (Cones.jl:157-161,185-186; caller search.jl:121-123), then update_feas / update_grad and the
 * per-cone factorisations (K9).  primal/dual: GLOBAL q-vectors. */
int hyp_cones_load_point(hyp_ctx* ctx, const double* primal, const double* dual, double scal);
/* is_feas / is_dual_feas for every cone (K bytes each, 1 = feasible) */
int hyp_cones_feas(hyp_ctx* ctx, uint8_t* is_feas, uint8_t* is_dual_feas);
/* grad(cone) for every cone (q) */
int hyp_cones_grad(hyp_ctx* ctx, double* grad);
/* prod[:, 0:ncols] = oracle(arr[:, 0:ncols]) for every cone block; arr/prod are q x ncols */
int hyp_cones_hess_prod(hyp_ctx* ctx, double* prod, const double* arr, int64_t ncols,
                        int64_t ld_prod, int64_t ld_arr, int mode);
/* dder3(cone, dir_k) for every cone (q) */
int hyp_cones_dder3(hyp_ctx* ctx, double* out, const double* dir);
/* check_numerics (Cones.jl:273-290) and get_proxsqr (Cones.jl:294-310, nonnegative.jl:137-145) */
int hyp_cones_proxsqr(hyp_ctx* ctx, double irtmu, int use_max, double* proxsqr,
                      uint8_t* numerics_ok);

/* explicit hess(cone) (inverse = 0) / inv_hess(cone) (inverse = 1) of every cone, Cones.jl:79-93 and
 * the per-cone update_hess / update_inv_hess: K column-major dim_k x dim_k blocks packed one after
 * the other (block k starts at sum_{j<k} dim_j^2). */
int hyp_cones_hess_blocks(hyp_ctx* ctx, double* blocks, int inverse);

/* ---- cone oracles, ONE cone per handle (plugin slot 2, SURVEY.md 8(b) "per-cone single-block variants") -------
 * What a `B200Cone <: Cones.Cone{Float64}` object of the Julia shim forwards its methods to: the per-cone oracle API of
 * src/Cones/Cones.jl:34-310 with the reference's lazy evaluation (the loads only copy, the first query evaluates).
 * Each handle owns a private one-cone context and runs the same kernels as the batched calls above (a batch of one);
 * the batched calls remain the hot path of a solve. */
typedef struct hyp_cone hyp_cone;
/* constructor + setup_data!(cone) (Cones.jl:139-152).  iparam / dparam: the integer / real parameter of
 * hyp_set_cone_params for this cone type (HYP_SSF_* code and power of EpiPerSepSpectral, d1 of EpiNormSpectral /
 * MatrixEpiPerSquare, R of the WSOS matrix / norm cones; 0 otherwise); alpha[0 .. nalpha): the per-cone array of
 * hyp_set_cone_alpha (NULL / 0 for cone types without one).  NULL on failure (message on stderr). */
hyp_cone* hyp_cone_create(int device, int cone_type, int64_t dim, int use_dual_barrier, int iparam, double dparam,
                          const double* alpha, int64_t nalpha);
void hyp_cone_destroy(hyp_cone* cone);
const char* hyp_cone_last_error(hyp_cone* cone);
int64_t hyp_cone_dimension(hyp_cone* cone);      /* dimension(cone), Cones.jl:34 */
double hyp_cone_nu(hyp_cone* cone);              /* get_nu(cone), Cones.jl:41 */
int hyp_cone_use_dual_barrier(hyp_cone* cone);   /* use_dual_barrier(cone), Cones.jl:138 */
/* load_point(cone, point, scal): cone.point = scal * point (Cones.jl:157-161; scal = 1 for the two-argument form
 * :163-166) followed by reset_data(cone); load_dual_point(cone, point) (:168-171); reset_data(cone) (:185-186) */
int hyp_cone_load_point(hyp_cone* cone, const double* point, double scal);
int hyp_cone_load_dual_point(hyp_cone* cone, const double* dual_point);
int hyp_cone_reset_data(hyp_cone* cone);
/* is_feas(cone) / is_dual_feas(cone) (Cones.jl:56-69); either output pointer may be NULL */
int hyp_cone_is_feas(hyp_cone* cone, int* is_feas, int* is_dual_feas);
/* grad(cone) (Cones.jl:71-77): dim values */
int hyp_cone_grad(hyp_cone* cone, double* grad);
/* hess(cone) (inverse = 0, Cones.jl:79-84) / inv_hess(cone) (inverse = 1, :86-93): dim x dim, column-major */
int hyp_cone_hess(hyp_cone* cone, double* H, int inverse);
/* hess_prod! / inv_hess_prod! / sqrt_hess_prod! / inv_sqrt_hess_prod! (mode = HYP_PROD_*; Cones.jl:101-118,198-218
 * and the per-cone files) on a dim x ncols array */
int hyp_cone_hess_prod(hyp_cone* cone, double* prod, const double* arr, int64_t ncols, int64_t ld_prod,
                       int64_t ld_arr, int mode);
/* use_sqrt_hess_oracles(arr_dim, cone) of the cone types with closed-form square-root oracles (1 / 0) */
int hyp_cone_use_sqrt_hess_oracles(hyp_cone* cone);
/* dder3(cone, dir) (Cones.jl:134 and the per-cone files) */
int hyp_cone_dder3(hyp_cone* cone, double* out, const double* dir);
/* get_proxsqr(cone, irtmu, use_max_prox) (Cones.jl:294-310) and check_numerics(cone) (:273-290) from one sweep;
 * either output pointer may be NULL */
int hyp_cone_proxsqr(hyp_cone* cone, double irtmu, int use_max_prox, double* proxsqr, int* numerics_ok);

/* ---- system solver (plugin slot 1) ------------------------------------------------------ */
/* which reference system solver the context restates: 0 = QRCholDenseSystemSolver (default,
 * qrchol.jl:104-257), 1 = SymIndefDenseSystemSolver (symindef.jl:203-271: dense (n+p+q)^2
 * symmetric-indefinite LHS, rook Bunch-Kaufman).  Call after hyp_load_model. */
int hyp_set_syssolver(hyp_ctx* ctx, int kind);
/* Column sharding for models dominated by ONE cone (BASELINE config 5's natvsext shape: one
 * HypoPerLogdetTri of side 1000).  Call after hyp_comm_init and BEFORE hyp_load_model, then load the model with ALL
 * rows on every rank (cone_lo = 0, cone_hi = K, G_local = G).  hess_prod! is independent per column of G_k
 * (src/Cones/hypoperlogdettri.jl:196-237), so rank r assembles the column panel S[:, J_r] = GQ2' (H GQ2)[:, J_r]
 * (the branch src/Solvers/systemsolvers/qrchol.jl:240-246) and one ncclAllGather completes S; solves and oracles
 * run replicated with no further exchange. */
int hyp_set_column_sharding(hyp_ctx* ctx, int on);
/* how the Schur SYRK runs: 0 = FP64 DMMA (mma.sync), 1 = FP64-accurate digit slicing on the int8
 * tcgen05 pipe (csrc/ozaki.cu).  Default 1 (0 when the environment has HYP_SCHUR_SYRK=dmma).  Models that
 * mix square-root and non-square-root cones take the two-operand product S = P'(HG), digit-sliced on tcgen05
 * as well since round 2 (HYP_K2_DMMA=1 restores FP64 DMMA for it). */
int hyp_set_syrk_mode(hyp_ctx* ctx, int mode);
/* mu and tau of the current iterate (solver.mu, solver.point.tau[]) used by
 * solve_subsystem4 / solve_system / apply_lhs (common.jl:117,171-175,147) */
int hyp_set_mu_tau(hyp_ctx* ctx, double mu, double tau_bar);
/* update_lhs(syssolver, solver), qrchol.jl:181-257: Schur assembly (K8+K1+K2), factorisation
 * with the posdef_fact_copy! chain (dense.jl:194-215), constant column solve.
 * fact_kind: 0 Cholesky, 1 Bunch-Kaufman, 2 shifted Bunch-Kaufman.
 * returns 0 ok, 1 Cholesky failed but a fallback succeeded, 2 every factorisation failed. */
int hyp_update_lhs(hyp_ctx* ctx, int* fact_kind);
/* solve_subsystem3(syssolver, solver, sol, rhs), qrchol.jl:39-85 (sub Points) */
int hyp_solve_subsystem3(hyp_ctx* ctx, double* sol, const double* rhs);
/* solve_system(syssolver, solver, sol, rhs), common.jl:129-151 (full Points) */
int hyp_solve_system(hyp_ctx* ctx, double* sol, const double* rhs);
/* apply_lhs(stepper, solver) restricted to its data flow: res = LHS6x6 * dir, common.jl:79-121 */
int hyp_apply_lhs(hyp_ctx* ctx, double* res, const double* dir);
/* Multi-column forms of the two calls above: column j is the full Point at offset j * ld (ld >= n+p+2q+2).
 * The stepper's data flow (src/Solvers/steppers/combined.jl:67-79) allows {cent, pred} and then
 * {centadj, predadj} to be solved together, so that the triangular sweeps and the passes over G are shared
 * (SURVEY.md 8(d)); results are identical to ncols calls of hyp_solve_system / hyp_apply_lhs. */
int hyp_solve_system_multi(hyp_ctx* ctx, double* sol, const double* rhs, int ncols, int64_t ld);
int hyp_apply_lhs_multi(hyp_ctx* ctx, double* res, const double* dir, int ncols, int64_t ld);

/* residual step on the other side of the path: the vectors and norms of
 * calc_convergence_params(solver), Solvers.jl:425-483, for the full Point `point`, with the two passes
 * over G done on the device.  x_residual (n) = -(A'y + G'z + c tau), y_residual (p) = A x - b tau,
 * z_residual (q) = s + G x - h tau; stats (10) = |A'y + G'z|_inf, |A'y + G'z + c tau|_inf, |A x|_inf,
 * |A x - b tau|_inf, |s + G x|_inf, |s + G x - h tau|_inf, c'x, b'y, h'z, z's. */
int hyp_calc_residuals(hyp_ctx* ctx, const double* point, double* x_residual, double* y_residual,
                       double* z_residual, double* stats);

/* ---- introspection used by tests / bench ----------------------------------------------- */
/* upper triangle of the assembled Schur matrix (n-p x n-p, leading dim ld) */
int hyp_get_schur(hyp_ctx* ctx, double* S, int64_t ld);
/* kernels launched by this context so far */
int64_t hyp_launch_count(hyp_ctx* ctx);
/* per-phase device timers: enable, then read accumulated ms / calls per slot and reset */
int hyp_timing_enable(hyp_ctx* ctx, int on);
int hyp_timing_get(hyp_ctx* ctx, int slot, double* total_ms, int64_t* calls);
int hyp_timing_reset(hyp_ctx* ctx);
int hyp_timing_slots(void);
const char* hyp_timing_name(int slot);

/* ---- building blocks exported for unit tests (tests/test_gpu_kernels.py) ---------------- */
/* C(upper 128-tiles) = alpha * P' R + beta * C; P, R: klen x ncols col-major (device or host) */
int hyp_test_atb_upper(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                       int64_t klen, int64_t ncols, double* C, int64_t ldc, double alpha,
                       double beta);
/* C = alpha * P' R + beta * C, full mrows x ncols */
int hyp_test_gemm_tn(hyp_ctx* ctx, const double* P, int64_t ldp, const double* R, int64_t ldr,
                     int64_t klen, int64_t mrows, int64_t ncols, double* C, int64_t ldc,
                     double alpha, double beta);
/* in-place upper Cholesky A = U'U; *info = 0 or index (1-based) of the failing pivot block */
int hyp_test_potrf(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, int* info);
/* profiling aid: clock64 phase timestamps of the one-CTA factor-and-invert kernel on an m x m block (m <= 128) */
int hyp_test_panel_clocks(hyp_ctx* ctx, double* A, int64_t lda, int64_t m, double* cycles16);
/* protocol of the triangular solves (process-wide): 0 = block flags (default), 1 = {value half, epoch} packets polled by
 * the consumer threads themselves (csrc/chol_kernels.cuh trsv_pkt_kernel; also HYP_TRSV_PKT=1).  Same arithmetic, same bits. */
int hyp_test_set_trsv_pkt(int on);
/* x = (U'U)^-1 x with the factor of the last hyp_test_potrf */
int hyp_test_potrs(hyp_ctx* ctx, const double* F, int64_t ldf, int64_t m, double* x);
/* y = alpha * op(M) x + beta * y */
int hyp_test_gemv(hyp_ctx* ctx, int trans, int64_t rows, int64_t cols, const double* M,
                  int64_t ld, const double* x, double alpha, double beta, double* y);
/* rook-pivoted LDL' factor + solve of a symmetric matrix given by its upper triangle
 * (device restatement of symm_fact!, dense.jl:164-165); returns info */
int hyp_test_ldlt_solve(hyp_ctx* ctx, const double* A, int64_t lda, int64_t m, double* x,
                        int* info);

/* experimental tcgen05 (kind::i8) building blocks of the FP64-by-slicing Schur SYRK (csrc/ozaki.cu):
 * C (int32) = A' B with int8 K-major operands (A: K x M, B: K x N);  signed 7-bit digit slices and
 * per-column exponents of a K x ncols FP64 matrix */
int hyp_test_i8_gemm_tn(hyp_ctx* ctx, const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb,
                        int64_t K, int64_t M, int64_t N, int32_t* C, int64_t ldc);
int hyp_test_ozaki_slices(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols,
                          int nslices, int8_t* digits, int* expo);
/* C(upper 128-tiles) = A' A through slicing + tcgen05 (FP64-accurate) */
/* profiling aid: issue rate of tcgen05.mma kind::i8 from shared memory (swz 0/1/2 = SWIZZLE_32B/64B/128B K-major,
 * N = 128 / 256, cg2 = cta_group::2): out[0] = SM cycles per MMA, out[1] = ms of the launch on `ctas` CTAs */
int hyp_test_mma_rate(hyp_ctx* ctx, int swz, int N, int cg2, int nmma, int ctas, double* out);
int hyp_test_ozaki_syrk(hyp_ctx* ctx, const double* A, int64_t lda, int64_t K, int64_t ncols, double* C,
                        int64_t ldc);

#ifdef __cplusplus
}
#endif
#endif /* HYPATIA_B200_H */

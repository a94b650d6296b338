// Per-cone single-block entry points (plug-in slot 2): one device cone behind the reference's
// per-cone oracle API, src/Cones/Cones.jl:34-310 - what a `B200Cone <: Cones.Cone{Float64}` object of the Julia shim
// forwards its methods to (julia/HypatiaB200.jl), and what hypatia.jl_b200/cones.py `DeviceCone` binds.
//
// A hyp_cone owns a private context that holds a ONE-cone table and no G columns, so every oracle below runs the
// same kernels as the batched hyp_cones_* calls (a batch of one).  The reference's cones are lazy - load_point /
// load_dual_point only copy, reset_data clears the *_updated flags, and the first oracle call after that evaluates
// (Cones.jl:56-77,157-186); the handle keeps that contract: the loads only stage the vectors, the first query after a
// load evaluates feasibility / gradient / factorisations in one device sweep.
//
// The batched calls stay the hot path (one launch serves all K cones of a model); these entry points exist so that
// generic reference code that touches ONE cone (tests, initialize_cone_point, user callbacks) has a drop-in object.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/hypatia_b200.h"

struct hyp_cone {
    hyp_ctx* ctx = nullptr;
    int device = 0;
    int type = 0;
    int use_dual = 0;
    int64_t dim = 0;
    double nu = 0.0;
    double* d_in = nullptr;      // [raw primal point (dim); dual point (dim)] as handed to load_point / load_dual_point
    double scal = 1.0;           // scal of the last load_point(cone, point, scal)
    bool have_point = false;
    bool stale = true;           // reset_data: the device state does not belong to the staged vectors
    std::string err;
};

namespace {

int fail(hyp_cone* c, const std::string& what) {
    c->err = what;
    return -1;
}

int inner(hyp_cone* c, int rc, const char* what) {
    if (rc < 0) {
        const char* m = hyp_last_error(c->ctx);
        c->err = std::string(what) + ": " + (m ? m : "?");
    }
    return rc;
}

int stage(hyp_cone* c, double* dst, const double* src) {
    cudaStream_t s = (cudaStream_t)hyp_stream(c->ctx);
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, (size_t)c->dim * 8, cudaMemcpyDefault, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // the caller may reuse `src` on return
    if (e != cudaSuccess) return fail(c, std::string("staging copy: ") + cudaGetErrorString(e));
    return 0;
}

// update_feas / update_grad / factorisations of the staged point (the lazy evaluation of Cones.jl:56-77)
int ensure_state(hyp_cone* c) {
    if (!c->have_point) return fail(c, "no point loaded (call hyp_cone_load_point first)");
    if (!c->stale) return 0;
    int rc = inner(c, hyp_cones_load_point(c->ctx, c->d_in, c->d_in + c->dim, c->scal), "hyp_cones_load_point");
    if (rc < 0) return rc;
    c->stale = false;
    return 0;
}

// barrier parameter: get_nu of the per-cone files (the same table hyp_load_model fills for the batched path)
double cone_nu(int t, int64_t d, int iparam, const double* alpha, int64_t nalpha) {
    auto side_of = [](int64_t len) { return (double)(int64_t)((std::sqrt(1.0 + 8.0 * (double)len) - 1.0) / 2.0 + 0.5); };
    auto wsos_nu = [&]() {   // sum of the L_k of the packed interpolation data [nP, L_1 .. L_nP, ...]
        double s = 0;
        if (alpha && nalpha >= 1)
            for (int64_t i = 0; i < (int64_t)alpha[0] && 1 + i < nalpha; i++) s += alpha[1 + i];
        return s;
    };
    switch (t) {
        case HYP_CONE_NONNEGATIVE: return (double)d;                       // nonnegative.jl:40
        case HYP_CONE_EPINORMEUCL: return 2.0;                             // epinormeucl.jl:42
        case HYP_CONE_POSSEMIDEFTRI: return side_of(d);                    // possemideftri.jl:67
        case HYP_CONE_HYPOPERLOGDETTRI: return 2.0 + side_of(d - 2);       // hypoperlogdettri.jl:80
        case HYP_CONE_HYPOROOTDETTRI: return 1.0 + side_of(d - 1);         // hyporootdettri.jl:80
        case HYP_CONE_EPIPERSEPSPECTRAL_MAT: return 2.0 + side_of(d - 2);  // epipersepspectral.jl:79
        case HYP_CONE_EPIPERSQUARE: return 2.0;                            // epipersquare.jl:50
        case HYP_CONE_EPINORMSPECTRAL:
        case HYP_CONE_MATRIXEPIPERSQUARE: return (double)iparam + 1.0;     // epinormspectral.jl:95, matrixepipersquare.jl:101
        case HYP_CONE_EPITRRELENTROPYTRI:                                  // epitrrelentropytri.jl:119
            return 2.0 * (double)(int64_t)((std::sqrt(1.0 + 4.0 * (double)(d - 1)) - 1.0) / 2.0 + 0.5) + 1.0;
        case HYP_CONE_WSOSINTERPNONNEGATIVE: return wsos_nu();             // wsosinterpnonnegative.jl:62
        case HYP_CONE_WSOSINTERPEPINORMEUCL: return 2.0 * wsos_nu();       // wsosinterpepinormeucl.jl:68
        case HYP_CONE_WSOSINTERPPOSSEMIDEFTRI:
        case HYP_CONE_WSOSINTERPEPINORMONE: return (double)iparam * wsos_nu();
        case HYP_CONE_LINMATRIXINEQ:
        case HYP_CONE_POSSEMIDEFTRISPARSE: return (alpha && nalpha >= 1) ? alpha[0] : 0.0;   // side
        case HYP_CONE_GENERALIZEDPOWER: return (double)nalpha + 1.0;       // generalizedpower.jl:38
        default: return (double)d;   // HypoPerLog, EpiNormInf, EpiPerSepSpectral{VectorCSqr}, HypoGeoMean, HypoPowerMean, ...
    }
}

}   // namespace

extern "C" {

hyp_cone* hyp_cone_create(int device, int cone_type, int64_t dim, int use_dual_barrier, int iparam, double dparam,
                          const double* alpha, int64_t nalpha) {
    if (dim < 1 || nalpha < 0) return nullptr;
    hyp_ctx* ctx = hyp_create(device);
    if (!ctx) return nullptr;
    hyp_cone* c = new hyp_cone;
    c->ctx = ctx;
    c->device = device;
    c->type = cone_type;
    c->use_dual = use_dual_barrier ? 1 : 0;
    c->dim = dim;
    const int t1 = cone_type, dual1 = c->use_dual, ip1 = iparam;
    const double dp1 = dparam, zero = 0.0;
    const int64_t d1 = dim, aoff[2] = {0, nalpha};
    std::vector<double> h(dim, 0.0);
    int rc = hyp_set_cone_params(ctx, 1, &ip1, &dp1);
    if (rc >= 0) rc = hyp_set_cone_alpha(ctx, 1, aoff, nalpha ? alpha : &zero);
    // a model with the cone table only: n = p = 0, G is q x 0 (the same stand-alone container DeviceConeBlock uses)
    if (rc >= 0)
        rc = hyp_load_model(ctx, 0, 0, dim, &zero, dim, nullptr, 1, &zero, &zero, h.data(), 1, &t1, &d1, &dual1, 0, 1,
                            nullptr, nullptr);
    if (rc >= 0 && cudaSetDevice(device) != cudaSuccess) rc = -1;
    if (rc >= 0 && cudaMalloc(&c->d_in, (size_t)dim * 16) != cudaSuccess) rc = -1;
    if (rc >= 0 && cudaMemset(c->d_in, 0, (size_t)dim * 16) != cudaSuccess) rc = -1;   // setup_data!: zero point / dual point
    if (rc < 0) {
        // creation failures have no handle to ask: the message goes to stderr, like the reference's constructor asserts
        const char* m = hyp_last_error(ctx);
        fprintf(stderr, "hyp_cone_create: %s\n", m && *m ? m : "device allocation failed");
        if (c->d_in) cudaFree(c->d_in);
        hyp_destroy(ctx);
        delete c;
        return nullptr;
    }
    c->nu = cone_nu(cone_type, dim, iparam, alpha, nalpha);
    return c;
}

void hyp_cone_destroy(hyp_cone* c) {
    if (!c) return;
    if (c->d_in) cudaFree(c->d_in);
    hyp_destroy(c->ctx);
    delete c;
}

const char* hyp_cone_last_error(hyp_cone* c) { return c ? c->err.c_str() : "null cone handle"; }
int64_t hyp_cone_dimension(hyp_cone* c) { return c ? c->dim : -1; }
double hyp_cone_nu(hyp_cone* c) { return c ? c->nu : 0.0; }
int hyp_cone_use_dual_barrier(hyp_cone* c) { return c ? c->use_dual : -1; }

int hyp_cone_load_point(hyp_cone* c, const double* point, double scal) {
    if (!c || !point) return -1;
    if (stage(c, c->d_in, point) < 0) return -1;
    c->scal = scal;
    c->have_point = true;
    c->stale = true;
    return 0;
}

int hyp_cone_load_dual_point(hyp_cone* c, const double* dual_point) {
    if (!c || !dual_point) return -1;
    if (stage(c, c->d_in + c->dim, dual_point) < 0) return -1;
    c->stale = true;
    return 0;
}

int hyp_cone_reset_data(hyp_cone* c) {
    if (!c) return -1;
    c->stale = true;
    return 0;
}

int hyp_cone_is_feas(hyp_cone* c, int* is_feas, int* is_dual_feas) {
    if (!c) return -1;
    if (ensure_state(c) < 0) return -1;
    uint8_t f = 0, d = 0;
    if (inner(c, hyp_cones_feas(c->ctx, &f, &d), "hyp_cones_feas") < 0) return -1;
    if (is_feas) *is_feas = f;
    if (is_dual_feas) *is_dual_feas = d;
    return 0;
}

int hyp_cone_grad(hyp_cone* c, double* grad) {
    if (!c) return -1;
    if (ensure_state(c) < 0) return -1;
    return inner(c, hyp_cones_grad(c->ctx, grad), "hyp_cones_grad");
}

int hyp_cone_hess(hyp_cone* c, double* H, int inverse) {
    if (!c) return -1;
    if (ensure_state(c) < 0) return -1;
    return inner(c, hyp_cones_hess_blocks(c->ctx, H, inverse), "hyp_cones_hess_blocks");
}

int hyp_cone_hess_prod(hyp_cone* c, double* prod, const double* arr, int64_t ncols, int64_t ld_prod, int64_t ld_arr,
                       int mode) {
    if (!c) return -1;
    if (ensure_state(c) < 0) return -1;
    return inner(c, hyp_cones_hess_prod(c->ctx, prod, arr, ncols, ld_prod, ld_arr, mode), "hyp_cones_hess_prod");
}

int hyp_cone_use_sqrt_hess_oracles(hyp_cone* c) {
    if (!c) return -1;
    return c->type == HYP_CONE_NONNEGATIVE || c->type == HYP_CONE_EPINORMEUCL || c->type == HYP_CONE_POSSEMIDEFTRI ||
           c->type == HYP_CONE_EPIPERSQUARE;
}

int hyp_cone_dder3(hyp_cone* c, double* out, const double* dir) {
    if (!c) return -1;
    if (ensure_state(c) < 0) return -1;
    return inner(c, hyp_cones_dder3(c->ctx, out, dir), "hyp_cones_dder3");
}

int hyp_cone_proxsqr(hyp_cone* c, double irtmu, int use_max_prox, double* proxsqr, int* numerics_ok) {
    if (!c) return -1;
    if (ensure_state(c) < 0) return -1;
    double p = 0;
    uint8_t ok = 0;
    if (inner(c, hyp_cones_proxsqr(c->ctx, irtmu, use_max_prox, &p, &ok), "hyp_cones_proxsqr") < 0) return -1;
    if (proxsqr) *proxsqr = p;
    if (numerics_ok) *numerics_ok = ok;
    return 0;
}

}   // extern "C"

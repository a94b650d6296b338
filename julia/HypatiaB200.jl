#=
HypatiaB200.jl - the ccall shim that plugs libhypatia_b200.so (include/hypatia_b200.h) into
Hypatia.jl's two plug-in slots for the per-iteration hot path:

  * Solvers.SystemSolver{Float64}   ->  B200QRCholSystemSolver      (load / update_lhs / solve_system /
                                                                      solve_subsystem3 / free_memory)
  * the per-cone oracle loops of the stepper and the line search  ->  batched hyp_cones_* calls: the five
    reference functions that contain those loops (apply_lhs, update_rhs_cent, update_rhs_centadj,
    update_rhs_predadj, check_cone_points) get Float64 methods here that take the batched path when the
    solver carries a B200QRCholSystemSolver and `invoke` the stock generic method otherwise - no edit of
    the Hypatia.jl sources is needed (INTEGRATION.md lists each override next to the lines it replaces)

  * Cones.Cone{Float64}             ->  B200Cone                    (one device cone behind the per-cone oracle API of
    Cones.jl:34-310, bound to the single-block entry points hyp_cone_*: the drop-in for code that handles ONE cone -
    cone tests, initialize_cone_point, user callbacks, the stock system solvers; end of this file)

Everything else (Solver, steppers, preprocessing, MOI) stays stock Hypatia.jl; usage:

    using Hypatia, HypatiaB200
    solver = Solvers.Solver{Float64}(syssolver = HypatiaB200.B200QRCholSystemSolver())
    Solvers.load(solver, model); Solvers.solve(solver)

NOTE: Julia is not available in the build image, so this file has not been executed there; it is the
binding a maintainer would add, and the same C entry points are exercised call-for-call by the Python
host mirror (hypatia.jl_b200/syssolver.py, cones.py) in the GPU test-suite.  Reference line numbers
refer to Hypatia.jl v0.5.1.
=#
module HypatiaB200

using LinearAlgebra
import Hypatia
import Hypatia.Cones
import Hypatia.Models
import Hypatia.Solvers
import Hypatia.Solvers: Solver, Point, SystemSolver, QRCholSystemSolver

const LIB = get(ENV, "HYPATIA_B200_LIB", "libhypatia_b200")
const Ctx = Ptr{Cvoid}

# cone type codes of include/hypatia_b200.h
cone_code(::Cones.Nonnegative) = Cint(0)
cone_code(::Cones.EpiNormEucl) = Cint(1)
cone_code(::Cones.PosSemidefTri{Float64, Float64}) = Cint(2)
cone_code(::Cones.HypoPerLogdetTri{Float64, Float64}) = Cint(3)
cone_code(::Cones.HypoRootdetTri{Float64, Float64}) = Cint(4)
cone_code(::Cones.EpiPerSepSpectral{Cones.MatrixCSqr{Float64, Float64}, Float64}) = Cint(5)
cone_code(::Cones.EpiPerSquare) = Cint(6)
cone_code(::Cones.HypoPerLog) = Cint(7)
cone_code(::Cones.EpiNormInf{Float64, Float64}) = Cint(8)
cone_code(::Cones.EpiPerSepSpectral{Cones.VectorCSqr{Float64}, Float64}) = Cint(9)
cone_code(::Cones.HypoGeoMean) = Cint(10)
cone_code(::Cones.GeneralizedPower) = Cint(11)
cone_code(::Cones.HypoPowerMean) = Cint(12)
cone_code(::Cones.EpiRelEntropy) = Cint(13)
cone_code(::Cones.EpiNormSpectral{Float64, Float64}) = Cint(14)
cone_code(::Cones.WSOSInterpNonnegative{Float64, Float64}) = Cint(15)
cone_code(::Cones.LinMatrixIneq{Float64}) = Cint(16)
cone_code(::Cones.DoublyNonnegativeTri{Float64}) = Cint(17)
cone_code(::Cones.MatrixEpiPerSquare{Float64, Float64}) = Cint(18)
cone_code(::Cones.WSOSInterpPosSemidefTri{Float64}) = Cint(19)
cone_code(::Cones.WSOSInterpEpiNormEucl{Float64}) = Cint(20)
cone_code(::Cones.EpiTrRelEntropyTri{Float64}) = Cint(23)
cone_code(::Cones.WSOSInterpEpiNormOne{Float64}) = Cint(21)
cone_code(::Cones.PosSemidefTriSparse{<:Cones.PSDSparseImpl, Float64, Float64}) = Cint(22)
cone_alpha(c::Cones.PosSemidefTriSparse{<:Cones.PSDSparseImpl, Float64, Float64}) =
    vcat(Float64(c.side), Float64.(c.row_idxs .- 1), Float64.(c.col_idxs .- 1))      # 0-based pattern
cone_alpha(c::Cones.WSOSInterpEpiNormOne{Float64}) =
    vcat(Float64(length(c.Ps)), Float64[size(P, 2) for P in c.Ps], (vec(P) for P in c.Ps)...)
cone_ssf(c::Cones.WSOSInterpEpiNormOne) = (Cint(c.R), 0.0)
cone_alpha(c::Cones.WSOSInterpEpiNormEucl{Float64}) =
    vcat(Float64(length(c.Ps)), Float64[size(P, 2) for P in c.Ps], (vec(P) for P in c.Ps)...)
cone_ssf(c::Cones.WSOSInterpEpiNormEucl) = (Cint(c.R), 0.0)
cone_alpha(c::Cones.WSOSInterpPosSemidefTri{Float64}) =
    vcat(Float64(length(c.Ps)), Float64[size(P, 2) for P in c.Ps], (vec(P) for P in c.Ps)...)
# packed matrices [side, vec(A_1) .. vec(A_dim)] (dense real symmetric A_i; UniformScaling entries are materialised)
cone_alpha(c::Cones.LinMatrixIneq{Float64}) =
    vcat(Float64(c.side), (vec(Matrix{Float64}(A isa UniformScaling ? A(c.side) : A)) for A in c.As)...)
# packed interpolation data [nP, L_1 .. L_nP, vec(P_1) .. vec(P_nP)]
cone_alpha(c::Cones.WSOSInterpNonnegative{Float64, Float64}) =
    vcat(Float64(length(c.Ps)), Float64[size(P, 2) for P in c.Ps], (vec(P) for P in c.Ps)...)
cone_alpha(c::Cones.GeneralizedPower) = Vector{Float64}(c.α)
cone_alpha(c::Cones.HypoPowerMean) = Vector{Float64}(c.α)
cone_alpha(::Cones.Cone) = Float64[]
cone_code(c::Cones.Cone) = error("cone $(typeof(c)) is not on the B200 hot path")

# HYP_SSF_* code and parameter of the `h` field of EpiPerSepSpectral (sepspectralfun.jl:17-116)
ssf_code(::Cones.InvSSF) = (Cint(0), 0.0)
ssf_code(::Cones.NegLogSSF) = (Cint(1), 0.0)
ssf_code(::Cones.NegEntropySSF) = (Cint(2), 0.0)
ssf_code(h::Cones.Power12SSF) = (Cint(3), Float64(h.p))
cone_ssf(c::Cones.EpiPerSepSpectral) = ssf_code(c.h)
cone_ssf(c::Cones.EpiNormSpectral) = (Cint(c.d1), 0.0)    # integer parameter = number of rows of W
cone_ssf(c::Cones.MatrixEpiPerSquare) = (Cint(c.d1), 0.0)
cone_ssf(c::Cones.WSOSInterpPosSemidefTri) = (Cint(c.R), 0.0)
cone_ssf(::Cones.Cone) = (Cint(0), 0.0)

function check(ctx::Ctx, rc::Cint, what::String)
    rc < 0 && error("$what: " * unsafe_string(ccall((:hyp_last_error, LIB), Cstring, (Ctx,), ctx)))
    return rc
end

# ---------------------------------------------------------------------------------------------
# plug-in slot 1: the system solver (replaces QRCholDenseSystemSolver, qrchol.jl:104-257)
# ---------------------------------------------------------------------------------------------
mutable struct B200QRCholSystemSolver <: QRCholSystemSolver{Float64}
    ctx::Ctx
    device::Int
    fact_kind::Cint
    # fields the generic elimination code expects (common.jl:184-208)
    rhs_sub::Point{Float64}
    sol_sub::Point{Float64}
    rhs_const::Point{Float64}
    sol_const::Point{Float64}
    # batched cone-oracle state (plug-in slot 2): the scaled primal point the device cones are loaded at
    # (= cone.point of every cone, Cones.jl:157-161) and q-vector scratch
    cone_point::Vector{Float64}
    vq1::Vector{Float64}
    vq2::Vector{Float64}
    vq3::Vector{Float64}
    B200QRCholSystemSolver(; device::Int = 0) = (s = new(); s.ctx = C_NULL; s.device = device; s)
end

# check_cone_points(model, stepper) (search.jl:74) is not handed the solver: the context is found by model
const CTX_OF_MODEL = IdDict{Any, B200QRCholSystemSolver}()
b200(solver::Solver{Float64}) = solver.syssolver isa B200QRCholSystemSolver ? solver.syssolver : nothing

# load(syssolver, solver): qrchol.jl:138-179.  G is uploaded once; Ap_Q / Ap_R only when p > 0.
function Solvers.load(syssolver::B200QRCholSystemSolver, solver::Solver{Float64})
    model = solver.model
    (n, p, q) = (model.n, model.p, model.q)
    syssolver.ctx = ccall((:hyp_create, LIB), Ctx, (Cint,), syssolver.device)
    syssolver.ctx == C_NULL && error("hyp_create failed: no sm_100 CUDA device (no CPU fallback)")
    G = Matrix{Float64}(model.G)                       # dense column-major, as load() densifies GQ
    A = Matrix{Float64}(model.A)
    K = length(model.cones)
    ctype = Cint[cone_code(c) for c in model.cones]
    cdim = Int64[Cones.dimension(c) for c in model.cones]
    cdual = Cint[Cones.use_dual_barrier(c) for c in model.cones]
    # Ap_Q is a QR "Q" object (or UniformScaling when p = 0), Ap_R an UpperTriangular (Solvers.jl:104-105):
    # materialise both as dense column-major matrices bound to locals, so that the ccall below roots them
    Qm = iszero(p) ? zeros(Float64, 0, 0) : Matrix{Float64}(solver.Ap_Q * Matrix{Float64}(I, n, n))
    Rm = iszero(p) ? zeros(Float64, 0, 0) : Matrix{Float64}(solver.Ap_R)
    ssf = [cone_ssf(c) for c in model.cones]
    (hkind, hparam) = (Cint[first(t) for t in ssf], Float64[last(t) for t in ssf])
    check(syssolver.ctx, ccall((:hyp_set_cone_params, LIB), Cint, (Ctx, Cint, Ptr{Cint}, Ptr{Float64}),
        syssolver.ctx, K, hkind, hparam), "hyp_set_cone_params")
    alphas = [cone_alpha(c) for c in model.cones]
    aoff = Int64[0; cumsum(length.(alphas))]
    check(syssolver.ctx, ccall((:hyp_set_cone_alpha, LIB), Cint, (Ctx, Cint, Ptr{Int64}, Ptr{Float64}),
        syssolver.ctx, K, aoff, vcat(alphas..., Float64[0])), "hyp_set_cone_alpha")
    GC.@preserve G A ctype cdim cdual Qm Rm begin
        ApQ = iszero(p) ? Ptr{Float64}(C_NULL) : pointer(Qm)
        ApR = iszero(p) ? Ptr{Float64}(C_NULL) : pointer(Rm)
        rc = ccall((:hyp_load_model, LIB), Cint,
            (Ctx, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Cint}, Ptr{Int64}, Ptr{Cint}, Cint, Cint,
             Ptr{Float64}, Ptr{Float64}),
            syssolver.ctx, n, p, q, G, max(q, 1), A, max(p, 1), model.c, model.b, model.h, K,
            ctype, cdim, cdual, 0, K, ApQ, ApR)
        check(syssolver.ctx, rc, "hyp_load_model")
    end
    Solvers.setup_point_sub(syssolver, model)
    syssolver.cone_point = zeros(q)
    (syssolver.vq1, syssolver.vq2, syssolver.vq3) = (zeros(q), zeros(q), zeros(q))
    CTX_OF_MODEL[model] = syssolver
    return syssolver
end

# update_lhs(syssolver, solver): qrchol.jl:181-257.  The cone state the reference reuses lazily
# (Cones.jl:56-93) is (re)loaded on the device from the current iterate scaled by 1/sqrt(mu).
function Solvers.update_lhs(syssolver::B200QRCholSystemSolver, solver::Solver{Float64})
    ctx = syssolver.ctx
    point = solver.point
    irtmu = inv(sqrt(solver.mu))
    (primal, dual) = primal_dual_vectors(solver.model, point)
    check(ctx, ccall((:hyp_cones_load_point, LIB), Cint, (Ctx, Ptr{Float64}, Ptr{Float64}, Float64),
        ctx, primal, dual, irtmu), "hyp_cones_load_point")
    @. syssolver.cone_point = irtmu * primal
    check(ctx, ccall((:hyp_set_mu_tau, LIB), Cint, (Ctx, Float64, Float64), ctx, solver.mu,
        point.tau[]), "hyp_set_mu_tau")
    kind = Ref{Cint}(0)
    solver.time_upfact += @elapsed rc = check(ctx, ccall((:hyp_update_lhs, LIB), Cint,
        (Ctx, Ptr{Cint}), ctx, kind), "hyp_update_lhs")
    syssolver.fact_kind = kind[]
    # same message, same non-throwing behaviour as qrchol.jl:252-254
    rc == 2 && println("positive definite linear system factorization failed")
    return syssolver
end

# solve_system(syssolver, solver, sol, rhs): common.jl:129-151 (4x4 -> 3x3 reductions included)
function Solvers.solve_system(syssolver::B200QRCholSystemSolver, solver::Solver{Float64},
    sol::Point{Float64}, rhs::Point{Float64})
    check(syssolver.ctx, ccall((:hyp_solve_system, LIB), Cint, (Ctx, Ptr{Float64}, Ptr{Float64}),
        syssolver.ctx, sol.vec, rhs.vec), "hyp_solve_system")
    return sol
end

# solve_subsystem3(syssolver, solver, sol, rhs): qrchol.jl:39-85
function Solvers.solve_subsystem3(syssolver::B200QRCholSystemSolver, solver::Solver{Float64},
    sol::Point{Float64}, rhs::Point{Float64})
    dim3 = solver.model.n + solver.model.p + solver.model.q
    GC.@preserve sol rhs check(syssolver.ctx, ccall((:hyp_solve_subsystem3, LIB), Cint,
        (Ctx, Ptr{Float64}, Ptr{Float64}), syssolver.ctx, pointer(sol.vec), pointer(rhs.vec)),
        "hyp_solve_subsystem3")
    return sol
end

# apply_lhs(stepper, solver): common.jl:79-121.  The reference's function is generic in T and not dispatched
# on the system solver (Solver{T} has ONE type parameter, Solvers.jl:62), so the shim adds the more specific
# Float64 method and falls through to the stock method for every other system solver.  It fills
# stepper.temp from stepper.dir exactly like common.jl:84-85.
const GENERIC_APPLY_LHS = Tuple{Solvers.Stepper{T}, Solver{T}} where {T <: Real}
function Solvers.apply_lhs(stepper::Solvers.Stepper{Float64}, solver::Solver{Float64})
    sys = b200(solver)
    isnothing(sys) && return invoke(Solvers.apply_lhs, GENERIC_APPLY_LHS, stepper, solver)
    ctx = sys.ctx
    check(ctx, ccall((:hyp_set_mu_tau, LIB), Cint, (Ctx, Float64, Float64), ctx, solver.mu,
        solver.point.tau[]), "hyp_set_mu_tau")
    check(ctx, ccall((:hyp_apply_lhs, LIB), Cint, (Ctx, Ptr{Float64}, Ptr{Float64}),
        ctx, stepper.temp.vec, stepper.dir.vec), "hyp_apply_lhs")
    return stepper.temp
end

# calc_convergence_params(solver): Solvers.jl:425-483.  Optional: the residual vectors and norms with the
# two passes over G done on the device (hyp_calc_residuals); the remaining scalar bookkeeping of the
# reference's function (x_feas .. improv, primal_obj, dual_obj, gap) is unchanged host code.
function device_residuals!(solver::Solver{Float64})
    sys = b200(solver)
    isnothing(sys) && error("device_residuals! needs a B200QRCholSystemSolver")
    ctx = sys.ctx
    stats = zeros(10)
    check(ctx, ccall((:hyp_calc_residuals, LIB), Cint,
        (Ctx, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ctx, solver.point.vec, solver.x_residual, solver.y_residual, solver.z_residual, stats),
        "hyp_calc_residuals")
    tau = solver.point.tau[]
    (solver.x_norm_res_t, solver.x_norm_res) = (stats[1], stats[2] / tau)
    (solver.y_norm_res_t, solver.y_norm_res) = (stats[3], stats[4] / tau)
    (solver.z_norm_res_t, solver.z_norm_res) = (stats[5], stats[6] / tau)
    solver.primal_obj_t = stats[7]
    solver.dual_obj_t = -stats[8] - stats[9]
    solver.gap = stats[10]
    return solver
end

# free_memory(syssolver): Solvers.jl:582-584
function Solvers.free_memory(syssolver::B200QRCholSystemSolver)
    syssolver.ctx == C_NULL || ccall((:hyp_destroy, LIB), Cvoid, (Ctx,), syssolver.ctx)
    syssolver.ctx = C_NULL
    for (m, s) in collect(CTX_OF_MODEL)
        s === syssolver && delete!(CTX_OF_MODEL, m)
    end
    return
end

# ---------------------------------------------------------------------------------------------
# plug-in slot 2: batched cone oracles.  The reference loops `for k in eachindex(cones)` over
# per-cone objects (steppers/common.jl:26-118, search.jl:112-135); on the device one call covers
# all K cones, so the shim exposes q-vector versions of the oracles and the three callers use them.
# ---------------------------------------------------------------------------------------------
function primal_dual_vectors(model::Models.Model{Float64}, point::Point{Float64})
    primal = copy(point.s)
    dual = copy(point.z)
    for (k, cone) in enumerate(model.cones)
        if Cones.use_dual_barrier(cone)
            idxs = model.cone_idxs[k]
            primal[idxs] .= point.z[idxs]
            dual[idxs] .= point.s[idxs]
        end
    end
    return (primal, dual)
end

cones_load_point(ctx::Ctx, primal::Vector{Float64}, dual::Vector{Float64}, scal::Float64) =
    check(ctx, ccall((:hyp_cones_load_point, LIB), Cint, (Ctx, Ptr{Float64}, Ptr{Float64}, Float64),
        ctx, primal, dual, scal), "hyp_cones_load_point")

function cones_feas(ctx::Ctx, K::Int)
    (f, d) = (zeros(UInt8, K), zeros(UInt8, K))
    check(ctx, ccall((:hyp_cones_feas, LIB), Cint, (Ctx, Ptr{UInt8}, Ptr{UInt8}), ctx, f, d),
        "hyp_cones_feas")
    return (all(!iszero, f), all(!iszero, d))
end

cones_grad!(grad::Vector{Float64}, ctx::Ctx) = (check(ctx, ccall((:hyp_cones_grad, LIB), Cint,
    (Ctx, Ptr{Float64}), ctx, grad), "hyp_cones_grad"); grad)

# mode: 0 hess_prod!, 1 inv_hess_prod!, 2 sqrt_hess_prod!, 3 inv_sqrt_hess_prod!, 4 block_hess_prod!
function cones_hess_prod!(prod::VecOrMat{Float64}, arr::VecOrMat{Float64}, ctx::Ctx, mode::Int)
    q = size(arr, 1)
    check(ctx, ccall((:hyp_cones_hess_prod, LIB), Cint,
        (Ctx, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Cint),
        ctx, prod, arr, size(arr, 2), max(q, 1), max(q, 1), mode), "hyp_cones_hess_prod")
    return prod
end

cones_dder3!(out::Vector{Float64}, dir::Vector{Float64}, ctx::Ctx) = (check(ctx,
    ccall((:hyp_cones_dder3, LIB), Cint, (Ctx, Ptr{Float64}, Ptr{Float64}), ctx, out, dir),
    "hyp_cones_dder3"); out)

function cones_proxsqr(ctx::Ctx, K::Int, irtmu::Float64, use_max_prox::Bool)
    (prox, ok) = (zeros(K), zeros(UInt8, K))
    check(ctx, ccall((:hyp_cones_proxsqr, LIB), Cint, (Ctx, Float64, Cint, Ptr{Float64}, Ptr{UInt8}),
        ctx, irtmu, use_max_prox, prox, ok), "hyp_cones_proxsqr")
    return (prox, all(!iszero, ok))
end


# ---------------------------------------------------------------------------------------------
# The four caller loops of plug-in slot 2, batched.  Each method has the reference's name and a Float64
# signature (more specific than the stock `where {T <: Real}` method), runs the batched device oracles when the
# solver carries the B200 system solver, and otherwise `invoke`s the stock method - Hypatia.jl is not edited.
# ---------------------------------------------------------------------------------------------
const GENERIC_RHS2 = Tuple{Solver{T}, Point{T}} where {T <: Real}
const GENERIC_RHS3 = Tuple{Solver{T}, Point{T}, Point{T}} where {T <: Real}
const GENERIC_CHECK = Tuple{Models.Model{T}, Solvers.Stepper{T}} where {T <: Real}

seg_dot(model::Models.Model{Float64}, a::Vector{Float64}, b::Vector{Float64}) =
    Float64[dot(view(a, idxs), view(b, idxs)) for idxs in model.cone_idxs]

# update_rhs_cent: steppers/common.jl:62-82 (rhs.s_k = -dual_k - sqrt(mu) grad_k)
function Solvers.update_rhs_cent(solver::Solver{Float64}, rhs::Point{Float64})
    sys = b200(solver)
    isnothing(sys) && return invoke(Solvers.update_rhs_cent, GENERIC_RHS2, solver, rhs)
    rhs.x .= 0
    rhs.y .= 0
    rhs.z .= 0
    rhs.tau[] = 0
    rtmu = sqrt(solver.mu)
    grad = cones_grad!(sys.vq1, sys.ctx)
    (_, dual) = primal_dual_vectors(solver.model, solver.point)
    @. rhs.s = -dual - rtmu * grad
    rhs.kap[] = -solver.point.kap[] + solver.mu / solver.point.tau[]
    return rhs
end

# shared body of update_rhs_predadj (steppers/common.jl:26-59) and update_rhs_centadj (:85-118)
function adj_rhs!(sys::B200QRCholSystemSolver, solver::Solver{Float64}, rhs::Point{Float64},
    dir::Point{Float64}, pred::Bool)
    model = solver.model
    rhs.vec .= 0
    rteps = sqrt(eps(Float64))
    irtrtmu = inv(sqrt(sqrt(solver.mu)))
    (prim_dir, _) = primal_dual_vectors(model, dir)
    prim_scal = sys.vq1
    @. prim_scal = irtrtmu * prim_dir
    # hess_prod_slow!(H_prim_dir_k, prim_dir_k) for pred (:39), of the scaled direction for cent (:98)
    H_prim = cones_hess_prod!(sys.vq2, pred ? prim_dir : prim_scal, sys.ctx, 0)
    dder3 = cones_dder3!(sys.vq3, prim_scal, sys.ctx)
    dot1 = seg_dot(model, dder3, sys.cone_point)
    dot2 = seg_dot(model, prim_scal, H_prim)
    pred && (dot2 .*= irtrtmu)
    for (k, cone_k) in enumerate(model.cones)
        Cones.use_dder3(cone_k) || continue
        dder3_viol = abs(dot1[k] - dot2[k]) / (rteps + abs(dot2[k]))
        if dder3_viol < 1e-4
            idxs = model.cone_idxs[k]
            if pred
                @views @. rhs.s[idxs] = H_prim[idxs] + dder3[idxs]
            else
                @views rhs.s[idxs] .= dder3[idxs]
            end
        end
    end
    taubar = solver.point.tau[]
    tau_dir_tau = dir.tau[] / taubar
    rhs.kap[] = tau_dir_tau * solver.mu / taubar * (pred ? 1 + tau_dir_tau : tau_dir_tau)
    return rhs
end

function Solvers.update_rhs_predadj(solver::Solver{Float64}, rhs::Point{Float64}, dir::Point{Float64})
    sys = b200(solver)
    isnothing(sys) && return invoke(Solvers.update_rhs_predadj, GENERIC_RHS3, solver, rhs, dir)
    return adj_rhs!(sys, solver, rhs, dir, true)
end

function Solvers.update_rhs_centadj(solver::Solver{Float64}, rhs::Point{Float64}, dir::Point{Float64})
    sys = b200(solver)
    isnothing(sys) && return invoke(Solvers.update_rhs_centadj, GENERIC_RHS3, solver, rhs, dir)
    return adj_rhs!(sys, solver, rhs, dir, false)
end

# check_cone_points: search.jl:74-138.  The cheap scalar tests are the reference's; the per-cone oracle sweep
# (:112-135) is ONE batched sweep - all cones are evaluated instead of leaving at the first failing cone, the
# Boolean outcome is identical.  Julia's `max` propagates NaN, so a NaN proximity rejects the candidate.
function Solvers.check_cone_points(model::Models.Model{Float64}, stepper::Solvers.Stepper{Float64})
    sys = get(CTX_OF_MODEL, model, nothing)
    isnothing(sys) && return invoke(Solvers.check_cone_points, GENERIC_CHECK, model, stepper)
    searcher = stepper.searcher
    cand = stepper.temp
    szk = searcher.szk
    cones = model.cones
    min_prox = searcher.min_prox
    use_max_prox = searcher.use_max_prox
    proxsqr_bound = abs2(searcher.prox_bound)
    taukap = cand.tau[] * cand.kap[]
    (min(cand.tau[], cand.kap[], taukap) < eps(Float64)) && return false
    for k in eachindex(cones)
        szk[k] = dot(cand.primal_views[k], cand.dual_views[k])
        (szk[k] < eps(Float64)) && return false
    end
    mu = (sum(szk) + taukap) / searcher.nup1
    (mu < eps(Float64)) && return false
    taukap_rel = taukap / mu
    (taukap_rel < min_prox) && return false
    taukap_proxsqr = abs2(taukap_rel - 1)
    (taukap_proxsqr > proxsqr_bound) && return false
    for k in eachindex(cones)
        nu_k = Cones.get_nu(cones[k])
        sz_rel_k = szk[k] / (mu * nu_k)
        if (sz_rel_k < min_prox) || (nu_k * abs2(sz_rel_k - 1) > proxsqr_bound)
            return false
        end
    end
    irtmu = inv(sqrt(mu))
    (primal, dual) = primal_dual_vectors(model, cand)
    cones_load_point(sys.ctx, primal, dual, irtmu)
    @. sys.cone_point = irtmu * primal
    (feas, dual_feas) = cones_feas(sys.ctx, length(cones))
    (feas && dual_feas) || return false
    (proxsqr, numerics_ok) = cones_proxsqr(sys.ctx, length(cones), irtmu, use_max_prox)
    numerics_ok || return false
    agg_proxsqr = taukap_proxsqr
    aggfun = (use_max_prox ? max : +)
    for proxsqr_k in proxsqr
        agg_proxsqr = aggfun(agg_proxsqr, proxsqr_k)
    end
    (agg_proxsqr < proxsqr_bound) || return false
    searcher.prox = sqrt(agg_proxsqr)
    return true
end

# ---------------------------------------------------------------------------------------------
# plug-in slot 2, per-cone form: B200Cone <: Cones.Cone{Float64}.  One device cone behind the reference's per-cone
# oracle API (Cones.jl:34-310), bound to the single-block entry points hyp_cone_* (SURVEY.md 8(b)).  It carries every
# field the generic code of Cones.jl touches and the reference's lazy evaluation (the loads write cone.point /
# cone.dual_point on the host, reset_data clears the flags, the first query uploads and evaluates), so stock code that
# handles ONE cone - tests, initialize_cone_point (Solvers.jl:530-548), user callbacks, the stock system solvers -
# runs unchanged on it:  model.cones .= HypatiaB200.B200Cone.(model.cones).
# With the B200 system solver the batched hyp_cones_* path above serves all K cones per launch and remains the hot path.
# ---------------------------------------------------------------------------------------------
const ConeHandle = Ptr{Cvoid}

mutable struct B200Cone <: Cones.Cone{Float64}
    inner::Cones.Cone{Float64}            # the stock cone this object stands in for (type, parameters, initial point)
    device::Int
    handle::ConeHandle
    use_dual_barrier::Bool
    dim::Int
    nu::Float64
    # fields of the generic Cone code (Cones.jl:34-310)
    point::Vector{Float64}
    dual_point::Vector{Float64}
    grad::Vector{Float64}
    dder3::Vector{Float64}
    vec1::Vector{Float64}
    vec2::Vector{Float64}
    feas_updated::Bool
    grad_updated::Bool
    hess_updated::Bool
    inv_hess_updated::Bool
    hess_fact_updated::Bool
    is_feas::Bool
    use_hess_prod_slow::Bool
    use_hess_prod_slow_updated::Bool
    hess::Symmetric{Float64, Matrix{Float64}}
    inv_hess::Symmetric{Float64, Matrix{Float64}}
    hess_fact_mat::Symmetric{Float64, Matrix{Float64}}
    hess_fact::Factorization{Float64}

    function B200Cone(inner::Cones.Cone{Float64}; device::Int = 0)
        cone = new()
        cone.inner = inner
        cone.device = device
        cone.handle = C_NULL
        cone.use_dual_barrier = Cones.use_dual_barrier(inner)
        cone.dim = Cones.dimension(inner)
        cone.nu = Cones.get_nu(inner)
        cone.use_hess_prod_slow = cone.use_hess_prod_slow_updated = false
        return cone
    end
end

cone_check(cone::B200Cone, rc::Cint, what::String) = (rc < 0 &&
    error("$what: " * unsafe_string(ccall((:hyp_cone_last_error, LIB), Cstring, (ConeHandle,), cone.handle))); rc)

# setup_data!(cone) (Cones.jl:139-152) allocates the generic fields and calls this hook: create the device cone
function Cones.setup_extra_data!(cone::B200Cone)
    free_cone!(cone)
    (iparam, dparam) = cone_ssf(cone.inner)
    alpha = cone_alpha(cone.inner)
    cone.handle = ccall((:hyp_cone_create, LIB), ConeHandle,
        (Cint, Cint, Int64, Cint, Cint, Float64, Ptr{Float64}, Int64),
        cone.device, cone_code(cone.inner), cone.dim, cone.use_dual_barrier, iparam, dparam, alpha, length(alpha))
    cone.handle == C_NULL && error("hyp_cone_create failed for $(typeof(cone.inner)) (no CPU fallback)")
    finalizer(free_cone!, cone)
    return cone
end

function free_cone!(cone::B200Cone)
    cone.handle == C_NULL || ccall((:hyp_cone_destroy, LIB), Cvoid, (ConeHandle,), cone.handle)
    cone.handle = C_NULL
    return
end

Cones.set_initial_point!(arr::AbstractVector, cone::B200Cone) = Cones.set_initial_point!(arr, cone.inner)

# update_feas(cone): the per-cone files read cone.point here for the first time after load_point / reset_data
# (e.g. epinormeucl.jl:54-68); the device cone receives cone.point and cone.dual_point at the same moment
function Cones.update_feas(cone::B200Cone)
    @assert !cone.feas_updated
    cone_check(cone, ccall((:hyp_cone_load_point, LIB), Cint, (ConeHandle, Ptr{Float64}, Float64),
        cone.handle, cone.point, 1.0), "hyp_cone_load_point")
    cone_check(cone, ccall((:hyp_cone_load_dual_point, LIB), Cint, (ConeHandle, Ptr{Float64}),
        cone.handle, cone.dual_point), "hyp_cone_load_dual_point")
    feas = Ref{Cint}(0)
    cone_check(cone, ccall((:hyp_cone_is_feas, LIB), Cint, (ConeHandle, Ptr{Cint}, Ptr{Cint}),
        cone.handle, feas, C_NULL), "hyp_cone_is_feas")
    cone.is_feas = !iszero(feas[])
    cone.feas_updated = true
    return cone.is_feas
end

# is_dual_feas(cone) is not cached by the reference (Cones.jl:69 and the per-cone overrides read cone.dual_point)
function Cones.is_dual_feas(cone::B200Cone)
    cone.feas_updated || Cones.update_feas(cone)
    dual_feas = Ref{Cint}(0)
    cone_check(cone, ccall((:hyp_cone_is_feas, LIB), Cint, (ConeHandle, Ptr{Cint}, Ptr{Cint}),
        cone.handle, C_NULL, dual_feas), "hyp_cone_is_feas")
    return !iszero(dual_feas[])
end

function Cones.update_grad(cone::B200Cone)
    @assert Cones.is_feas(cone)
    cone_check(cone, ccall((:hyp_cone_grad, LIB), Cint, (ConeHandle, Ptr{Float64}), cone.handle, cone.grad),
        "hyp_cone_grad")
    cone.grad_updated = true
    return cone.grad
end

function Cones.update_hess(cone::B200Cone)
    Cones.grad(cone)
    isdefined(cone, :hess) || Cones.alloc_hess!(cone)
    cone_check(cone, ccall((:hyp_cone_hess, LIB), Cint, (ConeHandle, Ptr{Float64}, Cint),
        cone.handle, cone.hess.data, 0), "hyp_cone_hess")
    cone.hess_updated = true
    return cone.hess
end

function Cones.update_inv_hess(cone::B200Cone)
    Cones.grad(cone)
    isdefined(cone, :inv_hess) || Cones.alloc_inv_hess!(cone)
    cone_check(cone, ccall((:hyp_cone_hess, LIB), Cint, (ConeHandle, Ptr{Float64}, Cint),
        cone.handle, cone.inv_hess.data, 1), "hyp_cone_hess")
    cone.inv_hess_updated = true
    return cone.inv_hess
end

# prod / arr may be views of columns of a larger matrix (qrchol.jl:219-246): unit stride down a column is required
# (as for the reference's BLAS calls), the column stride travels as the leading dimension
function cone_prod!(prod::AbstractVecOrMat{Float64}, arr::AbstractVecOrMat{Float64}, cone::B200Cone, mode::Int)
    Cones.grad(cone)
    @assert size(prod) == size(arr) && size(arr, 1) == cone.dim
    @assert stride(arr, 1) == 1 && stride(prod, 1) == 1
    ncols = size(arr, 2)
    ld_arr = (ncols > 1 ? stride(arr, 2) : cone.dim)
    ld_prod = (ncols > 1 ? stride(prod, 2) : cone.dim)
    GC.@preserve prod arr cone_check(cone, ccall((:hyp_cone_hess_prod, LIB), Cint,
        (ConeHandle, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Cint),
        cone.handle, pointer(prod), pointer(arr), ncols, ld_prod, ld_arr, mode), "hyp_cone_hess_prod")
    return prod
end

Cones.hess_prod!(prod::AbstractVecOrMat{Float64}, arr::AbstractVecOrMat{Float64}, cone::B200Cone) =
    cone_prod!(prod, arr, cone, 0)
Cones.inv_hess_prod!(prod::AbstractVecOrMat{Float64}, arr::AbstractVecOrMat{Float64}, cone::B200Cone) =
    cone_prod!(prod, arr, cone, 1)
Cones.sqrt_hess_prod!(prod::AbstractVecOrMat{Float64}, arr::AbstractVecOrMat{Float64}, cone::B200Cone) =
    cone_prod!(prod, arr, cone, 2)
Cones.inv_sqrt_hess_prod!(prod::AbstractVecOrMat{Float64}, arr::AbstractVecOrMat{Float64}, cone::B200Cone) =
    cone_prod!(prod, arr, cone, 3)

# closed-form square-root oracles only (nonnegative.jl:38, epinormeucl.jl:40, possemideftri.jl:54, epipersquare.jl:48);
# the generic Cholesky-of-the-Hessian route of Cones.jl:189-196 is not offered, callers take the hess_prod! branch
Cones.use_sqrt_hess_oracles(arr_dim::Int, cone::B200Cone) =
    !iszero(ccall((:hyp_cone_use_sqrt_hess_oracles, LIB), Cint, (ConeHandle,), cone.handle))

function Cones.dder3(cone::B200Cone, dir::AbstractVector{Float64})
    Cones.grad(cone)
    dirv = Vector{Float64}(dir)
    cone_check(cone, ccall((:hyp_cone_dder3, LIB), Cint, (ConeHandle, Ptr{Float64}, Ptr{Float64}),
        cone.handle, cone.dder3, dirv), "hyp_cone_dder3")
    return cone.dder3
end

function cone_prox(cone::B200Cone, irtmu::Float64, use_max_prox::Bool)
    Cones.grad(cone)
    (proxsqr, ok) = (Ref{Float64}(0.0), Ref{Cint}(0))
    cone_check(cone, ccall((:hyp_cone_proxsqr, LIB), Cint, (ConeHandle, Float64, Cint, Ptr{Float64}, Ptr{Cint}),
        cone.handle, irtmu, use_max_prox, proxsqr, ok), "hyp_cone_proxsqr")
    return (proxsqr[], !iszero(ok[]))
end

Cones.check_numerics(cone::B200Cone) = last(cone_prox(cone, 1.0, true))
Cones.get_proxsqr(cone::B200Cone, irtmu::Float64, use_max_prox::Bool) = first(cone_prox(cone, irtmu, use_max_prox))

end # module

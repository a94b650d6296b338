"""Host-side caller of the hot path: RHS builders, direction refinement, combined stepper,
step-length search.

The north star keeps "src/Solvers/stepper" on the host (Julia in the reference).  Julia is
not available in this image, so this module is the Python stand-in for that caller: it
holds control flow and O(q) vector bookkeeping only, and reaches the hot path exclusively
through two plug-in objects:

  * solver.syssolver - load / update_lhs / solve_system / apply_lhs   (SystemSolver slot)
  * solver.cones     - a host.coneblock.ConeBlock                       (Cone oracle slot)

reference: src/Solvers/steppers/common.jl:7-118 (RHS builders),
           src/Solvers/systemsolvers/common.jl:15-76 (get_directions),
           src/Solvers/steppers/combined.jl:34-171 (CombinedStepper),
           src/Solvers/search.jl:5-138 (StepSearcher, search_alpha, check_cone_points).
"""
from __future__ import annotations

import time

import numpy as np

from .point import Point

EPS = np.finfo(np.float64).eps

# reference: search.jl:41-43
DEFAULT_ALPHA_SCHED = np.array([
    0.9999, 0.999, 0.99, 0.97, 0.95, 0.9, 0.85, 0.8, 0.7, 0.6, 0.5,
    0.3, 0.1, 0.05, 0.01, 0.005, 0.001, 0.0005])


# --------------------------------------------------------------------------------------
# RHS builders (steppers/common.jl)
# --------------------------------------------------------------------------------------
def update_rhs_pred(solver, rhs: Point):
    """reference: steppers/common.jl:7-23"""
    rhs.x[:] = solver.x_residual
    rhs.y[:] = solver.y_residual
    rhs.z[:] = solver.z_residual
    rhs.tau = solver.tau_residual
    _, dual = solver.point.primal_dual(solver.cones.dual_mask)
    rhs.s[:] = -dual
    rhs.kap = -solver.point.kap
    return rhs


def _adj_common(solver, rhs: Point, direction: Point, pred: bool):
    """Shared body of update_rhs_predadj / update_rhs_centadj
    (reference: steppers/common.jl:26-59 and :85-118)."""
    cones = solver.cones
    rhs.vec[:] = 0.0
    rteps = np.sqrt(EPS)
    irtrtmu = 1.0 / np.sqrt(np.sqrt(solver.mu))
    use = cones.use_dder3()
    if use.any():
        prim_dir, _ = direction.primal_dual(cones.dual_mask)
        prim_scal = irtrtmu * prim_dir
        if pred:
            h_prim = cones.hess_prod(prim_dir)           # H * prim_dir
        else:
            h_prim = cones.hess_prod(prim_scal)          # H * (irtrtmu * prim_dir)
        d3 = cones.dder3(prim_scal)
        dot1 = cones.seg_dot(d3, cones.point)
        dot2 = cones.seg_dot(prim_scal, h_prim)
        if pred:
            dot2 = irtrtmu * dot2
        with np.errstate(invalid="ignore", divide="ignore"):
            viol = np.abs(dot1 - dot2) / (rteps + np.abs(dot2))
        ok = use & (viol < 1e-4)  # NaN compares False, as in the reference
        okq = cones.expand(ok)
        if pred:
            rhs.s[okq] = (h_prim + d3)[okq]
        else:
            rhs.s[okq] = d3[okq]
    taubar = solver.point.tau
    tau_dir_tau = direction.tau / taubar
    if pred:
        rhs.kap = tau_dir_tau * solver.mu / taubar * (1 + tau_dir_tau)
    else:
        rhs.kap = tau_dir_tau * solver.mu / taubar * tau_dir_tau
    return rhs


def update_rhs_predadj(solver, rhs, direction):
    return _adj_common(solver, rhs, direction, True)


def update_rhs_cent(solver, rhs: Point):
    """reference: steppers/common.jl:62-82"""
    rhs.x[:] = 0.0
    rhs.y[:] = 0.0
    rhs.z[:] = 0.0
    rhs.tau = 0.0
    rtmu = np.sqrt(solver.mu)
    _, dual = solver.point.primal_dual(solver.cones.dual_mask)
    rhs.s[:] = -dual - rtmu * solver.cones.grad()
    rhs.kap = -solver.point.kap + solver.mu / solver.point.tau
    return rhs


def update_rhs_centadj(solver, rhs, direction):
    return _adj_common(solver, rhs, direction, False)


# --------------------------------------------------------------------------------------
# direction solve + iterative refinement (systemsolvers/common.jl:15-76)
# --------------------------------------------------------------------------------------
def get_directions(stepper, solver, min_impr_tol: float = 0.5):
    rhs, direction, res = stepper.rhs, stepper.dir, stepper.temp
    sys = solver.syssolver
    sys.solve_system(solver, direction, rhs)
    solver.n_solve_system += 1
    if solver.max_ref_steps == 0:
        return direction

    dir_temp = stepper.dir_temp
    dir_temp[:] = direction.vec
    sys.apply_lhs(solver, direction, res)
    solver.n_apply_lhs += 1
    res.vec -= rhs.vec
    res_norm = np.linalg.norm(res.vec, np.inf)

    if res_norm > solver.res_norm_cutoff:
        is_prev_slow = False
        prev_res_norm = res_norm
        for _ in range(solver.max_ref_steps):
            sys.solve_system(solver, direction, res)
            solver.n_solve_system += 1
            direction.vec[:] = dir_temp - direction.vec
            sys.apply_lhs(solver, direction, res)
            solver.n_apply_lhs += 1
            res.vec -= rhs.vec
            res_norm_new = np.linalg.norm(res.vec, np.inf)
            if not (res_norm_new < res_norm):
                direction.vec[:] = dir_temp  # residual has not improved
                break
            dir_temp[:] = direction.vec
            res_norm = res_norm_new
            if res_norm < solver.res_norm_cutoff:
                break
            is_curr_slow = res_norm > min_impr_tol * prev_res_norm
            if is_prev_slow and is_curr_slow:
                break
            prev_res_norm = res_norm
            is_prev_slow = is_curr_slow

    if np.isnan(res_norm):
        raise FloatingPointError("NaN residual in get_directions")
    solver.worst_dir_res = max(solver.worst_dir_res, res_norm)
    return direction


# --------------------------------------------------------------------------------------
# step search (search.jl)
# --------------------------------------------------------------------------------------
class StepSearcher:
    def __init__(self, model, min_prox=0.01, prox_bound=0.99, use_max_prox=True,
                 alpha_sched=DEFAULT_ALPHA_SCHED):
        self.min_prox = min_prox
        self.prox_bound = prox_bound
        self.use_max_prox = use_max_prox
        self.alpha_sched = np.asarray(alpha_sched, dtype=np.float64)
        self.nup1 = model.nu + 1.0
        self.prev_sched = 0
        self.prox = 0.0
        self.n_oracle_sweeps = 0


def check_cone_points(solver, stepper) -> bool:
    """reference: search.jl:74-138.  The reference visits cones one at a time in a
    timing-sorted order and leaves at the first failure; here all cones are evaluated in one
    batched sweep and the flags are reduced - the Boolean outcome is identical."""
    searcher = stepper.searcher
    cones = solver.cones
    cand = stepper.temp
    proxsqr_bound = searcher.prox_bound ** 2
    tau, kap = cand.tau, cand.kap
    taukap = tau * kap
    if min(tau, kap, taukap) < EPS:
        return False
    primal, dual = cand.primal_dual(cones.dual_mask)
    szk = cones.seg_dot(primal, dual)
    if (szk < EPS).any():
        return False
    mu = (szk.sum() + taukap) / searcher.nup1
    if mu < EPS:
        return False
    taukap_rel = taukap / mu
    if taukap_rel < searcher.min_prox:
        return False
    taukap_proxsqr = (taukap_rel - 1.0) ** 2
    if taukap_proxsqr > proxsqr_bound:
        return False
    sz_rel = szk / (mu * cones.nus)
    if ((sz_rel < searcher.min_prox) | (cones.nus * (sz_rel - 1.0) ** 2 > proxsqr_bound)).any():
        return False

    irtmu = 1.0 / np.sqrt(mu)
    searcher.n_oracle_sweeps += 1
    cones.load_point(primal, dual, irtmu)
    if not cones.is_feas().all():
        return False
    if not cones.is_dual_feas().all():
        return False
    if not cones.check_numerics(irtmu, searcher.use_max_prox).all():
        return False
    proxsqr = cones.get_proxsqr(irtmu, searcher.use_max_prox)
    if searcher.use_max_prox:
        # Julia's max propagates NaN (search.jl:112-135: a NaN proximity must reject the candidate);
        # Python's builtin max(x, nan) returns x, so aggregate with np.maximum instead
        agg = float(np.maximum(taukap_proxsqr, proxsqr.max())) if proxsqr.size else taukap_proxsqr
    else:
        agg = taukap_proxsqr + float(proxsqr.sum())
    if not (agg < proxsqr_bound):
        return False
    searcher.prox = np.sqrt(agg)
    return True


def search_alpha(solver, stepper, sched=None) -> float:
    """reference: search.jl:46-69"""
    searcher = stepper.searcher
    if sched is None:
        sched = stepper.start_sched()
    while sched <= len(searcher.alpha_sched):
        alpha = float(searcher.alpha_sched[sched - 1])
        stepper.update_stepper_points(alpha, solver.point, True)
        if check_cone_points(solver, stepper):
            searcher.prev_sched = sched
            return alpha
        sched += 1
    searcher.prev_sched = sched
    return 0.0


# --------------------------------------------------------------------------------------
# combined stepper (steppers/combined.jl)
# --------------------------------------------------------------------------------------
class CombinedStepper:
    def __init__(self, shift_sched: int = 0, **searcher_options):
        self.shift_sched = shift_sched
        self.searcher_options = searcher_options

    def load(self, solver):
        """reference: combined.jl:34-51"""
        model = solver.model
        self.prev_alpha = 1.0
        self.rhs = Point(model)
        self.dir = Point(model)
        self.temp = Point(model)
        self.dir_cent = Point(model)
        self.dir_pred = Point(model)
        self.dir_centadj = Point(model)
        self.dir_predadj = Point(model)
        self.dir_temp = np.zeros_like(self.rhs.vec)
        self.searcher = StepSearcher(model, **self.searcher_options)
        self.unadj_only = self.cent_only = False
        return self

    def start_sched(self):
        if self.shift_sched <= 0:
            return 1
        return max(1, self.searcher.prev_sched - self.shift_sched)

    def compute_directions(self, solver):
        """update_lhs + the four direction solves (reference: combined.jl:64-80).  This is
        exactly the unit BASELINE.json's metric counts: time_upsys + time_getdir."""
        rhs, d = self.rhs, self.dir
        t0 = time.perf_counter()
        solver.syssolver.update_lhs(solver)
        t1 = time.perf_counter()
        solver.time_upsys += t1 - t0

        update_rhs_cent(solver, rhs)
        t2 = time.perf_counter()
        get_directions(self, solver)
        t3 = time.perf_counter()
        self.dir_cent.vec[:] = d.vec
        update_rhs_centadj(solver, rhs, d)
        t4 = time.perf_counter()
        get_directions(self, solver)
        t5 = time.perf_counter()
        self.dir_centadj.vec[:] = d.vec

        update_rhs_pred(solver, rhs)
        t6 = time.perf_counter()
        get_directions(self, solver)
        t7 = time.perf_counter()
        self.dir_pred.vec[:] = d.vec
        update_rhs_predadj(solver, rhs, d)
        t8 = time.perf_counter()
        get_directions(self, solver)
        t9 = time.perf_counter()
        self.dir_predadj.vec[:] = d.vec
        solver.time_uprhs += (t2 - t1) + (t4 - t3) + (t6 - t5) + (t8 - t7)
        solver.time_getdir += (t3 - t2) + (t5 - t4) + (t7 - t6) + (t9 - t8)

    def step(self, solver) -> bool:
        """reference: combined.jl:53-120"""
        self.compute_directions(solver)
        t0 = time.perf_counter()
        self.unadj_only = self.cent_only = False
        alpha = search_alpha(solver, self)
        if alpha == 0.0:
            self.unadj_only = True                       # combined without adjustment
            alpha = search_alpha(solver, self)
            if alpha == 0.0:
                self.cent_only, self.unadj_only = True, False   # centering with adjustment
                alpha = search_alpha(solver, self)
                if alpha == 0.0:
                    self.unadj_only = True               # centering without adjustment
                    alpha = search_alpha(solver, self)
                    if alpha == 0.0:
                        solver.status = "NumericalFailure"
                        self.prev_alpha = alpha
                        solver.time_search += time.perf_counter() - t0
                        return False
        solver.time_search += time.perf_counter() - t0
        self.update_stepper_points(alpha, solver.point, False)
        self.prev_alpha = alpha
        return True

    def update_stepper_points(self, alpha, point, ztsk_only: bool):
        """reference: combined.jl:124-171"""
        if ztsk_only:
            cand = self.temp.ztsk
            cand[:] = point.ztsk
            sel = lambda P: P.ztsk
        else:
            cand = point.vec
            sel = lambda P: P.vec
        dc, dp = sel(self.dir_cent), sel(self.dir_pred)
        if self.unadj_only:
            if self.cent_only:
                cand += alpha * dc
            else:
                cand += alpha * dp + (1 - alpha) * dc
        else:
            dca = sel(self.dir_centadj)
            if self.cent_only:
                cand += alpha * dc + alpha ** 2 * dca
            else:
                dpa = sel(self.dir_predadj)
                am1 = 1 - alpha
                cand += alpha * dp + alpha ** 2 * dpa + am1 * dc + am1 ** 2 * dca

    def step_label(self):
        if self.cent_only:
            return "cent" if self.unadj_only else "ce-a"
        return "comb" if self.unadj_only else "co-a"

    def expect_improvement(self):
        return True                     # combined.jl:122


# --------------------------------------------------------------------------------------
# predict-or-center stepper (steppers/predorcent.jl)
# --------------------------------------------------------------------------------------
class PredOrCentStepper:
    """reference: predorcent.jl:5-203.  One direction pair per iteration - a prediction when the last accepted point
    was close to the central path (or after max_cent_steps centerings), a centering otherwise - so an iteration costs
    update_lhs + 1 or 2 solve_system calls instead of the combined stepper's 4."""

    def __init__(self, use_adjustment: bool = True, use_curve_search=None, max_cent_steps: int = 4,
                 pred_prox_bound: float = 0.0332, **searcher_options):
        if use_curve_search is None:
            use_curve_search = use_adjustment
        assert use_adjustment or not use_curve_search     # curve search needs the adjustment direction
        self.use_adjustment = use_adjustment
        self.use_curve_search = use_curve_search
        self.max_cent_steps = max_cent_steps
        self.pred_prox_bound = pred_prox_bound
        self.searcher_options = searcher_options

    def load(self, solver):
        """reference: predorcent.jl:46-70"""
        model = solver.model
        if self.use_adjustment and not solver.cones.use_dder3().any():
            self.use_adjustment = self.use_curve_search = False
        self.prev_alpha = 1.0
        self.cent_count = 0
        self.rhs = Point(model)
        self.dir = Point(model)
        self.temp = Point(model)
        self.dir_noadj = Point(model)
        self.dir_adj = Point(model)
        self.dir_temp = np.zeros_like(self.rhs.vec)
        self.searcher = StepSearcher(model, **self.searcher_options)
        self.unadj_only = False
        self.unadj_alpha = 0.0
        self.is_pred = True
        return self

    def start_sched(self):
        return 1                          # search.jl:72

    def step(self, solver) -> bool:
        """reference: predorcent.jl:72-163"""
        point, rhs, d = solver.point, self.rhs, self.dir
        t0 = time.perf_counter()
        solver.syssolver.update_lhs(solver)
        solver.time_upsys += time.perf_counter() - t0

        is_pred = self.cent_count >= self.max_cent_steps or self.searcher.prox < self.pred_prox_bound
        self.cent_count = 0 if is_pred else self.cent_count + 1
        self.is_pred = is_pred

        t0 = time.perf_counter()
        (update_rhs_pred if is_pred else update_rhs_cent)(solver, rhs)
        t1 = time.perf_counter()
        get_directions(self, solver)
        t2 = time.perf_counter()
        solver.time_uprhs += t1 - t0
        solver.time_getdir += t2 - t1
        self.dir_noadj.vec[:] = d.vec
        try_noadj = True
        alpha = 0.0

        if self.use_adjustment:
            t0 = time.perf_counter()
            (update_rhs_predadj if is_pred else update_rhs_centadj)(solver, rhs, d)
            t1 = time.perf_counter()
            get_directions(self, solver)
            t2 = time.perf_counter()
            solver.time_uprhs += t1 - t0
            solver.time_getdir += t2 - t1
            self.dir_adj.vec[:] = d.vec

            t0 = time.perf_counter()
            if self.use_curve_search:
                self.unadj_only = False                  # one curve search with the adjustment
                alpha = search_alpha(solver, self)
                solver.time_search += time.perf_counter() - t0
                if alpha != 0.0:
                    self.update_stepper_points(alpha, point, False)
                    self.prev_alpha = alpha
                    return True
            else:
                try_noadj = False                        # two line searches: unadjusted alpha, then the corrected one
                self.unadj_only = True
                alpha = search_alpha(solver, self)
                self.unadj_alpha = alpha
                unadj_sched = self.searcher.prev_sched
                if alpha != 0.0:
                    self.unadj_only = False
                    alpha = search_alpha(solver, self)
                    if alpha == 0.0:
                        self.unadj_only = True           # fall back to the unadjusted direction at the alpha found
                        alpha = search_alpha(solver, self, sched=unadj_sched)
                        assert self.searcher.prev_sched == unadj_sched
                    solver.time_search += time.perf_counter() - t0
                    self.update_stepper_points(alpha, point, False)
                    self.prev_alpha = alpha
                    return True
                solver.time_search += time.perf_counter() - t0

        if try_noadj:
            t0 = time.perf_counter()
            self.unadj_only = True
            alpha = search_alpha(solver, self)
            solver.time_search += time.perf_counter() - t0

        if alpha == 0.0:
            solver.status = "NumericalFailure"
            self.prev_alpha = alpha
            return False
        self.update_stepper_points(alpha, point, False)
        self.prev_alpha = alpha
        return True

    def expect_improvement(self):
        return self.cent_count == 0       # predorcent.jl:165

    def update_stepper_points(self, alpha, point, ztsk_only: bool):
        """reference: predorcent.jl:167-191"""
        if ztsk_only:
            cand = self.temp.ztsk
            cand[:] = point.ztsk
            sel = lambda P: P.ztsk
        else:
            cand = point.vec
            sel = lambda P: P.vec
        cand += alpha * sel(self.dir_noadj)
        if not self.unadj_only:
            adj_factor = alpha ** 2 if self.use_curve_search else alpha * self.unadj_alpha
            cand += adj_factor * sel(self.dir_adj)

    def step_label(self):
        return "pred" if self.cent_count == 0 else "cent"

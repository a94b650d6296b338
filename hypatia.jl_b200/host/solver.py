"""Interior-point driver: preprocessing, the solve loop, convergence tests, postprocessing.

Host-side stand-in for the reference's Julia driver (out of the hot path by the north
star: it *calls* the hot path once per iteration through solver.syssolver and
solver.cones).  Dense A/G only.

reference: src/Solvers/Solvers.jl:162-240 (options), :245-416 (solve), :418-528
           (calc_mu, calc_convergence_params, check_convergence), :530-548
           (initialize_cone_point); src/Solvers/process.jl:13-60 (rescale_data),
           :64-178 (find_initial_x), :182-365 (find_initial_y), :385-458 (postprocess).
"""
from __future__ import annotations

import time

import numpy as np
import scipy.linalg as sla

from .models import Model, CONE_NONNEGATIVE
from .point import Point
from .stepper import CombinedStepper

EPS = np.finfo(np.float64).eps


class Solver:
    def __init__(self, model: Model, syssolver, cone_factory, *, verbose=False, iter_limit=1000,
                 time_limit=np.inf, tol_rel_opt=None, tol_abs_opt=None, tol_feas=None,
                 tol_infeas=None, tol_illposed=None, default_tol_power=0.5,
                 default_tol_relax=None, tol_slow=1e-3, preprocess=True, reduce=True,
                 rescale=True, init_tol_qr=1000 * EPS, stepper=None, max_ref_steps=5):
        """cone_factory(model) -> ConeBlock for the (preprocessed) model; syssolver is the
        SystemSolver plug-in instance (reference: Solvers.jl:180, `syssolver` kwarg)."""
        if reduce:
            assert preprocess
        loose = EPS ** default_tol_power
        tight = EPS ** (1.5 * default_tol_power)
        if default_tol_relax is not None:
            loose *= default_tol_relax
            tight *= default_tol_relax
        self.tol_rel_opt = loose if tol_rel_opt is None else tol_rel_opt
        self.tol_abs_opt = tight if tol_abs_opt is None else tol_abs_opt
        self.tol_feas = loose if tol_feas is None else tol_feas
        self.tol_infeas = tight if tol_infeas is None else tol_infeas
        self.tol_illposed = tight / 100 if tol_illposed is None else tol_illposed
        self.tol_slow = tol_slow
        self.verbose = verbose
        self.iter_limit = iter_limit
        self.time_limit = time_limit
        self.preprocess = preprocess
        self.reduce = reduce
        self.rescale = rescale
        self.init_tol_qr = init_tol_qr
        self.stepper = stepper if stepper is not None else CombinedStepper()
        self.syssolver = syssolver
        self.cone_factory = cone_factory
        self.orig_model = model
        self.max_ref_steps = max_ref_steps
        self.status = "Loaded"
        self.record_iterates = None  # optional list: (point.vec copy, mu) per iteration

    # ------------------------------------------------------------------ preprocessing
    def _rescale_data(self):
        """reference: process.jl:13-60"""
        if not self.rescale:
            return False
        m = self.model
        minval = np.sqrt(EPS)

        def colmax(M):
            return np.maximum(np.abs(M).max(axis=0), minval) if M.shape[0] else np.full(M.shape[1], minval)

        def rowmax(M):
            return np.maximum(np.abs(M).max(axis=1), minval) if M.shape[1] else np.full(M.shape[0], minval)

        c_scale = np.sqrt(np.maximum(np.abs(m.c), np.maximum(colmax(m.A), colmax(m.G))))
        b_scale = np.sqrt(np.maximum(np.abs(m.b), rowmax(m.A)))
        h_scale = np.ones(m.q)
        g_rowmax = rowmax(m.G)
        for ck, sl in zip(m.cones, m.cone_idxs):
            if ck.ctype == CONE_NONNEGATIVE:
                h_scale[sl] = np.sqrt(np.maximum(np.abs(m.h[sl]), g_rowmax[sl]))
            else:
                hk = max(np.abs(m.h[sl]).max(), minval)
                h_scale[sl] = np.sqrt(max(hk, g_rowmax[sl].max()))
        self.c_scale, self.b_scale, self.h_scale = c_scale, b_scale, h_scale
        m.c = m.c / c_scale
        m.A = np.asfortranarray(m.A / c_scale[None, :] / b_scale[:, None])
        m.G = np.asfortranarray(m.G / c_scale[None, :] / h_scale[:, None])
        m.b = m.b / b_scale
        m.h = m.h / h_scale
        return True

    @staticmethod
    def _rank_est(Rfac, tol):
        d = np.abs(np.diag(Rfac))
        return int((d > tol).sum())

    def _find_initial_x(self, init_s):
        """reference: process.jl:64-178 (dense direct branch)"""
        if self.status != "SolveCalled":
            return np.zeros(0)
        m = self.model
        n, p, q = m.n, m.p, m.q
        if n == 0:
            self.x_keep_idxs = np.zeros(0, dtype=np.int64)
            return np.zeros(0)
        self.x_keep_idxs = np.arange(n)
        rhs = np.concatenate((m.b, m.h - init_s))
        AG = np.vstack((m.A, m.G)) if p else m.G.copy()
        Qf, Rf, piv = sla.qr(AG, mode="economic", pivoting=True)
        rank = self._rank_est(Rf, self.init_tol_qr)
        if (not self.preprocess) or rank == n:
            # least squares via the pivoted QR (rank-truncated like Julia's `\`)
            y = Qf[:, :rank].T @ rhs
            xs = sla.solve_triangular(Rf[:rank, :rank], y)
            x = np.zeros(n)
            x[piv[:rank]] = xs
            return x
        keep = piv[:rank]
        R1 = Rf[:rank, :rank]
        c_sub = m.c[keep]
        yz1 = sla.solve_triangular(R1, c_sub, trans="T")
        yz_sub = Qf[:, :rank] @ yz1
        residual = np.linalg.norm(m.A.T @ yz_sub[:p] + m.G.T @ yz_sub[p:] - m.c, np.inf)
        if residual > self.init_tol_qr:
            self.status = "DualInconsistent"
            return np.zeros(0)
        m.c = c_sub
        m.A = np.asfortranarray(m.A[:, keep])
        m.G = np.asfortranarray(m.G[:, keep])
        m.n = rank
        self.x_keep_idxs = keep
        temp = Qf[:, :rank].T @ np.concatenate((m.b, m.h - init_s))
        return sla.solve_triangular(R1, temp)

    def _find_initial_y(self, init_z, reduce):
        """reference: process.jl:182-365 (dense direct branch)"""
        if self.status != "SolveCalled":
            return np.zeros(0)
        m = self.model
        p = m.p
        self.Ap_R = np.zeros((0, 0))
        self.Ap_Q = None  # None stands for the identity (UniformScaling I)
        if p == 0:
            self.y_keep_idxs = np.zeros(0, dtype=np.int64)
            return np.zeros(0)
        n = m.n
        self.y_keep_idxs = np.arange(p)
        Qf, Rf, piv = sla.qr(m.A.T, mode="full", pivoting=True)
        rank = self._rank_est(Rf, self.init_tol_qr)
        if (not reduce) and (not self.preprocess):
            rhs = -m.c - m.G.T @ init_z
            yy = sla.solve_triangular(Rf[:rank, :rank], Qf[:, :rank].T @ rhs)
            y = np.zeros(p)
            y[piv[:rank]] = yy
            return y
        Ap_R = np.triu(Rf[:rank, :rank])
        keep = piv[:rank]
        b_sub = m.b[keep]
        if rank < p:
            x1 = sla.solve_triangular(Ap_R, b_sub, trans="T")
            x_sub = Qf[:, :rank] @ x1
            residual = np.linalg.norm(m.A @ x_sub - m.b, np.inf)
            if residual > self.init_tol_qr:
                self.status = "PrimalInconsistent"
                return np.zeros(0)
        if reduce:
            # eliminate the equalities: n <- n - p, p <- 0, G <- G*Q2 (process.jl:274-338)
            cQ = m.c @ Qf
            self.reduce_cQ1 = cQ[:rank]
            m.c = cQ[rank:].copy()
            m.n = m.c.size
            Rpib0 = sla.solve_triangular(Ap_R, b_sub, trans="T")
            self.reduce_Rpib0 = Rpib0
            m.obj_offset += float(self.reduce_cQ1 @ Rpib0)
            GQ = m.G @ Qf
            self.reduce_GQ1 = GQ[:, :rank]
            m.h = m.h - self.reduce_GQ1 @ Rpib0
            m.G = np.asfortranarray(GQ[:, rank:])
            m.p = 0
            m.A = np.asfortranarray(np.zeros((0, m.n)))
            m.b = np.zeros(0)
            self.reduce_Ap_R = Ap_R
            self.reduce_Ap_Q = Qf
            self.reduce_y_keep_idxs = keep
            return np.zeros(0)
        temp = Qf.T @ (m.c + m.G.T @ init_z)
        init_y = sla.solve_triangular(Ap_R, -temp[:rank])
        m.A = np.asfortranarray(m.A[keep, :])
        m.b = b_sub
        m.p = rank
        self.y_keep_idxs = keep
        self.Ap_R = Ap_R
        self.Ap_Q = Qf
        return init_y

    def _postprocess(self):
        """reference: process.jl:385-458"""
        point, result, om = self.point, self.result, self.orig_model
        infeas = self.status in ("PrimalInfeasible", "DualInfeasible")
        tau = 1.0 if infeas else point.tau
        if tau <= 0:
            result.vec[:] = np.nan
            return
        result.s[:] = point.s / tau
        result.z[:] = point.z / tau
        if self.preprocess and om.n and not np.isnan(point.x).any():
            if self.reduce and om.p:
                k = self.reduce_Rpib0.size
                xa = np.zeros(om.n - k)
                xa[self.x_keep_idxs] = point.x / tau
                r0 = np.zeros(k) if infeas else self.reduce_Rpib0
                result.x[:] = self.reduce_Ap_Q @ np.concatenate((r0, xa))
            else:
                result.x[self.x_keep_idxs] = point.x / tau
        else:
            result.x[:] = point.x / tau
        if self.preprocess and om.p and not np.isnan(point.y).any():
            if self.reduce:
                ya = self.reduce_GQ1.T @ result.z
                if not infeas:
                    ya = ya + self.reduce_cQ1
                ya = sla.solve_triangular(self.reduce_Ap_R, ya)
                result.y[self.reduce_y_keep_idxs] = -ya
            else:
                result.y[self.y_keep_idxs] = point.y / tau
        else:
            result.y[:] = point.y / tau
        if self.used_rescaling:
            result.s *= self.h_scale
            result.z /= self.h_scale
            result.y /= self.b_scale
            result.x /= self.c_scale

    # ------------------------------------------------------------------ iteration helpers
    def calc_mu(self):
        """reference: Solvers.jl:418-423"""
        pt = self.point
        self.mu = (float(pt.z @ pt.s) + pt.tau * pt.kap) / (self.model.nu + 1)
        return self.mu

    def calc_convergence_params(self):
        """reference: Solvers.jl:425-483"""
        m, pt = self.model, self.point
        tau = pt.tau
        # the reference computes these residuals in the (host) solver; a device system solver can take the
        # two passes over G instead (hyp_calc_residuals) when asked to: syssolver.device_residuals = True
        dev = getattr(self.syssolver, "calc_residuals", None) \
            if getattr(self.syssolver, "device_residuals", False) else None
        if dev is not None and getattr(self.syssolver, "ctx", None) is not None:
            # device plug-in: the two passes over G (and A) stay on the GPU (hyp_calc_residuals)
            xres, yres, zres, st = dev(self)
            self.x_norm_res_t, self.x_norm_res = st[0], st[1] / tau
            self.y_norm_res_t, self.y_norm_res = st[2], st[3] / tau
            self.z_norm_res_t, self.z_norm_res = st[4], st[5] / tau
            self.x_residual, self.y_residual, self.z_residual = xres, yres, zres
            x_feas = self.x_norm_res * self.x_conv_tol
            y_feas = self.y_norm_res * self.y_conv_tol
            z_feas = self.z_norm_res * self.z_conv_tol
            return self._finish_convergence_params(x_feas, y_feas, z_feas, float(st[6]),
                                                   -float(st[7]) - float(st[8]), float(st[9]))
        xr = m.G.T @ pt.z
        if m.p:
            xr += m.A.T @ pt.y
        self.x_norm_res_t = np.linalg.norm(xr, np.inf) if xr.size else 0.0
        xr += m.c * tau
        self.x_norm_res = (np.linalg.norm(xr, np.inf) if xr.size else 0.0) / tau
        self.x_residual = -xr
        x_feas = self.x_norm_res * self.x_conv_tol

        yr = m.A @ pt.x if m.p else np.zeros(0)
        self.y_norm_res_t = np.linalg.norm(yr, np.inf) if yr.size else 0.0
        yr = yr - m.b * tau
        self.y_norm_res = (np.linalg.norm(yr, np.inf) if yr.size else 0.0) / tau
        self.y_residual = yr
        y_feas = self.y_norm_res * self.y_conv_tol

        zr = m.G @ pt.x + pt.s
        self.z_norm_res_t = np.linalg.norm(zr, np.inf) if zr.size else 0.0
        zr = zr - m.h * tau
        self.z_norm_res = (np.linalg.norm(zr, np.inf) if zr.size else 0.0) / tau
        self.z_residual = zr
        z_feas = self.z_norm_res * self.z_conv_tol

        return self._finish_convergence_params(x_feas, y_feas, z_feas, float(m.c @ pt.x),
                                               -float(m.b @ pt.y) - float(m.h @ pt.z), float(pt.z @ pt.s))

    def _finish_convergence_params(self, x_feas, y_feas, z_feas, primal_obj_t, dual_obj_t, gap):
        m, pt = self.model, self.point
        tau = pt.tau
        self.primal_obj_t = primal_obj_t
        self.dual_obj_t = dual_obj_t
        self.tau_residual = self.primal_obj_t - self.dual_obj_t + pt.kap
        tau_feas = abs(self.tau_residual)

        improv = 0.0
        for curr, prev in ((x_feas, self.x_feas), (y_feas, self.y_feas),
                           (z_feas, self.z_feas), (tau_feas, self.tau_feas)):
            if np.isnan(prev) or np.isnan(curr):
                continue
            improv = max(improv, (prev - curr) / (abs(prev) + EPS))
        self.x_feas, self.y_feas, self.z_feas, self.tau_feas = x_feas, y_feas, z_feas, tau_feas
        self.primal_obj = self.primal_obj_t / tau + m.obj_offset
        self.dual_obj = self.dual_obj_t / tau + m.obj_offset
        self.gap = gap
        return improv

    def check_convergence(self):
        """reference: Solvers.jl:485-528"""
        tau = self.point.tau
        po, do = self.primal_obj_t, self.dual_obj_t
        is_feas = max(self.x_feas, self.y_feas, self.z_feas) <= self.tol_feas
        is_abs_opt = self.gap <= self.tol_abs_opt
        is_rel_opt = min(self.gap / tau, abs(po - do)) <= \
            self.tol_rel_opt * max(tau, min(abs(po), abs(do)))
        if is_feas and (is_abs_opt or is_rel_opt):
            self.status = "Optimal"
            return True
        if do > EPS and self.x_norm_res_t <= self.tol_infeas * do:
            self.status = "PrimalInfeasible"
            self.primal_obj, self.dual_obj = po, do
            return True
        if po < -EPS and max(self.y_norm_res_t, self.z_norm_res_t) <= self.tol_infeas * -po:
            self.status = "DualInfeasible"
            self.primal_obj, self.dual_obj = po, do
            return True
        if self.mu <= self.tol_illposed and tau <= self.tol_illposed * min(1.0, self.point.kap):
            self.status = "IllPosed"
            return True
        return False

    # ------------------------------------------------------------------ main loop
    def solve(self):
        """reference: Solvers.jl:245-416"""
        self.status = "SolveCalled"
        start = time.perf_counter()
        self.num_iters = 0
        self.time_upsys = self.time_uprhs = self.time_getdir = self.time_search = 0.0
        self.time_loadsys = 0.0
        self.n_solve_system = self.n_apply_lhs = 0
        self.res_norm_cutoff = 0.0
        self.worst_dir_res = 0.0
        self.x_feas = self.y_feas = self.z_feas = self.tau_feas = np.nan
        self.x_norm_res = self.y_norm_res = self.z_norm_res = np.nan
        self.primal_obj = self.dual_obj = self.gap = np.nan

        om = self.orig_model
        self.result = Point(om)
        model = self.model = om.copy()

        # initialize_cone_point (Solvers.jl:530-548): primal = central point, dual = -grad
        cones0 = self.cone_factory(model)
        prim0 = cones0.initial_point()
        cones0.load_point(prim0, prim0, 1.0)
        assert cones0.is_feas().all()
        dual0 = -cones0.grad()
        cones0.load_point(prim0, dual0, 1.0)
        assert cones0.is_dual_feas().all()
        if cones0.dual_mask is not None:
            init_s = np.where(cones0.dual_mask, dual0, prim0)
            init_z = np.where(cones0.dual_mask, prim0, dual0)
        else:
            init_s, init_z = prim0, dual0
        if hasattr(cones0, "free"):
            cones0.free()

        self.used_rescaling = self._rescale_data()
        if self.reduce:
            init_y = self._find_initial_y(init_z, True)
            init_x = self._find_initial_x(init_s)
        else:
            init_x = self._find_initial_x(init_s)
            init_y = self._find_initial_y(init_z, False)

        if self.status == "SolveCalled":
            model._index_cones()
            point = self.point = Point(model)
            point.x[:] = init_x
            point.y[:] = init_y
            point.z[:] = init_z
            point.s[:] = init_s
            point.tau = 1.0
            point.kap = 1.0
            self.calc_mu()
            self.syssolver.load(self)           # uploads G once; owns cone state on device
            self.cones = self.syssolver.cones if getattr(self.syssolver, "cones", None) is not None \
                else self.cone_factory(model)
            primal, dual = point.primal_dual(self.cones.dual_mask)
            self.cones.load_point(primal, dual, 1.0)

            self.x_conv_tol = 1.0 / (1 + (np.abs(model.c).max() if model.n else 0.0))
            self.y_conv_tol = 1.0 / (1 + (np.abs(model.b).max() if model.p else 0.0))
            self.z_conv_tol = 1.0 / (1 + (np.abs(model.h).max() if model.q else 0.0))
            prev_is_slow = prev2_is_slow = False

            stepper = self.stepper
            stepper.load(self)
            if self.verbose:
                print(f"{'iter':>5} {'p_obj':>12} {'d_obj':>12} {'gap':>9} {'x_feas':>9} {'y_feas':>9} "
                      f"{'z_feas':>9} {'tau':>9} {'kap':>9} {'mu':>9} {'dir_res':>9} {'prox':>9} "
                      f"{'step':>5} {'alpha':>9}")
            while True:
                improv = self.calc_convergence_params()
                if self.verbose:
                    line = (f"{self.num_iters:5d} {self.primal_obj:12.4e} {self.dual_obj:12.4e} "
                            f"{self.gap:9.2e} {self.x_feas:9.2e} {self.y_feas:9.2e} {self.z_feas:9.2e} "
                            f"{point.tau:9.2e} {point.kap:9.2e} {self.mu:9.2e}")
                    if self.num_iters:
                        line += (f" {self.worst_dir_res:9.2e} {stepper.searcher.prox:9.2e} "
                                 f"{stepper.step_label():>5} {stepper.prev_alpha:9.2e}")
                    print(line, flush=True)
                if self.record_iterates is not None:
                    self.record_iterates.append((point.vec.copy(), self.mu))
                if self.check_convergence():
                    break
                if self.num_iters == self.iter_limit:
                    self.status = "IterationLimit"
                    break
                if time.perf_counter() - start >= self.time_limit:
                    self.status = "TimeLimit"
                    break
                if stepper.expect_improvement():              # Solvers.jl:362-377
                    if improv < self.tol_slow:
                        if prev_is_slow and prev2_is_slow:
                            self.status = "SlowProgress"
                            break
                        prev2_is_slow, prev_is_slow = prev_is_slow, True
                    else:
                        prev2_is_slow, prev_is_slow = prev_is_slow, False

                self.res_norm_cutoff = 1e-4 * max(self.x_norm_res, self.y_norm_res,
                                                  self.z_norm_res, self.tau_feas)
                self.worst_dir_res = 0.0
                if not stepper.step(self):
                    break
                self.calc_mu()
                if min(point.tau, point.kap, self.mu) <= 0:
                    self.status = "NumericalFailure"
                    break
                self.num_iters += 1
            self._postprocess()
        self.solve_time = time.perf_counter() - start
        if hasattr(self.syssolver, "free_memory"):
            self.syssolver.free_memory()
        if self.verbose:
            print(f"status is {self.status} after {self.num_iters} iterations and "
                  f"{self.solve_time:.3f} seconds")
        return self

    # getters (reference: Solvers.jl:550-564)
    def get_x(self):
        return self.result.x.copy()

    def get_y(self):
        return self.result.y.copy()

    def get_s(self):
        return self.result.s.copy()

    def get_z(self):
        return self.result.z.copy()

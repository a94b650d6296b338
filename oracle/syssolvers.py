"""CPU restatement of the reference's dense system solvers (oracle; test infrastructure).

reference: src/Solvers/systemsolvers/common.jl:79-121 (apply_lhs), :129-151 (solve_system),
:154-182 (solve_subsystem4), :184-211 (setup_point_sub, dot_obj);
qrchol.jl:16-37 (setup_rhs3), :39-85 (solve_subsystem3), :138-179 (load), :181-257
(update_lhs, update_lhs_fact); symindef.jl:31-52 (setup_rhs3), :203-271 (dense solver);
naive.jl:144-216 (NaiveDense, the reference's own cross-check solver).

All solvers use `solver.cones` (an oracle.cones.OracleConeBlock), `solver.model`,
`solver.point`, `solver.mu`, `solver.Ap_Q` / `solver.Ap_R` like the reference.
"""
import numpy as np
import scipy.linalg as sla
from scipy.linalg import blas as _blas

from .layout import OracleSubPoint as SubPoint
from . import linalg as la


def dot_obj(model, pt):
    return float(model.c @ pt.x) + float(model.b @ pt.y) + float(model.h @ pt.z)


def apply_lhs(solver, direction, res):
    """6x6 operator applied to `direction` (common.jl:79-121)."""
    m = solver.model
    cones = solver.cones
    tau_dir, kap_dir = direction.tau, direction.kap
    res.x[:] = m.G.T @ direction.z + m.c * tau_dir
    res.z[:] = m.h * tau_dir - direction.s - m.G @ direction.x
    rt = -float(m.c @ direction.x) - float(m.h @ direction.z) - kap_dir
    if m.p:
        res.x += m.A.T @ direction.y
        res.y[:] = m.b * tau_dir - m.A @ direction.x
        rt -= float(m.b @ direction.y)
    res.tau = rt
    primal, dual = direction.primal_dual(cones.dual_mask)
    res.s[:] = cones.hess_prod(primal) + dual
    tau = solver.point.tau
    res.kap = solver.mu / tau * tau_dir / tau + kap_dir
    return res


class _ElimSolver:
    """Shared 6x6 -> 4x4 -> 3x3 reductions (common.jl:129-208)."""
    cones = None  # oracle solvers use solver.cones

    def setup_point_sub(self, model):
        n, p, q = model.n, model.p, model.q
        self.sol_sub, self.rhs_sub = SubPoint(n, p, q), SubPoint(n, p, q)
        self.rhs_const, self.sol_const = SubPoint(n, p, q), SubPoint(n, p, q)
        self.rhs_const.x[:] = -model.c
        self.rhs_const.y[:] = model.b
        self.rhs_const.z[:] = model.h

    def apply_lhs(self, solver, direction, res):
        return apply_lhs(solver, direction, res)

    def solve_subsystem4(self, solver, sol, rhs):
        model = solver.model
        rhs_sub, sol_sub = self.rhs_sub, self.sol_sub
        rhs_sub.x[:] = rhs.x
        rhs_sub.y[:] = -rhs.y
        self.setup_rhs3(solver, rhs, sol, rhs_sub)
        self.solve_subsystem3(solver, sol_sub, rhs_sub)
        tau_num = rhs.tau + rhs.kap + dot_obj(model, sol_sub)
        taubar = solver.point.tau
        tau_denom = solver.mu / taubar / taubar - dot_obj(model, self.sol_const)
        sol_tau = tau_num / tau_denom
        dim3 = sol_sub.vec.size
        sol.vec[:dim3] = sol_sub.vec + sol_tau * self.sol_const.vec
        sol.tau = sol_tau
        return sol

    def solve_system(self, solver, sol, rhs):
        model = solver.model
        self.solve_subsystem4(solver, sol, rhs)
        tau = sol.tau
        sol.s[:] = model.h * tau - rhs.z - model.G @ sol.x
        taubar = solver.point.tau
        sol.kap = -solver.mu / taubar / taubar * tau + rhs.kap
        return sol

    def free_memory(self):
        pass


class QRCholDenseSystemSolver(_ElimSolver):
    """qrchol.jl:104-257"""

    def load(self, solver):
        model = solver.model
        n, p, q = model.n, model.p, model.q
        self.nmp = n - p
        Q = solver.Ap_Q
        GQ = model.G if Q is None else model.G @ Q
        self.GQ2 = np.asfortranarray(GQ[:, p:])
        self.GQ1 = np.asfortranarray(GQ[:, :p]) if p else None
        self.HGQ2 = np.zeros((q, self.nmp), order="F")
        self.setup_point_sub(model)
        self.fact = None
        self.fact_kind = 0
        self.lhs = None
        return self

    def setup_rhs3(self, solver, rhs, sol, rhs_sub):
        cones = solver.cones
        if cones.dual_mask is None:
            rhs_sub.z[:] = -cones.hess_prod(rhs.z) - rhs.s
            return
        for ck, sl in zip(cones.cones, cones.slices):
            if ck.use_dual_barrier:
                rhs_sub.z[sl] = ck.inv_hess_prod(-rhs.z[sl] - rhs.s[sl])
            else:
                rhs_sub.z[sl] = -ck.hess_prod(rhs.z[sl]) - rhs.s[sl]

    def solve_subsystem3(self, solver, sol, rhs):
        model = solver.model
        p, n = model.p, model.n
        cones = solver.cones
        sol.vec[:] = rhs.vec
        x, y, z = sol.x, sol.y, sol.z
        Q, R = solver.Ap_Q, solver.Ap_R
        t = x + model.G.T @ z
        if Q is not None:
            t = Q.T @ t
        if p:
            y[:] = sla.solve_triangular(R, y, trans="T")
            sol.vec[:p] = y
            if self.nmp:
                HGQ1x = cones.block_hess_prod(self.GQ1 @ y)
                t[p:] -= self.GQ2.T @ HGQ1x
        if self.nmp:
            sol.vec[p:n] = self.fact.solve(t[p:])
        if Q is not None:
            x[:] = Q @ x
        HGx = cones.block_hess_prod(model.G @ x)
        z[:] = HGx - z
        if p:
            y[:] = sla.solve_triangular(R, t[:p] - self.GQ1.T @ HGx)
        return sol

    def update_lhs(self, solver):
        model = solver.model
        cones = solver.cones
        if self.nmp:
            self.update_lhs_fact(solver)
        self.rhs_const.z[:] = cones.block_hess_prod(model.h)
        self.solve_subsystem3(solver, self.sol_const, self.rhs_const)
        return self

    def update_lhs_fact(self, solver):
        cones = solver.cones
        nmp = self.nmp
        use_sqrt = [ck.use_sqrt_hess_oracles(nmp) for ck in cones.cones]
        lhs = np.zeros((nmp, nmp), order="F")
        if any(use_sqrt):
            idx = 0
            for k, (ck, sl) in enumerate(zip(cones.cones, cones.slices)):
                if not use_sqrt[k]:
                    continue
                arr = self.GQ2[sl]
                qk = arr.shape[0]
                self.HGQ2[idx:idx + qk] = ck.inv_sqrt_hess_prod(arr) if ck.use_dual_barrier \
                    else ck.sqrt_hess_prod(arr)
                idx += qk
            # outer_prod! = BLAS.syrk!('U','T',...) (dense.jl:80-86); upper triangle only
            lhs = _blas.dsyrk(1.0, self.HGQ2[:idx], trans=1, lower=0)
            lhs = np.asfortranarray(lhs)
        for k, (ck, sl) in enumerate(zip(cones.cones, cones.slices)):
            if use_sqrt[k]:
                continue
            arr = self.GQ2[sl]
            prod = ck.inv_hess_prod(arr) if ck.use_dual_barrier else ck.hess_prod(arr)
            lhs += arr.T @ prod
        self.use_sqrt_hess_cones = use_sqrt
        self.lhs = lhs
        self.fact = la.posdef_fact_copy(lhs)
        self.fact_kind = self.fact.kind
        return self.fact.issuccess()

    def lhs_full(self):
        """Symmetric Schur matrix (test helper): mirrors the upper triangle."""
        U = np.triu(self.lhs)
        return U + np.triu(self.lhs, 1).T


class SymIndefDenseSystemSolver(_ElimSolver):
    """symindef.jl:203-271: lower-triangular [0; A 0; G 0 -Hinv], Bunch-Kaufman rook."""

    def load(self, solver):
        model = solver.model
        n, p, q = model.n, model.p, model.q
        npq = n + p + q
        lhs = np.zeros((npq, npq), order="F")
        lhs[n:n + p, :n] = model.A
        lhs[n + p:, :n] = model.G
        self.lhs_sub = lhs
        self.setup_point_sub(model)
        return self

    def setup_rhs3(self, solver, rhs, sol, rhs_sub):
        cones = solver.cones
        for ck, sl in zip(cones.cones, cones.slices):
            if ck.use_dual_barrier:
                rhs_sub.z[sl] = -rhs.z[sl] - rhs.s[sl]
            else:
                rhs_sub.z[sl] = -ck.inv_hess_prod(rhs.s[sl]) - rhs.z[sl]

    def update_lhs(self, solver):
        model = solver.model
        z0 = model.n + model.p
        cones = solver.cones
        for ck, sl in zip(cones.cones, cones.slices):
            Hk = ck.hess() if ck.use_dual_barrier else ck.inv_hess()
            rows = slice(z0 + sl.start, z0 + sl.stop)
            self.lhs_sub[rows, rows] = -Hk
        self.fact = la.symm_fact_copy(self.lhs_sub, b"L")
        self.solve_subsystem3(solver, self.sol_const, self.rhs_const)
        return self

    def solve_subsystem3(self, solver, sol, rhs):
        sol.vec[:] = self.fact.solve(rhs.vec)
        return sol


class NaiveDenseSystemSolver:
    """naive.jl:144-216: the unreduced 6x6 system with LU (cross-check only)."""
    cones = None

    def load(self, solver):
        m = solver.model
        n, p, q = m.n, m.p, m.q
        dim = n + p + 2 * q + 2
        t = self.tau_row = n + p + q
        lhs = np.zeros((dim, dim))
        lhs[:n, n:n + p] = m.A.T
        lhs[:n, n + p:t] = m.G.T
        lhs[:n, t] = m.c
        lhs[n:n + p, :n] = -m.A
        lhs[n:n + p, t] = m.b
        lhs[n + p:t, :n] = -m.G
        lhs[n + p:t, t] = m.h
        lhs[n + p:t, t + 1:t + 1 + q] = -np.eye(q)
        lhs[t, :n] = -m.c
        lhs[t, n:n + p] = -m.b
        lhs[t, n + p:t] = -m.h
        lhs[t, -1] = -1.0
        lhs[t + 1:t + 1 + q, n + p:t] = np.eye(q)
        lhs[t + 1:t + 1 + q, t + 1:t + 1 + q] = np.eye(q)
        lhs[-1, t] = 1.0
        lhs[-1, -1] = 1.0
        self.lhs = lhs
        return self

    def update_lhs(self, solver):
        m = solver.model
        t = self.tau_row
        z0 = m.n + m.p
        cones = solver.cones
        for ck, sl in zip(cones.cones, cones.slices):
            rows = slice(t + 1 + sl.start, t + 1 + sl.stop)
            cols = slice(z0 + sl.start, z0 + sl.stop) if ck.use_dual_barrier else rows
            self.lhs[rows, cols] = ck.hess()
        tau = solver.point.tau
        self.lhs[-1, t] = solver.mu / tau / tau
        self.lu = sla.lu_factor(self.lhs)
        return self

    def solve_system(self, solver, sol, rhs):
        sol.vec[:] = sla.lu_solve(self.lu, rhs.vec)
        return sol

    def apply_lhs(self, solver, direction, res):
        return apply_lhs(solver, direction, res)

    def free_memory(self):
        pass

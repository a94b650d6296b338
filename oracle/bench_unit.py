"""CPU timing harness of the oracle for bench.py's `cpu_baseline` / `--impl reference` legs
(test infrastructure; never on the product path).

One unit = what bench.py times on the GPU: load_point of every cone, update_lhs (H^{1/2}G pre-pass,
dsyrk, dpotrf, constant column) and 4 x (solve_system + apply_lhs), following
src/Solvers/steppers/combined.jl:64-79 with zero refinement rounds.
"""
import time

import numpy as np
from scipy.linalg import blas as _blas

from hypatia_b200.host import stepper as st
from .layout import OraclePoint
from . import linalg as la
from . import syssolvers as osys


class _SampledQRChol(osys.QRCholDenseSystemSolver):
    """QRChol whose dsyrk can be timed on a row sample: the Schur matrix of the (fixed) iterate is
    computed in full once, later calls run dsyrk on the first `frac` of the rows and report the
    time that the remaining rows would have added (dsyrk is linear in the rows)."""
    frac = 1.0
    _lhs_cache = None
    extra_time = 0.0

    def update_lhs_fact(self, solver):
        if self.frac >= 1.0 or self._lhs_cache is None:
            ok = super().update_lhs_fact(solver)
            self._lhs_cache = self.lhs
            return ok
        cones = solver.cones
        idx = 0
        for ck, sl in zip(cones.cones, cones.slices):
            arr = self.GQ2[sl]
            qk = arr.shape[0]
            self.HGQ2[idx:idx + qk] = ck.sqrt_hess_prod(arr)
            idx += qk
        rows = max(1, int(round(self.frac * idx)))
        t0 = time.perf_counter()
        _blas.dsyrk(1.0, self.HGQ2[:rows], trans=1, lower=0)
        dt = time.perf_counter() - t0
        self.extra_time += dt * (idx / rows - 1.0)
        self.lhs = self._lhs_cache
        self.fact = la.posdef_fact_copy(self.lhs)
        self.fact_kind = self.fact.kind
        return self.fact.issuccess()


class Shell:
    pass


def iterate_shell(model, s0, z0, x0, mu, cone_cls, syrk_row_fraction=1.0):
    sh = Shell()
    sh.model, sh.mu = model, mu
    sh.Ap_Q, sh.Ap_R = None, np.zeros((0, 0))
    pt = sh.point = OraclePoint(model.n, model.p, model.q)
    pt.x[:] = x0
    pt.z[:] = z0
    pt.s[:] = s0
    pt.tau = pt.kap = 1.0
    sh.x_residual = np.zeros(model.n)
    sh.y_residual = np.zeros(model.p)
    sh.z_residual = np.zeros(model.q)
    sh.tau_residual = float(model.c @ pt.x) + float(model.h @ pt.z) + pt.kap
    sh.cones = cone_cls(model)
    sys_ = sh.syssolver = _SampledQRChol()
    sys_.load(sh)
    irtmu = 1.0 / np.sqrt(mu)
    sh.cones.load_point(pt.s, pt.z, irtmu)
    sys_.update_lhs(sh)              # full (also fills the Schur cache of the sampled variant)
    sys_.frac = syrk_row_fraction
    rhs, d = (OraclePoint(model.n, model.p, model.q) for _ in range(2))
    rhs_list = []
    st.update_rhs_cent(sh, rhs)
    rhs_list.append(rhs.vec.copy())
    sys_.solve_system(sh, d, rhs)
    st.update_rhs_centadj(sh, rhs, d)
    rhs_list.append(rhs.vec.copy())
    st.update_rhs_pred(sh, rhs)
    rhs_list.append(rhs.vec.copy())
    sys_.solve_system(sh, d, rhs)
    st.update_rhs_predadj(sh, rhs, d)
    rhs_list.append(rhs.vec.copy())
    sh.rhs_list = rhs_list

    sol, res, r = (OraclePoint(model.n, model.p, model.q) for _ in range(3))

    def unit(rhs_vecs):
        """Runs one unit; returns the extrapolated extra seconds of a sampled dsyrk."""
        sys_.extra_time = 0.0
        sh.cones.load_point(pt.s, pt.z, irtmu)
        sys_.update_lhs(sh)
        sh.sols = []
        for v in rhs_vecs:
            r.vec[:] = v
            sys_.solve_system(sh, sol, r)
            sys_.apply_lhs(sh, sol, res)
            sh.sols.append(sol.vec.copy())      # bench.py's full-size direction parity
        return sys_.extra_time

    sh.unit = unit
    sh.sol = sol
    return sh

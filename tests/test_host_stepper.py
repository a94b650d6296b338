"""Host-side line-search logic (hypatia.jl_b200/host/stepper.py), CPU tier.

reference: src/Solvers/search.jl:74-138.  Julia's `max` propagates NaN, so a cone whose proximity
comes back NaN makes `agg_proxsqr < proxsqr_bound` false and the candidate point is rejected; the
host mirror must do the same on both aggregation branches."""
import numpy as np
import pytest

from hypatia_b200.host import models as M
from hypatia_b200.host import stepper as st
from hypatia_b200.host.point import Point


class _Cones:
    """Two Nonnegative(2) cones whose oracle sweep always passes; get_proxsqr is scripted."""

    def __init__(self, model, prox):
        self.nus = np.asarray(model.cone_nus, dtype=np.float64)
        self.dual_mask = None
        self.slices = model.cone_idxs
        self.prox = np.asarray(prox, dtype=np.float64)

    def seg_dot(self, a, b):
        return np.array([float(a[sl] @ b[sl]) for sl in self.slices])

    def load_point(self, primal, dual, scal):
        pass

    def is_feas(self):
        return np.ones(len(self.slices), dtype=bool)

    is_dual_feas = is_feas

    def check_numerics(self, irtmu, use_max):
        return np.ones(len(self.slices), dtype=bool)

    def get_proxsqr(self, irtmu, use_max):
        return self.prox


class _Stepper:
    pass


def _setup(prox, use_max):
    model = M.Model(np.zeros(2), None, np.zeros(0), -np.eye(4, 2), np.zeros(4),
                    [M.Nonnegative(2), M.Nonnegative(2)])
    solver = _Stepper()
    solver.cones = _Cones(model, prox)
    stp = _Stepper()
    stp.searcher = st.StepSearcher(model, use_max_prox=use_max)
    stp.temp = Point(model)
    stp.temp.s[:] = 1.0
    stp.temp.z[:] = 1.0
    stp.temp.tau = stp.temp.kap = 1.0
    return solver, stp


@pytest.mark.parametrize("use_max", [True, False])
def test_finite_small_proximity_accepts(use_max):
    solver, stp = _setup([0.01, 0.02], use_max)
    assert st.check_cone_points(solver, stp)
    assert stp.searcher.prox == pytest.approx(np.sqrt(0.02 if use_max else 0.03))


@pytest.mark.parametrize("use_max", [True, False])
@pytest.mark.parametrize("prox", [[np.nan, 0.01], [0.01, np.nan], [np.nan, np.nan]])
def test_nan_proximity_rejects_the_candidate(use_max, prox):
    solver, stp = _setup(prox, use_max)
    assert not st.check_cone_points(solver, stp)


@pytest.mark.parametrize("use_max", [True, False])
def test_large_proximity_rejects(use_max):
    solver, stp = _setup([0.5, 0.99], use_max)
    assert not st.check_cone_points(solver, stp)


def test_oracle_layout_agrees_with_the_host_layout():
    """oracle/layout.py restates point.jl:24-54 independently of hypatia_b200.host.point; the two must
    place every block of the flat vector at the same offsets (tests hand host Points to oracle solvers)."""
    from hypatia_b200.host.coneblock import ConeBlock
    from hypatia_b200.host.point import SubPoint
    from oracle.layout import OracleConeBlockBase, OraclePoint, OracleSubPoint
    cones = [M.Nonnegative(3), M.EpiNormEucl(4), M.HypoPerLogdetTri(5, use_dual=True), M.PosSemidefTri(6)]
    q = sum(c.dim for c in cones)
    model = M.Model(np.zeros(5), np.zeros((2, 5)), np.zeros(2), np.zeros((q, 5)), np.zeros(q), cones)
    hp, op = Point(model), OraclePoint(model.n, model.p, model.q)
    hp.vec[:] = np.arange(hp.vec.size)
    op.vec[:] = hp.vec
    for name in ("x", "y", "z", "s", "ztsk"):
        assert np.array_equal(getattr(hp, name), getattr(op, name)), name
    assert (hp.tau, hp.kap) == (op.tau, op.kap)
    op.tau, op.kap = -1.0, -2.0
    assert op.vec[model.n + model.p + q] == -1.0 and op.vec[-1] == -2.0
    hb, ob = ConeBlock(model), OracleConeBlockBase(model)
    assert np.array_equal(hb.offsets, ob.offsets) and np.array_equal(hb.dims, ob.dims)
    assert np.array_equal(hb.dual_mask, ob.dual_mask) and np.array_equal(hb.nus, ob.nus)
    for a, b in zip(hp.primal_dual(hb.dual_mask), op.primal_dual(ob.dual_mask)):
        assert np.array_equal(a, b)
    assert np.allclose(hb.seg_dot(hp.s, hp.z), ob.seg_dot(hp.s, hp.z), rtol=1e-15)
    hs, os_ = SubPoint(model.n, model.p, q), OracleSubPoint(model.n, model.p, q)
    hs.vec[:] = np.arange(hs.vec.size)
    os_.vec[:] = hs.vec
    for name in ("x", "y", "z"):
        assert np.array_equal(getattr(hs, name), getattr(os_, name))

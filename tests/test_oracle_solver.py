"""Pins the oracle's system-solver restatements and the host driver with the reference's
deterministic known-answer instances (test/nativeinstances.jl, tol eps^(1/4)) and with the
reference's own cross-solver equivalence check (test/runnativetests.jl:101-118: every system
solver must solve `inst_minimal` with reduce=false)."""
import numpy as np
import pytest

import kat_instances as kat
from hypatia_b200.host import instances as inst
from hypatia_b200.host.point import Point
from hypatia_b200.host.solver import Solver
from oracle import syssolvers as osys
from oracle.cones import OracleConeBlock


def solve(model, syssolver, **kw):
    # the reference runs its instance tests with default_tol_relax = 10 (test/runnativetests.jl:13-18)
    kw.setdefault("default_tol_relax", 10)
    s = Solver(model, syssolver, OracleConeBlock, **kw)
    s.solve()
    return s


@pytest.mark.parametrize("build", kat.ALL + kat.EXTRA + kat.NEW_CONES + kat.SPECTRAL + kat.SPECTRAL_VEC, ids=lambda f: f.__name__)
def test_kat_qrchol_default(build):
    model, expected = build()
    s = solve(model, osys.QRCholDenseSystemSolver())
    kat.check_solution(s, model, expected)


@pytest.mark.parametrize("build", kat.ALL, ids=lambda f: f.__name__)
@pytest.mark.parametrize("sys_cls", [osys.QRCholDenseSystemSolver, osys.SymIndefDenseSystemSolver,
                                     osys.NaiveDenseSystemSolver], ids=lambda c: c.__name__)
def test_kat_all_syssolvers_no_reduce(build, sys_cls):
    model, expected = build()
    s = solve(model, sys_cls(), reduce=False)
    kat.check_solution(s, model, expected)


@pytest.mark.parametrize("build", [kat.epipersquare4, kat.hypoperlog1, kat.hypoperlog4, kat.hypoperlog7,
                                   kat.SPECTRAL[2], kat.SPECTRAL[9], kat.SPECTRAL[14]],
                         ids=lambda f: f.__name__)
@pytest.mark.parametrize("sys_cls", [osys.SymIndefDenseSystemSolver, osys.NaiveDenseSystemSolver],
                         ids=lambda c: c.__name__)
def test_kat_new_cones_other_syssolvers(build, sys_cls):
    model, expected = build()
    s = solve(model, sys_cls(), reduce=False)
    kat.check_solution(s, model, expected)


def test_kat_no_preprocess_symindef():
    # reference: runnativetests.jl:80-88 (no preprocess => SymIndefDense)
    for build in (kat.nonnegative4, kat.epinormeucl1, kat.hyporootdettri4):
        model, expected = build()
        s = solve(model, osys.SymIndefDenseSystemSolver(), preprocess=False, reduce=False)
        kat.check_solution(s, model, expected)


def test_linearopt_c1_symindef_vs_qrchol():
    """BASELINE config 1 (examples/linearopt native, SymIndefDense, CPU only) at reduced size;
    QRChol must reach the same optimum."""
    model = inst.linearopt(40, 80, seed=7)
    s1 = solve(model, osys.SymIndefDenseSystemSolver(), reduce=False)
    s2 = solve(model, osys.QRCholDenseSystemSolver())
    assert s1.status == s2.status == "Optimal"
    assert abs(s1.primal_obj - s2.primal_obj) <= 1e-6 * (1 + abs(s1.primal_obj))
    kat.check_solution(s1, model, dict(status="Optimal"))
    kat.check_solution(s2, model, dict(status="Optimal"))


def _iterate_solver(instance, sys_cls, Ap=None):
    """A Solver shell positioned at the instance's planted iterate (no solve loop)."""
    model = instance.model
    s = Solver(model, sys_cls(), OracleConeBlock)
    s.model = model
    s.point = instance.point
    s.mu = instance.mu
    s.Ap_Q, s.Ap_R = (None, np.zeros((0, 0))) if Ap is None else Ap
    s.cones = OracleConeBlock(model)
    primal, dual = s.point.primal_dual(s.cones.dual_mask)
    s.cones.load_point(primal, dual, 1 / np.sqrt(s.mu))
    s.syssolver.load(s)
    s.syssolver.update_lhs(s)
    return s


@pytest.mark.parametrize("p", [0, 3])
def test_directions_qrchol_vs_naive_mixed_cones(p):
    """Direction-level cross-check on a mixed-cone instance: the reduced QRChol solve must agree
    with the unreduced 6x6 LU solve (the reference's notion of solver equivalence)."""
    from hypatia_b200.host import models as M
    import scipy.linalg as sla
    cones = [M.Nonnegative(5), M.EpiNormEucl(4), M.PosSemidefTri(6), M.HypoPerLogdetTri(8),
             M.HypoRootdetTri(7), M.EpiNormEucl(3), M.HypoPerLogdetTri(5, use_dual=True)]
    I = inst.synthetic("mix", 12, p, cones, seed=11)
    Ap = None
    if p:
        Qf, Rf = sla.qr(I.model.A.T, mode="full")
        Ap = (Qf, np.triu(Rf[:p, :p]))
    a = _iterate_solver(I, osys.QRCholDenseSystemSolver, Ap)
    b = _iterate_solver(I, osys.NaiveDenseSystemSolver)
    c = _iterate_solver(I, osys.SymIndefDenseSystemSolver)
    rng = np.random.default_rng(5)
    rhs = Point(I.model)
    rhs.vec[:] = rng.standard_normal(rhs.vec.size)
    sa, sb, sc = Point(I.model), Point(I.model), Point(I.model)
    a.syssolver.solve_system(a, sa, rhs)
    b.syssolver.solve_system(b, sb, rhs)
    c.syssolver.solve_system(c, sc, rhs)
    nrm = np.linalg.norm(sb.vec)
    assert np.linalg.norm(sa.vec - sb.vec) <= 1e-9 * nrm
    assert np.linalg.norm(sc.vec - sb.vec) <= 1e-9 * nrm
    res = Point(I.model)
    a.syssolver.apply_lhs(a, sa, res)
    assert np.linalg.norm(res.vec - rhs.vec) <= 1e-9 * np.linalg.norm(rhs.vec)

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


# ---- stepper tests of the reference (test/runnativetests.jl:120-158) on inst_minimal (test/nativesets.jl:21-26) ----
INST_MINIMAL = [kat.primalinfeas3, kat.dualinfeas3, kat.epinorminf4, kat.hyporootdettri4]


@pytest.mark.parametrize("adj,curv", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("build", INST_MINIMAL, ids=lambda f: f.__name__)
def test_predorcent_stepper(build, adj, curv):
    from hypatia_b200.host.stepper import PredOrCentStepper
    model, expected = build()
    s = solve(model, osys.QRCholDenseSystemSolver(), stepper=PredOrCentStepper(use_adjustment=adj, use_curve_search=curv))
    kat.check_solution(s, model, expected)


@pytest.mark.parametrize("build", INST_MINIMAL, ids=lambda f: f.__name__)
def test_predorcent_stepper_other_options(build):
    from hypatia_b200.host.stepper import PredOrCentStepper
    model, expected = build()
    stepper = PredOrCentStepper(use_adjustment=False, use_curve_search=False, max_cent_steps=8, pred_prox_bound=0.0332,
                                min_prox=0.0, prox_bound=0.2844, use_max_prox=False,
                                alpha_sched=[0.9999 * 0.7 ** i for i in range(23)])
    kat.check_solution(solve(model, osys.QRCholDenseSystemSolver(), stepper=stepper), model, expected)


@pytest.mark.parametrize("shift", [0, 2])
@pytest.mark.parametrize("build", INST_MINIMAL, ids=lambda f: f.__name__)
def test_combined_stepper_shift_sched(build, shift):
    from hypatia_b200.host.stepper import CombinedStepper
    model, expected = build()
    kat.check_solution(solve(model, osys.QRCholDenseSystemSolver(), stepper=CombinedStepper(shift_sched=shift)), model, expected)


def test_predorcent_needs_fewer_system_solves_per_iteration():
    """One direction pair per iteration instead of the combined stepper's four directions (predorcent.jl:72-110)."""
    from hypatia_b200.host.stepper import CombinedStepper, PredOrCentStepper
    model, expected = kat.epinormeucl1()
    a = solve(model, osys.QRCholDenseSystemSolver(), stepper=CombinedStepper(), max_ref_steps=0)
    model, expected = kat.epinormeucl1()
    b = solve(model, osys.QRCholDenseSystemSolver(), stepper=PredOrCentStepper(), max_ref_steps=0)
    kat.check_solution(b, model, expected)
    assert a.n_solve_system == 4 * a.num_iters and b.n_solve_system == 2 * b.num_iters

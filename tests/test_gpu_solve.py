"""-m gpu end-to-end solves with the device system solver AND the device cone oracles plugged into
the host driver (the stand-in for the reference's Julia stepper): the reference's deterministic
known-answer instances (test/nativeinstances.jl, tol eps^(1/4)) must come out exactly as they do
with the CPU oracle plug-ins - same status, same closed-form optimum."""
import numpy as np
import pytest

import kat_instances as kat
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from hypatia_b200.host.solver import Solver

pytestmark = pytest.mark.gpu


def _solve_dev(model, **kw):
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    # the reference runs its instance tests with default_tol_relax = 10 (test/runnativetests.jl:13-18)
    kw.setdefault("default_tol_relax", 10)
    s = Solver(model, DevQRChol(), DeviceConeBlock, **kw)
    s.solve()
    return s


def _solve_ora(model, **kw):
    from oracle.cones import OracleConeBlock
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    kw.setdefault("default_tol_relax", 10)
    s = Solver(model, OraQRChol(), OracleConeBlock, **kw)
    s.solve()
    return s


@pytest.mark.parametrize("build", kat.ALL, ids=lambda f: f.__name__)
def test_kat_device_default(build):
    model, expected = build()
    kat.check_solution(_solve_dev(model), model, expected)


# the reference's instances for the cones added in the widening step (EpiPerSquare, HypoPerLog,
# EpiPerSepSpectral{MatrixCSqr} with every separable spectral function, primal and dual barrier)
NEW_CONE_KATS = kat.GPOW + kat.HPM + kat.RELENT + kat.NORMSPEC + kat.WSOS + [kat.hypogeomean1, kat.hypogeomean2_dual, kat.hypogeomean4, kat.hypogeomean6, kat.epinorminf1, kat.epinorminf2, kat.epinorminf4, kat.dualinfeas1, kat.primalinfeas3, kat.dualinfeas2, kat.epipersquare1, kat.epipersquare2, kat.epipersquare4,
                 kat.hypoperlog1, kat.hypoperlog4, kat.hypoperlog5, kat.hypoperlog7] + \
    [f for f in kat.SPECTRAL if "_d3_" in f.__name__ or "_d2_" in f.__name__] + \
    [f for f in kat.SPECTRAL_VEC if "_d3_" in f.__name__ or "_d4_" in f.__name__ or "vector3" in f.__name__
     or "vector4" in f.__name__]


@pytest.mark.parametrize("build", NEW_CONE_KATS, ids=lambda f: f.__name__)
def test_kat_device_new_cones(build):
    model, expected = build()
    kat.check_solution(_solve_dev(model), model, expected)


@pytest.mark.parametrize("build", [kat.nonnegative4, kat.epinormeucl1, kat.possemideftri8,
                                   kat.hyporootdettri4, kat.hypoperlogdettri4],
                         ids=lambda f: f.__name__)
def test_kat_device_no_reduce(build):
    # reduce=false keeps p > 0: exercises the Ap_Q / Ap_R branch of solve_subsystem3
    model, expected = build()
    kat.check_solution(_solve_dev(model, reduce=False), model, expected)


def test_linearopt_and_soc_solves_match_oracle_iterates():
    """Same status / objective and (nearly) the same iteration count as the oracle-driven solve: the two
    plug-in pairs see the same host driver, so any difference comes from the device arithmetic.  The
    last iterations run at the numerical limit of the ill-conditioned Schur system, where a rounding-
    level difference can flip one line-search decision, hence the +-3 iterations allowance."""
    cases = [inst.linearopt(40, 80, seed=7),
             inst.synthetic("soc", 60, 0, [M.EpiNormEucl(25) for _ in range(8)], seed=21).model,
             inst.synthetic("mix", 30, 4, [M.Nonnegative(10), M.PosSemidefTri(15), M.EpiNormEucl(6),
                                           M.HypoPerLogdetTri(8)], seed=22).model]
    for model in cases:
        sd, so = _solve_dev(model), _solve_ora(model)
        assert sd.status == so.status
        assert abs(sd.num_iters - so.num_iters) <= 3
        if so.status == "Optimal":
            assert abs(sd.primal_obj - so.primal_obj) <= 1e-6 * (1 + abs(so.primal_obj))


def test_solves_with_device_residuals():
    """Full solves with the residual step of calc_convergence_params on the device (hyp_calc_residuals,
    SURVEY.md 8(f) rank 3): same status / optimum as with the host residuals."""
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    cases = [kat.nonnegative4()[0], kat.epinormeucl1()[0], kat.possemideftri8()[0], kat.hypoperlog5()[0],
             kat.epinorminf2()[0], inst.linearopt(40, 80, seed=7),
             inst.synthetic("soc", 60, 0, [M.EpiNormEucl(25) for _ in range(8)], seed=21).model]
    for model in cases:
        s1 = Solver(model.copy(), DevQRChol(device_residuals=True), DeviceConeBlock, default_tol_relax=10)
        s1.solve()
        s0 = _solve_dev(model.copy())
        assert s1.status == s0.status
        assert abs(s1.num_iters - s0.num_iters) <= 3
        if s0.status == "Optimal":
            assert abs(s1.primal_obj - s0.primal_obj) <= 1e-6 * (1 + abs(s0.primal_obj))


def test_c1_linearopt_symindef_device():
    """BASELINE config 1 (examples/linearopt native, dense A 200 x 400, Nonnegative, SymIndefDense,
    no reduction) solved with the device SymIndefDense plug-in; same optimum as the oracle's."""
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.syssolver import SymIndefDenseSystemSolver as DevSym
    from oracle.cones import OracleConeBlock
    from oracle.syssolvers import SymIndefDenseSystemSolver as OraSym
    model = inst.linearopt(60, 120, seed=3)
    sd = Solver(model, DevSym(), DeviceConeBlock, reduce=False)
    sd.solve()
    so = Solver(model, OraSym(), OracleConeBlock, reduce=False)
    so.solve()
    assert sd.status == so.status == "Optimal"
    assert abs(sd.primal_obj - so.primal_obj) <= 1e-6 * (1 + abs(so.primal_obj))
    kat.check_solution(sd, model, dict(status="Optimal"))

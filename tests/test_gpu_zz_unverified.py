"""-m gpu tests of the cone types added at the end of round 1 (LinMatrixIneq, DoublyNonnegativeTri, MatrixEpiPerSquare,
PosSemidefTriSparse, EpiTrRelEntropyTri, the three matrix / norm WSOS cones), the late reference instances (kat.EXTRA) and
PredOrCentStepper with the device plug-ins.  All of them passed on a B200 in the round-1 driver run (GPUTEST_r01: 67 XPASS)
and are ordinary tests since round 2: a regression here fails the suite.  The file name still sorts last."""
import numpy as np
import pytest

import kat_instances as kat
from gpu_util import iterate_solver, rel
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from hypatia_b200.host.point import Point

pytestmark = pytest.mark.gpu



def _lmi(rng, side, dim, use_dual=False):
    As = []
    for i in range(dim):
        X = rng.random((side, side))
        As.append(X @ X.T + np.eye(side) if i == 0 else (X + X.T) / 2 - 0.5)
    return M.LinMatrixIneq(As, use_dual=use_dual)


def test_linmatrixineq_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    rng = np.random.default_rng(7)
    cones = [_lmi(rng, 2, 2), _lmi(rng, 3, 2), _lmi(rng, 4, 3), _lmi(rng, 3, 6), _lmi(rng, 12, 40),
             _lmi(rng, 5, 4, use_dual=True), _lmi(rng, 33, 20), M.Nonnegative(3)]
    I = inst.synthetic("lmi", 4, 0, cones, seed=77)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-9
    assert rel(dev.block_hess_prod(arr[:, 0]), ora.block_hess_prod(arr[:, 0])) <= 1e-9
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


def test_linmatrixineq_in_the_system_solve():
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    rng = np.random.default_rng(8)
    cones = [_lmi(rng, 4, 5), M.EpiNormEucl(5), _lmi(rng, 3, 3, use_dual=True), M.Nonnegative(4)]
    I = inst.synthetic("lmimix", 9, 0, cones, seed=31)
    dev, ora = iterate_solver(I, DevQRChol()), iterate_solver(I, OraQRChol())
    try:
        assert rel(dev.syssolver.lhs_full(), ora.syssolver.lhs_full()) <= 1e-11
        rhs = Point(I.model)
        rhs.vec[:] = np.random.default_rng(3).standard_normal(rhs.vec.size)
        sd, so = Point(I.model), Point(I.model)
        dev.syssolver.solve_system(dev, sd, rhs)
        ora.syssolver.solve_system(ora, so, rhs)
        assert rel(sd.vec, so.vec) <= 1e-8
    finally:
        dev.syssolver.free_memory()


def _solve_dev(model):
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.host.solver import Solver
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    s = Solver(model, DevQRChol(), DeviceConeBlock, default_tol_relax=10)    # test/runnativetests.jl:13-18
    s.solve()
    return s


@pytest.mark.parametrize("build", kat.EXTRA, ids=lambda f: f.__name__)
def test_kat_device_extra(build):
    model, expected = build()
    kat.check_solution(_solve_dev(model), model, expected)


# the DoublyNonnegativeTri kernels have not run on a GPU at all: keep them after everything else
def test_doublynonnegativetri_oracles_match_cpu_oracle():
    """Not yet run on a GPU (emulation tier: tests/test_emu_gpow.py)."""
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    cones = [M.DoublyNonnegativeTri(M.svec_length(sd)) for sd in (1, 2, 3, 5, 10, 15)] + \
        [M.DoublyNonnegativeTri(M.svec_length(4), use_dual=True), M.EpiNormEucl(4)]
    I = inst.synthetic("dnn", 4, 0, cones, seed=78)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-9
    assert rel(dev.block_hess_prod(arr[:, 0]), ora.block_hess_prod(arr[:, 0])) <= 1e-9
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


def test_doublynonnegativetri_in_the_system_solve():
    """Not yet run on a GPU."""
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    cones = [M.DoublyNonnegativeTri(10), M.Nonnegative(3), M.DoublyNonnegativeTri(6, use_dual=True), M.EpiNormEucl(4)]
    I = inst.synthetic("dnnmix", 8, 0, cones, seed=32)
    dev, ora = iterate_solver(I, DevQRChol()), iterate_solver(I, OraQRChol())
    try:
        assert rel(dev.syssolver.lhs_full(), ora.syssolver.lhs_full()) <= 1e-11
        rhs = Point(I.model)
        rhs.vec[:] = np.random.default_rng(3).standard_normal(rhs.vec.size)
        sd, so = Point(I.model), Point(I.model)
        dev.syssolver.solve_system(dev, sd, rhs)
        ora.syssolver.solve_system(ora, so, rhs)
        assert rel(sd.vec, so.vec) <= 1e-8
    finally:
        dev.syssolver.free_memory()


# the predict-or-center stepper (host control flow added late; it drives the same, GPU-verified, C-ABI calls in a new order)
@pytest.mark.parametrize("adj,curv", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("build", [kat.primalinfeas3, kat.dualinfeas3, kat.epinorminf4, kat.hyporootdettri4],
                         ids=lambda f: f.__name__)
def test_predorcent_stepper_device(build, adj, curv):
    """test/runnativetests.jl:120-132 with the device plug-ins."""
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.host.solver import Solver
    from hypatia_b200.host.stepper import PredOrCentStepper
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    model, expected = build()
    s = Solver(model, DevQRChol(), DeviceConeBlock, default_tol_relax=10,
               stepper=PredOrCentStepper(use_adjustment=adj, use_curve_search=curv))
    s.solve()
    kat.check_solution(s, model, expected)


# the MatrixEpiPerSquare kernels have not run on a GPU either (emulation tier: tests/test_emu_gpow.py)
def test_matrixepipersquare_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    cones = [M.MatrixEpiPerSquare(a, b) for a, b in ((1, 1), (1, 2), (2, 2), (2, 4), (3, 4), (1, 100), (8, 11), (5, 9))] + \
        [M.MatrixEpiPerSquare(2, 3, use_dual=True), M.Nonnegative(2)]
    I = inst.synthetic("meps", 4, 0, cones, seed=79)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    assert (dev.is_dual_feas() == ora.is_dual_feas()).all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-9
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


# WSOSInterpPosSemidefTri: emulation tier only so far (tests/test_emu_gpow.py)
def test_wsosinterppossemideftri_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    from wsos_util import interpolate_box

    def wpsd(R, n, halfdeg, use_dual=False):
        U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
        return M.WSOSInterpPosSemidefTri(R, U, Ps, use_dual=use_dual)
    cones = [wpsd(1, 1, 1), wpsd(2, 1, 2), wpsd(3, 1, 1), wpsd(2, 2, 1), wpsd(3, 2, 1), wpsd(2, 2, 2),
             wpsd(2, 1, 3, use_dual=True), wpsd(4, 1, 2), M.Nonnegative(2)]
    I = inst.synthetic("wsospsd", 4, 0, cones, seed=80)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-8
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


# WSOSInterpEpiNormEucl: emulation tier only so far (tests/test_emu_gpow.py)
def test_wsosinterpepinormeucl_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    from wsos_util import interpolate_box

    def weuc(R, n, halfdeg, use_dual=False):
        U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
        return M.WSOSInterpEpiNormEucl(R, U, Ps, use_dual=use_dual)
    cones = [weuc(2, 1, 1), weuc(2, 1, 2), weuc(3, 1, 2), weuc(3, 2, 1), weuc(4, 2, 1), weuc(2, 2, 2),
             weuc(3, 1, 3, use_dual=True), weuc(8, 2, 2), M.Nonnegative(2)]
    I = inst.synthetic("wsoseucl", 4, 0, cones, seed=81)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-8
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


# WSOSInterpEpiNormOne: emulation tier only so far (tests/test_emu_gpow.py)
def test_wsosinterpepinormone_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    from wsos_util import interpolate_box

    def wone(R, n, halfdeg, use_dual=False):
        U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
        return M.WSOSInterpEpiNormOne(R, U, Ps, use_dual=use_dual)
    cones = [wone(2, 1, 1), wone(2, 1, 2), wone(3, 1, 2), wone(3, 2, 1), wone(4, 2, 1), wone(2, 2, 2),
             wone(3, 1, 3, use_dual=True), wone(8, 2, 2), M.Nonnegative(2)]
    I = inst.synthetic("wsosone", 4, 0, cones, seed=82)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-8
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


# PosSemidefTriSparse: emulation tier only so far (tests/test_emu_gpow.py)
def test_possemideftrisparse_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    rng = np.random.default_rng(9)

    def sps(side, use_dual=False):
        mask = np.tril(rng.random((side, side)) < 1 / np.sqrt(side)) | np.eye(side, dtype=bool)
        rows, cols = np.nonzero(mask)
        if rows.size > 128:
            keep = np.sort(np.concatenate((np.nonzero(rows == cols)[0], np.nonzero(rows != cols)[0][:128 - side])))
            rows, cols = rows[keep], cols[keep]
        return M.PosSemidefTriSparse(side, rows, cols, use_dual=use_dual)
    cones = [sps(sd) for sd in (1, 2, 5, 10, 25, 40)] + [sps(6, use_dual=True), M.Nonnegative(2)]
    I = inst.synthetic("psdsparse", 4, 0, cones, seed=83)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-8
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()


# EpiTrRelEntropyTri: emulation tier only so far (tests/test_emu_gpow.py)
def test_epitrrelentropytri_oracles_match_cpu_oracle():
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    cones = [M.EpiTrRelEntropyTri(1 + 2 * M.svec_length(d)) for d in (1, 2, 3, 4, 6, 10)] + \
        [M.EpiTrRelEntropyTri(1 + 2 * M.svec_length(3), use_dual=True), M.Nonnegative(2)]
    I = inst.synthetic("epitrrelent", 4, 0, cones, seed=84)
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    arr = np.random.default_rng(1).standard_normal((I.model.q, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-8
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-9
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # test/cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # test/cone.jl:71
    dev.free()

"""CPU-tier checks of the device code of GeneralizedPower, HypoPowerMean, EpiNormSpectral and of the generic inverse-Hessian product
(csrc/cones_gpow_kernels.cuh, compiled for the host by tests/emu/) against the CPU oracle."""
import numpy as np
import pytest

import emu_util as eu
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from oracle.cones import OracleConeBlock


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def _alpha(rng, m):
    a = rng.random(m) + 0.05
    return a / a.sum()


def _bal(rng, d):
    a = rng.random(d) + 1        # balanced powers as in the reference's tests (test/cone.jl:291-296)
    return a / a.sum()


def _wsos(n, halfdeg, use_dual=False):
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpNonnegative(U, Ps, use_dual=use_dual)


def _wpsd(R, n, halfdeg, use_dual=False):
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpPosSemidefTri(R, U, Ps, use_dual=use_dual)


def _weuc(R, n, halfdeg, use_dual=False):
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpEpiNormEucl(R, U, Ps, use_dual=use_dual)


def _wone(R, n, halfdeg, use_dual=False):
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpEpiNormOne(R, U, Ps, use_dual=use_dual)


def _sps(rng, side, use_dual=False):
    mask = np.tril(rng.random((side, side)) < 1 / np.sqrt(side)) | np.eye(side, dtype=bool)
    rows, cols = np.nonzero(mask)
    if rows.size > 128:            # keep within the batched path: drop off-diagonal entries
        keep = np.concatenate((np.nonzero(rows == cols)[0], np.nonzero(rows != cols)[0][:128 - side]))
        rows, cols = rows[np.sort(keep)], cols[np.sort(keep)]
    return M.PosSemidefTriSparse(side, rows, cols, use_dual=use_dual)


def _lmi(rng, side, dim, use_dual=False):
    As = []
    for i in range(dim):
        X = rng.random((side, side))
        As.append(X @ X.T + np.eye(side) if i == 0 else (X + X.T) / 2 - 0.5)
    return M.LinMatrixIneq(As, use_dual=use_dual)


def _sets():
    rng = np.random.default_rng(7)
    return {
        "dnn": [M.DoublyNonnegativeTri(M.svec_length(sd)) for sd in (1, 2, 3, 5, 10, 15)] +
               [M.DoublyNonnegativeTri(M.svec_length(4), use_dual=True)],
        "meps": [M.MatrixEpiPerSquare(a, b) for a, b in ((1, 1), (1, 2), (2, 2), (2, 4), (3, 4), (1, 100), (8, 11), (5, 9))] +
                [M.MatrixEpiPerSquare(2, 3, use_dual=True)],
        "wsospsd": [_wpsd(1, 1, 1), _wpsd(2, 1, 2), _wpsd(3, 1, 1), _wpsd(2, 2, 1), _wpsd(3, 2, 1), _wpsd(2, 2, 2),
                    _wpsd(2, 1, 3, use_dual=True), _wpsd(4, 1, 2)],
        "wsoseucl": [_weuc(2, 1, 1), _weuc(2, 1, 2), _weuc(3, 1, 2), _weuc(3, 2, 1), _weuc(4, 2, 1), _weuc(2, 2, 2),
                     _weuc(3, 1, 3, use_dual=True), _weuc(8, 2, 2)],
        "wsosone": [_wone(2, 1, 1), _wone(2, 1, 2), _wone(3, 1, 2), _wone(3, 2, 1), _wone(4, 2, 1), _wone(2, 2, 2),
                    _wone(3, 1, 3, use_dual=True), _wone(8, 2, 2)],
        "psdsparse": [_sps(rng, sd) for sd in (1, 2, 5, 10, 25, 40)] + [_sps(rng, 6, use_dual=True)],
        "epitrrelent": [M.EpiTrRelEntropyTri(1 + 2 * M.svec_length(d)) for d in (1, 2, 3, 4, 6, 10)] +
                       [M.EpiTrRelEntropyTri(1 + 2 * M.svec_length(3), use_dual=True)],
        "lmi": [_lmi(rng, 2, 2), _lmi(rng, 3, 2), _lmi(rng, 4, 3), _lmi(rng, 3, 6), _lmi(rng, 12, 40),
                _lmi(rng, 5, 4, use_dual=True), _lmi(rng, 33, 20)],
        "wsos": [_wsos(1, 1), _wsos(1, 3), _wsos(2, 2), _wsos(3, 1), _wsos(2, 4), _wsos(1, 2, use_dual=True),
                 _wsos(3, 2)],
        "gpow": [M.GeneralizedPower(_alpha(rng, m), n) for m, n in ((2, 1), (3, 2), (4, 1), (2, 4), (20, 30), (40, 5))],
        "hpm": [M.HypoPowerMean(_bal(rng, d)) for d in (1, 2, 5, 33, 70)],
        "hpm_dual": [M.HypoPowerMean(_bal(rng, 3), use_dual=True), M.HypoPowerMean(_bal(rng, 6)),
                     M.HypoPowerMean(_bal(rng, 40), use_dual=True)],
        "normspec": [M.EpiNormSpectral(a, b) for a, b in ((1, 1), (1, 2), (2, 2), (2, 4), (3, 4), (5, 9), (1, 127), (11, 11))],
        "normspec_dual": [M.EpiNormSpectral(2, 3, use_dual=True), M.EpiNormSpectral(3, 5),
                          M.EpiNormSpectral(4, 20, use_dual=True)],
        "gpow_dual": [M.GeneralizedPower(_alpha(rng, 2), 1, use_dual=True), M.GeneralizedPower(_alpha(rng, 3), 3),
                      M.GeneralizedPower(_alpha(rng, 5), 33, use_dual=True)],
    }


NAMES = ["gpow", "gpow_dual", "hpm", "hpm_dual", "normspec", "normspec_dual", "wsos", "lmi", "dnn", "meps", "wsospsd", "wsoseucl", "wsosone", "psdsparse", "epitrrelent"]


@pytest.mark.parametrize("name", NAMES)
def test_gpow_kernels_match_oracle(name):
    cones = _sets()[name]
    I = inst.synthetic(name, 3, 0, cones, seed=600 + NAMES.index(name))
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    assert ora.is_feas().all() and ora.is_dual_feas().all()
    dev = eu.EmuGpowGroup(cones)
    dev.load_point(scal * prim, dual)
    assert dev.feas.all() and dev.dual_feas.all()
    assert rel(dev.grad, ora.grad()) <= 1e-13
    for c, ck in enumerate(ora.cones):
        assert rel(dev.lay.get(dev.H, c), np.asarray(ck.hess())) <= 1e-12      # explicit Hessian
    rng = np.random.default_rng(5)
    arr = rng.standard_normal((I.model.q, 3))
    assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-10
    assert rel(dev.prod(arr, 1, in_place=True), ora.inv_hess_prod(arr)) <= 1e-10
    assert rel(dev.prod(arr, 4), ora.block_hess_prod(arr)) <= 1e-10
    # EpiTrRelEntropyTri: third divided differences of log over Jacobi (device) vs LAPACK (oracle) eigenvalues
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= (1e-10 if name == "epitrrelent" else 1e-12)
    pt = scal * prim
    assert rel(dev.prod(pt, 0), -dev.grad) <= 1e-12
    assert rel(dev.prod(dev.grad, 1), -pt) <= 1e-10
    assert rel(-dev.dder3(pt), dev.grad) <= 1e-11


@pytest.mark.parametrize("name", [n for n in NAMES if not n.endswith("_dual")])
def test_gpow_group_reused_across_iterates(name):
    """The per-cone scratch of a device group (g.d_vecs, the explicit Hessian, its factor) persists from one
    interior-point iterate to the next, and an infeasible trial point of the line search comes in between: a group
    that has seen another point and an infeasible point must give bit-identical results to a fresh one."""
    full = _sets()[name]
    cones = full[:2] + full[-1:] if len(full) > 3 else full     # state reuse does not depend on the sizes: three cones keep
    I = inst.synthetic(name, 3, 0, cones, seed=600 + NAMES.index(name))   # the CPU tier short (one pthread per CUDA thread)
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    rng = np.random.default_rng(11)
    first = prim * (1.0 + 0.02 * rng.standard_normal(prim.shape))     # a nearby interior point
    dev = eu.EmuGpowGroup(cones)
    dev.load_point(first, dual)
    assert dev.feas.all()
    dev.load_point(-prim, dual)                                          # every cone rejects it
    assert not dev.feas.any()
    dev.load_point(prim, dual)
    fresh = eu.EmuGpowGroup(cones)
    fresh.load_point(prim, dual)
    assert dev.feas.all() and fresh.feas.all()
    assert np.array_equal(dev.grad, fresh.grad)
    arr = rng.standard_normal((I.model.q, 2))
    for mode in (0, 1):
        assert np.array_equal(dev.prod(arr, mode), fresh.prod(arr, mode))
    assert np.array_equal(dev.dder3(arr[:, 0]), fresh.dder3(arr[:, 0]))


def test_gpow_kernels_flag_infeasible_points():
    rng = np.random.default_rng(8)
    cones = [M.GeneralizedPower(_alpha(rng, 2), 2), M.GeneralizedPower(_alpha(rng, 3), 1),
             M.GeneralizedPower(_alpha(rng, 2), 1)]
    I = inst.synthetic("gpinf", 2, 0, cones, seed=14)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    prim[0] = -1.0             # u_1 < 0
    prim[4 + 3] = 50.0         # |w| above prod u^alpha
    dual[8 + 2] = 30.0         # dual |w| too large
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuGpowGroup(cones)
    dev.load_point(prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all() and not dev.feas[:2].any() and dev.feas[2]
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all() and not dev.dual_feas[2]


def test_normspec_kernels_flag_infeasible_points():
    """u must exceed the largest singular value (primal) / the nuclear norm (dual): the device Cholesky of
    u^2 I - W W' and the one-sided Jacobi singular values against the oracle (LAPACK) at the boundary."""
    cones = [M.EpiNormSpectral(2, 3), M.EpiNormSpectral(3, 4), M.EpiNormSpectral(2, 2), M.EpiNormSpectral(1, 4)]
    I = inst.synthetic("nsinf", 2, 0, cones, seed=13)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    rng = np.random.default_rng(2)
    W = rng.standard_normal((2, 3))
    sv = np.linalg.svd(W, compute_uv=False)
    prim[0] = sv[0] * (1 - 1e-9)                    # cone 0: u just below sigma_max
    prim[1:7] = W.ravel(order="F")
    o1 = 7
    Wd = rng.standard_normal((3, 4))
    dual[o1] = np.linalg.svd(Wd, compute_uv=False).sum() * (1 - 1e-9)      # cone 1: dual u just below the nuclear norm
    dual[o1 + 1:o1 + 13] = Wd.ravel(order="F")
    o2 = o1 + 13
    W2 = rng.standard_normal((2, 2))
    prim[o2] = np.linalg.svd(W2, compute_uv=False)[0] * (1 + 1e-6)         # cone 2: just inside
    prim[o2 + 1:o2 + 5] = W2.ravel(order="F")
    dual[o2] = np.linalg.svd(W2, compute_uv=False).sum() * (1 + 1e-9)
    dual[o2 + 1:o2 + 5] = W2.ravel(order="F")
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuGpowGroup(cones)
    dev.load_point(prim, dual)
    assert (ora.is_feas() == np.array([False, True, True, True])).all()
    assert (ora.is_dual_feas() == np.array([True, False, True, True])).all()
    # the explicit Hessian of cone 2 (1e-6 from the boundary) may fail its Cholesky on either side: compare the cone's own flag
    assert (dev.feas.astype(bool)[[0, 1, 3]] == ora.is_feas()[[0, 1, 3]]).all()
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all()


def test_dnn_kernels_flag_infeasible_points():
    """DoublyNonnegativeTri leaves the cone through a nonpositive svec entry or through an indefinite matrix."""
    cones = [M.DoublyNonnegativeTri(6), M.DoublyNonnegativeTri(6), M.DoublyNonnegativeTri(3), M.DoublyNonnegativeTri(10)]
    I = inst.synthetic("dnninf", 2, 0, cones, seed=14)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    prim[1] = -0.01                                   # cone 0: a negative off-diagonal entry
    prim[6:12] = [1, 3 * np.sqrt(2), 1, 0.1, 0.1, 1]  # cone 1: entrywise positive but indefinite leading 2 x 2
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuGpowGroup(cones)
    dev.load_point(prim, dual)
    assert (ora.is_feas() == np.array([False, False, True, True])).all()
    assert (dev.feas.astype(bool) == ora.is_feas()).all()

"""-m gpu parity of the batched device cone oracles against the CPU oracle (oracle/cones.py) on the
same seeded points, plus the reference's own oracle identities (test/cone.jl:23-114) evaluated
on the device results.  FP64 tolerance: 1e-11 relative per q-vector unless noted (the oracle
identities themselves hold to 1e3*eps on well-scaled points)."""
import numpy as np
import pytest

from gpu_util import rel
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M

pytestmark = pytest.mark.gpu


def _blocks(model):
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    return DeviceConeBlock(model), OracleConeBlock(model)


def _wsos(n, halfdeg, use_dual=False):
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpNonnegative(U, Ps, use_dual=use_dual)


CONE_SETS = {
    "nonneg": [M.Nonnegative(1), M.Nonnegative(6), M.Nonnegative(700)],
    "soc": [M.EpiNormEucl(2), M.EpiNormEucl(3), M.EpiNormEucl(25), M.EpiNormEucl(33), M.EpiNormEucl(70)],
    "vecmix": [M.EpiNormEucl(25), M.Nonnegative(9), M.EpiNormEucl(5), M.Nonnegative(1)],
    "psd": [M.PosSemidefTri(1), M.PosSemidefTri(3), M.PosSemidefTri(6), M.PosSemidefTri(15),
            M.PosSemidefTri(M.svec_length(40)), M.PosSemidefTri(M.svec_length(129))],
    "logdet": [M.HypoPerLogdetTri(3), M.HypoPerLogdetTri(5), M.HypoPerLogdetTri(12),
               M.HypoPerLogdetTri(2 + M.svec_length(33)), M.HypoPerLogdetTri(8, use_dual=True)],
    # BASELINE config 5 (natvsext D-optimal design, SURVEY.md 8(d) "C5a"): THE side-1000 log-det cone (dim 500502),
    # reference oracles hypoperlogdettri.jl:196-368; and two of config 4's side-100 PSD cones
    "logdet_side1000": [M.HypoPerLogdetTri(2 + M.svec_length(1000))],
    "psd_side100": [M.PosSemidefTri(M.svec_length(100)), M.PosSemidefTri(M.svec_length(100))],
    # sides above 512: the congruences run on the int8 tensor pipe (hyp_ozaki_gemm_tn; sides not multiples of 16 or 64:
    # padded digit rows, zero-filled TMA edges, grouped right-hand operand)
    "big_sides_i8": [M.PosSemidefTri(M.svec_length(520)), M.HypoPerLogdetTri(2 + M.svec_length(531)),
                     M.HypoRootdetTri(1 + M.svec_length(513))],
    "rootdet": [M.HypoRootdetTri(2), M.HypoRootdetTri(4), M.HypoRootdetTri(11),
                M.HypoRootdetTri(1 + M.svec_length(33)), M.HypoRootdetTri(7, use_dual=True)],
    "sepspec": [M.EpiPerSepSpectralMat(2 + M.svec_length(1), M.SSF_NEGLOG),
                M.EpiPerSepSpectralMat(2 + M.svec_length(3), M.SSF_NEGENTROPY),
                M.EpiPerSepSpectralMat(2 + M.svec_length(6), M.SSF_INV),
                M.EpiPerSepSpectralMat(2 + M.svec_length(12), M.SSF_POWER12, 1.5),
                M.EpiPerSepSpectralMat(2 + M.svec_length(33), M.SSF_NEGLOG),
                M.EpiPerSepSpectralMat(2 + M.svec_length(5), M.SSF_NEGENTROPY, use_dual=True),
                M.EpiPerSepSpectralMat(2 + M.svec_length(100), M.SSF_NEGENTROPY)],
    "sepspec_big": [M.EpiPerSepSpectralMat(2 + M.svec_length(130), M.SSF_NEGLOG),
                    M.EpiPerSepSpectralMat(2 + M.svec_length(4), M.SSF_POWER12, 2.0)],
    "epipersquare": [M.EpiPerSquare(3), M.EpiPerSquare(4), M.EpiPerSquare(25), M.EpiPerSquare(34),
                     M.EpiPerSquare(70)],
    "hypoperlog": [M.HypoPerLog(3), M.HypoPerLog(7), M.HypoPerLog(34), M.HypoPerLog(80),
                   M.HypoPerLog(6, use_dual=True)],
    "epinorminf": [M.EpiNormInf(2), M.EpiNormInf(6), M.EpiNormInf(34), M.EpiNormInf(70),
                   M.EpiNormInf(9, use_dual=True)],
    "sepspec_vec": [M.EpiPerSepSpectralVec(3, M.SSF_NEGLOG), M.EpiPerSepSpectralVec(8, M.SSF_NEGENTROPY),
                    M.EpiPerSepSpectralVec(35, M.SSF_INV), M.EpiPerSepSpectralVec(72, M.SSF_POWER12, 1.5),
                    M.EpiPerSepSpectralVec(7, M.SSF_NEGENTROPY, use_dual=True)],
    "hypogeomean": [M.HypoGeoMean(2), M.HypoGeoMean(6), M.HypoGeoMean(34), M.HypoGeoMean(70),
                    M.HypoGeoMean(9, use_dual=True)],
    "epirelentropy": [M.EpiRelEntropy(3), M.EpiRelEntropy(9), M.EpiRelEntropy(35), M.EpiRelEntropy(81),
                      M.EpiRelEntropy(7, use_dual=True)],
    "normspec": [M.EpiNormSpectral(1, 1), M.EpiNormSpectral(2, 2), M.EpiNormSpectral(3, 4), M.EpiNormSpectral(5, 9),
                 M.EpiNormSpectral(1, 127), M.EpiNormSpectral(11, 11), M.EpiNormSpectral(2, 3, use_dual=True),
                 M.EpiNormSpectral(4, 20, use_dual=True)],
    "wsos": [_wsos(1, 1), _wsos(1, 3), _wsos(2, 2), _wsos(3, 1), _wsos(2, 4), _wsos(1, 2, use_dual=True), _wsos(3, 2),
             _wsos(2, 6)],
    "gpow": [M.GeneralizedPower([0.5, 0.5], 1), M.GeneralizedPower([0.2, 0.3, 0.5], 2),
             M.GeneralizedPower(np.full(20, 0.05), 30), M.GeneralizedPower([0.7, 0.3], 1, use_dual=True),
             M.GeneralizedPower(np.full(4, 0.25), 60)],
    "hpm": [M.HypoPowerMean([1.0]), M.HypoPowerMean([0.4, 0.6]), M.HypoPowerMean(np.full(33, 1 / 33)),
            M.HypoPowerMean([0.3, 0.3, 0.4], use_dual=True), M.HypoPowerMean(np.full(70, 1 / 70))],
    "allmix": [M.Nonnegative(5), M.EpiNormEucl(4), M.PosSemidefTri(6), M.HypoPerLogdetTri(8),
               M.HypoRootdetTri(7), M.EpiNormEucl(3), M.HypoPerLogdetTri(5, use_dual=True),
               M.EpiPerSepSpectralMat(2 + M.svec_length(4), M.SSF_NEGENTROPY), M.EpiPerSquare(6),
               M.HypoPerLog(5), M.HypoPerLog(4, use_dual=True), M.EpiNormInf(5), M.EpiNormInf(4, use_dual=True),
               M.EpiPerSepSpectralVec(6, M.SSF_NEGENTROPY), M.HypoGeoMean(5), M.HypoGeoMean(4, use_dual=True),
               M.GeneralizedPower([0.3, 0.7], 2), M.GeneralizedPower([0.5, 0.5], 1, use_dual=True),
               M.HypoPowerMean([0.25, 0.35, 0.4]), M.HypoPowerMean([0.5, 0.5], use_dual=True),
               M.EpiRelEntropy(7), M.EpiRelEntropy(5, use_dual=True), M.EpiNormSpectral(2, 3),
               M.EpiNormSpectral(2, 2, use_dual=True), _wsos(2, 2), _wsos(1, 2, use_dual=True)],
}


def _instance(name):
    return inst.synthetic(name, 4, 0, CONE_SETS[name], seed=100 + sorted(CONE_SETS).index(name))


@pytest.mark.parametrize("name", list(CONE_SETS))
def test_cone_oracles_match_cpu_oracle(name):
    I = _instance(name)
    dev, ora = _blocks(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    dev.load_point(prim, dual, scal)
    ora.load_point(prim, dual, scal)
    assert dev.is_feas().all() and ora.is_feas().all()
    assert (dev.is_dual_feas() == ora.is_dual_feas()).all()
    g = dev.grad()
    assert rel(g, ora.grad()) <= 1e-11
    rng = np.random.default_rng(1)
    arr = rng.standard_normal((I.model.q, 3))
    # side-130 spectral cone: its eigenvalues cluster near 1 (spacing ~1e-3), and the second divided
    # differences of dder3 (matrixcsqr.jl:449-502) divide by those gaps, amplifying the O(side * eps)
    # difference between the Jacobi and LAPACK eigen-decompositions (measured 2e-10 .. 1.3e-9)
    tol = 1e-9 if name in ("logdet_side1000", "big_sides_i8") else 1e-10
    d3tol = 1e-8 if name == "sepspec_big" else tol
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= tol
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= tol
    assert rel(dev.block_hess_prod(arr[:, 0]), ora.block_hess_prod(arr[:, 0])) <= tol
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= d3tol
    irtmu = 0.9
    assert np.allclose(dev.get_proxsqr(irtmu, True), ora.get_proxsqr(irtmu, True), rtol=1e-8, atol=1e-12)
    assert np.allclose(dev.get_proxsqr(irtmu, False), ora.get_proxsqr(irtmu, False), rtol=1e-8, atol=1e-12)
    assert (dev.check_numerics(irtmu, True) == ora.check_numerics()).all()
    # identities of test/cone.jl on the device results
    pt = scal * prim
    assert rel(dev.hess_prod(pt), -g) <= 1e-10                      # cone.jl:50,78
    assert abs(float(pt @ g) + I.model.nu) <= 1e-9 * I.model.nu     # cone.jl:71
    assert rel(dev.inv_hess_prod(g), -pt) <= 1e-9                   # cone.jl:79
    assert rel(dev.inv_hess_prod(dev.hess_prod(arr)), arr) <= 1e-8  # cone.jl:81-83
    dev.free()


@pytest.mark.parametrize("name", ["nonneg", "soc", "vecmix", "psd", "epipersquare"])
def test_sqrt_oracles(name):
    I = _instance(name)
    dev, ora = _blocks(I.model)
    prim, dual = I.point.primal_dual(None)
    dev.load_point(prim, dual, 1.0)
    ora.load_point(prim, dual, 1.0)
    rng = np.random.default_rng(2)
    arr = rng.standard_normal((I.model.q, 4))
    assert rel(dev.sqrt_hess_prod(arr), ora.sqrt_hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_sqrt_hess_prod(arr), ora.inv_sqrt_hess_prod(arr)) <= 1e-10
    # cone.jl:97-102: (H^{1/2})' H^{1/2} = H and (H^{-1/2})' H^{-1/2} = H^{-1}
    s = dev.sqrt_hess_prod(arr)
    assert rel(s.T @ s, arr.T @ dev.hess_prod(arr)) <= 1e-9
    t = dev.inv_sqrt_hess_prod(arr)
    assert rel(t.T @ t, arr.T @ dev.inv_hess_prod(arr)) <= 1e-9
    dev.free()


def test_infeasible_points_are_flagged():
    cones = [M.Nonnegative(4), M.EpiNormEucl(5), M.PosSemidefTri(6), M.EpiNormEucl(3)]
    I = inst.synthetic("infeas", 3, 0, cones, seed=5)
    dev, ora = _blocks(I.model)
    prim = I.point.s.copy()
    dual = I.point.z.copy()
    prim[1] = -1.0                      # nonnegative block leaves the cone
    prim[4] = -5.0                      # SOC: u < 0
    prim[9:15] = [1, 3 * np.sqrt(2), 1, 0, 0, 1]   # PSD: indefinite leading 2x2
    dual[15] = 0.0                      # last SOC dual: u = 0
    dev.load_point(prim, dual, 1.0)
    ora.load_point(prim, dual, 1.0)
    assert (dev.is_feas() == ora.is_feas()).all()
    assert (dev.is_feas() == np.array([False, False, False, True])).all()
    assert (dev.is_dual_feas() == ora.is_dual_feas()).all()
    dev.free()

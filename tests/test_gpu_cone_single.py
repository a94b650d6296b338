"""-m gpu parity of the per-cone single-block entry points hyp_cone_* (the methods of a
`B200Cone <: Cones.Cone{Float64}`, src/Cones/Cones.jl:34-310) against the CPU oracle's per-cone objects
(oracle/cones.py make_cone) on the same seeded points, method by method with the reference's names.
FP64 tolerance 1e-10 relative per vector (the batched tests of tests/test_gpu_cones.py use the same bound)."""
import numpy as np
import pytest

from gpu_util import rel
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M

pytestmark = pytest.mark.gpu

SPECS = {
    "Nonnegative": M.Nonnegative(9),
    "EpiNormEucl": M.EpiNormEucl(25),
    "PosSemidefTri": M.PosSemidefTri(M.svec_length(12)),
    "PosSemidefTri_side100": M.PosSemidefTri(M.svec_length(100)),
    "HypoPerLogdetTri": M.HypoPerLogdetTri(2 + M.svec_length(9)),
    "HypoPerLogdetTri_dual": M.HypoPerLogdetTri(8, use_dual=True),
    "HypoRootdetTri": M.HypoRootdetTri(1 + M.svec_length(7)),
    "EpiPerSepSpectralMat": M.EpiPerSepSpectralMat(2 + M.svec_length(6), M.SSF_NEGENTROPY),
    "EpiPerSepSpectralMat_power": M.EpiPerSepSpectralMat(2 + M.svec_length(5), M.SSF_POWER12, 1.5),
    "EpiPerSquare": M.EpiPerSquare(7),
    "HypoPerLog": M.HypoPerLog(6),
    "EpiNormInf_dual": M.EpiNormInf(9, use_dual=True),
    "EpiPerSepSpectralVec": M.EpiPerSepSpectralVec(8, M.SSF_NEGLOG),
    "HypoGeoMean": M.HypoGeoMean(6),
    "GeneralizedPower": M.GeneralizedPower([0.2, 0.3, 0.5], 2),
    "HypoPowerMean": M.HypoPowerMean([0.4, 0.6]),
    "EpiRelEntropy": M.EpiRelEntropy(9),
    "EpiNormSpectral": M.EpiNormSpectral(3, 4),
}


def _pair(spec, seed):
    """(device cone, oracle cone, primal point, dual point) at a planted interior iterate of the one-cone model"""
    from hypatia_b200.cones import DeviceCone
    from oracle.cones import OracleConeBlock, make_cone
    I = inst.synthetic("single", 3, 0, [spec], seed=seed)
    prim, dual = I.point.primal_dual(OracleConeBlock(I.model).dual_mask)
    return DeviceCone(spec), make_cone(spec), prim, dual


@pytest.mark.parametrize("name", list(SPECS))
def test_single_cone_oracles_match_cpu_oracle(name):
    spec = SPECS[name]
    dev, ora, prim, dual = _pair(spec, 300 + list(SPECS).index(name))
    assert dev.dimension() == spec.dim and dev.use_dual_barrier == bool(spec.use_dual)
    assert dev.get_nu() == pytest.approx(ora.nu if not callable(ora.nu) else ora.nu(), abs=0)
    scal = 0.8
    for c in (dev, ora):
        c.load_point(prim, scal)
        c.load_dual_point(dual)
        c.reset_data()
    assert dev.is_feas() and ora.is_feas()
    assert dev.is_dual_feas() == bool(ora.is_dual_feas())
    assert rel(dev.grad(), ora.grad()) <= 1e-11
    rng = np.random.default_rng(5)
    arr = rng.standard_normal((spec.dim, 3))
    assert rel(dev.hess_prod(arr), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.inv_hess_prod(arr), ora.inv_hess_prod(arr)) <= 1e-10
    assert rel(dev.hess_prod(arr[:, 0]), ora.hess_prod(arr[:, 0])) <= 1e-10          # vector form
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10
    H, Hi = dev.hess(), dev.inv_hess()
    assert rel(H, np.asarray(ora.hess())) <= 1e-10
    assert rel(Hi, np.asarray(ora.inv_hess())) <= 1e-9
    assert rel(H @ arr, dev.hess_prod(arr)) <= 1e-10                                # explicit block = the operator
    assert dev.use_sqrt_hess_oracles(spec.dim) == bool(ora.use_sqrt_hess_oracles(spec.dim)) or \
        not dev.use_sqrt_hess_oracles(spec.dim)       # the device only claims the closed-form square roots
    if dev.use_sqrt_hess_oracles(spec.dim):
        # sqrt_hess_prod!' sqrt_hess_prod! = hess_prod! (test/cone.jl:86-96); closed forms match the oracle's
        assert rel(dev.sqrt_hess_prod(arr), ora.sqrt_hess_prod(arr)) <= 1e-10
        assert rel(dev.inv_sqrt_hess_prod(arr), ora.inv_sqrt_hess_prod(arr)) <= 1e-10
        S = dev.sqrt_hess_prod(np.eye(spec.dim))
        assert rel(S.T @ S, H) <= 1e-10
    irtmu = 0.9
    for use_max in (True, False):
        assert np.isclose(dev.get_proxsqr(irtmu, use_max), ora.get_proxsqr(irtmu, use_max), rtol=1e-8, atol=1e-12)
    assert dev.check_numerics() == bool(ora.check_numerics())
    # identities of test/cone.jl:60-84 on the device results
    nu = dev.get_nu()
    point = scal * prim
    assert abs(np.dot(point, dev.grad()) + nu) <= 1e-9 * max(nu, 1)
    assert rel(dev.hess_prod(point), -dev.grad()) <= 1e-9
    assert rel(dev.inv_hess_prod(dev.grad()), -point) <= 1e-9
    dev.free()


def test_single_cone_is_lazy_and_reloadable_like_the_reference():
    """load_point / load_dual_point only copy (Cones.jl:157-171); the next query sees the new point; an infeasible
    point answers is_feas = false without an error (SURVEY.md 8(b) error convention)."""
    spec = M.EpiNormEucl(5)
    dev, ora, prim, dual = _pair(spec, 77)
    dev.load_point(prim)                      # two-argument form: scal = 1
    dev.load_dual_point(dual)
    ora.load_point(prim)
    ora.load_dual_point(dual)
    ora.reset_data()
    g1 = dev.grad()
    assert rel(g1, ora.grad()) <= 1e-12
    dev.load_point(prim, 2.0)                 # reload: the barrier gradient is homogeneous of degree -1
    assert rel(dev.grad(), g1 / 2.0) <= 1e-12
    bad = prim.copy()
    bad[0] = -1.0
    dev.load_point(bad)
    assert dev.is_feas() is False
    dev.load_point(prim)
    assert dev.is_feas() is True and rel(dev.grad(), g1) <= 1e-15
    dev.free()


def test_single_cone_matches_the_batched_block_bit_for_bit():
    """A handle is a batch of one: same kernels, same results as the same cone inside a DeviceConeBlock."""
    from hypatia_b200.cones import DeviceCone, DeviceConeBlock
    spec = M.PosSemidefTri(M.svec_length(20))
    I = inst.synthetic("single", 3, 0, [spec], seed=9)
    blk = DeviceConeBlock(I.model)
    prim, dual = I.point.primal_dual(blk.dual_mask)
    blk.load_point(prim, dual, 0.7)
    one = DeviceCone(spec)
    one.load_point(prim, 0.7)
    one.load_dual_point(dual)
    arr = np.random.default_rng(2).standard_normal((spec.dim, 2))
    assert np.array_equal(one.grad(), blk.grad())
    assert np.array_equal(one.hess_prod(arr), blk.hess_prod(arr))
    assert np.array_equal(one.inv_hess_prod(arr), blk.inv_hess_prod(arr))
    assert np.array_equal(one.dder3(arr[:, 0]), blk.dder3(arr[:, 0]))
    one.free()
    blk.free()


def test_single_cone_refuses_bad_arguments_without_throwing():
    from hypatia_b200 import capi
    lib = capi.load_library()
    assert not lib.hyp_cone_create(0, M.CONE_EPINORMEUCL, 0, 0, 0, 0.0, None, 0)          # dim < 1
    assert not lib.hyp_cone_create(0, M.CONE_EPINORMEUCL, 5, 1, 0, 0.0, None, 0)          # no dual barrier for this type
    h = lib.hyp_cone_create(0, M.CONE_EPINORMEUCL, 5, 0, 0, 0.0, None, 0)
    assert h
    g = np.zeros(5)
    assert lib.hyp_cone_grad(h, capi.ptr(g)) < 0                                          # no point loaded
    assert b"no point loaded" in lib.hyp_cone_last_error(h)
    lib.hyp_cone_destroy(h)

"""Property-based CPU-tier checks (hypothesis) of the emulated device kernels of the vector cones: random cone
lists (types, dimensions, dual flags, spectral functions) at random interior points against the CPU oracle, plus
the size-independent identities of test/cone.jl:50-79 (H point = -grad, <point, grad> = -nu, H^-1 grad = -point)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import emu_util as eu
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from oracle.cones import OracleConeBlock


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def _cone(kind, dim, dual, hk, hp):
    if kind == "epipersquare":
        return M.EpiPerSquare(max(dim, 3))
    if kind == "hypoperlog":
        return M.HypoPerLog(max(dim, 3), use_dual=dual)
    if kind == "epinorminf":
        return M.EpiNormInf(dim, use_dual=dual)
    if kind == "hypogeomean":
        return M.HypoGeoMean(dim, use_dual=dual)
    if kind == "epirelentropy":
        return M.EpiRelEntropy(1 + 2 * max(dim // 2, 1), use_dual=dual)
    return M.EpiPerSepSpectralVec(max(dim, 3), hk, hp, use_dual=dual)


KINDS = ["epipersquare", "hypoperlog", "epinorminf", "hypogeomean", "sepspec_vec", "epirelentropy"]


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(kind=st.sampled_from(KINDS), dims=st.lists(st.integers(2, 75), min_size=1, max_size=5),
       duals=st.lists(st.booleans(), min_size=5, max_size=5), hk=st.integers(0, 3),
       hp=st.floats(1.05, 2.0), seed=st.integers(0, 10 ** 6))
def test_vec3_kernels_random_cones(kind, dims, duals, hk, hp, seed):
    cones = [_cone(kind, d, duals[i], hk, hp) for i, d in enumerate(dims)]
    I = inst.synthetic("prop", 2, 0, cones, seed=seed)
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    dev = eu.EmuVec3Group(cones)
    dev.load_point(scal * prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all()
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all()
    if not ora.is_feas().all():
        return
    g = ora.grad()
    assert rel(dev.grad, g) <= 1e-11
    arr = np.random.default_rng(seed).standard_normal((I.model.q, 2))
    assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-10
    assert rel(dev.dder3(arr[:, 0]), ora.dder3(arr[:, 0])) <= 1e-10
    pt = scal * prim
    assert rel(dev.prod(pt, 0), -dev.grad) <= 1e-10
    assert abs(float(pt @ dev.grad) + I.model.nu) <= 1e-10 * I.model.nu
    assert rel(dev.prod(dev.grad, 1), -pt) <= 1e-9


@settings(max_examples=10, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(sides=st.lists(st.integers(1, 14), min_size=1, max_size=3), ctype=st.sampled_from([2, 3, 4]),
       seed=st.integers(0, 10 ** 6))
def test_small_matrix_kernels_random_cones(sides, ctype, seed):
    mk = {2: lambda s: M.PosSemidefTri(M.svec_length(s)), 3: lambda s: M.HypoPerLogdetTri(2 + M.svec_length(s)),
          4: lambda s: M.HypoRootdetTri(1 + M.svec_length(s))}[ctype]
    cones = [mk(s) for s in sides]
    I = inst.synthetic("propmat", 2, 0, cones, seed=seed)
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(None)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    if not ora.is_feas().all():
        return
    dev = eu.EmuMatGroup(cones, ora.cones, scal * prim)
    arr = np.random.default_rng(seed).standard_normal((I.model.q, 2))
    assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-10
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-10


# ---- the cones on the generic inverse-Hessian path (csrc/cones_gpow_kernels.cuh + the batched Cholesky kernel) ----
def _genfact_cone(kind, a, b, dual, rng):
    if kind == "gpow":
        m, n = 1 + a % 6, 1 + b % 9
        n = max(n, 3 - m)             # generalizedpower.jl: dim >= 3
        al = rng.random(m) + 0.05
        return M.GeneralizedPower(al / al.sum(), n, use_dual=dual)
    if kind == "hpm":
        d = 1 + a % 12
        al = rng.random(d) + 1
        return M.HypoPowerMean(al / al.sum(), use_dual=dual)
    if kind == "normspec":
        d1 = 1 + a % 5
        d2 = d1 + b % 7
        return M.EpiNormSpectral(d1, d2, use_dual=dual)
    if kind == "epitrrelent":
        return M.EpiTrRelEntropyTri(1 + 2 * M.svec_length(1 + a % 6), use_dual=dual)
    if kind == "psdsparse":
        side = 1 + a % 20
        mask = np.tril(rng.random((side, side)) < 1 / np.sqrt(side)) | np.eye(side, dtype=bool)
        rows, cols = np.nonzero(mask)
        return M.PosSemidefTriSparse(side, rows, cols, use_dual=dual)
    if kind == "wsosone":
        from wsos_util import interpolate_box
        Rr = 2 + a % 4
        n, halfdeg = ((1, 1 + b % 3), (2, 1))[b % 2]
        U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
        return M.WSOSInterpEpiNormOne(Rr, U, Ps, use_dual=dual)
    if kind == "wsoseucl":
        from wsos_util import interpolate_box
        Rr = 2 + a % 4
        n, halfdeg = ((1, 1 + b % 3), (2, 1))[b % 2]
        U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
        return M.WSOSInterpEpiNormEucl(Rr, U, Ps, use_dual=dual)
    if kind == "wsospsd":
        from wsos_util import interpolate_box
        Rr = 1 + a % 3
        n, halfdeg = ((1, 1 + b % 3), (2, 1))[b % 2]
        U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
        return M.WSOSInterpPosSemidefTri(Rr, U, Ps, use_dual=dual)
    if kind == "meps":
        d1 = 1 + a % 4
        d2 = d1 + b % 6
        return M.MatrixEpiPerSquare(d1, d2, use_dual=dual)
    if kind == "dnn":
        return M.DoublyNonnegativeTri(M.svec_length(1 + a % 9), use_dual=dual)
    if kind == "lmi":
        side = 2 + a % 5
        dim = 2 + b % min(6, side * (side + 1) // 2 - 1)
        As = []
        for i in range(dim):
            X = rng.random((side, side))
            As.append(X @ X.T + np.eye(side) if i == 0 else (X + X.T) / 2 - 0.5)
        return M.LinMatrixIneq(As, use_dual=dual)
    from wsos_util import interpolate_box
    n, halfdeg = ((1, 1 + a % 4), (2, 1 + a % 3), (3, 1))[b % 3]
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpNonnegative(U, Ps, use_dual=dual)


GENFACT_KINDS = ["gpow", "hpm", "normspec", "dnn", "lmi", "wsos", "meps", "wsospsd", "wsoseucl", "wsosone", "psdsparse", "epitrrelent"]


@settings(max_examples=30, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(kind=st.sampled_from(GENFACT_KINDS), shapes=st.lists(st.tuples(st.integers(0, 50), st.integers(0, 50)), min_size=1,
                                                             max_size=4),
       duals=st.lists(st.booleans(), min_size=4, max_size=4), seed=st.integers(0, 10 ** 6))
def test_genfact_kernels_random_cones(kind, shapes, duals, seed):
    rng = np.random.default_rng(seed)
    cones = [_genfact_cone(kind, a, b, duals[i], rng) for i, (a, b) in enumerate(shapes)]
    I = inst.synthetic("prop", 2, 0, cones, seed=seed)
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    dev = eu.EmuGpowGroup(cones)
    dev.load_point(scal * prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all()
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all()
    if not ora.is_feas().all():
        return
    g = ora.grad()
    assert rel(dev.grad, g) <= 1e-11
    arr = rng.standard_normal((I.model.q, 2))
    assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-10
    assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-8
    assert rel(dev.dder3(arr[:, 0]), ora.dder3(arr[:, 0])) <= 1e-10
    pt = scal * prim
    assert rel(dev.prod(pt, 0), -dev.grad) <= 1e-10                    # test/cone.jl:50,78
    assert abs(float(pt @ dev.grad) + I.model.nu) <= 1e-10 * I.model.nu   # test/cone.jl:71
    assert rel(dev.prod(dev.grad, 1), -pt) <= 1e-8                      # test/cone.jl:79

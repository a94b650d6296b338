"""CPU-tier checks of the device code of the spectral cones: the CUDA kernels of
csrc/eig_kernels.cuh and csrc/cones_spec_kernels.cuh, compiled for the host by tests/emu/, against
LAPACK (numpy.linalg.eigh) and the CPU oracle (oracle/cones_sepspec.py)."""
import numpy as np
import pytest

import emu_util as eu
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from oracle.cones import OracleConeBlock


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def _pack(mats, lay):
    buf = np.zeros(lay.total)
    for c, m in enumerate(mats):
        lay.get(buf, c)[:] = m
    return buf


@pytest.mark.parametrize("smem", [True, False])
@pytest.mark.parametrize("want_vectors", [True, False])
def test_syevj_kernel_matches_lapack(smem, want_vectors):
    rng = np.random.default_rng(0)
    mats = []
    for d in (1, 2, 3, 5, 6, 12, 17):
        B = rng.standard_normal((d, d))
        mats.append(B @ B.T + 0.1 * np.eye(d))      # positive definite
        mats.append(B + B.T)                        # indefinite
    Q, _ = np.linalg.qr(rng.standard_normal((8, 8)))
    mats.append((Q * np.array([1, 1, 1, 1, 2, 2, 2, 3.0])) @ Q.T)   # repeated eigenvalues
    lay = eu.MatLayout([m.shape[0] for m in mats])
    lam, lam_off, V = eu.syevj(lay, _pack(mats, lay), want_vectors, smem=smem)
    for c, m in enumerate(mats):
        d = m.shape[0]
        l = lam[lam_off[c]:lam_off[c] + d]
        ref = np.linalg.eigvalsh(m)
        scale = np.abs(ref).max()
        assert np.abs(l - ref).max() <= 1e-13 * scale
        if want_vectors:
            Vc = lay.get(V, c)
            assert np.abs(Vc.T @ Vc - np.eye(d)).max() <= 1e-13
            assert np.abs(Vc.T @ m @ Vc - np.diag(l)).max() <= 1e-13 * scale


def test_syevj_kernel_divisor():
    rng = np.random.default_rng(1)
    B = rng.standard_normal((5, 5))
    m = B @ B.T + np.eye(5)
    lay = eu.MatLayout([5, 5])
    divv = np.array([9.0, 4.0, 0.0])
    lam, lam_off, _ = eu.syevj(lay, _pack([m, m], lay), False, divv=divv, div_off=np.array([1, 2], dtype=np.int64))
    ref = np.linalg.eigvalsh(m)
    assert np.allclose(lam[:5], ref / 4.0, rtol=1e-13)
    assert np.allclose(lam[5:], ref, rtol=1e-13)        # a divisor <= eps is replaced by 1


SPEC_SETS = {
    "neglog": [M.EpiPerSepSpectralMat(2 + M.svec_length(s), M.SSF_NEGLOG) for s in (1, 2, 3, 6)],
    "negentropy": [M.EpiPerSepSpectralMat(2 + M.svec_length(s), M.SSF_NEGENTROPY) for s in (1, 2, 5)],
    "inv": [M.EpiPerSepSpectralMat(2 + M.svec_length(s), M.SSF_INV) for s in (1, 3, 6)],
    "power": [M.EpiPerSepSpectralMat(2 + M.svec_length(s), M.SSF_POWER12, hp)
              for s, hp in ((2, 1.5), (4, 2.0), (6, 1.1))],
    "mixed": [M.EpiPerSepSpectralMat(2 + M.svec_length(7), M.SSF_NEGENTROPY),
              M.EpiPerSepSpectralMat(2 + M.svec_length(4), M.SSF_INV),
              M.EpiPerSepSpectralMat(2 + M.svec_length(9), M.SSF_POWER12, 1.7),
              M.EpiPerSepSpectralMat(2 + M.svec_length(12), M.SSF_NEGLOG)],
}


@pytest.mark.parametrize("name", list(SPEC_SETS))
def test_spec_kernels_match_oracle(name):
    cones = SPEC_SETS[name]
    I = inst.synthetic(name, 3, 0, cones, seed=300 + sorted(SPEC_SETS).index(name))
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(None)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    assert ora.is_feas().all() and ora.is_dual_feas().all()
    dev = eu.EmuSpecGroup(cones)
    dev.load_point(scal * prim, dual)
    assert dev.feas.all() and dev.dual_feas.all()
    assert rel(dev.grad, ora.grad()) <= 1e-12
    rng = np.random.default_rng(2)
    arr = rng.standard_normal((I.model.q, 3))
    assert rel(dev.prod(arr, False), ora.hess_prod(arr)) <= 1e-11
    assert rel(dev.prod(arr, True), ora.inv_hess_prod(arr)) <= 1e-11
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-11
    # the fused few-column kernel (one launch for all cones)
    assert rel(dev.small_prod(arr, 0), ora.hess_prod(arr)) <= 1e-11
    assert rel(dev.small_prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-11
    assert rel(dev.small_prod(arr, 4), ora.block_hess_prod(arr)) <= 1e-11
    assert rel(dev.small_prod(arr[:, 2], 1, in_place=True), ora.inv_hess_prod(arr[:, 2])) <= 1e-11
    # identities of test/cone.jl:50-83 on the emulated device results
    pt = scal * prim
    assert rel(dev.prod(pt, False), -dev.grad) <= 1e-11
    assert rel(dev.prod(dev.grad, True), -pt) <= 1e-10
    assert abs(float(pt @ dev.grad) + I.model.nu) <= 1e-11 * I.model.nu
    assert rel(-dev.dder3(pt), dev.grad) <= 1e-10


def test_spec_kernels_flag_infeasible_points():
    cones = [M.EpiPerSepSpectralMat(2 + M.svec_length(3), k) for k in
             (M.SSF_NEGLOG, M.SSF_NEGENTROPY, M.SSF_INV, M.SSF_NEGLOG, M.SSF_NEGENTROPY)]
    I = inst.synthetic("specinf", 2, 0, cones, seed=9)
    prim, dual = I.point.primal_dual(None)
    prim, dual = prim.copy(), dual.copy()
    o = I.model.cone_offsets
    prim[o[0]] = -50.0              # epigraph variable far too small: zeta < 0
    prim[o[1] + 1] = -1.0           # perspective variable negative
    prim[o[2] + 2] = -1.0           # W not positive definite
    dual[o[3] + 2] = -3.0           # dual W indefinite with a conjugate domain that needs W > 0
    dual[o[4]] = 0.0                # dual u = 0
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuSpecGroup(cones)
    dev.load_point(prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all()
    assert (dev.feas.astype(bool) == np.array([False, False, False, True, True])).all()
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all()
    assert not dev.dual_feas[3] and not dev.dual_feas[4]

"""CPU-tier check of the fused small-matrix kernels of the matrix cones (mat_small_prod_kernel, mat_small_dder3_kernel,
csrc/cones_mat_kernels.cuh, compiled for the host by tests/emu/) against the CPU oracle."""
import numpy as np
import pytest

import emu_util as eu
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from oracle.cones import OracleConeBlock


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


SETS = {
    "psd": [M.PosSemidefTri(M.svec_length(s)) for s in (1, 2, 3, 5, 8, 13)],
    "logdet": [M.HypoPerLogdetTri(2 + M.svec_length(s)) for s in (1, 2, 4, 7, 12)],
    "logdet_dual": [M.HypoPerLogdetTri(2 + M.svec_length(3), use_dual=True), M.HypoPerLogdetTri(2 + M.svec_length(6))],
    "rootdet": [M.HypoRootdetTri(1 + M.svec_length(s)) for s in (1, 3, 6, 11)],
    "rootdet_dual": [M.HypoRootdetTri(1 + M.svec_length(4), use_dual=True), M.HypoRootdetTri(1 + M.svec_length(5))],
}


@pytest.mark.parametrize("name", list(SETS))
def test_small_prod_kernel_matches_oracle(name):
    cones = SETS[name]
    I = inst.synthetic(name, 3, 0, cones, seed=500 + sorted(SETS).index(name))
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    assert ora.is_feas().all()
    dev = eu.EmuMatGroup(cones, ora.cones, scal * prim)
    rng = np.random.default_rng(4)
    arr = rng.standard_normal((I.model.q, 2))
    assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr, 4), ora.block_hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr[:, 0], 0, in_place=True), ora.hess_prod(arr[:, 0])) <= 1e-12
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-12
    if name == "psd":
        assert rel(dev.prod(arr, 2), ora.sqrt_hess_prod(arr)) <= 1e-12
        assert rel(dev.prod(arr, 3), ora.inv_sqrt_hess_prod(arr)) <= 1e-12


STATE_SETS = {
    "psd": [M.PosSemidefTri(M.svec_length(s)) for s in (1, 2, 3, 33, 64, 100, 128)],
    "logdet": [M.HypoPerLogdetTri(2 + M.svec_length(s)) for s in (1, 2, 4, 33, 70)] +
              [M.HypoPerLogdetTri(2 + M.svec_length(5), use_dual=True)],
    "rootdet": [M.HypoRootdetTri(1 + M.svec_length(s)) for s in (1, 3, 6, 40, 97)] +
               [M.HypoRootdetTri(1 + M.svec_length(4), use_dual=True)],
}


@pytest.mark.parametrize("name", list(STATE_SETS))
def test_device_state_pipeline_matches_oracle(name):
    """The whole state update of hyp_mat_update_state for sides <= 128 - unpack_state_kernel, the batched
    factor-and-invert kernel of chol_kernels.cuh, mat_post_kernel, and the dual-feasibility pass - run on the host, then
    the fused product kernels on THAT state (possemideftri.jl:80-107, hypoperlogdettri.jl:96-151,
    hyporootdettri.jl:100-145)."""
    cones = STATE_SETS[name]
    I = inst.synthetic(name, 3, 0, cones, seed=520 + sorted(STATE_SETS).index(name))
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    assert ora.is_feas().all()
    dev = eu.EmuMatGroup(cones, ora.cones, scal * prim)
    grad, feas, dfeas = dev.device_state(scal * prim, dual)
    # (the planted dual point of a large log-det cone may sit outside the dual cone: the flags must agree either way)
    assert feas.all() and (dfeas.astype(bool) == ora.is_dual_feas()).all()
    assert rel(grad, ora.grad()) <= 1e-12
    pt = scal * prim
    assert abs(float(pt @ grad) + I.model.nu) <= 1e-10 * I.model.nu        # test/cone.jl:71
    small = [c for c, ck in enumerate(cones) if ck.side <= 13]
    if len(small) == len(cones):
        arr = np.random.default_rng(4).standard_normal((I.model.q, 2))
        assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-11
        assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-11


def test_device_state_pipeline_flags_infeasible_points():
    cones = [M.HypoPerLogdetTri(2 + M.svec_length(3)), M.HypoPerLogdetTri(2 + M.svec_length(2)),
             M.HypoPerLogdetTri(2 + M.svec_length(4))]
    I = inst.synthetic("ldinf", 2, 0, cones, seed=9)
    ora = OracleConeBlock(I.model)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    prim[0] = 50.0                      # cone 0: u above the hypograph
    prim[8 + 1] = -1.0                  # cone 1: perspective variable v < 0
    dual[8 + 5 + 0] = 1.0               # cone 2: dual u > 0
    ora.load_point(prim, dual, 1.0)
    grp = eu.EmuMatGroup.__new__(eu.EmuMatGroup)
    grp.type, grp.K = cones[0].ctype, len(cones)
    grp.dims = np.array([c.dim for c in cones], dtype=np.int64)
    grp.off = np.concatenate(([0], np.cumsum(grp.dims)))[:-1].astype(np.int64)
    grp.q = int(grp.dims.sum())
    grp.lay = eu.MatLayout([c.side for c in cones])
    _, feas, dfeas = grp.device_state(prim, dual)
    assert (feas.astype(bool) == ora.is_feas()).all() and list(feas) == [0, 0, 1]
    assert (dfeas.astype(bool) == ora.is_dual_feas()).all() and list(dfeas) == [1, 1, 0]

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

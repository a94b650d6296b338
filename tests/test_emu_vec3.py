"""CPU-tier checks of the device code of EpiPerSquare, HypoPerLog and EpiNormInf (csrc/cones_vec3_kernels.cuh,
compiled for the host by tests/emu/) against the CPU oracle (oracle/cones_vec3.py)."""
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
    "epipersquare": [M.EpiPerSquare(d) for d in (3, 4, 6, 25, 33, 34, 70)],
    "hypoperlog": [M.HypoPerLog(d) for d in (3, 4, 7, 12, 34, 35, 80)],
    "epinorminf": [M.EpiNormInf(d) for d in (2, 3, 6, 33, 34, 70)],
    "epinorminf_dual": [M.EpiNormInf(4, use_dual=True), M.EpiNormInf(9), M.EpiNormInf(40, use_dual=True)],
    "sepspec_vec": [M.EpiPerSepSpectralVec(2 + d, hk, hp) for d, hk, hp in
                    ((1, M.SSF_NEGLOG, 0), (3, M.SSF_NEGENTROPY, 0), (6, M.SSF_INV, 0), (33, M.SSF_POWER12, 1.5),
                     (40, M.SSF_NEGLOG, 0), (70, M.SSF_NEGENTROPY, 0))],
    "sepspec_vec_dual": [M.EpiPerSepSpectralVec(6, M.SSF_NEGENTROPY, use_dual=True),
                         M.EpiPerSepSpectralVec(9, M.SSF_INV), M.EpiPerSepSpectralVec(40, M.SSF_POWER12, 2.0, use_dual=True)],
    "hypogeomean": [M.HypoGeoMean(d) for d in (2, 3, 6, 33, 34, 70)],
    "hypogeomean_dual": [M.HypoGeoMean(4, use_dual=True), M.HypoGeoMean(9), M.HypoGeoMean(40, use_dual=True)],
    "epirelentropy": [M.EpiRelEntropy(1 + 2 * d) for d in (1, 2, 4, 16, 33, 40)],
    "epirelentropy_dual": [M.EpiRelEntropy(5, use_dual=True), M.EpiRelEntropy(9), M.EpiRelEntropy(71, use_dual=True)],
    "hypoperlog_dual": [M.HypoPerLog(5, use_dual=True), M.HypoPerLog(9), M.HypoPerLog(40, use_dual=True)],
}


@pytest.mark.parametrize("name", list(SETS))
def test_vec3_kernels_match_oracle(name):
    cones = SETS[name]
    I = inst.synthetic(name, 3, 0, cones, seed=400 + sorted(SETS).index(name))
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(ora.dual_mask)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    assert ora.is_feas().all() and ora.is_dual_feas().all()
    dev = eu.EmuVec3Group(cones)
    dev.load_point(scal * prim, dual)
    assert dev.feas.all() and dev.dual_feas.all()
    assert rel(dev.grad, ora.grad()) <= 1e-13
    rng = np.random.default_rng(3)
    arr = rng.standard_normal((I.model.q, 3))
    assert rel(dev.prod(arr, 0), ora.hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr, 1), ora.inv_hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr, 1, in_place=True), ora.inv_hess_prod(arr)) <= 1e-12
    assert rel(dev.prod(arr, 4), ora.block_hess_prod(arr)) <= 1e-12
    assert rel(dev.dder3(arr[:, 1]), ora.dder3(arr[:, 1])) <= 1e-12
    if name == "epipersquare":
        assert rel(dev.prod(arr, 2), ora.sqrt_hess_prod(arr)) <= 1e-12
        assert rel(dev.prod(arr, 3), ora.inv_sqrt_hess_prod(arr)) <= 1e-12
        assert rel(dev.prod(arr, 2, in_place=True), ora.sqrt_hess_prod(arr)) <= 1e-12
    pt = scal * prim
    assert rel(dev.prod(pt, 0), -dev.grad) <= 1e-12
    assert rel(-dev.dder3(pt), dev.grad) <= 1e-11


def test_vec3_kernels_flag_infeasible_points():
    cones = [M.EpiPerSquare(5), M.EpiPerSquare(4), M.EpiPerSquare(3)]
    I = inst.synthetic("v3inf", 2, 0, cones, seed=11)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    prim[0] = -1.0             # u < 0
    prim[5 + 2] = 9.0          # 2uv < |w|^2
    dual[9 + 1] = 0.0          # dual v = 0
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuVec3Group(cones)
    dev.load_point(prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all() and not dev.feas[:2].any() and dev.feas[2]
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all() and not dev.dual_feas[2]
    cones = [M.HypoPerLog(5), M.HypoPerLog(4), M.HypoPerLog(6)]
    I = inst.synthetic("v3inf2", 2, 0, cones, seed=12)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    prim[3] = -0.5             # w_2 < 0
    prim[5] = 50.0             # u above the hypograph
    dual[9] = 0.5              # dual u > 0
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuVec3Group(cones)
    dev.load_point(prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all() and not dev.feas[:2].any() and dev.feas[2]
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all() and not dev.dual_feas[2]


def test_epinorminf_kernels_flag_infeasible_points():
    cones = [M.EpiNormInf(5), M.EpiNormInf(4), M.EpiNormInf(3)]
    I = inst.synthetic("eniinf", 2, 0, cones, seed=13)
    prim, dual = (x.copy() for x in I.point.primal_dual(None))
    prim[0] = -1.0             # u < 0
    prim[5 + 2] = 9.0          # |w|_inf > u
    dual[9 + 1] = dual[9] + 1  # |dual w|_1 > dual u
    ora = OracleConeBlock(I.model)
    ora.load_point(prim, dual, 1.0)
    dev = eu.EmuVec3Group(cones)
    dev.load_point(prim, dual)
    assert (dev.feas.astype(bool) == ora.is_feas()).all() and not dev.feas[:2].any() and dev.feas[2]
    assert (dev.dual_feas.astype(bool) == ora.is_dual_feas()).all() and not dev.dual_feas[2]

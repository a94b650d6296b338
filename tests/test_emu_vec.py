"""CPU-tier checks of the device code of the benchmarked path's two vector cones - Nonnegative and EpiNormEucl - and of
the proximity / numerics reductions (csrc/cones_vec_kernels.cuh, compiled for the host by tests/emu/) against the CPU
oracle.  Mirrors the launches of csrc/cones.cu: thread per element for the orthant, warp per (cone, column) and the
many-column chunk kernel of the Schur pre-pass for second-order cones, one CTA per cone for cone_prox_kernel."""
import numpy as np
import pytest

import emu_util as eu
from emu_util import i64, lib, p
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from oracle.cones import OracleConeBlock


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


class Table:
    def __init__(self, cones):
        self.dims = np.array([c.dim for c in cones], dtype=np.int32)
        self.off = np.concatenate(([0], np.cumsum(self.dims)))[:-1].astype(np.int64)
        self.q = int(self.dims.sum())
        self.K = len(cones)
        self.kidx = np.arange(self.K, dtype=np.int32)


def _setup(cones, seed):
    I = inst.synthetic("vec", 3, 0, cones, seed=seed)
    ora = OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(None)
    scal = 1 / np.sqrt(I.mu)
    ora.load_point(prim, dual, scal)
    return I, ora, np.ascontiguousarray(scal * prim), np.ascontiguousarray(dual)


def _cols(arr, q):
    return np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(q, -1, order="F")).copy(order="F")


def test_nonnegative_kernels_match_oracle():
    cones = [M.Nonnegative(1), M.Nonnegative(6), M.Nonnegative(130)]
    I, ora, pt, dual = _setup(cones, 1)
    T = Table(cones)
    rows = np.arange(T.q, dtype=np.int32)
    rowcone = np.repeat(T.kidx, T.dims).astype(np.int32)
    grad = np.zeros(T.q)
    feas, dfeas = np.ones(T.K, dtype=np.uint8), np.ones(T.K, dtype=np.uint8)
    lib().emu_nn_state(i64(T.q), p(rows), p(rowcone), p(pt), p(dual), p(grad), p(feas), p(dfeas))
    assert feas.all() and dfeas.all() and rel(grad, ora.grad()) <= 1e-15
    arr = _cols(np.random.default_rng(2).standard_normal((T.q, 3)), T.q)
    for mode, ref in ((0, ora.hess_prod), (1, ora.inv_hess_prod), (2, ora.sqrt_hess_prod), (3, ora.inv_sqrt_hess_prod)):
        out = np.zeros_like(arr, order="F")
        lib().emu_nn_prod(mode, i64(T.q), p(rows), p(pt), p(arr), i64(T.q), p(out), i64(T.q), i64(3), i64(0))
        assert rel(out, ref(arr)) <= 1e-14
    out = np.zeros(T.q)
    d = np.ascontiguousarray(arr[:, 0])
    lib().emu_nn_dder3(i64(T.q), p(rows), p(pt), p(d), p(out))
    assert rel(out, ora.dder3(d)) <= 1e-14
    # infeasible entries flag their own cone only (nonnegative.jl:44-60)
    bad = pt.copy()
    bad[3] = -1.0
    dbad = dual.copy()
    dbad[0] = 0.0
    feas[:], dfeas[:] = 1, 1
    lib().emu_nn_state(i64(T.q), p(rows), p(rowcone), p(bad), p(dbad), p(grad), p(feas), p(dfeas))
    assert list(feas) == [1, 0, 1] and list(dfeas) == [0, 1, 1]


SOC_DIMS = (2, 3, 25, 25, 33, 70, 25)


def _soc_state(T, pt, dual):
    grad, scal = np.zeros(T.q), np.zeros(8 * T.K)
    feas, dfeas = np.ones(T.K, dtype=np.uint8), np.ones(T.K, dtype=np.uint8)
    lib().emu_soc_state(T.K, p(T.off), p(T.dims), p(T.kidx), p(pt), p(dual), p(grad), p(scal), p(feas), p(dfeas))
    return grad, scal, feas, dfeas


def test_epinormeucl_kernels_match_oracle():
    cones = [M.EpiNormEucl(d) for d in SOC_DIMS]
    I, ora, pt, dual = _setup(cones, 2)
    T = Table(cones)
    grad, scal, feas, dfeas = _soc_state(T, pt, dual)
    assert feas.all() and dfeas.all() and rel(grad, ora.grad()) <= 1e-14
    arr = _cols(np.random.default_rng(3).standard_normal((T.q, 3)), T.q)
    refs = ((0, ora.hess_prod), (1, ora.inv_hess_prod), (2, ora.sqrt_hess_prod), (3, ora.inv_sqrt_hess_prod))
    for mode, ref in refs:
        out = np.zeros_like(arr, order="F")
        lib().emu_soc_prod(mode, T.K, p(T.off), p(T.dims), p(scal), p(pt), p(arr), i64(T.q), p(out), i64(T.q), i64(3),
                           i64(0))
        assert rel(out, ref(arr)) <= 1e-13
        inplace = arr.copy(order="F")
        lib().emu_soc_prod(mode, T.K, p(T.off), p(T.dims), p(scal), p(pt), p(inplace), i64(T.q), p(inplace), i64(T.q),
                           i64(3), i64(0))
        assert rel(inplace, ref(arr)) <= 1e-13
    # the many-column chunk kernel of the Schur pre-pass: chunks of consecutive cones staged in shared memory
    crow0 = np.array([0, T.off[3], T.off[5]], dtype=np.int64)
    ccone0 = np.array([0, 3, 5], dtype=np.int32)
    ccount = np.array([3, 2, 2], dtype=np.int32)
    crows = np.array([int(T.dims[a:a + n].sum()) for a, n in zip(ccone0, ccount)], dtype=np.int32)
    for mode, ref in refs:
        out = np.zeros_like(arr, order="F")
        lib().emu_soc_prod_chunk(mode, 3, 2 * 3000 * 8, p(crow0), p(crows), p(ccone0), p(ccount), p(T.off),
                                 p(T.dims), p(scal), p(pt), p(arr), i64(T.q), p(out), i64(T.q), i64(3), i64(0))
        assert rel(out, ref(arr)) <= 1e-13
    # the same kernel on a rank's LOCAL row panel (row sharding): the group holds the local cones only, offsets
    # and the point stay global, the panel starts at row_shift = first local row
    lo = int(T.off[3])
    off_g = np.ascontiguousarray(T.off[3:])
    dims_g = np.ascontiguousarray(T.dims[3:])
    scal_g = np.ascontiguousarray(scal[8 * 3:])
    crow0_g = np.array([T.off[3], T.off[5]], dtype=np.int64)
    ccone0_g = np.array([0, 2], dtype=np.int32)
    ccount_g = np.array([2, 2], dtype=np.int32)
    crows_g = np.array([int(T.dims[3:5].sum()), int(T.dims[5:7].sum())], dtype=np.int32)
    arr_loc = np.asfortranarray(arr[lo:])
    for mode, ref in refs:
        out = np.full_like(arr_loc, np.nan, order="F")
        lib().emu_soc_prod_chunk(mode, 2, 2 * 3000 * 8, p(crow0_g), p(crows_g), p(ccone0_g), p(ccount_g), p(off_g),
                                 p(dims_g), p(scal_g), p(pt), p(arr_loc), i64(T.q - lo), p(out), i64(T.q - lo), i64(3),
                                 i64(lo))
        assert rel(out, ref(arr)[lo:]) <= 1e-13
    out = np.zeros(T.q)
    d = np.ascontiguousarray(arr[:, 1])
    lib().emu_soc_dder3(T.K, p(T.off), p(T.dims), p(scal), p(pt), p(d), p(out))
    assert rel(out, ora.dder3(d)) <= 1e-12
    # identities of test/cone.jl:50-79 on the emulated device results
    hp = np.zeros((T.q, 1), order="F")
    lib().emu_soc_prod(0, T.K, p(T.off), p(T.dims), p(scal), p(pt), p(_cols(pt, T.q)), i64(T.q), p(hp), i64(T.q), i64(1),
                       i64(0))
    assert rel(hp[:, 0], -grad) <= 1e-13 and abs(pt @ grad + I.model.nu) <= 1e-12 * I.model.nu


def test_epinormeucl_kernels_flag_infeasible_points():
    cones = [M.EpiNormEucl(4), M.EpiNormEucl(5), M.EpiNormEucl(3)]
    I, ora, pt, dual = _setup(cones, 4)
    pt, dual = pt.copy(), dual.copy()
    pt[0] = -abs(pt[0])             # u < 0
    pt[4 + 1] = 10 * pt[4]          # |w| > u
    dual[9] = 0.0                   # dual u = 0
    ora.load_point(pt, dual, 1.0)
    T = Table(cones)
    _, _, feas, dfeas = _soc_state(T, pt, dual)
    assert (feas.astype(bool) == ora.is_feas()).all() and list(feas) == [0, 0, 1]
    assert (dfeas.astype(bool) == ora.is_dual_feas()).all() and list(dfeas) == [1, 1, 0]


@pytest.mark.parametrize("use_max", [True, False])
def test_prox_and_numerics_reductions_match_oracle(use_max):
    """cone_prox_kernel (check_numerics Cones.jl:273-290, get_proxsqr Cones.jl:294-310, nonnegative.jl:137-145) fed with
    the oracle's v1 = irtmu * dual + grad, v2 = H^-1 v1, v3 = H^-1 grad."""
    cones = [M.Nonnegative(7), M.EpiNormEucl(25), M.EpiNormEucl(3), M.Nonnegative(1), M.EpiNormEucl(130)]
    I, ora, pt, dual = _setup(cones, 5)
    T = Table(cones)
    irtmu = 0.9
    grad = ora.grad()
    v1 = irtmu * dual + grad
    v2, v3 = ora.inv_hess_prod(v1), ora.inv_hess_prod(grad)
    ctype = np.array([c.ctype for c in cones], dtype=np.int32)
    cdim = T.dims.astype(np.int64)
    cnu = np.array([c.nu for c in cones])
    prox, ok = np.zeros(T.K), np.zeros(T.K, dtype=np.uint8)
    lib().emu_cone_prox(T.K, p(ctype), p(T.off), p(cdim), p(cnu), p(pt), p(dual), p(np.ascontiguousarray(grad)),
                        p(np.ascontiguousarray(v1)), p(np.ascontiguousarray(v2)), p(np.ascontiguousarray(v3)),
                        eu.C.c_double(irtmu), int(use_max), p(prox), p(ok))
    assert np.allclose(prox, ora.get_proxsqr(irtmu, use_max), rtol=1e-10, atol=1e-14)
    assert (ok.astype(bool) == ora.check_numerics()).all()


def test_soc_chunk_kernel_several_columns_per_cta_at_c3_shape():
    """The Schur pre-pass layout of BASELINE config 3 in small: chunks of 50 x EpiNormEucl(25) (and a ragged last chunk),
    five columns dealt to gridDim.y = 2 CTAs per chunk, so every CTA reuses its staged points and per-cone constants for
    two or three columns (the emulation wrapper launches 64 threads per CTA: at most 64 cones per chunk)."""
    cones = [M.EpiNormEucl(25) for _ in range(110)] + [M.EpiNormEucl(7), M.EpiNormEucl(31)]
    I, ora, pt, dual = _setup(cones, 5)
    T = Table(cones)
    grad, scal, feas, dfeas = _soc_state(T, pt, dual)
    assert feas.all()
    starts = [0, 50, 100]
    counts = [50, 50, 12]
    crow0 = np.array([T.off[a] for a in starts], dtype=np.int64)
    ccone0 = np.array(starts, dtype=np.int32)
    ccount = np.array(counts, dtype=np.int32)
    crows = np.array([int(T.dims[a:a + n].sum()) for a, n in zip(starts, counts)], dtype=np.int32)
    assert crows.max() <= 3000
    arr = _cols(np.random.default_rng(8).standard_normal((T.q, 5)), T.q)
    refs = ((0, ora.hess_prod), (1, ora.inv_hess_prod), (2, ora.sqrt_hess_prod), (3, ora.inv_sqrt_hess_prod))
    for mode, ref in refs:
        out = np.full_like(arr, np.nan, order="F")
        lib().emu_soc_prod_chunk(mode, 3, 2 * 3000 * 8, p(crow0), p(crows), p(ccone0), p(ccount), p(T.off), p(T.dims),
                                 p(scal), p(pt), p(arr), i64(T.q), p(out), i64(T.q), i64(5), i64(0))
        assert rel(out, ref(arr)) <= 1e-13

"""-m gpu unit tests of the building-block kernels, called through the C ABI test entry points
(hyp_test_*) and checked against NumPy / LAPACK on the same seeded inputs.
Tolerances: FP64 contractions of length k are compared at 50*k*eps relative (Frobenius)."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import ctx, rel

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def cx():
    c = ctx()
    yield c
    c.close()


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("klen,ncols", [(1, 1), (7, 5), (16, 128), (100, 130), (333, 257), (2050, 400)])
@pytest.mark.parametrize("same", [True, False])
def test_atb_upper(cx, klen, ncols, same):
    rng = np.random.default_rng(klen * 1000 + ncols)
    P = np.asfortranarray(rng.standard_normal((klen, ncols)))
    R = P if same else np.asfortranarray(rng.standard_normal((klen, ncols)))
    C0 = np.asfortranarray(rng.standard_normal((ncols, ncols)))
    for alpha, beta in ((1.0, 0.0), (-1.0, 1.0)):
        Cm = C0.copy(order="F")
        rc = cx.lib.hyp_test_atb_upper(cx.h, _p(P), klen, _p(R), klen, klen, ncols, _p(Cm), ncols,
                                       alpha, beta)
        cx.check(rc, "atb_upper")
        ref = alpha * (P.T @ R) + beta * C0
        iu = np.triu_indices(ncols)
        assert rel(Cm[iu], ref[iu]) <= 50 * klen * EPS


@pytest.mark.parametrize("klen,mrows,ncols", [(3, 2, 9), (100, 100, 700), (128, 128, 1000), (50, 260, 129)])
def test_gemm_tn(cx, klen, mrows, ncols):
    rng = np.random.default_rng(klen + mrows + ncols)
    P = np.asfortranarray(rng.standard_normal((klen, mrows)))
    R = np.asfortranarray(rng.standard_normal((klen, ncols)))
    Cm = np.asfortranarray(rng.standard_normal((mrows, ncols)))
    C0 = Cm.copy()
    rc = cx.lib.hyp_test_gemm_tn(cx.h, _p(P), klen, _p(R), klen, klen, mrows, ncols, _p(Cm), mrows, 2.0, -1.0)
    cx.check(rc, "gemm_tn")
    assert rel(Cm, 2.0 * (P.T @ R) - C0) <= 50 * klen * EPS


@pytest.mark.parametrize("m", [1, 5, 128, 129, 300, 1000, 1025, 1537, 2500, 3210])
def test_potrf_potrs(cx, m):
    rng = np.random.default_rng(m)
    B = rng.standard_normal((m + 20, m))
    A = np.asfortranarray(B.T @ B + 0.5 * np.eye(m))
    F = A.copy(order="F")
    info = C.c_int(-1)
    cx.check(cx.lib.hyp_test_potrf(cx.h, _p(F), m, m, C.byref(info)), "potrf")
    assert info.value == 0
    U = np.triu(F)
    assert rel(U.T @ U, A) <= 100 * m * EPS
    Uref = np.linalg.cholesky(A).T
    assert rel(U, Uref) <= 1e-9
    b = rng.standard_normal(m)
    x = b.copy()
    cx.check(cx.lib.hyp_test_potrs(cx.h, _p(F), m, m, _p(x)), "potrs")
    xref = np.linalg.solve(A, b)
    assert rel(x, xref) <= 1e-9
    assert rel(A @ x, b) <= 1e-10 * np.linalg.cond(A)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("m", [129, 1025, 3210, 10000])
def test_packet_triangular_solves_equal_the_flag_protocol_bit_for_bit(cx, m):
    """trsv_pkt_kernel (opt-in): the solution blocks travel between CTAs as {32 bits of the double, epoch} words that the
    consumer threads poll themselves - the inter-CTA protocol the CPU emulation cannot exercise.  Same arithmetic in the
    same order as the flag protocol, so the solves must agree bit for bit; repeated so that epochs and the packet buffer
    are reused."""
    rng = np.random.default_rng(m)
    B = rng.standard_normal((m + 20, m))
    A = np.asfortranarray(B.T @ B + 0.5 * np.eye(m))
    F = A.copy(order="F")
    info = C.c_int(-1)
    cx.check(cx.lib.hyp_test_potrf(cx.h, _p(F), m, m, C.byref(info)), "potrf")
    assert info.value == 0
    b = rng.standard_normal(m)
    x_flag = b.copy()
    cx.check(cx.lib.hyp_test_potrs(cx.h, _p(F), m, m, _p(x_flag)), "potrs")
    try:
        cx.lib.hyp_test_set_trsv_pkt(1)
        for _ in range(3):
            x_pkt = b.copy()
            cx.check(cx.lib.hyp_test_potrs(cx.h, _p(F), m, m, _p(x_pkt)), "potrs")
            assert np.array_equal(x_pkt, x_flag)
    finally:
        cx.lib.hyp_test_set_trsv_pkt(0)
    assert rel(A @ x_flag, b) <= 1e-10 * np.linalg.cond(A)


def test_potrf_not_posdef(cx):
    m = 300
    rng = np.random.default_rng(3)
    B = rng.standard_normal((m, m))
    A = np.asfortranarray(B + B.T)       # indefinite
    info = C.c_int(0)
    cx.check(cx.lib.hyp_test_potrf(cx.h, _p(A), m, m, C.byref(info)), "potrf")
    assert info.value > 0


def test_potrf_repeated_factorisations_are_bit_identical(cx):
    """Steady state of the two-stream Cholesky (chol.cu, potrf_upper_i8): the first factorisation of a process builds
    the tile / pair lists with stream synchronisations that serialise the streams, later ones do not.  Every kernel is
    deterministic, so repeated factorisations of one matrix must agree bit for bit - a difference is a race between
    the chain and the bulk stream (profiles/r02_potrf_race.md: found through bench.py's parity block, invisible to
    single-shot tests)."""
    import torch
    from hypatia_b200 import capi
    m = 3000
    rng = np.random.default_rng(11)
    B = rng.standard_normal((m + 30, m))
    A = np.asfortranarray(B.T @ B + 0.5 * np.eye(m))
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()        # symmetric: row-major view = column-major matrix
    ref = None
    info = C.c_int(-1)
    for rep in range(25):
        F = dA.clone()
        torch.cuda.synchronize()
        cx.check(cx.lib.hyp_test_potrf(cx.h, capi.ptr(F), m, m, C.byref(info)), "potrf")
        cx.sync()
        assert info.value == 0
        U = torch.tril(F)
        if ref is None:
            ref = U.clone()
            Un = U.cpu().numpy().T
            assert rel(Un.T @ Un, A) <= 100 * m * EPS
        else:
            assert torch.equal(U, ref), f"factorisation {rep} differs from the first one"


def test_potrf_large_not_posdef_and_dag_agreement(cx, monkeypatch):
    """m > 1024 takes the blocked Cholesky with tcgen05 (digit-sliced) trailing updates (chol.cu, potrf_upper_i8): a
    matrix that loses definiteness deep inside the trailing part is reported, and on a definite one the factor agrees
    with the task-graph FP64-DMMA kernel (HYP_POTRF=dag) to rounding."""
    m = 2200
    rng = np.random.default_rng(5)
    B = rng.standard_normal((m + 30, m))
    A = np.asfortranarray(B.T @ B + 0.5 * np.eye(m))
    F1 = A.copy(order="F")
    info = C.c_int(-1)
    cx.check(cx.lib.hyp_test_potrf(cx.h, _p(F1), m, m, C.byref(info)), "potrf")
    assert info.value == 0
    monkeypatch.setenv("HYP_POTRF", "dag")
    F2 = A.copy(order="F")
    cx.check(cx.lib.hyp_test_potrf(cx.h, _p(F2), m, m, C.byref(info)), "potrf")
    monkeypatch.delenv("HYP_POTRF")
    assert info.value == 0
    assert rel(np.triu(F1), np.triu(F2)) <= 1e-11
    Abad = A.copy(order="F")
    Abad[1800, 1800] = -1.0
    cx.check(cx.lib.hyp_test_potrf(cx.h, _p(Abad), m, m, C.byref(info)), "potrf")
    assert info.value == 1801


@pytest.mark.parametrize("rows,cols", [(1, 1), (25, 7), (5000, 33), (4097, 300), (300, 4097), (20000, 64)])
@pytest.mark.parametrize("trans", [0, 1])
def test_gemv(cx, rows, cols, trans):
    rng = np.random.default_rng(rows + cols)
    M = np.asfortranarray(rng.standard_normal((rows, cols)))
    x = rng.standard_normal(rows if trans else cols)
    y = rng.standard_normal(cols if trans else rows)
    y0 = y.copy()
    cx.check(cx.lib.hyp_test_gemv(cx.h, trans, rows, cols, _p(M), rows, _p(x), 1.5, -0.5, _p(y)), "gemv")
    ref = 1.5 * (M.T @ x if trans else M @ x) - 0.5 * y0
    assert rel(y, ref) <= 50 * max(rows, cols) * EPS


@pytest.mark.parametrize("m", [1, 2, 7, 60, 300])
@pytest.mark.parametrize("kind", ["indefinite", "posdef", "saddle"])
def test_ldlt_rook_solve(cx, m, kind):
    """Device Bunch-Kaufman (rook) factor + solve against numpy.linalg.solve."""
    rng = np.random.default_rng(m * 7 + len(kind))
    B = rng.standard_normal((m, m))
    if kind == "indefinite":
        A = B + B.T
    elif kind == "posdef":
        A = B @ B.T + np.eye(m)
    else:   # zero diagonal block forces 2x2 pivots
        A = B + B.T
        h = m // 2
        A[:h, :h] = 0.0
    A = np.asfortranarray(A)
    Au = np.asfortranarray(np.triu(A))          # only the upper triangle is read
    b = rng.standard_normal(m)
    x = b.copy()
    info = C.c_int(-1)
    cx.check(cx.lib.hyp_test_ldlt_solve(cx.h, _p(Au), m, m, _p(x), C.byref(info)), "ldlt")
    assert info.value == 0
    xref = np.linalg.solve(A, b)
    assert rel(x, xref) <= 1e-8 * max(1.0, np.linalg.cond(A) * 1e-6)
    assert rel(A @ x, b) <= 1e-9 * np.linalg.cond(A)

"""-m gpu tests of the experimental tcgen05 (kind::i8) building blocks of the FP64-by-slicing SYRK:
the int8 TN GEMM must be bit-exact against NumPy integer arithmetic; the digit slices must
reconstruct the FP64 input to 2^-55 of the column maximum."""
import ctypes as C

import numpy as np
import pytest

from gpu_util import ctx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cx():
    c = ctx()
    yield c
    c.close()


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("K,M,N", [(128, 128, 128), (32, 128, 128), (300, 128, 128), (1000, 200, 130),
                                   (4096, 384, 256), (77, 5, 9)])
def test_i8_gemm_tn_bit_exact(cx, K, M, N):
    rng = np.random.default_rng(K + M + N)
    A = np.asfortranarray(rng.integers(-64, 65, size=(K, M), dtype=np.int8))
    B = np.asfortranarray(rng.integers(-64, 65, size=(K, N), dtype=np.int8))
    Cm = np.zeros((M, N), dtype=np.int32, order="F")
    cx.check(cx.lib.hyp_test_i8_gemm_tn(cx.h, _p(A), K, _p(B), K, K, M, N, _p(Cm), M), "i8 gemm")
    ref = A.astype(np.int64).T @ B.astype(np.int64)
    assert (Cm.astype(np.int64) == ref).all()


def test_ozaki_slices_reconstruct(cx):
    rng = np.random.default_rng(0)
    K, n, S = 500, 37, 8
    A = np.asfortranarray(rng.standard_normal((K, n)) * np.exp(rng.uniform(-8, 8, size=(1, n))))
    A[:, 3] = 0.0
    D = np.zeros((S, n, K), dtype=np.int8)      # slice-major, then column-major K x n
    e = np.zeros(n, dtype=np.int32)
    cx.check(cx.lib.hyp_test_ozaki_slices(cx.h, _p(A), K, K, n, S, _p(D), _p(e)), "slices")
    Dk = D.transpose(0, 2, 1).astype(np.float64)        # (S, K, n)
    assert np.abs(Dk).max() <= 64
    rec = np.zeros((K, n))
    for s in range(S):
        rec += Dk[s] * 2.0 ** -(6 + 7 * s)
    rec *= 2.0 ** e[None, :].astype(np.float64)
    colmax = np.abs(A).max(axis=0)
    assert (np.abs(A).max(axis=0) < 2.0 ** e.astype(np.float64) + (colmax == 0)).all()
    err = np.abs(rec - A).max(axis=0)
    assert (err <= 2.0 ** -55 * np.maximum(colmax, 1e-300) * 4).all()


def test_ozaki_radix256_slices_reconstruct(cx):
    """Seven balanced radix-256 digits (HYP_OZAKI_RADIX=256): every digit in [-128, 127] after the carry pass,
    a = 2^e sum_s 2^-(7 + 8 s) d_s to 2^-55 of the column maximum, including values that force carries."""
    rng = np.random.default_rng(1)
    K, n, S = 600, 29, 7
    A = np.asfortranarray(rng.standard_normal((K, n)) * np.exp(rng.uniform(-8, 8, size=(1, n))))
    A[:, 3] = 0.0
    A[:40, 5] = np.ldexp(1.0, -3) * (126.498046875 + np.arange(40) * 2.0 ** -20)   # remainders just below 1/2: carries
    A[:, 6] = 1.0 - 2.0 ** -52 * np.arange(K)                                       # mantissas next to 1: exponent bump
    D = np.zeros((S, n, K), dtype=np.int8)
    e = np.zeros(n, dtype=np.int32)
    cx.check(cx.lib.hyp_test_ozaki_slices(cx.h, _p(A), K, K, n, -S, _p(D), _p(e)), "slices256")
    Dk = D.transpose(0, 2, 1).astype(np.float64)
    rec = np.zeros((K, n))
    for s in range(S):
        rec += Dk[s] * 2.0 ** -(7 + 8 * s)
    rec *= 2.0 ** e[None, :].astype(np.float64)
    colmax = np.abs(A).max(axis=0)
    assert (colmax * 2.0 ** (7 - e.astype(np.float64)) <= 127).all()
    err = np.abs(rec - A).max(axis=0)
    assert (err <= 2.0 ** -56 * 2.0 ** e.astype(np.float64)).all()


@pytest.mark.parametrize("K,n", [(64, 128), (1000, 130), (5000, 300), (40000, 256), (33000, 7), (18689, 140),
                                 (50000, 129), (70001, 40)])
def test_ozaki_syrk_matches_fp64(cx, K, n):
    """Sliced int8 tcgen05 SYRK vs an extended-precision reference: the error must be at the level of a
    correctly-rounded FP64 result (a few ulp of sum |a_ki a_kj|), i.e. at least as accurate as dsyrk."""
    rng = np.random.default_rng(K + n)
    A = np.asfortranarray(rng.standard_normal((K, n)) * np.exp(rng.uniform(-3, 3, size=(1, n))))
    Cm = np.zeros((n, n), order="F")
    cx.check(cx.lib.hyp_test_ozaki_syrk(cx.h, _p(A), K, K, n, _p(Cm), n), "ozaki syrk")
    Al = A.astype(np.longdouble)
    ref = (Al.T @ Al)
    bound = (np.abs(Al).T @ np.abs(Al)).astype(np.float64)
    iu = np.triu_indices(n)
    err = np.abs(Cm[iu] - ref[iu].astype(np.float64))
    # fp64 dsyrk has error up to ~K * eps * bound; the sliced product must stay below 8 ulp of `bound`
    assert (err <= 8 * np.finfo(np.float64).eps * bound[iu]).all(), float((err / bound[iu]).max())
    ref64 = A.T @ A
    assert np.abs(Cm[iu] - ref64[iu]).max() <= 1e-12 * np.abs(ref64).max()

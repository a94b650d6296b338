"""CPU-tier checks of the rook-pivoted symmetric-indefinite factorisation and solve - the fallback of the reference's
posdef_fact_copy! chain (src/linearalgebra/dense.jl:164-215: bunchkaufman!(A, true) = LAPACK dsytrf_rook / dsytrs_rook,
increase_diag!) - with the device kernels of csrc/ldlt_kernels.cuh compiled for the host by tests/emu/ and driven by the
same launch sequence as ldlt.cu."""
import numpy as np
import pytest

from emu_util import i64, lib, p

EPS = np.finfo(np.float64).eps


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _matrix(kind, m, rng):
    B = rng.standard_normal((m, m))
    if kind == "indefinite":
        return B + B.T
    if kind == "posdef":
        return B @ B.T + np.eye(m)
    A = B + B.T                      # "saddle": a zero diagonal block forces 2 x 2 pivots
    A[:m // 2, :m // 2] = 0.0
    return A


def _factor(A):
    m = A.shape[0]
    F = np.asfortranarray(np.triu(A))          # only the upper triangle is read
    ipiv = np.zeros(3 * m, dtype=np.int32)
    info = lib().emu_ldlt_factor(p(F), i64(m), i64(m), p(ipiv))
    return F, ipiv, info


@pytest.mark.parametrize("m", [1, 2, 7, 19])
@pytest.mark.parametrize("kind", ["indefinite", "posdef", "saddle"])
def test_rook_factor_and_solve(m, kind):
    rng = np.random.default_rng(m * 7 + len(kind))
    A = _matrix(kind, m, rng)
    F, ipiv, info = _factor(A)
    assert info == 0
    b = rng.standard_normal(m)
    x = b.copy()
    lib().emu_ldlt_solve(p(F), i64(m), i64(m), p(ipiv), p(x))
    assert rel(A @ x, b) <= 1e-10 * np.linalg.cond(A)
    assert rel(x, np.linalg.solve(A, b)) <= 1e-9 * max(1.0, np.linalg.cond(A) * 1e-6)
    # pivot structure: 1 x 1 steps and 2 x 2 steps tile the columns; posdef matrices never need a 2 x 2 pivot
    steps = ipiv[0::3]
    k, n2 = 0, 0
    while k < m:
        assert steps[k] in (1, 2)
        if steps[k] == 2:
            assert steps[k + 1] == 0
            n2 += 1
        k += steps[k]
    assert k == m
    if kind == "posdef":
        assert n2 == 0
    # rook pivoting bounds the multipliers: |L_ij| <= 1 / (1 - alpha) with alpha = (1 + sqrt 17) / 8
    L = np.tril(F, -1)
    k = 0
    while k < m:                      # the sub-diagonal entry of a 2 x 2 block belongs to D, not to L
        if steps[k] == 2:
            L[k + 1, k] = 0.0
        k += steps[k]
    assert np.abs(L).max(initial=0.0) <= 1.0 / (1.0 - (1 + np.sqrt(17.0)) / 8) + 1e-12


def test_inertia_matches_the_eigenvalues():
    """Sylvester: the inertia of D (1 x 1 and 2 x 2 blocks) equals the inertia of A."""
    rng = np.random.default_rng(5)
    m = 16
    A = _matrix("saddle", m, rng)
    F, ipiv, info = _factor(A)
    assert info == 0
    steps = ipiv[0::3]
    pos = neg = 0
    k = 0
    while k < m:
        if steps[k] == 1:
            pos, neg = pos + (F[k, k] > 0), neg + (F[k, k] < 0)
        else:
            ev = np.linalg.eigvalsh(np.array([[F[k, k], F[k + 1, k]], [F[k + 1, k], F[k + 1, k + 1]]]))
            pos, neg = pos + int((ev > 0).sum()), neg + int((ev < 0).sum())
        k += steps[k]
    w = np.linalg.eigvalsh(A)
    assert (pos, neg) == (int((w > 0).sum()), int((w < 0).sum()))


def test_exactly_singular_column_is_reported():
    A = np.diag([2.0, 0.0, 3.0])      # dsytf2_rook: info = index of the zero pivot, the factorisation continues
    F, ipiv, info = _factor(A)
    assert info == 2


def test_increase_diag():
    # dense.jl:106-113: A_jj = (1 + 1e-5) max(A_jj, 1000 eps)
    A = np.asfortranarray(np.diag([2.0, -1.0, 0.0, 1e-20]) + np.triu(np.ones((4, 4)), 1))
    A0 = A.copy()
    lib().emu_increase_diag(p(A), i64(4), i64(4))
    want = (1 + 1e-5) * np.maximum(np.diag(A0), 1000 * EPS)
    assert np.allclose(np.diag(A), want, rtol=1e-15) and np.array_equal(np.triu(A, 1), np.triu(A0, 1))

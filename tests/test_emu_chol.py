"""CPU-tier checks of the factor-and-invert kernel of the blocked Cholesky (csrc/chol_kernels.cuh, compiled for the host
by tests/emu/): the 128 x 128 diagonal-block panel of hyp_potrf_upper (dpotrf, dense.jl:191-192), the inversion of the
diagonal blocks of a given factor (hyp_trtri_diag) and the batched per-cone cholesky! + inv_fact! of the matrix cones
and of the generic Hessian factorisation (Cones.jl:239-259), against LAPACK."""
import numpy as np
import pytest

import emu_util as eu
from emu_util import i64, lib, p

NB = 128


def _spd(rng, n, cond=1e3):
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return (Q * np.logspace(0, -np.log10(cond), n)) @ Q.T


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("m,blk", [(128, 0), (200, 1), (5, 0), (33, 0), (97, 0), (300, 2)])
def test_panel_kernel_factors_and_inverts_a_diagonal_block(m, blk):
    rng = np.random.default_rng(m + blk)
    S = _spd(rng, m)
    A = np.asfortranarray(np.triu(S) + np.tril(rng.standard_normal((m, m)), -1))   # the lower triangle is never read
    A0 = A.copy(order="F")
    nblk = (m + NB - 1) // NB
    dinv = np.zeros(nblk * NB * NB)
    info = np.zeros(1, dtype=np.int32)
    lib().emu_panel_factor(p(A), i64(m), i64(m), i64(blk), p(dinv), p(info))
    k0 = blk * NB
    nb = min(NB, m - k0)
    U = np.linalg.cholesky(S[k0:k0 + nb, k0:k0 + nb]).T
    assert info[0] == 0
    assert rel(np.triu(A[k0:k0 + nb, k0:k0 + nb]), U) <= 1e-13
    X = dinv[blk * NB * NB:(blk + 1) * NB * NB].reshape(NB, NB, order="F")
    assert rel(X[:nb, :nb], np.linalg.inv(U)) <= 1e-11
    assert np.array_equal(X[nb:, nb:], np.eye(NB - nb)) and not X[:nb, nb:].any() and not X[nb:, :nb].any()
    # nothing outside the diagonal block (and nothing below its diagonal) is written
    mask = np.ones((m, m), dtype=bool)
    mask[k0:k0 + nb, k0:k0 + nb] = np.tril(np.ones((nb, nb), dtype=bool), -1)
    assert np.array_equal(A[mask], A0[mask])


def test_panel_kernel_reports_the_first_bad_pivot():
    rng = np.random.default_rng(3)
    m = 60
    S = _spd(rng, m)
    S[40, 40] = -1.0                     # leading 40 x 40 minor positive definite, the 41st pivot is not
    A = np.asfortranarray(np.triu(S))
    dinv = np.zeros(NB * NB)
    info = np.zeros(1, dtype=np.int32)
    lib().emu_panel_factor(p(A), i64(m), i64(m), i64(0), p(dinv), p(info))
    assert info[0] == 41                 # LAPACK dpotrf convention (1-based order of the failing minor)


@pytest.mark.parametrize("m", [1, 64, 128, 129, 300])
def test_panel_kernel_inverts_the_diagonal_blocks_of_a_factor(m):
    rng = np.random.default_rng(m)
    U = np.asfortranarray(np.linalg.cholesky(_spd(rng, m)).T)
    nblk = (m + NB - 1) // NB
    dinv = np.zeros(nblk * NB * NB)
    lib().emu_panel_invert(p(U), i64(m), i64(m), p(dinv))
    for b in range(nblk):
        k0, nb = b * NB, min(NB, m - b * NB)
        X = dinv[b * NB * NB:(b + 1) * NB * NB].reshape(NB, NB, order="F")
        assert rel(X[:nb, :nb], np.linalg.inv(U[k0:k0 + nb, k0:k0 + nb])) <= 1e-11


def test_batched_cholesky_of_cone_groups():
    rng = np.random.default_rng(5)
    sides = np.array([1, 2, 7, 32, 33, 100, 128, 15], dtype=np.int32)
    lay = eu.MatLayout(sides)
    U, Ui = np.zeros(lay.total), np.zeros(lay.total)
    mats = []
    for c, sd in enumerate(sides):
        S = _spd(rng, int(sd), cond=1e2)
        if c == 3:
            S[10, 10] = -0.5             # cone 3 is not positive definite
        mats.append(S)
        lay.get(U, c)[:] = S
    kidx = np.arange(len(sides), dtype=np.int32)
    flag = np.ones(len(sides), dtype=np.uint8)
    lib().emu_chol_batched(len(sides), p(sides), p(lay.moff), p(kidx), p(U), p(Ui), p(flag))
    assert list(flag) == [1, 1, 1, 0, 1, 1, 1, 1]
    for c, S in enumerate(mats):
        if c == 3:
            continue
        R = np.linalg.cholesky(S).T
        got = lay.get(U, c)
        assert rel(np.triu(got), R) <= 1e-13 and not np.tril(got, -1).any()      # zero_lower
        assert rel(lay.get(Ui, c), np.linalg.inv(R)) <= 1e-11


@pytest.mark.parametrize("m", [1, 100, 128, 129, 300, 390])
def test_blocked_triangular_solves_give_potrs(m):
    """dpotrs = two sweeps of trsv_kernel over 128-blocks with the inverted diagonal blocks of the factor
    (ldiv!(x, fact, rhs), qrchol.jl:68).  In the emulation the first CTA takes every ticket, so this checks the
    arithmetic of the sweeps, not the inter-CTA flag protocol (that is covered by the -m gpu tier)."""
    rng = np.random.default_rng(m)
    S = _spd(rng, m, cond=1e2)
    U = np.asfortranarray(np.linalg.cholesky(S).T)
    nblk = (m + NB - 1) // NB
    dinv = np.zeros(nblk * NB * NB)
    lib().emu_panel_invert(p(U), i64(m), i64(m), p(dinv))
    b = rng.standard_normal(m)
    y = b.copy()
    lib().emu_trsv_upper(p(U), i64(m), i64(m), p(dinv), p(y), 1)          # y = U^-T b
    assert rel(y, np.linalg.solve(U.T, b)) <= 1e-11
    x = y.copy()
    lib().emu_trsv_upper(p(U), i64(m), i64(m), p(dinv), p(x), 0)          # x = U^-1 y
    assert rel(x, np.linalg.solve(S, b)) <= 1e-10
    assert rel(S @ x, b) <= 1e-11


@pytest.mark.parametrize("m", [1, 100, 128, 129, 300, 390])
@pytest.mark.parametrize("nrhs", [1, 2])
def test_packet_triangular_solves_equal_the_flag_protocol_bit_for_bit(m, nrhs):
    """trsv_pkt_kernel (default): the solution blocks travel as {32 bits of the double, epoch} words that the consumer
    threads poll themselves.  Same arithmetic in the same order as trsv_kernel => identical bits; the packet buffer is
    reused by both sweeps (epochs 1, 2), as in the library; entries past m of the last block are published as zeros."""
    import ctypes as C
    rng = np.random.default_rng(7 * m + nrhs)
    S = _spd(rng, m, cond=1e2)
    U = np.asfortranarray(np.linalg.cholesky(S).T)
    nblk = (m + NB - 1) // NB
    dinv = np.zeros(nblk * NB * NB)
    lib().emu_panel_invert(p(U), i64(m), i64(m), p(dinv))
    stride = m + 3
    b = rng.standard_normal((nrhs, stride))
    ref = b.copy()
    for v in range(nrhs):
        col = np.ascontiguousarray(ref[v, :m])
        lib().emu_trsv_upper(p(U), i64(m), i64(m), p(dinv), p(col), 1)
        lib().emu_trsv_upper(p(U), i64(m), i64(m), p(dinv), p(col), 0)
        ref[v, :m] = col
    x = np.ascontiguousarray(b.copy())
    pkt = np.zeros(nrhs * nblk * NB * 2, dtype=np.uint64)
    lib().emu_trsv_upper_pkt(p(U), i64(m), i64(m), p(dinv), p(x), i64(stride), nrhs, 1, p(pkt), 1)
    lib().emu_trsv_upper_pkt(p(U), i64(m), i64(m), p(dinv), p(x), i64(stride), nrhs, 0, p(pkt), 2)
    assert np.array_equal(x, ref)                     # the padding entries past m are untouched as well
    assert rel(S @ x[0, :m], b[0, :m]) <= 1e-11
    assert ((pkt >> np.uint64(32)) == 2).all()        # every entry of every block was published in the second sweep


@pytest.mark.parametrize("m,seg", [(1, 8), (129, 1), (390, 1), (390, 2), (700, 2), (700, 8), (641, 3)])
def test_segmented_triangular_solves_give_potrs(m, seg):
    """The default sweeps (trsv_seg_kernel): block columns cut into runs of `seg` tiles, partial sums through global
    scratch (NaN-filled here), the final run of a column adds them up.  Ticket order = trsv_build_tasks (the same
    builder the library uses); the first emulated CTA takes every ticket in that order, so a ticket whose inputs came
    later in the list would spin forever - the test also checks the order is topological."""
    rng = np.random.default_rng(m + seg)
    S = _spd(rng, m, cond=1e2)
    U = np.asfortranarray(np.linalg.cholesky(S).T)
    nblk = (m + NB - 1) // NB
    dinv = np.zeros(nblk * NB * NB)
    lib().emu_panel_invert(p(U), i64(m), i64(m), p(dinv))
    b = rng.standard_normal(m)
    y = b.copy()
    lib().emu_trsv_upper_seg(p(U), i64(m), i64(m), p(dinv), p(y), 1, seg)
    assert rel(y, np.linalg.solve(U.T, b)) <= 1e-11
    x = y.copy()
    lib().emu_trsv_upper_seg(p(U), i64(m), i64(m), p(dinv), p(x), 0, seg)
    assert rel(x, np.linalg.solve(S, b)) <= 1e-10
    assert rel(S @ x, b) <= 1e-11

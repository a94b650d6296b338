"""CPU-tier checks of the digit slicing behind the FP64-accurate tcgen05 SYRK (csrc/ozaki_slice_kernels.cuh, compiled
for the host by tests/emu/): every step of the decomposition is exact, so it is checked in exact integer / rational
arithmetic - digit ranges, reconstruction to the last kept bit, and the recombination formula of the SYRK epilogue
C_ij = 2^(e_i + e_j) sum_d 2^-(2 w0 + w d) sum_{s + t = d} (D_s' D_t)_ij  (ozaki.cu; w = 8, w0 = 7 for radix 256 and
w = 7, w0 = 6 for radix 128) against the exact Gram matrix."""
from fractions import Fraction

import numpy as np
import pytest

from emu_util import i64, lib, p


def slice_emu(A, radix, nslices):
    K, n = A.shape
    A = np.asfortranarray(A, dtype=np.float64)
    ldd = ((max(K, 16) + 15) // 16) * 16
    D = np.zeros((nslices, n, ldd), dtype=np.int8)          # D[s][j][k] = digit s of A[k, j]
    expo = np.zeros(n, dtype=np.int32)
    dscale = np.zeros(n)
    lib().emu_ozaki_slice(radix, i64(K), i64(n), p(A), i64(K), nslices, p(expo), p(dscale), p(D), i64(ldd), i64(ldd * n))
    return D, expo, dscale, ldd


def _matrix(rng, K, n, spread):
    A = rng.standard_normal((K, n)) * np.exp(spread * rng.standard_normal((K, n)))
    A[:, 0] *= 1e-7
    A[:, -1] *= 3e5
    if n > 2:
        A[:, 1] = 0.0                                         # an all-zero column
    A[0, -1] = 0.99999 * 2.0 ** np.ceil(np.log2(np.abs(A[:, -1]).max()))   # just below a power of two: the carry case
    return A


@pytest.mark.parametrize("radix,nslices,w,w0", [(256, 7, 8, 7), (128, 8, 7, 6)])
@pytest.mark.parametrize("K", [1, 7, 64, 203])
def test_digits_are_in_range_and_reconstruct_exactly(radix, nslices, w, w0, K):
    rng = np.random.default_rng(K + radix)
    A = _matrix(rng, K, 5, 2.0)
    D, expo, dscale, ldd = slice_emu(A, radix, nslices)
    lim = 128 if radix == 256 else 64
    Di = D.astype(np.int64)
    assert Di.min() >= -lim and Di.max() <= (127 if radix == 256 else 64)
    assert np.abs(Di[0]).max() <= (127 if radix == 256 else 64)
    assert (Di[:, :, K:] == 0).all()                          # padding rows carry zero digits
    for j in range(A.shape[1]):
        mx = np.abs(A[:, j]).max()
        e = int(expo[j])
        assert dscale[j] == 2.0 ** e
        if mx == 0:
            assert e == 0 and (Di[:, j] == 0).all()
            continue
        assert mx < 2.0 ** e and mx >= 2.0 ** (e - 2)         # smallest admissible exponent (+1 in the carry case)
        if radix == 256:
            assert mx <= 127.0 / 128.0 * 2.0 ** e
        for k in range(K):
            recon = sum(Fraction(int(Di[s, j, k])) * Fraction(2) ** (e - w0 - w * s) for s in range(nslices))
            err = abs(Fraction(float(A[k, j])) - recon)
            assert err <= Fraction(2) ** (e - w0 - w * (nslices - 1) - 1)      # half a unit of the last digit kept


@pytest.mark.parametrize("radix,nslices,w,w0", [(256, 7, 8, 7), (128, 8, 7, 6)])
def test_recombined_digit_products_give_the_gram_matrix(radix, nslices, w, w0):
    rng = np.random.default_rng(9)
    K, n = 70, 5
    A = _matrix(rng, K, n, 1.0)
    D, expo, _, _ = slice_emu(A, radix, nslices)
    Di = [[[int(x) for x in D[s, j, :K]] for j in range(n)] for s in range(nslices)]
    exact = [[sum(Fraction(float(A[k, i])) * Fraction(float(A[k, j])) for k in range(K)) for j in range(n)] for i in range(n)]
    absum = np.abs(A).T @ np.abs(A)
    for i in range(n):
        for j in range(i, n):
            acc = Fraction(0)
            for d in range(nslices):                          # digit-sum groups d = s + t <= nslices - 1
                g = 0                                         # the exact int32 accumulator of group d
                for s in range(d + 1):
                    t = d - s
                    g += sum(a * b for a, b in zip(Di[s][i], Di[t][j]))
                assert abs(g) < 2 ** 31
                acc += Fraction(g) * Fraction(2) ** (-(2 * w0 + w * d))
            c = acc * Fraction(2) ** (int(expo[i]) + int(expo[j]))
            err = abs(c - exact[i][j])
            # dropped pairs (s + t >= nslices) and the last-digit rounding: a few units of 2^-(w0 + w (nslices - 1)) per factor
            bound = Fraction(2) ** (int(expo[i]) + int(expo[j])) * K * Fraction(2) ** (-(w0 + w * (nslices - 1)) + 2)
            assert err <= bound
            if absum[i, j] > 0 and i != 1 and j != 1:
                col_ratio = float(2.0 ** (int(expo[i]) + int(expo[j])) * K / absum[i, j])
                assert float(err) <= 2.0 ** -50 * col_ratio * absum[i, j]


@pytest.mark.parametrize("nslices", [1, 3, 5, 7, 8])
def test_radix256_digits_match_the_golden_digits_of_the_top_down_formulation(nslices):
    """tests/golden/slice256_digits.npz was written by the `rint`-per-digit formulation (make_slice256_golden.py); the
    byte-parallel formulation of slice256_pack8 must give the same digits and exponents bit for bit - including the
    rounding ties at every digit level, the +128 -> -128 carry chains and the range ends the fixture contains."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slice256_digits.npz"))
    D, expo, _, _ = slice_emu(g["A"], 256, nslices)
    assert np.array_equal(expo, g[f"expo{nslices}"])
    assert np.array_equal(D, g[f"D{nslices}"])

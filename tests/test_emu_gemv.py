"""CPU-tier check of the fused one-pass kernel for w = G x and y = G' z (csrc/gemv_kernels.cuh, compiled for the
host by tests/emu/) against NumPy, on ragged shapes (odd row counts, column counts that are not multiples of 8)."""
import ctypes as C

import numpy as np
import pytest

import emu_util as eu


@pytest.mark.parametrize("rows,ncols,nchunks", [(1, 1, 1), (2, 9, 1), (255, 8, 1), (257, 17, 2), (600, 40, 3),
                                                (1000, 7, 1)])
def test_gemv_nt_kernel(rows, ncols, nchunks):
    rng = np.random.default_rng(rows + ncols)
    ld = rows + (rows & 1)
    Mfull = np.zeros((ld, ncols), order="F")
    Mfull[:rows] = rng.standard_normal((rows, ncols))
    x = rng.standard_normal(ncols)
    z = rng.standard_normal(rows)
    w = rng.standard_normal(rows)
    y = rng.standard_normal(ncols)
    w0, y0 = w.copy(), y.copy()
    p = eu.p
    eu.lib().emu_gemv_nt(eu.i64(rows), eu.i64(ncols), p(Mfull), eu.i64(ld), p(x), p(z), nchunks,
                         C.c_double(-1.0), C.c_double(0.5), p(w), C.c_double(2.0), C.c_double(1.0), p(y))
    G = Mfull[:rows]
    assert np.allclose(w, -1.0 * (G @ x) + 0.5 * w0, rtol=1e-13, atol=1e-13)
    assert np.allclose(y, 2.0 * (G.T @ z) + y0, rtol=1e-13, atol=1e-13)

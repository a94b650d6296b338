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


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("rows,ncols", [(1, 1), (2, 5), (33, 9), (2049, 7), (2050, 4), (5000, 3)])
def test_gemv_t_kernels(kind, rows, ncols):
    """y = alpha G' x + beta y: the CTA-per-column kernel (vectorised and scalar loads) and the warp-per-column kernel
    of hyp_gemv_t (gemv.cu): mul!(.., G', z) of qrchol.jl:52 / common.jl:91."""
    rng = np.random.default_rng(rows + 7 * ncols)
    ld = rows + (rows & 1)
    Mfull = np.zeros((ld, ncols), order="F")
    Mfull[:rows] = rng.standard_normal((rows, ncols))
    x = rng.standard_normal(rows + 1)[:rows].copy()
    y = rng.standard_normal(ncols)
    y0 = y.copy()
    p = eu.p
    eu.lib().emu_gemv_t(kind, eu.i64(rows), eu.i64(ncols), p(Mfull), eu.i64(ld), p(x), C.c_double(-1.5), C.c_double(0.25),
                        p(y))
    assert np.allclose(y, -1.5 * (Mfull[:rows].T @ x) + 0.25 * y0, rtol=1e-12, atol=1e-12)
    # beta = 0 must not read y (NaN there stays out of the result)
    y[:] = np.nan
    eu.lib().emu_gemv_t(kind, eu.i64(rows), eu.i64(ncols), p(Mfull), eu.i64(ld), p(x), C.c_double(1.0), C.c_double(0.0),
                        p(y))
    assert np.allclose(y, Mfull[:rows].T @ x, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("vec", [1, 0])
@pytest.mark.parametrize("rows,ncols,nchunks", [(1, 1, 1), (2, 9, 1), (255, 8, 2), (257, 17, 3), (600, 40, 4)])
def test_gemv_n_kernels(vec, rows, ncols, nchunks):
    """y = alpha G x + beta y through the column-chunked partial sums and the fixed-order reduce of hyp_gemv_n
    (gemv.cu): mul!(Gx, G, x) of qrchol.jl:73 / common.jl:94,144."""
    rng = np.random.default_rng(rows + 3 * ncols)
    ld = rows + (rows & 1)
    Mfull = np.zeros((ld, ncols), order="F")
    Mfull[:rows] = rng.standard_normal((rows, ncols))
    x = rng.standard_normal(ncols)
    y = rng.standard_normal(rows)
    y0 = y.copy()
    p = eu.p
    eu.lib().emu_gemv_n(vec, eu.i64(rows), eu.i64(ncols), p(Mfull), eu.i64(ld), p(x), nchunks, C.c_double(2.0),
                        C.c_double(-0.5), p(y))
    assert np.allclose(y, 2.0 * (Mfull[:rows] @ x) - 0.5 * y0, rtol=1e-12, atol=1e-12)

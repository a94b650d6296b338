"""Dense factorisation helpers of the oracle (test infrastructure).

Restates src/linearalgebra/dense.jl with the same LAPACK routines Julia's stdlib calls,
taken from SciPy's bundled OpenBLAS: dpotrf / dpotrs / dpotri through scipy.linalg.lapack,
dsytrf_rook / dsytrs_rook / dsytri_rook through ctypes (SciPy does not wrap the rook
variants).  reference: dense.jl:15-65 (inv_fact!), :106-113 (increase_diag!), :164-184
(symm_fact!, symm_fact_copy!), :191-215 (posdef_fact!, posdef_fact_copy!).
"""
import ctypes
import glob
import os

import numpy as np
import scipy
from scipy.linalg import lapack as _lp

_EPS = np.finfo(np.float64).eps


def _load_openblas():
    libs = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                  "libscipy_openblas*.so"))
    for f in libs:
        lib = ctypes.CDLL(f)
        if hasattr(lib, "scipy_dsytrf_rook_"):
            return lib
    raise ImportError("scipy's bundled OpenBLAS with dsytrf_rook not found")


_OB = _load_openblas()
_i = ctypes.c_int
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def openblas_threads():
    return int(_OB.scipy_openblas_get_num_threads())


def _ptr(a, t=_dp):
    return a.ctypes.data_as(t)


class Cholesky:
    """Upper Cholesky factor A = U'U (LAPACK dpotrf 'U'), like Julia's cholesky!(Symmetric(A,:U))."""

    kind = 0

    def __init__(self, factors, info):
        self.factors = factors      # upper triangle holds U (lower triangle is junk)
        self.info = int(info)

    def issuccess(self):
        return self.info == 0

    @property
    def U(self):
        return np.triu(self.factors)

    def solve(self, rhs):
        x, info = _lp.dpotrs(self.factors, rhs, lower=0)
        assert info == 0
        return x

    def logdet(self):
        return 2.0 * np.log(np.diag(self.factors)).sum()

    def inverse(self):
        """potri: full symmetric inverse (reference: dense.jl:15-22 inv_fact!)."""
        inv, info = _lp.dpotri(self.factors, lower=0)
        assert info == 0
        iu = np.triu(inv)
        return iu + np.triu(inv, 1).T


class BunchKaufman:
    """Rook-pivoted LDL' (LAPACK dsytrf_rook), like Julia's bunchkaufman!(A, true)."""

    def __init__(self, LD, ipiv, info, uplo, kind=1):
        self.LD, self.ipiv, self.info, self.uplo, self.kind = LD, ipiv, int(info), uplo, kind

    def issuccess(self):
        return self.info == 0

    def solve(self, rhs):
        b = np.array(rhs, dtype=np.float64, order="F", copy=True)
        b2 = b.reshape(b.shape[0], -1, order="F")
        n, nrhs = b2.shape
        info = _i(0)
        _OB.scipy_dsytrs_rook_(ctypes.c_char_p(self.uplo), ctypes.byref(_i(n)), ctypes.byref(_i(nrhs)),
                               _ptr(self.LD), ctypes.byref(_i(n)), _ptr(self.ipiv, _ip), _ptr(b2),
                               ctypes.byref(_i(n)), ctypes.byref(info), 1)
        assert info.value == 0
        return b2.reshape(b.shape, order="F")


def posdef_fact(mat):
    """cholesky!(Symmetric(mat, :U), check=false) on a copy (dense.jl:191-192)."""
    c, info = _lp.dpotrf(mat, lower=0, clean=0, overwrite_a=0)
    return Cholesky(c, info)


def symm_fact(mat, uplo=b"U", kind=1):
    """bunchkaufman!(Symmetric(mat, uplo), rook=true, check=false) on a copy (dense.jl:164-165)."""
    a = np.array(mat, dtype=np.float64, order="F", copy=True)
    n = a.shape[0]
    ipiv = np.zeros(max(n, 1), dtype=np.int32)
    info = _i(0)
    lwork = _i(-1)
    wq = np.zeros(1)
    _OB.scipy_dsytrf_rook_(ctypes.c_char_p(uplo), ctypes.byref(_i(n)), _ptr(a), ctypes.byref(_i(max(n, 1))),
                           _ptr(ipiv, _ip), _ptr(wq), ctypes.byref(lwork), ctypes.byref(info), 1)
    lw = max(int(wq[0]), 1)
    work = np.zeros(lw)
    _OB.scipy_dsytrf_rook_(ctypes.c_char_p(uplo), ctypes.byref(_i(n)), _ptr(a), ctypes.byref(_i(max(n, 1))),
                           _ptr(ipiv, _ip), _ptr(work), ctypes.byref(_i(lw)), ctypes.byref(info), 1)
    return BunchKaufman(a, ipiv, info.value, uplo, kind)


def increase_diag(mat):
    """dense.jl:106-113"""
    d = np.diag(mat).copy()
    np.fill_diagonal(mat, (1 + 1e-5) * np.maximum(d, 1000 * _EPS))
    return mat


def symm_fact_copy(mat, uplo=b"U"):
    """dense.jl:170-184"""
    fact = symm_fact(mat, uplo)
    if not fact.issuccess():
        m2 = increase_diag(np.array(mat, order="F", copy=True))
        fact = symm_fact(m2, uplo, kind=2)
    return fact


def posdef_fact_copy(mat, try_shift=True):
    """Cholesky -> Bunch-Kaufman -> shifted Bunch-Kaufman chain (dense.jl:194-215).
    Returns an object with issuccess()/solve(); .kind = 0 chol, 1 BK, 2 shifted BK."""
    fact = posdef_fact(mat)
    if not fact.issuccess():
        fact = symm_fact(mat, b"U", kind=1)
        if try_shift and not fact.issuccess():
            m2 = increase_diag(np.array(mat, order="F", copy=True))
            fact = symm_fact(m2, b"U", kind=2)
    return fact

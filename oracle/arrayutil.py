"""svec/smat conversions and the symmetric Kronecker product (oracle; test infrastructure).

svec order: columns of the upper triangle (`for j in 1:side, i in 1:j`), off-diagonals
scaled by sqrt(2).  reference: src/Cones/arrayutilities.jl:71-116 (lengths / indices),
:136-156 (scale_svec!), :163-181 (smat_to_svec!), :218-236 (svec_to_smat!), :268-306
(symm_kron!).
"""
import numpy as np

RT2 = np.sqrt(2.0)


def svec_length(side):
    return side * (side + 1) // 2


def svec_side(length):
    side = int((np.sqrt(1 + 8 * length)) // 2)
    while side * (side + 1) < 2 * length:
        side += 1
    assert side * (side + 1) == 2 * length
    return side


_IDX_CACHE = {}


def _tri_idx(side):
    """(row, col) index arrays of the upper triangle in svec order and the off-diag mask."""
    if side not in _IDX_CACHE:
        cols, rows = np.tril_indices(side)  # enumerates (j, i) with i <= j, j outer
        # np.tril_indices yields (r, c) with c <= r in row-major order of r: r plays the role
        # of the svec column index j and c the row index i.
        _IDX_CACHE[side] = (rows.copy(), cols.copy(), rows != cols)
    return _IDX_CACHE[side]


def smat_to_svec(mat):
    """Upper triangle of a symmetric matrix -> svec (off-diagonals * sqrt 2)."""
    side = mat.shape[0]
    i, j, off = _tri_idx(side)
    vec = mat[i, j].astype(np.float64, copy=True)
    vec[off] *= RT2
    return vec


def svec_to_smat(vec):
    """svec -> full symmetric matrix (both triangles filled)."""
    side = svec_side(vec.shape[0])
    i, j, off = _tri_idx(side)
    vals = np.array(vec, dtype=np.float64, copy=True)
    vals[off] /= RT2
    mat = np.zeros((side, side))
    mat[i, j] = vals
    mat[j, i] = vals
    return mat


def svecs_to_smats(arr):
    """(dim, c) array of svec columns -> (c, side, side) stack of symmetric matrices."""
    dim, c = arr.shape
    side = svec_side(dim)
    i, j, off = _tri_idx(side)
    vals = np.array(arr.T, dtype=np.float64, copy=True)  # (c, dim)
    vals[:, off] /= RT2
    mats = np.zeros((c, side, side))
    mats[:, i, j] = vals
    mats[:, j, i] = vals
    return mats


def smats_to_svecs(mats):
    """(c, side, side) stack -> (dim, c) svec columns (uses the upper triangle)."""
    side = mats.shape[1]
    i, j, off = _tri_idx(side)
    vals = mats[:, i, j].copy()
    vals[:, off] *= RT2
    return np.ascontiguousarray(vals.T)


def scale_svec(arr, scal):
    """Scale the off-diagonal rows of svec-indexed rows of `arr` in place
    (reference: arrayutilities.jl:136-156)."""
    side = svec_side(arr.shape[0])
    _, _, off = _tri_idx(side)
    arr[off] *= scal
    return arr


def symm_kron(mat):
    """Symmetric Kronecker product: the matrix of M -> svec(mat * smat(M) * mat') in svec
    coordinates (reference: arrayutilities.jl:268-306; the reference fills the upper triangle
    entry by entry, here it is formed by applying the operator to the svec basis)."""
    side = mat.shape[0]
    dim = svec_length(side)
    basis = svecs_to_smats(np.eye(dim))           # (dim, side, side)
    prod = mat @ basis @ mat.T
    return smats_to_svecs(prod)

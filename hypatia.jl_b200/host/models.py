"""Conic model container: min c'x : b - Ax = 0, h - Gx in K.

Host-side mirror of the reference's data container (reference:
src/Models/Models.jl:14-66).  Its field layout (n, p, q, c, A, b, G, h, cones,
cone_idxs, nu) is the input contract of the system-solver boundary
(SURVEY.md section 8b).  Arrays are float64; A and G are dense, Fortran
(column-major) ordered exactly like Julia matrices so that the raw pointers
can be handed to the C ABI without a transpose.
"""
from __future__ import annotations

import numpy as np

# cone type codes shared with include/hypatia_b200.h (HYP_CONE_*)
CONE_NONNEGATIVE = 0
CONE_EPINORMEUCL = 1
CONE_POSSEMIDEFTRI = 2
CONE_HYPOPERLOGDETTRI = 3
CONE_HYPOROOTDETTRI = 4
CONE_EPIPERSEPSPECTRAL_MAT = 5
CONE_EPIPERSQUARE = 6
CONE_HYPOPERLOG = 7
CONE_EPINORMINF = 8
CONE_EPIPERSEPSPECTRAL_VEC = 9
CONE_HYPOGEOMEAN = 10
CONE_GENERALIZEDPOWER = 11
CONE_HYPOPOWERMEAN = 12
CONE_EPIRELENTROPY = 13
CONE_EPINORMSPECTRAL = 14
CONE_WSOSINTERPNONNEGATIVE = 15
CONE_LINMATRIXINEQ = 16
CONE_DOUBLYNONNEGATIVETRI = 17
CONE_MATRIXEPIPERSQUARE = 18
CONE_WSOSINTERPPOSSEMIDEFTRI = 19
CONE_WSOSINTERPEPINORMEUCL = 20
CONE_WSOSINTERPEPINORMONE = 21
CONE_POSSEMIDEFTRISPARSE = 22
CONE_EPITRRELENTROPYTRI = 23

# separable spectral functions of EpiPerSepSpectral (sepspectralfun.jl:17-116), HYP_SSF_*
SSF_INV, SSF_NEGLOG, SSF_NEGENTROPY, SSF_POWER12 = 0, 1, 2, 3

CONE_NAMES = {
    CONE_NONNEGATIVE: "Nonnegative",
    CONE_EPINORMEUCL: "EpiNormEucl",
    CONE_POSSEMIDEFTRI: "PosSemidefTri",
    CONE_HYPOPERLOGDETTRI: "HypoPerLogdetTri",
    CONE_HYPOROOTDETTRI: "HypoRootdetTri",
    CONE_EPIPERSEPSPECTRAL_MAT: "EpiPerSepSpectral{MatrixCSqr}",
    CONE_EPIPERSQUARE: "EpiPerSquare",
    CONE_HYPOPERLOG: "HypoPerLog",
    CONE_EPINORMINF: "EpiNormInf",
    CONE_EPIPERSEPSPECTRAL_VEC: "EpiPerSepSpectral{VectorCSqr}",
    CONE_HYPOGEOMEAN: "HypoGeoMean",
    CONE_GENERALIZEDPOWER: "GeneralizedPower",
    CONE_HYPOPOWERMEAN: "HypoPowerMean",
    CONE_EPIRELENTROPY: "EpiRelEntropy",
    CONE_EPINORMSPECTRAL: "EpiNormSpectral",
    CONE_WSOSINTERPNONNEGATIVE: "WSOSInterpNonnegative",
    CONE_LINMATRIXINEQ: "LinMatrixIneq",
    CONE_DOUBLYNONNEGATIVETRI: "DoublyNonnegativeTri",
    CONE_MATRIXEPIPERSQUARE: "MatrixEpiPerSquare",
    CONE_WSOSINTERPPOSSEMIDEFTRI: "WSOSInterpPosSemidefTri",
    CONE_WSOSINTERPEPINORMEUCL: "WSOSInterpEpiNormEucl",
    CONE_WSOSINTERPEPINORMONE: "WSOSInterpEpiNormOne",
    CONE_POSSEMIDEFTRISPARSE: "PosSemidefTriSparse",
    CONE_EPITRRELENTROPYTRI: "EpiTrRelEntropyTri",
}


def svec_length(side: int) -> int:
    """reference: src/Cones/arrayutilities.jl:71"""
    return side * (side + 1) // 2


def svec_side(length: int) -> int:
    """reference: src/Cones/arrayutilities.jl:87-91"""
    side = int((np.sqrt(1 + 8 * length)) // 2)
    while side * (side + 1) < 2 * length:
        side += 1
    while side * (side + 1) > 2 * length:
        side -= 1
    if side * (side + 1) != 2 * length:
        raise ValueError(f"{length} is not a triangular number")
    return side


class ConeSpec:
    """(type, dim) descriptor of one cone block; nu follows the reference's get_nu."""

    __slots__ = ("ctype", "dim", "use_dual", "hkind", "hparam", "alpha")

    def __init__(self, ctype: int, dim: int, use_dual: bool = False, hkind: int = 0,
                 hparam: float = 0.0, alpha=()):
        self.ctype = int(ctype)
        self.dim = int(dim)
        self.use_dual = bool(use_dual)
        self.hkind = int(hkind)        # EpiPerSepSpectral: which separable spectral function
        self.hparam = float(hparam)    # ... and its parameter (the power of Power12SSF)
        self.alpha = tuple(float(a) for a in alpha)   # GeneralizedPower: the powers (sum 1); dim = len(alpha) + n
        if ctype == CONE_NONNEGATIVE:
            assert dim >= 1
        elif ctype == CONE_EPINORMEUCL:
            assert dim >= 2
        elif ctype == CONE_POSSEMIDEFTRI:
            svec_side(dim)
        elif ctype == CONE_HYPOPERLOGDETTRI:
            assert dim >= 3
            svec_side(dim - 2)
        elif ctype == CONE_HYPOROOTDETTRI:
            assert dim >= 2
            svec_side(dim - 1)
        elif ctype == CONE_EPIPERSEPSPECTRAL_MAT:
            assert dim >= 3 and hkind in (SSF_INV, SSF_NEGLOG, SSF_NEGENTROPY, SSF_POWER12)
            assert hkind != SSF_POWER12 or 1 < hparam <= 2
            svec_side(dim - 2)
        elif ctype == CONE_EPIPERSQUARE:
            assert dim >= 3
        elif ctype == CONE_HYPOPERLOG:
            assert dim >= 3
        elif ctype in (CONE_WSOSINTERPEPINORMEUCL, CONE_WSOSINTERPEPINORMONE):
            # hkind = R >= 2; alpha = packed Ps; dim = U R
            Rr = hkind
            assert Rr >= 2 and dim % Rr == 0
            Uu = dim // Rr
            nP = int(self.alpha[0])
            Ls = [int(x) for x in self.alpha[1:1 + nP]]
            assert nP >= 1 and all(1 <= L <= Uu for L in Ls) and len(self.alpha) == 1 + nP + Uu * sum(Ls)
        elif ctype == CONE_WSOSINTERPPOSSEMIDEFTRI:
            # hkind = R; alpha = packed data [nP, L_1 .. L_nP, vec(P_1) .. vec(P_nP)], P_k of U x L_k, dim = U svec_length(R)
            Rr = hkind
            assert Rr >= 1 and dim % (Rr * (Rr + 1) // 2) == 0
            Uu = dim // (Rr * (Rr + 1) // 2)
            nP = int(self.alpha[0])
            Ls = [int(x) for x in self.alpha[1:1 + nP]]
            assert nP >= 1 and all(1 <= L <= Uu for L in Ls) and len(self.alpha) == 1 + nP + Uu * sum(Ls)
        elif ctype == CONE_MATRIXEPIPERSQUARE:
            # hkind = d1: dim = svec_length(d1) + 1 + d1 * d2 with d1 <= d2 (matrixepipersquare.jl:56-74)
            d1 = hkind
            rest = dim - d1 * (d1 + 1) // 2 - 1
            assert d1 >= 1 and rest >= d1 * d1 and rest % d1 == 0
        elif ctype == CONE_EPITRRELENTROPYTRI:
            assert dim > 2 and dim % 2 == 1       # epitrrelentropytri.jl:70-71
            svec_side((dim - 1) // 2)
        elif ctype == CONE_POSSEMIDEFTRISPARSE:
            # alpha = packed pattern [side, row_1 .. row_dim, col_1 .. col_dim] (0-based, col <= row, every diagonal present)
            side = int(self.alpha[0])
            assert len(self.alpha) == 1 + 2 * dim and 1 <= side <= dim
            rows = [int(x) for x in self.alpha[1:1 + dim]]
            cols = [int(x) for x in self.alpha[1 + dim:]]
            assert all(0 <= c <= r < side for r, c in zip(rows, cols))
            assert sorted(r for r, c in zip(rows, cols) if r == c) == list(range(side))
        elif ctype == CONE_DOUBLYNONNEGATIVETRI:
            assert dim >= 1
            svec_side(dim)
        elif ctype == CONE_LINMATRIXINEQ:
            # alpha = packed data [side, vec(A_1) .. vec(A_dim)], A_i symmetric side x side (linmatrixineq.jl:38-66)
            side = int(self.alpha[0])
            assert dim > 1 and side >= 1 and side * (side + 1) // 2 >= dim
            assert len(self.alpha) == 1 + dim * side * side
        elif ctype == CONE_WSOSINTERPNONNEGATIVE:
            # alpha = packed data [nP, L_1 .. L_nP, vec(P_1) .. vec(P_nP)] with P_k of dim x L_k, column-major
            nP = int(self.alpha[0])
            Ls = [int(x) for x in self.alpha[1:1 + nP]]
            assert dim >= 1 and nP >= 1 and all(1 <= L <= dim for L in Ls)
            assert len(self.alpha) == 1 + nP + dim * sum(Ls)
        elif ctype == CONE_EPINORMSPECTRAL:
            # hkind = d1 (rows), d2 = (dim - 1) / d1 columns, d1 <= d2 (epinormspectral.jl:55-66)
            assert dim >= 2 and hkind >= 1 and (dim - 1) % hkind == 0 and hkind <= (dim - 1) // hkind
        elif ctype == CONE_EPIRELENTROPY:
            assert dim >= 3 and dim % 2 == 1      # epirelentropy.jl:51-52
        elif ctype in (CONE_EPINORMINF, CONE_HYPOGEOMEAN):
            assert dim >= 2
        elif ctype == CONE_GENERALIZEDPOWER:
            assert dim >= 3 and 1 <= len(self.alpha) < dim and all(a > 0 for a in self.alpha)
            assert abs(sum(self.alpha) - 1) <= 1e-12
        elif ctype == CONE_HYPOPOWERMEAN:
            assert dim >= 2 and len(self.alpha) == dim - 1 and all(a > 0 for a in self.alpha)
            assert abs(sum(self.alpha) - 1) <= 1e-12
        elif ctype == CONE_EPIPERSEPSPECTRAL_VEC:
            assert dim >= 3 and hkind in (SSF_INV, SSF_NEGLOG, SSF_NEGENTROPY, SSF_POWER12)
            assert hkind != SSF_POWER12 or 1 < hparam <= 2
        else:
            raise ValueError(f"unknown cone type {ctype}")

    @property
    def side(self) -> int:
        if self.ctype == CONE_POSSEMIDEFTRI:
            return svec_side(self.dim)
        if self.ctype in (CONE_HYPOPERLOGDETTRI, CONE_EPIPERSEPSPECTRAL_MAT):
            return svec_side(self.dim - 2)
        if self.ctype == CONE_HYPOROOTDETTRI:
            return svec_side(self.dim - 1)
        return 0

    @property
    def nu(self) -> float:
        # nonnegative.jl:40, epinormeucl.jl:42, possemideftri.jl:67,
        # hypoperlogdettri.jl:80, hyporootdettri.jl:80, epipersepspectral.jl:79,
        # epipersquare.jl:50, hypoperlog.jl:54, epinorminf.jl:86
        if self.ctype == CONE_NONNEGATIVE:
            return float(self.dim)
        if self.ctype == CONE_EPINORMEUCL:
            return 2.0
        if self.ctype == CONE_POSSEMIDEFTRI:
            return float(self.side)
        if self.ctype in (CONE_HYPOPERLOGDETTRI, CONE_EPIPERSEPSPECTRAL_MAT):
            return 2.0 + self.side
        if self.ctype == CONE_EPIPERSQUARE:
            return 2.0
        if self.ctype == CONE_GENERALIZEDPOWER:
            return float(len(self.alpha) + 1)
        if self.ctype == CONE_EPITRRELENTROPYTRI:
            return float(2 * svec_side((self.dim - 1) // 2) + 1)      # epitrrelentropytri.jl:119
        if self.ctype in (CONE_EPINORMSPECTRAL, CONE_MATRIXEPIPERSQUARE):
            return float(self.hkind + 1)      # epinormspectral.jl:95, matrixepipersquare.jl:101
        if self.ctype in (CONE_LINMATRIXINEQ, CONE_POSSEMIDEFTRISPARSE):
            return float(int(self.alpha[0]))      # linmatrixineq.jl:72, possemideftrisparse.jl:101 (= side)
        if self.ctype == CONE_WSOSINTERPEPINORMEUCL:
            return float(2 * sum(self.alpha[1:1 + int(self.alpha[0])]))    # wsosinterpepinormeucl.jl:68
        if self.ctype in (CONE_WSOSINTERPPOSSEMIDEFTRI, CONE_WSOSINTERPEPINORMONE):   # R sum L (wsosinterpepinormone.jl:88)
            return float(self.hkind * sum(self.alpha[1:1 + int(self.alpha[0])]))    # wsosinterppossemideftri.jl:66
        if self.ctype == CONE_WSOSINTERPNONNEGATIVE:
            return float(sum(self.alpha[1:1 + int(self.alpha[0])]))    # wsosinterpnonnegative.jl:62
        if self.ctype in (CONE_HYPOPERLOG, CONE_EPINORMINF, CONE_EPIPERSEPSPECTRAL_VEC, CONE_HYPOGEOMEAN,
                          CONE_HYPOPOWERMEAN, CONE_EPIRELENTROPY, CONE_DOUBLYNONNEGATIVETRI):
            return float(self.dim)
        return 1.0 + self.side

    def clone(self):
        return ConeSpec(self.ctype, self.dim, self.use_dual, self.hkind, self.hparam, self.alpha)

    def __repr__(self):
        return f"{CONE_NAMES[self.ctype]}({self.dim})"


def Nonnegative(dim):
    return ConeSpec(CONE_NONNEGATIVE, dim)


def EpiNormEucl(dim):
    return ConeSpec(CONE_EPINORMEUCL, dim)


def PosSemidefTri(dim):
    return ConeSpec(CONE_POSSEMIDEFTRI, dim)


def HypoPerLogdetTri(dim, use_dual=False):
    return ConeSpec(CONE_HYPOPERLOGDETTRI, dim, use_dual)


def HypoRootdetTri(dim, use_dual=False):
    return ConeSpec(CONE_HYPOROOTDETTRI, dim, use_dual)


def EpiPerSepSpectralMat(dim, hkind=SSF_NEGLOG, hparam=1.5, use_dual=False):
    """EpiPerSepSpectral{MatrixCSqr{Float64, Float64}}(h, side), dim = 2 + svec_length(side)."""
    return ConeSpec(CONE_EPIPERSEPSPECTRAL_MAT, dim, use_dual, hkind,
                    hparam if hkind == SSF_POWER12 else 0.0)


def EpiPerSepSpectralVec(dim, hkind=SSF_NEGLOG, hparam=1.5, use_dual=False):
    """EpiPerSepSpectral{VectorCSqr{Float64}, Float64}(h, d), dim = 2 + d."""
    return ConeSpec(CONE_EPIPERSEPSPECTRAL_VEC, dim, use_dual, hkind, hparam if hkind == SSF_POWER12 else 0.0)


def EpiPerSquare(dim, use_dual=False):
    return ConeSpec(CONE_EPIPERSQUARE, dim, use_dual)


def HypoPerLog(dim, use_dual=False):
    return ConeSpec(CONE_HYPOPERLOG, dim, use_dual)


def GeneralizedPower(alpha, n, use_dual=False):
    """GeneralizedPower{Float64}(alpha, n): (u in R^m_+, w in R^n), prod u_i^alpha_i >= |w|; MOI's PowerCone(a) is
    GeneralizedPower([a, 1 - a], 1) (MathOptInterface/cones.jl:33-37)."""
    return ConeSpec(CONE_GENERALIZEDPOWER, len(alpha) + n, use_dual, alpha=alpha)


def WSOSInterpNonnegative(U, Ps, use_dual=False):
    """WSOSInterpNonnegative{Float64, Float64}(U, Ps, use_dual): Ps[k] is U x L_k.  As in the reference
    (wsosinterpnonnegative.jl:59) the barrier is the dual cone's, so the spec's use_dual (= use_dual_barrier) is
    `not use_dual`.  The matrices travel in the per-cone double array of hyp_set_cone_alpha."""
    Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
    assert all(P.ndim == 2 and P.shape[0] == U for P in Ps)
    packed = np.concatenate([[float(len(Ps))], [float(P.shape[1]) for P in Ps]] + [P.ravel(order="F") for P in Ps])
    return ConeSpec(CONE_WSOSINTERPNONNEGATIVE, U, not use_dual, alpha=packed)


def MatrixEpiPerSquare(d1, d2, use_dual=False):
    """MatrixEpiPerSquare{Float64, Float64}(d1, d2): (svec(U), v, vec(W)), U symmetric d1 x d1, W d1 x d2, d1 <= d2,
    2 v U - W W' psd."""
    assert 1 <= d1 <= d2
    return ConeSpec(CONE_MATRIXEPIPERSQUARE, d1 * (d1 + 1) // 2 + 1 + d1 * d2, use_dual, hkind=d1)


def PosSemidefTriSparse(side, row_idxs, col_idxs, use_dual=False):
    """PosSemidefTriSparse{PSDSparseDense, Float64, Float64}(side, row_idxs, col_idxs): 0-based lower-triangle pattern with
    every diagonal entry; the pattern travels in the per-cone double array of hyp_set_cone_alpha."""
    rows, cols = np.asarray(row_idxs, dtype=np.float64), np.asarray(col_idxs, dtype=np.float64)
    assert rows.size == cols.size
    return ConeSpec(CONE_POSSEMIDEFTRISPARSE, rows.size, use_dual, alpha=np.concatenate(([float(side)], rows, cols)))


def EpiTrRelEntropyTri(dim, use_dual=False):
    """EpiTrRelEntropyTri{Float64}(dim): (u, svec(V), svec(W)), u >= tr(W log W - W log V), dim = 1 + 2 svec_length(d)."""
    return ConeSpec(CONE_EPITRRELENTROPYTRI, dim, use_dual)


def DoublyNonnegativeTri(dim, use_dual=False):
    """DoublyNonnegativeTri{Float64}(dim): svec of a symmetric matrix that is psd and entrywise nonnegative."""
    return ConeSpec(CONE_DOUBLYNONNEGATIVETRI, dim, use_dual)


def LinMatrixIneq(As, use_dual=False):
    """LinMatrixIneq{Float64}(As): {w : sum_i w_i A_i psd}, dense real symmetric A_i, A_1 positive definite; the matrices
    travel in the per-cone double array of hyp_set_cone_alpha as [side, vec(A_1) .. vec(A_dim)]."""
    As = [np.asarray(A, dtype=np.float64) for A in As]
    side = As[0].shape[0]
    assert all(A.shape == (side, side) for A in As)
    packed = np.concatenate([[float(side)]] + [A.ravel(order="F") for A in As])
    return ConeSpec(CONE_LINMATRIXINEQ, len(As), use_dual, alpha=packed)


def lmi_unpack(spec):
    """The As matrices of a LinMatrixIneq spec."""
    side = int(spec.alpha[0])
    data = np.asarray(spec.alpha[1:], dtype=np.float64)
    return [data[i * side * side:(i + 1) * side * side].reshape(side, side, order="F") for i in range(spec.dim)]


def WSOSInterpPosSemidefTri(R, U, Ps, use_dual=False):
    """WSOSInterpPosSemidefTri{Float64}(R, U, Ps, use_dual): R x R matrix polynomials in svec block order, dim =
    U svec_length(R); dual barrier by default like WSOSInterpNonnegative.  R travels as the integer parameter of
    hyp_set_cone_params, the Ps in hyp_set_cone_alpha."""
    Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
    assert all(P.ndim == 2 and P.shape[0] == U for P in Ps)
    packed = np.concatenate([[float(len(Ps))], [float(P.shape[1]) for P in Ps]] + [P.ravel(order="F") for P in Ps])
    return ConeSpec(CONE_WSOSINTERPPOSSEMIDEFTRI, U * R * (R + 1) // 2, not use_dual, hkind=R, alpha=packed)


def WSOSInterpEpiNormEucl(R, U, Ps, use_dual=False):
    """WSOSInterpEpiNormEucl{Float64}(R, U, Ps, use_dual): R polynomials of U coefficients, the first one an epigraph of the
    Euclidean norm of the others; dual barrier by default.  R in hyp_set_cone_params, the Ps in hyp_set_cone_alpha."""
    Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
    assert R >= 2 and all(P.ndim == 2 and P.shape[0] == U for P in Ps)
    packed = np.concatenate([[float(len(Ps))], [float(P.shape[1]) for P in Ps]] + [P.ravel(order="F") for P in Ps])
    return ConeSpec(CONE_WSOSINTERPEPINORMEUCL, U * R, not use_dual, hkind=R, alpha=packed)


def WSOSInterpEpiNormOne(R, U, Ps, use_dual=False):
    """WSOSInterpEpiNormOne{Float64}(R, U, Ps, use_dual): like WSOSInterpEpiNormEucl with the l1 norm of the other
    polynomials."""
    spec = WSOSInterpEpiNormEucl(R, U, Ps, use_dual)
    return ConeSpec(CONE_WSOSINTERPEPINORMONE, spec.dim, spec.use_dual, hkind=R, alpha=spec.alpha)


def wsos_unpack(spec):
    """The Ps matrices of a WSOSInterpNonnegative / WSOSInterpPosSemidefTri spec."""
    U = spec.dim
    if spec.ctype == CONE_WSOSINTERPPOSSEMIDEFTRI:
        U = spec.dim // (spec.hkind * (spec.hkind + 1) // 2)
    if spec.ctype in (CONE_WSOSINTERPEPINORMEUCL, CONE_WSOSINTERPEPINORMONE):
        U = spec.dim // spec.hkind
    nP = int(spec.alpha[0])
    Ls = [int(x) for x in spec.alpha[1:1 + nP]]
    data = np.asarray(spec.alpha[1 + nP:], dtype=np.float64)
    out, o = [], 0
    for L in Ls:
        out.append(data[o:o + U * L].reshape(U, L, order="F"))
        o += U * L
    return out


def EpiNormSpectral(d1, d2, use_dual=False):
    """EpiNormSpectral{Float64, Float64}(d1, d2): (u, vec(W)), W of d1 <= d2 rows x columns, u >= sigma_max(W);
    use_dual = True gives the nuclear-norm epigraph (MOI NormSpectralCone / NormNuclearCone)."""
    assert 1 <= d1 <= d2
    return ConeSpec(CONE_EPINORMSPECTRAL, 1 + d1 * d2, use_dual, hkind=d1)


def EpiRelEntropy(dim, use_dual=False):
    """EpiRelEntropy{Float64}(dim): (u, v, w), u >= sum w_i log(w_i / v_i), dim = 1 + 2 d (MOI RelativeEntropyCone)."""
    return ConeSpec(CONE_EPIRELENTROPY, dim, use_dual)


def HypoPowerMean(alpha, use_dual=False):
    """HypoPowerMean{Float64}(alpha): (u, w in R^d_+), u <= prod w_i^alpha_i, dim = 1 + len(alpha)."""
    return ConeSpec(CONE_HYPOPOWERMEAN, 1 + len(alpha), use_dual, alpha=alpha)


def HypoGeoMean(dim, use_dual=False):
    return ConeSpec(CONE_HYPOGEOMEAN, dim, use_dual)


def EpiNormInf(dim, use_dual=False):
    """EpiNormInf{Float64, Float64}(dim); use_dual = True gives the l1-norm epigraph."""
    return ConeSpec(CONE_EPINORMINF, dim, use_dual)


class Model:
    """reference: src/Models/Models.jl:14-54 (fields and derived cone_idxs / nu)."""

    def __init__(self, c, A, b, G, h, cones, obj_offset: float = 0.0):
        self.c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1)
        self.b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        self.h = np.ascontiguousarray(h, dtype=np.float64).reshape(-1)
        self.n = self.c.size
        self.p = self.b.size
        self.q = self.h.size
        A = np.zeros((self.p, self.n)) if A is None else np.asarray(A, dtype=np.float64)
        self.A = np.asfortranarray(A.reshape(self.p, self.n))
        self.G = np.asfortranarray(np.asarray(G, dtype=np.float64).reshape(self.q, self.n))
        self.cones = list(cones)
        self.obj_offset = float(obj_offset)
        self._index_cones()

    def _index_cones(self):
        # reference: build_cone_idxs, Models.jl:56-66 (0-based half-open here)
        dims = np.array([ck.dim for ck in self.cones], dtype=np.int64)
        self.cone_dims = dims
        self.cone_offsets = np.concatenate(([0], np.cumsum(dims)))[:-1].astype(np.int64) \
            if dims.size else np.zeros(0, dtype=np.int64)
        if int(dims.sum()) != self.q:
            raise ValueError("cone dimensions do not sum to q")
        self.cone_idxs = [slice(int(o), int(o + d)) for o, d in zip(self.cone_offsets, dims)]
        self.cone_nus = np.array([ck.nu for ck in self.cones], dtype=np.float64)
        self.nu = float(self.cone_nus.sum()) if dims.size else 0.0

    def copy(self):
        return Model(self.c.copy(), self.A.copy(), self.b.copy(), self.G.copy(), self.h.copy(),
                     [ck.clone() for ck in self.cones],
                     self.obj_offset)

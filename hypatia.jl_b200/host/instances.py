"""Synthetic dense conic instances for the parity tests and bench.py.

All data float64 from numpy.random.Generator(PCG64(seed)).  Recipe: a
strictly feasible primal-dual pair is planted so that the first iterate is interior:
G ~ N(0,1)^{q x n} (rows of matrix-cone blocks svec-scaled: off-diagonals * sqrt 2, like
scale_svec!, reference src/Cones/arrayutilities.jl:136-156), x0 ~ N(0,1)^n, s0 / z0 = the
cone's central point perturbed like the reference's cone tests (test/cone.jl:236-248),
h = G x0 + s0, c = -G'z0 - A'y0, b = A x0.  Instances are stated post-preprocessing, i.e. as
the (n, p, q, G, A, cones) that load(syssolver, solver) sees (reference
src/Solvers/Solvers.jl:334).

`linearopt` restates examples/linearopt/native.jl:15-30 (BASELINE config 1).
"""
from __future__ import annotations

import numpy as np

from . import models as M
from .point import Point

RT2 = np.sqrt(2.0)

_RAYS = np.array([
    [-0.827838387, 0.805102007, 1.290927686], [-0.689607388, 0.724605082, 1.224617936],
    [-0.584372665, 0.68128058, 1.182421942], [-0.503499342, 0.65448622, 1.153053152],
    [-0.440285893, 0.636444224, 1.131466926], [-0.389979809, 0.623569352, 1.114979519],
    [-0.349255921, 0.613978276, 1.102013921], [-0.315769104, 0.606589839, 1.091577908],
    [-0.287837744, 0.600745284, 1.083013], [-0.264242734, 0.596019009, 1.075868782]])


def _central_ray_hypoperlog(d):
    # numerical data of the reference (src/Cones/hypoperlog.jl:289-319)
    if d <= 10:
        return _RAYS[d - 1]
    x = 1.0 / d
    if d <= 70:
        return np.array([4.657876 * x * x - 3.116192 * x + 0.000647, 0.424682 * x + 0.553392,
                         0.760412 * x + 1.001795])
    return np.array([-3.011166 * x - 0.000122, 0.395308 * x + 0.553955, 0.837545 * x + 1.000024])


def _svec_diag_idx(side):
    j = np.arange(side)
    return j * (j + 1) // 2 + j


def _svec_offdiag_mask(side):
    mask = np.ones(side * (side + 1) // 2, dtype=bool)
    mask[_svec_diag_idx(side)] = False
    return mask


def ssf_initial_point(hkind, d):
    """get_initial_point of the separable spectral functions (sepspectralfun.jl:29-32, :49-52,
    :69-72, :112-115): (u, v, w_ii)."""
    return (2.0 * d, 1.0, 1.0) if hkind in (M.SSF_INV, M.SSF_POWER12) else (1.0, 1.0, 1.0)


def ssf_eval(hkind, hparam, x):
    """(h, h', h'') of a separable spectral function at the scalar x (sepspectralfun.jl:17-110)."""
    if hkind == M.SSF_INV:
        return 1 / x, -x ** -2.0, 2 * x ** -3.0
    if hkind == M.SSF_NEGLOG:
        return -np.log(x), -1 / x, x ** -2.0
    if hkind == M.SSF_NEGENTROPY:
        return x * np.log(x), 1 + np.log(x), 1 / x
    return x ** hparam, hparam * x ** (hparam - 1), hparam * (hparam - 1) * x ** (hparam - 2)


def mat_offset(spec):
    """Number of scalar entries in front of the svec block of a matrix cone."""
    return {M.CONE_POSSEMIDEFTRI: 0, M.CONE_HYPOPERLOGDETTRI: 2, M.CONE_HYPOROOTDETTRI: 1,
            M.CONE_EPIPERSEPSPECTRAL_MAT: 2}[spec.ctype]


def cone_initial_point(spec):
    """Central primal point of one cone (set_initial_point!; nonnegative.jl:42,
    epinormeucl.jl:44-52, possemideftri.jl:69-78, hypoperlogdettri.jl:82-94,
    hyporootdettri.jl:82-98, matrixcsqr.jl:75-88 (not central), epipersquare.jl:59-64,
    hypoperlog.jl:62-69)."""
    arr = np.zeros(spec.dim)
    if spec.ctype == M.CONE_NONNEGATIVE:
        arr[:] = 1.0
    elif spec.ctype == M.CONE_EPINORMEUCL:
        arr[0] = RT2
    elif spec.ctype == M.CONE_POSSEMIDEFTRI:
        arr[_svec_diag_idx(spec.side)] = 1.0
    elif spec.ctype == M.CONE_HYPOPERLOGDETTRI:
        u, v, w = _central_ray_hypoperlog(spec.side)
        arr[0], arr[1] = u, v
        arr[2 + _svec_diag_idx(spec.side)] = w
    elif spec.ctype == M.CONE_HYPOROOTDETTRI:
        d = spec.side
        c1 = np.sqrt(5.0 * d * d + 2 * d + 1)
        c2 = arr[0] = -np.sqrt((3 * d + 1 - c1) / (2.0 * d + 2))
        arr[1 + _svec_diag_idx(d)] = -c2 * (d + 1 + c1) / (2.0 * d)
    elif spec.ctype == M.CONE_EPIPERSEPSPECTRAL_MAT:
        u, v, w = ssf_initial_point(spec.hkind, spec.side)
        arr[0], arr[1] = u, v
        arr[2 + _svec_diag_idx(spec.side)] = w
    elif spec.ctype == M.CONE_EPIPERSQUARE:
        arr[0] = arr[1] = 1.0
    elif spec.ctype == M.CONE_HYPOPERLOG:
        u, v, w = _central_ray_hypoperlog(spec.dim - 2)
        arr[0], arr[1] = u, v
        arr[2:] = w
    elif spec.ctype == M.CONE_EPINORMINF:
        arr[0] = np.sqrt(spec.dim)      # epinorminf.jl:88-95
    elif spec.ctype in (M.CONE_WSOSINTERPEPINORMEUCL, M.CONE_WSOSINTERPEPINORMONE):
        arr[:spec.dim // spec.hkind] = 1.0       # wsosinterpepinormeucl.jl:113-117
    elif spec.ctype == M.CONE_WSOSINTERPPOSSEMIDEFTRI:
        Rr = spec.hkind                         # wsosinterppossemideftri.jl:98-106: ones on the diagonal blocks
        Uu = spec.dim // (Rr * (Rr + 1) // 2)
        for p in range(Rr):
            b = p * (p + 1) // 2 + p
            arr[b * Uu:(b + 1) * Uu] = 1.0
    elif spec.ctype == M.CONE_EPITRRELENTROPYTRI:
        vw = (spec.dim - 1) // 2                # epitrrelentropytri.jl:121-135: diagonal V, W from the vector cone's ray
        d = M.svec_side(vw)
        u, v, w = _central_ray_epirelentropy(d)
        arr[0] = u
        arr[1 + _svec_diag_idx(d)] = v
        arr[1 + vw + _svec_diag_idx(d)] = w
    elif spec.ctype == M.CONE_POSSEMIDEFTRISPARSE:
        a = np.asarray(spec.alpha)              # possemideftrisparse.jl:103-116: the identity
        arr[:] = (a[1:1 + spec.dim] == a[1 + spec.dim:]).astype(float)
    elif spec.ctype == M.CONE_MATRIXEPIPERSQUARE:
        d1 = spec.hkind                         # matrixepipersquare.jl:103-116: U = I, v = 1, W = 0
        arr[_svec_diag_idx(d1)] = 1.0
        arr[d1 * (d1 + 1) // 2] = 1.0
    elif spec.ctype == M.CONE_DOUBLYNONNEGATIVETRI:
        side = M.svec_side(spec.dim)
        ond, offd = dnn_initial_point(side)
        arr[:] = offd
        arr[_svec_diag_idx(side)] = ond
    elif spec.ctype == M.CONE_LINMATRIXINEQ:
        arr[0] = 1.0                            # linmatrixineq.jl:74-82
    elif spec.ctype == M.CONE_WSOSINTERPNONNEGATIVE:
        arr[:] = 1.0                            # wsosinterpnonnegative.jl:89
    elif spec.ctype == M.CONE_EPINORMSPECTRAL:
        arr[0] = np.sqrt(spec.hkind + 1.0)      # epinormspectral.jl:97-105
    elif spec.ctype == M.CONE_EPIRELENTROPY:
        d = (spec.dim - 1) // 2           # epirelentropy.jl:84-89, :377-409
        u, v, w = _central_ray_epirelentropy(d)
        arr[0] = u
        arr[1:1 + d] = v
        arr[1 + d:] = w
    elif spec.ctype == M.CONE_HYPOPOWERMEAN:
        al = np.array(spec.alpha)         # hypopowermean.jl:58-72, :205-232 (fitted central ray)
        d = al.size
        if np.all(al == 1.0 / d):
            c = np.sqrt(5.0 * d * d + 2 * d + 1)
            arr[0] = -np.sqrt((-c + 3 * d + 1) / (2.0 + 2 * d))
            arr[1:] = (c - d + 1) / np.sqrt((1 + d) * (-2 * c + 6 * d + 2))
        else:
            if d == 1:
                w = np.full(1, 1.306563)
            elif d == 2:
                w = 1.0049885 + 0.2986276 * al
            elif d <= 5:
                w = 1.0040142949 - 0.0004885108 * d + 0.3016645951 * al
            elif d <= 20:
                w = 1.001168 - 4.547017e-05 * d + 3.032880e-01 * al
            elif d <= 100:
                w = 1.000069 - 5.469926e-07 * d + 3.074084e-01 * al
            else:
                w = 1 + 3.086535e-01 * al
            pw = np.exp(np.sum(al * np.log(w)))
            arr[0] = pw - pw / d * np.sum(al / (w * w - 1))
            arr[1:] = w
    elif spec.ctype == M.CONE_GENERALIZEDPOWER:
        arr[:len(spec.alpha)] = np.sqrt(1 + np.array(spec.alpha))      # generalizedpower.jl:71-75
    elif spec.ctype == M.CONE_HYPOGEOMEAN:
        d = spec.dim - 1                  # hypogeomean.jl:259-264
        c = np.sqrt(5.0 * d * d + 2 * d + 1)
        arr[0] = -np.sqrt((-c + 3 * d + 1) / (2.0 + 2 * d))
        arr[1:] = (c - d + 1) / np.sqrt((1 + d) * (-2 * c + 6 * d + 2))
    elif spec.ctype == M.CONE_EPIPERSEPSPECTRAL_VEC:
        u, v, w = ssf_initial_point(spec.hkind, spec.dim - 2)   # vectorcsqr.jl:52-59
        arr[0], arr[1] = u, v
        arr[2:] = w
    return arr


def dnn_initial_point(side):
    """doublynonnegativetri.jl:72-126: (on-diagonal, off-diagonal) value of the central point in svec coordinates."""
    if side == 1:
        return 1.0, 1.0
    if side == 2:
        return np.sqrt(5.0) / 2, 1 / np.sqrt(2.0)
    n, d = float(side), float(side * (side + 1) // 2)
    rt2 = np.sqrt(2.0)
    # roots of p1 (coefficients by ascending power in the reference) give the off-diagonal value
    p1 = [-n - 1, 0, n ** 2 + n + 7, 0, -2 * n ** 2 - 8, 0, n ** 2]
    for r in np.roots(p1[::-1]):
        offd = float(np.real(r))
        if offd > 0:
            temp = d - (d - n) * offd ** 2
            if temp > np.sqrt(np.finfo(np.float64).eps):
                ond = np.sqrt(temp / n)
                denom = ond ** 2 + (n - 2) / rt2 * ond * offd - (n - 1) * offd ** 2 / 2
                if np.isclose(ond * rt2 + (n - 2) * offd, ond * denom * rt2) and np.isclose(denom, offd ** 2 * (denom + 1)):
                    return ond, offd
    return n + 1, 1.0


_CENTRAL_EPIRELENTROPY = np.array([      # epirelentropy.jl:398-409
    [0.827838399, 1.290927714, 0.805102005], [0.708612491, 1.256859155, 0.818070438],
    [0.622618845, 1.231401008, 0.829317079], [0.558111266, 1.211710888, 0.838978357],
    [0.508038611, 1.196018952, 0.847300431], [0.468039614, 1.183194753, 0.854521307],
    [0.435316653, 1.172492397, 0.860840992], [0.408009282, 1.163403374, 0.866420017],
    [0.38483862, 1.155570329, 0.871385499], [0.364899122, 1.148735192, 0.875838068]])


def _central_ray_epirelentropy(d):
    # epirelentropy.jl:377-396
    if d <= 10:
        return _CENTRAL_EPIRELENTROPY[d - 1]
    rt = np.sqrt(d)
    if d <= 20:
        return np.array([1.2023 / rt - 0.015, 0.432 / rt + 1.0125, -0.3057 / rt + 0.972])
    return np.array([1.1513 / rt - 0.0069, 0.4873 / rt + 1.0008, -0.4247 / rt + 0.9961])


def _cone_dual_initial(spec, prim):
    """-grad at the central point, closed form per cone (dual of the central point)."""
    if spec.ctype in (M.CONE_WSOSINTERPEPINORMEUCL, M.CONE_WSOSINTERPEPINORMONE):
        # -grad at (1, 0, .., 0): the arrow matrices are block diagonal, so -g_1 = c diag(P (P'P)^-1 P') summed over k with
        # c = R - (R - 2) = 2 (Euclidean norm) or 2 (R - 1) - (R - 2) = R (l1 norm)
        Uu = spec.dim // spec.hkind
        cf = 2.0 if spec.ctype == M.CONE_WSOSINTERPEPINORMEUCL else float(spec.hkind)
        out = np.zeros_like(prim)
        for P in M.wsos_unpack(spec):
            out[:Uu] += cf * np.einsum("ij,ji->i", P, np.linalg.solve(P.T @ P, P.T))
        return out
    if spec.ctype == M.CONE_WSOSINTERPPOSSEMIDEFTRI:
        # -grad at the initial point: D = I, so every diagonal block gets diag(P_k (P_k' P_k)^-1 P_k') summed over k and
        # the off-diagonal blocks vanish (wsosinterppossemideftri.jl:144-188)
        Rr = spec.hkind
        Uu = spec.dim // (Rr * (Rr + 1) // 2)
        dg = np.zeros(Uu)
        for P in M.wsos_unpack(spec):
            dg += np.einsum("ij,ji->i", P, np.linalg.solve(P.T @ P, P.T))
        out = np.zeros_like(prim)
        for p in range(Rr):
            b = p * (p + 1) // 2 + p
            out[b * Uu:(b + 1) * Uu] = dg
        return out
    if spec.ctype == M.CONE_EPITRRELENTROPYTRI:
        # -grad at diagonal (u, v I, w I): as the vector cone (epirelentropy.jl:123-140) on the diagonals
        vw = (spec.dim - 1) // 2
        d = M.svec_side(vw)
        u, v, w = prim[0], prim[1], prim[1 + vw]
        z = u - d * w * np.log(w / v)
        out = np.zeros_like(prim)
        out[0] = 1.0 / z
        out[1 + _svec_diag_idx(d)] = w / v / z + 1.0 / v
        out[1 + vw + _svec_diag_idx(d)] = -(np.log(w / v) + 1) / z + 1.0 / w
        return out
    if spec.ctype == M.CONE_POSSEMIDEFTRISPARSE:
        return prim.copy()      # -grad at the identity is the identity
    if spec.ctype == M.CONE_MATRIXEPIPERSQUARE:
        # -grad at (I, 1, 0): Z = 2 I, Zi = I / 2 => -g_U = svec(I), -g_v = 2 tr(Zi U) - (d1 - 1) = 1, -g_W = 0
        return prim.copy()
    if spec.ctype == M.CONE_DOUBLYNONNEGATIVETRI:
        return prim.copy()      # the initial point satisfies s = -g(s) (doublynonnegativetri.jl:72-126)
    if spec.ctype == M.CONE_LINMATRIXINEQ:
        # -grad_i = tr(S^-1 A_i) with S = sum_j w_j A_j, linmatrixineq.jl:98-109
        As = M.lmi_unpack(spec)
        Si = np.linalg.inv(sum(w * A for w, A in zip(prim, As)))
        return np.array([np.sum(Si * A) for A in As])
    if spec.ctype == M.CONE_WSOSINTERPNONNEGATIVE:
        # -grad_j = sum_k (P_k (P_k' D P_k)^-1 P_k')_jj, wsosinterpnonnegative.jl:123-138
        out = np.zeros_like(prim)
        for P in M.wsos_unpack(spec):
            Lam = P.T @ (prim[:, None] * P)
            out += np.einsum("ij,ji->i", P, np.linalg.solve(Lam, P.T))
        return out
    if spec.ctype == M.CONE_EPINORMSPECTRAL:
        return prim.copy()      # -grad at (sqrt(d1 + 1), 0) = ((d1 + 1) / u, 0): the central point is self-dual
    if spec.ctype == M.CONE_EPIRELENTROPY:
        # -grad, epirelentropy.jl:123-140
        d = (spec.dim - 1) // 2
        u, v, w = prim[0], prim[1:1 + d], prim[1 + d:]
        lwv = np.log(w / v)
        z = u - w @ lwv
        out = np.empty_like(prim)
        out[0] = 1.0 / z
        out[1:1 + d] = w / v / z + 1.0 / v
        out[1 + d:] = -(lwv + 1) / z + 1.0 / w
        return out
    if spec.ctype == M.CONE_NONNEGATIVE:
        return 1.0 / prim
    if spec.ctype == M.CONE_EPINORMEUCL:
        return prim.copy()      # central point is self-dual: -g = (u, -w)/dist with dist = 1
    if spec.ctype == M.CONE_POSSEMIDEFTRI:
        return prim.copy()
    if spec.ctype == M.CONE_HYPOPOWERMEAN:
        # -grad, hypopowermean.jl:104-116
        al = np.array(spec.alpha)
        u, w = prim[0], prim[1:]
        phi = np.exp(np.sum(al * np.log(w)))
        zeta = phi - u
        out = np.zeros_like(prim)
        out[0] = -1.0 / zeta
        out[1:] = (phi / zeta * al + 1) / w
        return out
    if spec.ctype == M.CONE_GENERALIZEDPOWER:
        # generalizedpower.jl:107-120 at w = 0: zwzwi = 1, -g_u = (alpha + 1) / u = sqrt(1 + alpha) (central point)
        return prim.copy()
    if spec.ctype == M.CONE_HYPOGEOMEAN:
        # hypogeomean.jl:97-110 at w = w0 * 1: phi = w0
        d = spec.dim - 1
        u, w = prim[0], prim[1]
        zeta = w - u
        out = np.zeros_like(prim)
        out[0] = -1.0 / zeta
        out[1:] = (w / zeta / d + 1) / w
        return out
    if spec.ctype == M.CONE_EPIPERSEPSPECTRAL_VEC:
        # vectorcsqr.jl:97-114 at w = w0 * 1
        d = spec.dim - 2
        u, v, w = prim[0], prim[1], prim[2]
        lam = w / v
        hv, h1, _ = ssf_eval(spec.hkind, spec.hparam, lam)
        phi = d * hv
        zeta = u - v * phi
        sigma = phi - d * lam * h1
        out = np.zeros_like(prim)
        out[0] = 1.0 / zeta
        out[1] = 1.0 / v - sigma / zeta
        out[2:] = 1.0 / w - h1 / zeta
        return out
    if spec.ctype == M.CONE_EPINORMINF:
        return prim.copy()      # -g = ((n + 1) / u, 0...) = (sqrt(dim), 0...) at the central point
    if spec.ctype == M.CONE_EPIPERSQUARE:
        return prim.copy()      # (1, 1, 0...): -g = (v, u, -w) / dist with dist = u v - |w|^2 / 2 = 1
    out = np.zeros_like(prim)
    if spec.ctype == M.CONE_HYPOPERLOG:
        # hypoperlog.jl:99-113 at w = w0 * 1
        d = spec.dim - 2
        u, v, w = prim[0], prim[1], prim[2]
        phi = d * np.log(w / v)
        zeta = v * phi - u
        out[0] = -1.0 / zeta
        out[1] = 1.0 / v + (phi - d) / zeta
        out[2:] = (1 + v / zeta) / w
        return out
    d = spec.side
    if spec.ctype == M.CONE_EPIPERSEPSPECTRAL_MAT:
        # matrixcsqr.jl:140-165 at W = w0 * I
        u, v, w = prim[0], prim[1], prim[2]
        lam = w / v
        hv, h1, _ = ssf_eval(spec.hkind, spec.hparam, lam)
        phi = d * hv
        zeta = u - v * phi
        sigma = phi - d * lam * h1
        out[0] = 1.0 / zeta
        out[1] = 1.0 / v - sigma / zeta
        out[2 + _svec_diag_idx(d)] = -(h1 / zeta - 1.0 / w)
        return out
    if spec.ctype == M.CONE_HYPOPERLOGDETTRI:
        u, v, w = prim[0], prim[1], prim[2]
        phi = d * np.log(w) - d * np.log(v)
        zeta = v * phi - u
        out[0] = -1.0 / zeta
        out[1] = 1.0 / v + (phi - d) / zeta
        out[2 + _svec_diag_idx(d)] = (1 + v / zeta) / w
    else:
        u, w = prim[0], prim[1]
        phi = w
        zeta = phi - u
        out[0] = -1.0 / zeta
        out[1 + _svec_diag_idx(d)] = (phi / zeta / d + 1) / w
    return out


def _perturb(rng, spec, vec, noise):
    """vec += U(-noise, noise), with the noise on matrix blocks shrunk by 1/sqrt(side) so the
    perturbed matrix stays safely positive definite at any side."""
    if spec.ctype in (M.CONE_NONNEGATIVE, M.CONE_EPINORMEUCL):
        vec += noise * (2 * rng.random(vec.size) - 1)
        return vec
    if spec.ctype == M.CONE_EPIPERSEPSPECTRAL_VEC:
        vec += noise / (2.0 * (vec.size - 2)) * (2 * rng.random(vec.size) - 1)   # the initial point is not central
        return vec
    if spec.ctype == M.CONE_EPITRRELENTROPYTRI:
        vec += 0.5 * noise / np.sqrt(vec.size) * (2 * rng.random(vec.size) - 1)
        return vec
    if spec.ctype == M.CONE_POSSEMIDEFTRISPARSE:
        vec += 0.5 * noise / np.sqrt(vec.size) * (2 * rng.random(vec.size) - 1)
        return vec
    if spec.ctype == M.CONE_MATRIXEPIPERSQUARE:
        vec += 0.5 * noise / np.sqrt(vec.size) * (2 * rng.random(vec.size) - 1)
        return vec
    if spec.ctype == M.CONE_DOUBLYNONNEGATIVETRI:
        vec += 0.5 * noise / np.sqrt(vec.size) * (2 * rng.random(vec.size) - 1)
        return vec
    if spec.ctype == M.CONE_LINMATRIXINEQ:
        vec += 0.1 * noise / np.sqrt(vec.size) * (2 * rng.random(vec.size) - 1)    # test/cone.jl:426 uses noise 1e-2
        return vec
    if spec.ctype in (M.CONE_GENERALIZEDPOWER, M.CONE_WSOSINTERPNONNEGATIVE, M.CONE_WSOSINTERPPOSSEMIDEFTRI,
                      M.CONE_WSOSINTERPEPINORMEUCL, M.CONE_WSOSINTERPEPINORMONE):
        vec += 0.5 * noise / np.sqrt(vec.size) * (2 * rng.random(vec.size) - 1)
        return vec
    if spec.ctype in (M.CONE_HYPOGEOMEAN, M.CONE_HYPOPOWERMEAN, M.CONE_EPIRELENTROPY, M.CONE_EPINORMSPECTRAL):
        vec[0] += 0.5 * noise * (2 * rng.random() - 1)
        vec[1:] += noise / np.sqrt(vec.size) * (2 * rng.random(vec.size - 1) - 1)
        return vec
    if spec.ctype == M.CONE_EPINORMINF:
        vec[0] += 0.5 * noise * (2 * rng.random() - 1)
        vec[1:] += noise / np.sqrt(vec.size) * (2 * rng.random(vec.size - 1) - 1)
        return vec
    if spec.ctype in (M.CONE_EPIPERSQUARE, M.CONE_HYPOPERLOG):
        vec[:2] += 0.5 * noise * (2 * rng.random(2) - 1)
        vec[2:] += noise / np.sqrt(vec.size - 2) * (2 * rng.random(vec.size - 2) - 1)
        return vec
    if spec.ctype == M.CONE_EPIPERSEPSPECTRAL_MAT:
        noise = noise / (2.0 * spec.side)   # the initial point is not central: stay close to it
    off = mat_offset(spec)
    vec[:off] += 0.5 * noise * (2 * rng.random(off) - 1)
    vec[off:] += noise / np.sqrt(spec.side) * (2 * rng.random(vec.size - off) - 1)
    return vec


class Instance:
    """A model plus an interior starting iterate (x0, y0, z0, tau=1, s0, kap=1)."""

    def __init__(self, name, model, point, mu):
        self.name, self.model, self.point, self.mu = name, model, point, mu


def synthetic(name, n, p, cones, seed, noise=0.1, dtype_rows_chunk=4096):
    """Planted-feasible dense instance with the given cone list (post-preprocessing)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q = int(sum(ck.dim for ck in cones))
    G = np.empty((q, n), order="F")
    for r0 in range(0, q, dtype_rows_chunk):        # chunked so the transient stays small
        r1 = min(q, r0 + dtype_rows_chunk)
        G[r0:r1] = rng.standard_normal((r1 - r0, n))
    s0 = np.empty(q)
    z0 = np.empty(q)
    off = 0
    for ck in cones:
        sl = slice(off, off + ck.dim)
        prim = cone_initial_point(ck)
        dual = _cone_dual_initial(ck, prim)
        prim = _perturb(rng, ck, prim, noise)
        dual = _perturb(rng, ck, dual, noise)
        if ck.use_dual:
            s0[sl], z0[sl] = dual, prim
        else:
            s0[sl], z0[sl] = prim, dual
        if ck.side:
            mo = off + mat_offset(ck)
            rows = mo + np.nonzero(_svec_offdiag_mask(ck.side))[0]
            G[rows] *= RT2
        off += ck.dim
    A = rng.standard_normal((p, n)) if p else np.zeros((0, n))
    x0 = rng.standard_normal(n)
    y0 = rng.standard_normal(p)
    h = G @ x0 + s0
    c = -(G.T @ z0)
    if p:
        c -= A.T @ y0
    b = A @ x0
    model = M.Model(c, A, b, G, h, cones)
    pt = Point(model)
    pt.x[:] = x0
    pt.y[:] = y0
    pt.z[:] = z0
    pt.s[:] = s0
    pt.tau = 1.0
    pt.kap = 1.0
    mu = (float(z0 @ s0) + 1.0) / (model.nu + 1)
    return Instance(name, model, pt, mu)


def linearopt(m=200, n=400, seed=1001):
    """LinearOptNative(m, n, 1.0): A = 10 U(0,1)^{m x n}, b = A 1, c ~ U(0,1)^n, G = -I, h = 0,
    one Nonnegative(n) (reference: examples/linearopt/native.jl:15-30).  BASELINE config 1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A = 10.0 * rng.random((m, n))
    b = A.sum(axis=1)
    c = rng.random(n)
    return M.Model(c, A, b, -np.eye(n), np.zeros(n), [M.Nonnegative(n)])


# ---- BASELINE.json configs; `scale` < 1 shrinks every dimension for tests ----
def config(name, scale=1.0):
    def sc(v, lo=1):
        return max(lo, int(round(v * scale)))
    if name == "C2":     # dense LP after reduction: m = 4000, one Nonnegative(5000)
        return synthetic("C2", sc(4000), 0, [M.Nonnegative(sc(5000))], 1002)
    if name == "C3":     # n = 10000, 2000 x EpiNormEucl(25), q = 50000
        return synthetic("C3", sc(10000), 0, [M.EpiNormEucl(25) for _ in range(sc(2000))], 1003)
    if name == "C4":     # n = 20000, 50 x PosSemidefTri(side 100), q = 252500
        side = 100 if scale >= 1 else max(3, int(round(100 * np.sqrt(scale))))
        return synthetic("C4", sc(20000), 0,
                         [M.PosSemidefTri(M.svec_length(side)) for _ in range(sc(50, 2))], 1004)
    if name == "C5b":    # n = 20000: 40 LogDet(side 100) + 10000 SOC(25) + Nonneg(47920)
        side = 100 if scale >= 1 else max(3, int(round(100 * np.sqrt(scale))))
        cones = [M.HypoPerLogdetTri(2 + M.svec_length(side)) for _ in range(sc(40, 2))]
        cones += [M.EpiNormEucl(25) for _ in range(sc(10000))]
        cones += [M.Nonnegative(sc(47920))]
        return synthetic("C5b", sc(20000), 0, cones, 1005)
    raise ValueError(name)

"""Minimal stand-in for the reference's PolyUtils.interpolate on box domains (src/PolyUtils/: out of scope of the hot
path; the cone only needs the resulting matrices): a unisolvent point set for total degree 2*halfdeg, an orthonormalised
graded polynomial basis P0 evaluated at the points, and the box weights g_i = (x_i - lo_i)(hi_i - x_i) applied to the
basis of degree halfdeg - 1.  The optimum of a WSOS programme does not depend on which valid basis / point set is used,
so the reference's instances (test/nativeinstances.jl:2286-2343) can be reproduced without Julia's random sampling."""
import itertools

import numpy as np


def _exponents(n, deg):
    """Graded (total-degree ordered) exponent tuples of n variables up to degree deg."""
    out = []
    for d in range(deg + 1):
        out += [e for e in itertools.product(range(d + 1), repeat=n) if sum(e) == d]
    return out


def interpolate_box(lo, hi, halfdeg):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    n = lo.size
    deg = 2 * halfdeg
    # principal lattice of the simplex with `deg` subdivisions, mapped into the box: unisolvent for total degree `deg`
    lattice = np.array([e for e in itertools.product(range(deg + 1), repeat=n) if sum(e) <= deg], dtype=float)
    pts = lo + (hi - lo) * lattice / deg
    t = 2 * (pts - lo) / (hi - lo) - 1                      # [-1, 1]^n coordinates for the basis
    exps = _exponents(n, halfdeg)
    V = np.stack([np.prod(t ** np.array(e), axis=1) for e in exps], axis=1)
    P0 = np.linalg.qr(V)[0]
    L1 = len(_exponents(n, halfdeg - 1))
    Ps = [P0]
    for i in range(n):
        g = (pts[:, i] - lo[i]) * (hi[i] - pts[:, i])
        Ps.append(np.sqrt(np.maximum(g, 0.0))[:, None] * P0[:, :L1])
    return pts.shape[0], pts, Ps


def interpolate_free(n, halfdeg):
    """FreeDomain: no weights, Ps = [P0] (the points are those of the box [-1, 1]^n)."""
    U, pts, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return U, pts, Ps[:1]

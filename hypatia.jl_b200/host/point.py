"""Primal-dual point: one flat vector [x(n); y(p); z(q); tau; s(q); kap] with views.

Host-side mirror of the reference layout (reference: src/Solvers/point.jl:24-54);
the same flat layout is what crosses the C ABI in hyp_solve_system / hyp_apply_lhs.
"""
from __future__ import annotations

import numpy as np


class Point:
    def __init__(self, model=None, dims=None):
        if dims is None:
            dims = (model.n, model.p, model.q)
        n, p, q = (int(d) for d in dims)
        self.n, self.p, self.q = n, p, q
        self.tau_idx = n + p + q
        self.vec = np.zeros(n + p + 2 * q + 2)
        self._bind()

    def _bind(self):
        n, p, q, v = self.n, self.p, self.q, self.vec
        self.x = v[:n]
        self.y = v[n:n + p]
        self.z = v[n + p:n + p + q]
        self.s = v[self.tau_idx + 1:self.tau_idx + 1 + q]
        self.ztsk = v[n + p:]

    @property
    def tau(self) -> float:
        return float(self.vec[self.tau_idx])

    @tau.setter
    def tau(self, val):
        self.vec[self.tau_idx] = val

    @property
    def kap(self) -> float:
        return float(self.vec[-1])

    @kap.setter
    def kap(self, val):
        self.vec[-1] = val

    def primal_dual(self, dual_mask):
        """(primal, dual) q-vectors: primal_k = s_k (z_k for dual-barrier cones).

        reference: point.jl:46-51 (primal_views / dual_views).  dual_mask is a boolean
        q-vector that is True on rows of dual-barrier cones, or None."""
        if dual_mask is None or not dual_mask.any():
            return self.s, self.z
        return np.where(dual_mask, self.z, self.s), np.where(dual_mask, self.s, self.z)


class SubPoint:
    """(x, y, z) sub-vector used by the 3x3 subsystem (reference: common.jl:184-208)."""

    def __init__(self, n, p, q):
        self.n, self.p, self.q = n, p, q
        self.vec = np.zeros(n + p + q)
        self.x = self.vec[:n]
        self.y = self.vec[n:n + p]
        self.z = self.vec[n + p:]

"""Batched cone-oracle interface the host driver is written against.

The reference calls per-cone methods in `for k in eachindex(cones)` loops
(reference: src/Cones/Cones.jl:27-310 for the API; callers in
src/Solvers/steppers/common.jl:7-118 and src/Solvers/search.jl:74-138).  On a
GPU one launch per cone *type* replaces that loop, so the boundary is batched
("plugin slot 2"): every method below acts on all K cones
at once, on q-vectors laid out cone after cone exactly like the z/s blocks of a
Point.  Two implementations exist:

  * hypatia_b200.cones.DeviceConeBlock  - the product: C-ABI calls into the
    sm_100a kernels (hyp_cones_*).
  * oracle.cones.OracleConeBlock        - test infrastructure: loops over CPU
    restatements of the reference's per-cone code.
"""
from __future__ import annotations

import numpy as np


class ConeBlock:
    def __init__(self, model):
        self.specs = list(model.cones)
        self.K = len(self.specs)
        self.q = model.q
        self.offsets = np.asarray(model.cone_offsets, dtype=np.int64)
        self.dims = np.asarray(model.cone_dims, dtype=np.int64)
        self.nus = np.asarray(model.cone_nus, dtype=np.float64)
        dual = np.zeros(self.q, dtype=bool)
        for ck, sl in zip(self.specs, model.cone_idxs):
            if ck.use_dual:
                dual[sl] = True
        self.dual_mask = dual if dual.any() else None

    # ---- per-cone segmented reductions used by the stepper / line search ----
    def seg_dot(self, a, b):
        """dot(a_k, b_k) for every cone k (K-vector)."""
        if self.K == 0:
            return np.zeros(0)
        return np.add.reduceat(a * b, self.offsets)

    def expand(self, per_cone):
        """Broadcast a K-vector to a q-vector (one value per cone row)."""
        return np.repeat(per_cone, self.dims)

    # ---- oracle surface (implemented by subclasses) ----
    def load_point(self, primal, dual, scal=1.0):
        raise NotImplementedError

    def is_feas(self):
        raise NotImplementedError

    def is_dual_feas(self):
        raise NotImplementedError

    def grad(self):
        raise NotImplementedError

    def hess_prod(self, arr):
        raise NotImplementedError

    def inv_hess_prod(self, arr):
        raise NotImplementedError

    def block_hess_prod(self, arr):
        """hess_prod for primal-barrier cones, inv_hess_prod for dual-barrier cones
        (reference: qrchol.jl:87-98)."""
        raise NotImplementedError

    def use_dder3(self):
        """Boolean K-vector (reference: Cones.jl:126)."""
        raise NotImplementedError

    def dder3(self, direction):
        raise NotImplementedError

    def check_numerics(self, irtmu=None, use_max_prox=None):
        """irtmu / use_max_prox let a batched backend answer check_numerics and get_proxsqr
        (always called back to back, search.jl:126-128) from one device sweep."""
        raise NotImplementedError

    def get_proxsqr(self, irtmu, use_max_prox):
        raise NotImplementedError

    def initial_point(self):
        raise NotImplementedError

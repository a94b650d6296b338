"""Device cone oracles behind the reference's Cone API (plugin slot 2 of SURVEY.md 8(b)).

`DeviceConeBlock` implements host.coneblock.ConeBlock by calling the batched C entry points
hyp_cones_* (include/hypatia_b200.h); method names and argument meaning follow the reference's
per-cone functions (src/Cones/Cones.jl:34-134): load_point, is_feas, is_dual_feas, grad,
hess_prod!, inv_hess_prod!, sqrt_hess_prod!, inv_sqrt_hess_prod!, dder3, get_proxsqr,
check_numerics - applied to every cone block of a q-vector at once.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .host import instances as _inst
from .host.coneblock import ConeBlock
from .host.models import Model

PROD_HESS, PROD_INV_HESS, PROD_SQRT_HESS, PROD_INV_SQRT_HESS, PROD_BLOCK = range(5)


class DeviceConeBlock(ConeBlock):
    def __init__(self, model, ctx: capi.Context | None = None, device: int = 0):
        super().__init__(model)
        self._own = ctx is None
        if ctx is None:
            # stand-alone oracle container (e.g. initialize_cone_point, Solvers.jl:530-548):
            # a context that holds the cone table only (no G columns)
            ctx = capi.Context(device)
            stub = Model(np.zeros(0), None, np.zeros(0), np.zeros((model.q, 0)), np.zeros(model.q),
                         model.cones)
            ctx.load_model(stub)
        self.ctx = ctx
        self._prox_cache = None

    def free(self):
        if self._own and self.ctx is not None:
            self.ctx.close()
        self.ctx = None

    # ---- state ----
    def load_point(self, primal, dual, scal=1.0):
        self.point = scal * np.asarray(primal, dtype=np.float64)
        self.ctx.cones_load_point(np.ascontiguousarray(primal, dtype=np.float64),
                                  np.ascontiguousarray(dual, dtype=np.float64), scal)
        self._feas = None
        self._prox_cache = None

    def _flags(self):
        if self._feas is None:
            self._feas = self.ctx.cones_feas()
        return self._feas

    def is_feas(self):
        return self._flags()[0]

    def is_dual_feas(self):
        return self._flags()[1]

    def grad(self):
        return self.ctx.cones_grad()

    def hess(self):
        """hess(cone_k) for every cone (list of dense blocks; Cones.jl:79-84)."""
        return self.ctx.cones_hess_blocks(False, self.dims)

    def inv_hess(self):
        """inv_hess(cone_k) for every cone (Cones.jl:86-93)."""
        return self.ctx.cones_hess_blocks(True, self.dims)

    # ---- products ----
    def hess_prod(self, arr):
        return self.ctx.cones_hess_prod(arr, PROD_HESS)

    def inv_hess_prod(self, arr):
        return self.ctx.cones_hess_prod(arr, PROD_INV_HESS)

    def sqrt_hess_prod(self, arr):
        return self.ctx.cones_hess_prod(arr, PROD_SQRT_HESS)

    def inv_sqrt_hess_prod(self, arr):
        return self.ctx.cones_hess_prod(arr, PROD_INV_SQRT_HESS)

    def block_hess_prod(self, arr):
        return self.ctx.cones_hess_prod(arr, PROD_BLOCK)

    def use_dder3(self):
        return np.ones(self.K, dtype=bool)

    def dder3(self, direction):
        return self.ctx.cones_dder3(direction)

    # ---- line-search oracles: one device sweep yields both answers ----
    def _prox(self, irtmu, use_max):
        key = (float(irtmu), bool(use_max))
        if self._prox_cache is None or self._prox_cache[0] != key:
            self._prox_cache = (key, self.ctx.cones_proxsqr(irtmu, use_max))
        return self._prox_cache[1]

    def check_numerics(self, irtmu=1.0, use_max=True):
        return self._prox(irtmu, use_max)[1]

    def get_proxsqr(self, irtmu, use_max_prox):
        return self._prox(irtmu, use_max_prox)[0]

    def initial_point(self):
        """set_initial_point! for every cone (closed forms; host-side, not on the hot path)."""
        out = np.zeros(self.q)
        for spec, sl in zip(self.specs, self._slices()):
            out[sl] = _inst.cone_initial_point(spec)
        return out

    def _slices(self):
        return [slice(int(o), int(o + d)) for o, d in zip(self.offsets, self.dims)]

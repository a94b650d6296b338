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


class DeviceCone:
    """ONE device cone behind the reference's per-cone oracle API (src/Cones/Cones.jl:34-310): the Python mirror of
    `B200Cone <: Cones.Cone{Float64}` in julia/HypatiaB200.jl, bound to the per-cone single-block entry points
    hyp_cone_* (SURVEY.md 8(b)).  Method names, argument meaning and the lazy evaluation (loads only copy; the first
    query after a load / reset_data evaluates) are the reference's.  `spec` is a host.models cone spec."""

    def __init__(self, spec, device: int = 0):
        self.lib = capi.load_library()
        alpha = np.ascontiguousarray(getattr(spec, "alpha", ()), dtype=np.float64)
        self.h = self.lib.hyp_cone_create(int(device), int(spec.ctype), int(spec.dim), int(bool(spec.use_dual)),
                                          int(getattr(spec, "hkind", 0)), float(getattr(spec, "hparam", 0.0)),
                                          capi.ptr(alpha) if alpha.size else None, int(alpha.size))
        if not self.h:
            raise capi.HypatiaB200Error(
                f"hyp_cone_create({capi_name(spec)}) failed: no sm_100 CUDA device, or a cone the library refuses "
                "(no CPU fallback)")
        self.dim = int(self.lib.hyp_cone_dimension(self.h))
        self.use_dual_barrier = bool(self.lib.hyp_cone_use_dual_barrier(self.h))

    def free(self):
        if getattr(self, "h", None):
            self.lib.hyp_cone_destroy(self.h)
            self.h = None

    __del__ = free

    def _check(self, rc, what):
        if rc < 0:
            raise capi.HypatiaB200Error(f"{what}: {self.lib.hyp_cone_last_error(self.h).decode()}")
        return rc

    def _vec(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.shape == (self.dim,)
        return v

    # ---- Cones.jl:34-41,138 ----
    def dimension(self):
        return self.dim

    def get_nu(self):
        return float(self.lib.hyp_cone_nu(self.h))

    # ---- Cones.jl:157-186 ----
    def load_point(self, point, scal=1.0):
        v = self._vec(point)        # bound to a local: the (possibly copied) array must outlive the call
        self._check(self.lib.hyp_cone_load_point(self.h, capi.ptr(v), float(scal)), "hyp_cone_load_point")

    def load_dual_point(self, point):
        v = self._vec(point)
        self._check(self.lib.hyp_cone_load_dual_point(self.h, capi.ptr(v)), "hyp_cone_load_dual_point")

    def reset_data(self):
        self._check(self.lib.hyp_cone_reset_data(self.h), "hyp_cone_reset_data")

    # ---- Cones.jl:56-93 ----
    def _feas(self):
        import ctypes as C
        f, d = C.c_int(0), C.c_int(0)
        self._check(self.lib.hyp_cone_is_feas(self.h, C.byref(f), C.byref(d)), "hyp_cone_is_feas")
        return bool(f.value), bool(d.value)

    def is_feas(self):
        return self._feas()[0]

    def is_dual_feas(self):
        return self._feas()[1]

    def grad(self):
        g = np.empty(self.dim)
        self._check(self.lib.hyp_cone_grad(self.h, capi.ptr(g)), "hyp_cone_grad")
        return g

    def _hess(self, inverse):
        H = np.zeros((self.dim, self.dim), order="F")
        self._check(self.lib.hyp_cone_hess(self.h, capi.ptr(H), int(inverse)), "hyp_cone_hess")
        return H

    def hess(self):
        return self._hess(False)

    def inv_hess(self):
        return self._hess(True)

    # ---- Cones.jl:101-118,189-218 ----
    def _prod(self, arr, mode):
        a = np.asarray(arr, dtype=np.float64)
        was1d = a.ndim == 1
        a2 = np.asfortranarray(a.reshape(self.dim, -1, order="F"))
        prod = np.empty_like(a2, order="F")
        self._check(self.lib.hyp_cone_hess_prod(self.h, capi.ptr(prod), capi.ptr(a2), a2.shape[1], self.dim, self.dim,
                                                int(mode)), "hyp_cone_hess_prod")
        return prod[:, 0] if was1d else prod

    def hess_prod(self, arr):
        return self._prod(arr, PROD_HESS)

    hess_prod_slow = hess_prod

    def inv_hess_prod(self, arr):
        return self._prod(arr, PROD_INV_HESS)

    def use_sqrt_hess_oracles(self, arr_dim=None):
        return bool(self.lib.hyp_cone_use_sqrt_hess_oracles(self.h))

    def sqrt_hess_prod(self, arr):
        return self._prod(arr, PROD_SQRT_HESS)

    def inv_sqrt_hess_prod(self, arr):
        return self._prod(arr, PROD_INV_SQRT_HESS)

    # ---- Cones.jl:120-134 ----
    def use_dder3(self):
        return True

    def dder3(self, direction):
        out = np.empty(self.dim)
        d = self._vec(direction)
        self._check(self.lib.hyp_cone_dder3(self.h, capi.ptr(out), capi.ptr(d)), "hyp_cone_dder3")
        return out

    # ---- Cones.jl:273-310 ----
    def _prox(self, irtmu, use_max_prox):
        import ctypes as C
        p, ok = C.c_double(0.0), C.c_int(0)
        self._check(self.lib.hyp_cone_proxsqr(self.h, float(irtmu), int(bool(use_max_prox)), C.byref(p), C.byref(ok)),
                    "hyp_cone_proxsqr")
        return p.value, bool(ok.value)

    def check_numerics(self):
        return self._prox(1.0, True)[1]

    def get_proxsqr(self, irtmu, use_max_prox):
        return self._prox(irtmu, use_max_prox)[0]


def capi_name(spec):
    from .host.models import CONE_NAMES
    return f"{CONE_NAMES.get(spec.ctype, spec.ctype)}({spec.dim})"

"""ctypes binding of libhypatia_b200.so (include/hypatia_b200.h).

This is the same C ABI a Julia `ccall` shim binds (julia/HypatiaB200.jl); nothing here computes.
The library is mandatory: importing this module raises if the .so is missing, and creating a
context raises if there is no B200-class CUDA device - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)

# symbols declared in include/hypatia_b200.h: (restype, argtypes)
_SIGS = {
    "hyp_version": (C.c_int, []),
    "hyp_create": (C.c_void_p, [C.c_int]),
    "hyp_destroy": (None, [C.c_void_p]),
    "hyp_last_error": (C.c_char_p, [C.c_void_p]),
    "hyp_stream": (C.c_void_p, [C.c_void_p]),
    "hyp_sync": (C.c_int, [C.c_void_p]),
    "hyp_comm_unique_id": (C.c_int, [C.c_char_p]),
    "hyp_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "hyp_load_model": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 c_ip, c_i64p, c_ip, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hyp_set_cone_params": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_dp]),
    "hyp_set_cone_alpha": (C.c_int, [C.c_void_p, C.c_int, c_i64p, c_dp]),
    "hyp_cones_load_point": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "hyp_cones_feas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_cones_grad": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hyp_cones_hess_prod": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_int]),
    "hyp_cones_dder3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_cones_proxsqr": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "hyp_cones_hess_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hyp_cone_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int64]),
    "hyp_cone_destroy": (None, [C.c_void_p]),
    "hyp_cone_last_error": (C.c_char_p, [C.c_void_p]),
    "hyp_cone_dimension": (C.c_int64, [C.c_void_p]),
    "hyp_cone_nu": (C.c_double, [C.c_void_p]),
    "hyp_cone_use_dual_barrier": (C.c_int, [C.c_void_p]),
    "hyp_cone_load_point": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double]),
    "hyp_cone_load_dual_point": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hyp_cone_reset_data": (C.c_int, [C.c_void_p]),
    "hyp_cone_is_feas": (C.c_int, [C.c_void_p, c_ip, c_ip]),
    "hyp_cone_grad": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hyp_cone_hess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hyp_cone_hess_prod": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "hyp_cone_use_sqrt_hess_oracles": (C.c_int, [C.c_void_p]),
    "hyp_cone_dder3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_cone_proxsqr": (C.c_int, [C.c_void_p, C.c_double, C.c_int, c_dp, c_ip]),
    "hyp_set_syssolver": (C.c_int, [C.c_void_p, C.c_int]),
    "hyp_set_syrk_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "hyp_set_column_sharding": (C.c_int, [C.c_void_p, C.c_int]),
    "hyp_set_mu_tau": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "hyp_update_lhs": (C.c_int, [C.c_void_p, c_ip]),
    "hyp_solve_subsystem3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_solve_system": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_apply_lhs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_solve_system_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64]),
    "hyp_apply_lhs_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64]),
    "hyp_calc_residuals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hyp_get_schur": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "hyp_launch_count": (C.c_int64, [C.c_void_p]),
    "hyp_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "hyp_timing_get": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_i64p]),
    "hyp_timing_reset": (C.c_int, [C.c_void_p]),
    "hyp_timing_slots": (C.c_int, []),
    "hyp_timing_name": (C.c_char_p, [C.c_int]),
    "hyp_test_atb_upper": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_double]),
    "hyp_test_gemm_tn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_double]),
    "hyp_test_potrf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, c_ip]),
    "hyp_test_panel_clocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "hyp_test_set_trsv_pkt": (C.c_int, [C.c_int]),
    "hyp_test_potrs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "hyp_test_gemv": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    "hyp_test_i8_gemm_tn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    "hyp_test_ozaki_slices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "hyp_test_mma_rate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hyp_test_ozaki_syrk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_int64]),
    "hyp_test_ldlt_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, c_ip]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


class HypatiaB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load_library():
    """dlopen the in-tree library (building it is `__graft_entry__.build()`'s job)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise HypatiaB200Error(
            f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(libhypatia_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def ptr(a):
    """Raw address of a numpy array / torch tensor / int (device pointer) / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


def _f64(a):
    a = np.asarray(a, dtype=np.float64)
    return a if a.flags.c_contiguous or a.flags.f_contiguous else np.ascontiguousarray(a)


class Context:
    """One `hyp_ctx` (one GPU).  Methods are 1:1 with the C entry points."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = self.lib.hyp_create(int(device))
        if not self.h:
            raise HypatiaB200Error(
                f"hyp_create({device}) failed: no sm_100 CUDA device visible (no CPU fallback)")
        self.device = device
        self.nranks = 1
        self.rank = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.hyp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc < 0:
            raise HypatiaB200Error(f"{what}: {self.lib.hyp_last_error(self.h).decode()}")
        return rc

    # ---- multi-GPU ----
    def comm_init(self, rank, nranks, uid: bytes):
        self.check(self.lib.hyp_comm_init(self.h, rank, nranks, uid), "hyp_comm_init")
        self.rank, self.nranks = rank, nranks

    # ---- model ----
    def load_model(self, model, G_local=None, cone_lo=0, cone_hi=None, Ap_Q=None, Ap_R=None):
        K = len(model.cones)
        cone_hi = K if cone_hi is None else cone_hi
        ctype = np.array([ck.ctype for ck in model.cones], dtype=np.int32)
        cdim = np.array([ck.dim for ck in model.cones], dtype=np.int64)
        cdual = np.array([1 if ck.use_dual else 0 for ck in model.cones], dtype=np.int32)
        if G_local is None:
            lo = int(model.cone_offsets[cone_lo]) if cone_lo < K else model.q
            hi = int(model.cone_offsets[cone_hi]) if cone_hi < K else model.q
            G_local = model.G[lo:hi]
        if isinstance(G_local, np.ndarray):
            G_local = np.asfortranarray(G_local, dtype=np.float64)
            ldG = max(G_local.shape[0], 1)
        else:                       # torch tensor holding the column-major panel as (n, q_local) rows
            ldG = int(G_local.shape[1])
        A = np.asfortranarray(model.A, dtype=np.float64)
        Q = None if Ap_Q is None else np.asfortranarray(Ap_Q, dtype=np.float64)
        R = None if Ap_R is None else np.asfortranarray(Ap_R, dtype=np.float64)
        hkind = np.array([getattr(ck, "hkind", 0) for ck in model.cones], dtype=np.int32)
        hparam = np.array([getattr(ck, "hparam", 0.0) for ck in model.cones], dtype=np.float64)
        self.check(self.lib.hyp_set_cone_params(self.h, K, hkind.ctypes.data_as(c_ip),
                                                hparam.ctypes.data_as(c_dp)), "hyp_set_cone_params")
        alphas = [np.asarray(getattr(ck, "alpha", ()), dtype=np.float64) for ck in model.cones]
        aoff = np.concatenate(([0], np.cumsum([a.size for a in alphas]))).astype(np.int64)
        aval = np.concatenate(alphas) if K and aoff[-1] else np.zeros(1)
        self.check(self.lib.hyp_set_cone_alpha(self.h, K, aoff.ctypes.data_as(c_i64p), aval.ctypes.data_as(c_dp)),
                   "hyp_set_cone_alpha")
        self._keep = (G_local, A, Q, R, ctype, cdim, cdual)
        c, b, h = _f64(model.c), _f64(model.b), _f64(model.h)
        rc = self.lib.hyp_load_model(
            self.h, model.n, model.p, model.q, ptr(G_local), ldG, ptr(A) if model.p else None,
            max(model.p, 1), ptr(c), ptr(b), ptr(h), K,
            ctype.ctypes.data_as(c_ip), cdim.ctypes.data_as(c_i64p), cdual.ctypes.data_as(c_ip),
            cone_lo, cone_hi, ptr(Q), ptr(R))
        self.check(rc, "hyp_load_model")
        self._keep = None
        self.n, self.p, self.q, self.K = model.n, model.p, model.q, K

    # ---- cones ----
    def cones_load_point(self, primal, dual, scal=1.0):
        self.check(self.lib.hyp_cones_load_point(self.h, ptr(primal), ptr(dual), float(scal)),
                   "hyp_cones_load_point")

    def cones_feas(self):
        f = np.zeros(self.K, dtype=np.uint8)
        d = np.zeros(self.K, dtype=np.uint8)
        self.check(self.lib.hyp_cones_feas(self.h, ptr(f), ptr(d)), "hyp_cones_feas")
        return f.astype(bool), d.astype(bool)

    def cones_grad(self, out=None):
        out = np.empty(self.q) if out is None else out
        self.check(self.lib.hyp_cones_grad(self.h, ptr(out)), "hyp_cones_grad")
        return out

    def cones_hess_prod(self, arr, mode, out=None):
        a = np.asarray(arr, dtype=np.float64)
        was1d = a.ndim == 1
        a2 = np.asfortranarray(a.reshape(self.q, -1, order="F"))
        ncols = a2.shape[1]
        prod = np.empty_like(a2, order="F") if out is None else out
        ld = max(self.q, 1)
        self.check(self.lib.hyp_cones_hess_prod(self.h, ptr(prod), ptr(a2), ncols, ld, ld, int(mode)),
                   "hyp_cones_hess_prod")
        return prod[:, 0] if was1d else prod

    def cones_dder3(self, direction):
        d = _f64(direction)
        out = np.empty(self.q)
        self.check(self.lib.hyp_cones_dder3(self.h, ptr(out), ptr(d)), "hyp_cones_dder3")
        return out

    def cones_proxsqr(self, irtmu, use_max):
        prox = np.zeros(self.K)
        ok = np.zeros(self.K, dtype=np.uint8)
        self.check(self.lib.hyp_cones_proxsqr(self.h, float(irtmu), int(bool(use_max)), ptr(prox),
                                              ptr(ok)), "hyp_cones_proxsqr")
        return prox, ok.astype(bool)

    def cones_hess_blocks(self, inverse: bool, dims):
        """Explicit hess / inv_hess of every cone: list of (dim_k, dim_k) arrays."""
        dims = [int(d) for d in dims]
        buf = np.zeros(sum(d * d for d in dims))
        self.check(self.lib.hyp_cones_hess_blocks(self.h, ptr(buf), int(bool(inverse))), "hyp_cones_hess_blocks")
        out, o = [], 0
        for d in dims:
            out.append(buf[o:o + d * d].reshape(d, d, order="F"))
            o += d * d
        return out

    # ---- system solver ----
    def set_syssolver(self, kind: int):
        self.check(self.lib.hyp_set_syssolver(self.h, int(kind)), "hyp_set_syssolver")

    def set_column_sharding(self, on: bool = True):
        """Single-giant-cone models: every rank loads ALL rows, the Schur assembly is split by columns."""
        self.check(self.lib.hyp_set_column_sharding(self.h, 1 if on else 0), "hyp_set_column_sharding")

    def set_syrk_mode(self, mode: int):
        self.check(self.lib.hyp_set_syrk_mode(self.h, int(mode)), "hyp_set_syrk_mode")

    def set_mu_tau(self, mu, tau):
        self.check(self.lib.hyp_set_mu_tau(self.h, float(mu), float(tau)), "hyp_set_mu_tau")

    def update_lhs(self):
        kind = C.c_int(0)
        rc = self.check(self.lib.hyp_update_lhs(self.h, C.byref(kind)), "hyp_update_lhs")
        return rc, kind.value

    def solve_subsystem3(self, sol, rhs):
        self.check(self.lib.hyp_solve_subsystem3(self.h, ptr(sol), ptr(rhs)), "hyp_solve_subsystem3")

    def solve_system(self, sol, rhs):
        self.check(self.lib.hyp_solve_system(self.h, ptr(sol), ptr(rhs)), "hyp_solve_system")

    def apply_lhs(self, res, direction):
        self.check(self.lib.hyp_apply_lhs(self.h, ptr(res), ptr(direction)), "hyp_apply_lhs")

    def _ld(self, a, ncols):
        dim6 = self.n + self.p + 2 * self.q + 2
        if hasattr(a, "stride") and not isinstance(a, np.ndarray):
            return int(a.stride(0)) if ncols > 1 else dim6
        return int(a.strides[0] // 8) if ncols > 1 else dim6

    def has_multi(self):
        return True

    def solve_system_multi(self, sol, rhs, ncols):
        """sol / rhs: (ncols, n+p+2q+2) row-major arrays or tensors - one Point per row."""
        self.check(self.lib.hyp_solve_system_multi(self.h, ptr(sol), ptr(rhs), int(ncols), self._ld(rhs, ncols)),
                   "hyp_solve_system_multi")

    def apply_lhs_multi(self, res, direction, ncols):
        self.check(self.lib.hyp_apply_lhs_multi(self.h, ptr(res), ptr(direction), int(ncols), self._ld(direction, ncols)),
                   "hyp_apply_lhs_multi")

    def calc_residuals(self, point_vec):
        """(x_residual, y_residual, z_residual, stats[10]) of calc_convergence_params (Solvers.jl:425-483)."""
        xr, yr, zr, st = np.zeros(self.n), np.zeros(self.p), np.zeros(self.q), np.zeros(10)
        self.check(self.lib.hyp_calc_residuals(self.h, ptr(point_vec), ptr(xr), ptr(yr), ptr(zr), ptr(st)),
                   "hyp_calc_residuals")
        return xr, yr, zr, st

    def get_schur(self):
        m = self.n - self.p
        S = np.zeros((m, m), order="F")
        self.check(self.lib.hyp_get_schur(self.h, ptr(S), max(m, 1)), "hyp_get_schur")
        return S

    def sync(self):
        self.check(self.lib.hyp_sync(self.h), "hyp_sync")

    def stream(self):
        return self.lib.hyp_stream(self.h)

    def launch_count(self):
        return int(self.lib.hyp_launch_count(self.h))

    # ---- timers ----
    def timing_enable(self, on=True):
        self.lib.hyp_timing_enable(self.h, int(on))

    def timing_reset(self):
        self.lib.hyp_timing_reset(self.h)

    def timing(self):
        out = {}
        for s in range(self.lib.hyp_timing_slots()):
            ms, n = C.c_double(0), C.c_int64(0)
            self.lib.hyp_timing_get(self.h, s, C.byref(ms), C.byref(n))
            out[self.lib.hyp_timing_name(s).decode()] = (ms.value, n.value)
        return out


def comm_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.hyp_comm_unique_id(buf) != 0:
        raise HypatiaB200Error("hyp_comm_unique_id failed (NCCL not loadable)")
    return buf.raw

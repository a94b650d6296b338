"""Host emulation of the CUDA kernel headers (hypatia.jl_b200/csrc/*_kernels.cuh) for the CPU-only
test tier: tests/emu/ compiles the very same __global__ functions with g++ (one pthread per CUDA
thread) into tests/emu/_build/libemu.so; the wrappers below mirror the launch sequences of the
host code in csrc/ so that the device arithmetic is checked against the oracle without a GPU.
Test infrastructure only - the product library has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
BUILD = os.path.join(EMU, "_build")
# HYP_EMU_ASAN=1 builds the emulation with AddressSanitizer + UBSan (run the tests with
# LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0; see tools/emu_asan.sh)
ASAN = os.environ.get("HYP_EMU_ASAN", "0") == "1"
# HYP_EMU_TSAN=1: ThreadSanitizer build - one pthread per CUDA thread, so a missing __syncthreads / __syncwarp
# between a shared-memory write and another thread's read is a reported data race (tools/emu_tsan.sh)
TSAN = os.environ.get("HYP_EMU_TSAN", "0") == "1"
LIB = os.path.join(BUILD, "libemu_asan.so" if ASAN else "libemu_tsan.so" if TSAN else "libemu.so")
CSRC = os.path.join(os.path.dirname(HERE), "hypatia.jl_b200", "csrc")
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    os.makedirs(BUILD, exist_ok=True)
    srcs = [os.path.join(EMU, f) for f in ("emu_kernels.cpp", "cuda_emu.cpp")]
    deps = srcs + [os.path.join(EMU, "cuda_emu.h")] + \
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith("_kernels.cuh") or f == "devdefs.cuh"]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        opt = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-DHYP_EMU_PLAIN_SHARED"] if ASAN else \
            ["-O1", "-g", "-fsanitize=thread", "-fno-omit-frame-pointer"] if TSAN else ["-O2"]
        # built under a private name and renamed into place: with pytest-xdist several workers may get here at once,
        # and none of them may dlopen a half-written file
        tmp = f"{LIB}.{os.getpid()}.tmp"
        cmd = ["g++"] + opt + ["-std=c++17", "-fPIC", "-shared", "-pthread", "-o", tmp] + srcs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + r.stderr)
        os.replace(tmp, LIB)
    _lib = C.CDLL(LIB)
    return _lib


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def i64(x):
    return C.c_int64(int(x))


class MatLayout:
    """Per-cone d x d blocks with an even leading dimension, as hyp_mat_alloc_group lays them out."""

    def __init__(self, sides):
        self.sides = np.asarray(sides, dtype=np.int32)
        self.lde = (self.sides + 1) & ~1
        sizes = self.lde.astype(np.int64) * self.sides
        self.moff = np.concatenate(([0], np.cumsum(sizes)))[:-1].astype(np.int64)
        self.total = int(sizes.sum())

    def get(self, buf, c):
        d, lde = int(self.sides[c]), int(self.lde[c])
        return buf[self.moff[c]:self.moff[c] + lde * d].reshape(lde, d, order="F")[:d]


def syevj(layout, Ain, want_vectors=True, divv=None, div_off=None, div_idx=0, smem=True, threads=64):
    n = len(layout.sides)
    lam_off = np.concatenate(([0], np.cumsum(layout.sides)))[:-1].astype(np.int64)
    lam = np.zeros(int(layout.sides.sum()))
    V = np.zeros(layout.total) if want_vectors else None
    lib().emu_syevj(n, int(layout.sides.max()), p(layout.sides), p(layout.moff), p(Ain), p(V), p(lam_off), p(lam),
                    p(divv), p(div_off), div_idx, int(smem), threads)
    return lam, lam_off, V


class EmuSpecGroup:
    """Mirrors hyp_spec_update_state / hyp_spec_prod / hyp_spec_dder3 (csrc/cones_spec.cu) for a list
    of EpiPerSepSpectral{MatrixCSqr} cones laid out one after the other in a q-vector; the two
    congruences around the middle kernel are NumPy products here (TMA + DMMA GEMMs on the device)."""

    def __init__(self, specs, threads=64):
        self.specs = specs
        self.K = len(specs)
        self.threads = threads
        self.dims = np.array([s.dim for s in specs], dtype=np.int64)
        self.off = np.concatenate(([0], np.cumsum(self.dims)))[:-1].astype(np.int64)
        self.q = int(self.dims.sum())
        self.lay = MatLayout([s.side for s in specs])
        d = self.lay.sides.astype(np.int64)
        self.voff = np.concatenate(([0], np.cumsum(8 * d)))[:-1].astype(np.int64)
        self.voff7 = self.voff + 7 * d
        self.kidx = np.arange(self.K, dtype=np.int32)
        self.hkind = np.array([s.hkind for s in specs], dtype=np.int32)
        self.hparam = np.array([s.hparam for s in specs], dtype=np.float64)
        self.vecs = np.zeros(int(8 * d.sum()))
        self.scal = np.zeros(8 * self.K)

    def load_point(self, point, dual):
        L = lib()
        lay, K = self.lay, self.K
        self.point = np.ascontiguousarray(point, dtype=np.float64)
        self.dual = np.ascontiguousarray(dual, dtype=np.float64)
        self.feas = np.ones(K, dtype=np.uint8)
        self.dual_feas = np.ones(K, dtype=np.uint8)
        self.grad = np.full(self.q, np.nan)
        W = np.zeros(lay.total)
        Wc = np.zeros(lay.total)
        L.emu_unpack_state(K, p(self.off), p(lay.sides), p(lay.moff), 2, p(self.point), p(W), p(Wc), 2)
        for c in range(K):                          # Cholesky gate (hyp_chol_batched on the device)
            try:
                np.linalg.cholesky(lay.get(Wc, c))
            except np.linalg.LinAlgError:
                self.feas[c] = 0
        self.V = np.zeros(lay.total)
        L.emu_syevj(K, int(lay.sides.max()), p(lay.sides), p(lay.moff), p(W), p(self.V), p(self.voff), p(self.vecs),
                    p(self.point), p(self.off), 1, 1, self.threads)
        self.Vt = np.zeros(lay.total)
        self.theta = np.zeros(lay.total)
        self.Dh = np.zeros(lay.total)
        L.emu_spec_post(K, p(self.off), p(lay.sides), p(lay.moff), p(self.voff), p(self.kidx), p(self.hkind),
                        p(self.hparam), p(self.point), p(self.V), p(self.Vt), p(self.theta), p(self.Dh),
                        p(self.vecs), p(self.scal), p(self.grad), p(self.feas), self.threads)
        # dual feasibility
        Wd = np.zeros(lay.total)
        L.emu_unpack_state(K, p(self.off), p(lay.sides), p(lay.moff), 2, p(self.dual), p(Wd), None, 1)
        chol_ok = np.ones(K, dtype=np.uint8)
        for c in range(K):
            try:
                np.linalg.cholesky(lay.get(Wd, c))
            except np.linalg.LinAlgError:
                chol_ok[c] = 0
        L.emu_syevj(K, int(lay.sides.max()), p(lay.sides), p(lay.moff), p(Wd), None, p(self.voff7), p(self.vecs),
                    p(self.dual), p(self.off), 0, 1, self.threads)
        L.emu_spec_dualfeas(K, p(self.off), p(lay.sides), p(self.voff7), p(self.kidx), p(self.hkind),
                            p(self.hparam), p(self.dual), p(self.vecs), p(chol_ok), p(self.dual_feas))

    def _congr(self, Mall, X, d, lde, cc):
        """Y_j = X' M_j X in place for the cc matrices of Mall (what `congruence` does on the device)."""
        per = lde * lde
        for j in range(cc):
            Mj = Mall[j * per:(j + 1) * per].reshape(lde, lde, order="F")
            Mj[:d, :d] = X.T @ Mj[:d, :d] @ X

    def prod(self, arr, inverse):
        L = lib()
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(self.q, -1, order="F"))
        ncols = a.shape[1]
        out = np.full_like(a, np.nan, order="F")
        for c in range(self.K):
            d, lde = int(self.lay.sides[c]), int(self.lay.lde[c])
            ln = d * (d + 1) // 2
            Mall = np.zeros(lde * lde * ncols)
            a0 = a[self.off[c]:, :]
            base_a = a.ctypes.data + 8 * int(self.off[c])
            base_o = out.ctypes.data + 8 * int(self.off[c])
            L.emu_unpack_cols(d, lde, i64(ln), C.c_void_p(base_a + 16), i64(self.q), i64(ncols), p(Mall), 2)
            self._congr(Mall, self.lay.get(self.V, c), d, lde, ncols)
            vec_c = self.vecs[self.voff[c]:]
            L.emu_spec_mid(int(inverse), d, lde, p(self.scal[8 * c:]), p(vec_c), p(self.theta[self.lay.moff[c]:]),
                           p(self.Dh[self.lay.moff[c]:]), p(Mall), C.c_void_p(base_a), i64(self.q),
                           C.c_void_p(base_o), i64(self.q), i64(ncols), min(ncols, 3), self.threads)
            self._congr(Mall, self.lay.get(self.Vt, c), d, lde, ncols)
            L.emu_pack_cols(d, lde, i64(ln), p(Mall), i64(ncols), None, None, None, C.c_void_p(base_o + 16),
                            i64(self.q), 2)
            del a0
        return out[:, 0] if np.ndim(arr) == 1 else out

    def small_prod(self, arr, mode, in_place=False):
        """The fused few-column kernel (spec_small_prod_kernel): one launch for all cones and columns."""
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(self.q, -1, order="F")).copy(order="F")
        out = a if in_place else np.full_like(a, np.nan, order="F")
        lay = self.lay
        dualf = np.array([1 if s.use_dual else 0 for s in self.specs], dtype=np.int32)
        lib().emu_spec_small_prod(int(mode), self.K, int(lay.sides.max()), p(self.off), p(lay.sides), p(lay.moff),
                                  p(self.voff), p(dualf), p(self.V), p(self.Vt), p(self.theta), p(self.Dh),
                                  p(self.vecs), p(self.scal), p(a), i64(self.q), p(out), i64(self.q), i64(0),
                                  a.shape[1], self.threads)
        return out[:, 0] if np.ndim(arr) == 1 else out

    def dder3(self, direction):
        L = lib()
        dirv = np.ascontiguousarray(direction, dtype=np.float64)
        out = np.full(self.q, np.nan)
        for c in range(self.K):
            d, lde = int(self.lay.sides[c]), int(self.lay.lde[c])
            ln = d * (d + 1) // 2
            per = lde * lde
            E = np.zeros(per)
            X = np.zeros(per)
            OUT = np.zeros(per)
            base_d = dirv.ctypes.data + 8 * int(self.off[c])
            base_o = out.ctypes.data + 8 * int(self.off[c])
            L.emu_unpack_cols(d, lde, i64(ln), C.c_void_p(base_d + 16), i64(self.q), i64(1), p(E), 2)
            self._congr(E, self.lay.get(self.V, c), d, lde, 1)
            L.emu_spec_dder3(d, lde, p(self.scal[8 * c:]), p(self.vecs[self.voff[c]:]), p(self.Dh[self.lay.moff[c]:]),
                             p(E), p(X), p(OUT), C.c_void_p(base_d), C.c_void_p(base_o), self.threads)
            self._congr(OUT, self.lay.get(self.Vt, c), d, lde, 1)
            L.emu_pack_cols(d, lde, i64(ln), p(OUT), i64(1), None, None, None, C.c_void_p(base_o + 16), i64(self.q), 2)
        return out


class EmuVec3Group:
    """Mirrors the launches of cones.cu for a list of EpiPerSquare or HypoPerLog cones (one type)."""

    def __init__(self, specs):
        self.type = specs[0].ctype
        assert all(s.ctype == self.type for s in specs)
        self.K = len(specs)
        self.dims = np.array([s.dim for s in specs], dtype=np.int32)
        self.off = np.concatenate(([0], np.cumsum(self.dims)))[:-1].astype(np.int64)
        self.q = int(self.dims.sum())
        self.kidx = np.arange(self.K, dtype=np.int32)
        self.dualf = np.array([1 if s.use_dual else 0 for s in specs], dtype=np.int32)
        self.hkind = np.array([s.hkind for s in specs], dtype=np.int32)
        self.hparam = np.array([s.hparam for s in specs], dtype=np.float64)
        self.scal = np.zeros(8 * self.K)

    def load_point(self, point, dual):
        self.point = np.ascontiguousarray(point, dtype=np.float64)
        self.dual = np.ascontiguousarray(dual, dtype=np.float64)
        self.feas = np.ones(self.K, dtype=np.uint8)
        self.dual_feas = np.ones(self.K, dtype=np.uint8)
        self.grad = np.full(self.q, np.nan)
        lib().emu_v3_state(self.type, self.K, p(self.off), p(self.dims), p(self.kidx), p(self.hkind), p(self.hparam),
                           p(self.point), p(self.dual),
                           p(self.grad), p(self.scal), p(self.feas), p(self.dual_feas))

    def prod(self, arr, mode, in_place=False):
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(self.q, -1, order="F")).copy(order="F")
        out = a if in_place else np.full_like(a, np.nan, order="F")
        lib().emu_v3_prod(self.type, int(mode), self.K, p(self.off), p(self.dims), p(self.dualf), p(self.hkind),
                          p(self.hparam), p(self.scal),
                          p(self.point), p(a), i64(self.q), p(out), i64(self.q), i64(a.shape[1]), i64(0), 2)
        return out[:, 0] if np.ndim(arr) == 1 else out

    def dder3(self, direction):
        d = np.ascontiguousarray(direction, dtype=np.float64)
        out = np.full(self.q, np.nan)
        lib().emu_v3_dder3(self.type, self.K, p(self.off), p(self.dims), p(self.hkind), p(self.hparam), p(self.scal),
                           p(self.point), p(d), p(out))
        return out


class EmuMatGroup:
    """State of a group of PosSemidefTri / HypoPerLogdetTri / HypoRootdetTri cones laid out like
    hyp_mat_update_state leaves it (W, W^-1, U^-1, U' per cone, scal, svec(W^-1)), built here with NumPy from
    the oracle's cone objects, to drive the fused small-matrix product kernel (mat_small_prod_kernel)."""

    def __init__(self, specs, oracle_cones, point):
        self.type = specs[0].ctype
        self.K = len(specs)
        self.dims = np.array([s.dim for s in specs], dtype=np.int64)
        self.off = np.concatenate(([0], np.cumsum(self.dims)))[:-1].astype(np.int64)
        self.q = int(self.dims.sum())
        self.lay = MatLayout([s.side for s in specs])
        self.dualf = np.array([1 if s.use_dual else 0 for s in specs], dtype=np.int32)
        lead = {2: 0, 3: 2, 4: 1}[self.type]
        lay = self.lay
        self.W, self.Wi, self.Ui, self.Ut, self.Uit = (np.zeros(lay.total) for _ in range(5))
        self.scal = np.zeros(8 * self.K)
        self.wivec = np.zeros(self.q)
        self.point = np.ascontiguousarray(point, dtype=np.float64)
        from oracle import arrayutil as au
        for c, ck in enumerate(oracle_cones):
            ck.grad()
            Wm = np.array(ck.mat)
            U = np.linalg.cholesky(Wm).T
            lay.get(self.W, c)[:] = Wm
            lay.get(self.Wi, c)[:] = np.linalg.inv(Wm)
            lay.get(self.Ui, c)[:] = np.linalg.inv(U)
            lay.get(self.Ut, c)[:] = U.T
            lay.get(self.Uit, c)[:] = np.linalg.inv(U).T
            o = int(self.off[c])
            self.wivec[o + lead:o + self.dims[c]] = au.smat_to_svec(np.linalg.inv(Wm))
            if self.type == 3:
                self.scal[8 * c + 1], self.scal[8 * c + 2], self.scal[8 * c + 4] = ck.phi, ck.zeta, ck.point[1]
            elif self.type == 4:
                self.scal[8 * c + 1], self.scal[8 * c + 2], self.scal[8 * c + 5] = ck.phi, ck.zeta, ck.pzd

    def device_state(self, point, dual):
        """Replaces the NumPy-built state by the device pipeline of hyp_mat_update_state (sides <= 128): unpack ->
        batched Cholesky + inverse -> mat_post_kernel, and the dual-feasibility pass.  Returns (grad, feas, dual_feas)."""
        lay = self.lay
        self.point = np.ascontiguousarray(point, dtype=np.float64)
        dual = np.ascontiguousarray(dual, dtype=np.float64)
        self.W, self.Wi, self.Ui, self.Ut, self.Uit = (np.zeros(lay.total) for _ in range(5))
        U, U2, Ui2 = (np.zeros(lay.total) for _ in range(3))
        self.scal = np.zeros(8 * self.K)
        self.wivec = np.zeros(self.q)
        grad = np.zeros(self.q)
        kidx = np.arange(self.K, dtype=np.int32)
        feas, dfeas = np.ones(self.K, dtype=np.uint8), np.ones(self.K, dtype=np.uint8)
        lib().emu_mat_state(self.type, self.K, p(self.off), p(lay.sides), p(lay.moff), p(kidx), p(self.point), p(dual),
                            p(self.W), p(U), p(self.Ui), p(self.Ut), p(self.Uit), p(self.Wi), p(U2), p(Ui2), p(self.scal),
                            p(grad), p(self.wivec), p(feas), p(dfeas))
        return grad, feas, dfeas

    def dder3(self, direction, threads=64):
        d = np.ascontiguousarray(direction, dtype=np.float64)
        out = np.full(self.q, np.nan)
        lay = self.lay
        lib().emu_mat_small_dder3(self.type, self.K, int(lay.sides.max()), p(self.off), p(lay.sides), p(lay.moff),
                                  p(self.Ui), p(self.Uit), p(self.scal), p(d), p(out), threads)
        return out

    def prod(self, arr, mode, threads=64, in_place=False):
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(self.q, -1, order="F")).copy(order="F")
        out = a if in_place else np.full_like(a, np.nan, order="F")
        lay = self.lay
        lib().emu_mat_small_prod(self.type, int(mode), self.K, int(lay.sides.max()), p(self.off), p(lay.sides),
                                 p(lay.moff), p(self.dualf), p(self.W), p(self.Wi), p(self.Ui), p(self.Ut),
                                 p(self.scal), p(self.point), p(self.wivec), p(a), i64(self.q), p(out), i64(self.q),
                                 i64(0), a.shape[1], threads)
        return out[:, 0] if np.ndim(arr) == 1 else out


class EmuGpowGroup:
    """Mirrors hyp_gpow_update_state / hyp_gpow_prod / hyp_gpow_dder3 (csrc/cones_gpow.cu) for a list of
    GeneralizedPower cones, including the batched Cholesky + inverse of the explicit Hessians (chol_kernels.cuh)."""

    def __init__(self, specs):
        self.K = len(specs)
        self.hpm = specs[0].ctype == 12          # HypoPowerMean, else GeneralizedPower
        self.ens = specs[0].ctype == 14          # EpiNormSpectral: d1 per cone, workspace instead of powers
        self.etr = specs[0].ctype == 23          # EpiTrRelEntropyTri: state + dder3 scratch
        if self.etr:
            sizes = []
            for s in specs:
                d = int(round((np.sqrt(1 + 4 * (s.dim - 1)) - 1) / 2))
                n = d * d
                sizes.append(23 * n + 2 * d + 2 * d * n + n * n)
            self.voff = np.concatenate(([0], np.cumsum(sizes)))[:-1].astype(np.int64)
            self.vecs = np.zeros(int(sum(sizes)))
        self.sps = specs[0].ctype == 22          # PosSemidefTriSparse: packed pattern + workspace
        if self.sps:
            regions = [np.concatenate((np.asarray(s.alpha, dtype=np.float64), np.zeros(5 * int(s.alpha[0]) ** 2)))
                       for s in specs]
            self.voff = np.concatenate(([0], np.cumsum([r.size for r in regions])))[:-1].astype(np.int64)
            self.vecs = np.concatenate(regions)
        self.wone = specs[0].ctype == 21         # WSOSInterpEpiNormOne: R - 1 pair factorisations per P_k
        self.weuc = specs[0].ctype == 20         # WSOSInterpEpiNormEucl: like 19 with dim = U R and two more scratch blocks
        self.wpsd = specs[0].ctype == 19         # WSOSInterpPosSemidefTri: R per cone, packed Ps + workspace
        if self.wpsd or self.weuc or self.wone:
            self.Rs = np.array([s.hkind for s in specs], dtype=np.int32)
            regions = []
            for s in specs:
                Rr = s.hkind
                U = s.dim // (Rr if (self.weuc or self.wone) else Rr * (Rr + 1) // 2)
                nP = int(s.alpha[0])
                Ls = [int(x) for x in s.alpha[1:1 + nP]]
                if self.wone:
                    wsz = sum((Rr - 1) * (4 * L * U + 4 * L * L) for L in Ls) + 5 * U * U + 5 * max(Ls) ** 2 + \
                        5 * max(Ls) * U
                else:
                    wsz = sum(Rr * L * Rr * U + (Rr * L) ** 2 for L in Ls) + (Rr * U) ** 2 + (Rr * max(Ls)) ** 2 + \
                        Rr * max(Ls) * Rr * U + (U * U + max(Ls) ** 2 + max(Ls) * U if self.weuc else 0)
                regions.append(np.concatenate((np.asarray(s.alpha, dtype=np.float64), np.zeros(wsz))))
            self.voff = np.concatenate(([0], np.cumsum([r.size for r in regions])))[:-1].astype(np.int64)
            self.vecs = np.concatenate(regions)
        self.mep = specs[0].ctype == 18          # MatrixEpiPerSquare: d1 per cone, state + dder3 scratch
        if self.mep:
            self.d1 = np.array([s.hkind for s in specs], dtype=np.int32)
            sizes = []
            for s in specs:
                d1 = s.hkind
                d2 = (s.dim - d1 * (d1 + 1) // 2 - 1) // d1
                sizes.append(21 * d1 * d1 + 13 * d1 * d2 + 3 * d2 * d2)
            self.voff = np.concatenate(([0], np.cumsum(sizes)))[:-1].astype(np.int64)
            self.vecs = np.zeros(int(sum(sizes)))
        self.dnn = specs[0].ctype == 17          # DoublyNonnegativeTri: side per cone + workspace
        if self.dnn:
            self.sides = np.array([int(round((np.sqrt(1 + 8 * s.dim) - 1) / 2)) for s in specs], dtype=np.int32)
            sizes = [2 * int(sd) ** 2 for sd in self.sides]
            self.voff = np.concatenate(([0], np.cumsum(sizes)))[:-1].astype(np.int64)
            self.vecs = np.zeros(int(sum(sizes)))
        self.lmi = specs[0].ctype == 16          # LinMatrixIneq: packed As + workspace
        if self.lmi:
            regions = []
            for s in specs:
                sd = int(s.alpha[0])
                regions.append(np.concatenate((np.asarray(s.alpha, dtype=np.float64), np.zeros((s.dim + 3) * sd * sd))))
            self.voff = np.concatenate(([0], np.cumsum([r.size for r in regions])))[:-1].astype(np.int64)
            self.vecs = np.concatenate(regions)
        self.wsos = specs[0].ctype == 15         # WSOSInterpNonnegative: packed Ps + workspace
        if self.wsos:
            regions = []
            for s in specs:
                nP = int(s.alpha[0])
                Ls = [int(x) for x in s.alpha[1:1 + nP]]
                wsz = sum(L * s.dim + L * L for L in Ls) + max(Ls) ** 2
                regions.append(np.concatenate((np.asarray(s.alpha, dtype=np.float64), np.zeros(wsz))))
            self.voff = np.concatenate(([0], np.cumsum([r.size for r in regions])))[:-1].astype(np.int64)
            self.vecs = np.concatenate(regions)
        if self.ens:
            self.d1 = np.array([s.hkind for s in specs], dtype=np.int32)
            sizes = [3 * (s.dim - 1) + 2 * s.hkind ** 2 for s in specs]
            self.voff = np.concatenate(([0], np.cumsum(sizes)))[:-1].astype(np.int64)
            self.vecs = np.zeros(int(sum(sizes)))
        self.dims = np.array([s.dim for s in specs], dtype=np.int32)
        self.off = np.concatenate(([0], np.cumsum(self.dims)))[:-1].astype(np.int64)
        self.q = int(self.dims.sum())
        self.mu = np.array([len(s.alpha) for s in specs], dtype=np.int32)
        self.aoff = np.concatenate(([0], np.cumsum(self.mu)))[:-1].astype(np.int64)
        self.alpha = np.concatenate([np.asarray(s.alpha, dtype=np.float64) for s in specs] + [np.zeros(1)])
        self.kidx = np.arange(self.K, dtype=np.int32)
        self.dualf = np.array([1 if s.use_dual else 0 for s in specs], dtype=np.int32)
        self.lay = MatLayout(self.dims)
        self.scal = np.zeros(8 * self.K)

    def load_point(self, point, dual):
        self.point = np.ascontiguousarray(point, dtype=np.float64)
        self.dual = np.ascontiguousarray(dual, dtype=np.float64)
        self.feas = np.ones(self.K, dtype=np.uint8)
        self.dual_feas = np.ones(self.K, dtype=np.uint8)
        # outputs start out poisoned: on the device they hold the previous iterate's values (or whatever
        # cudaMalloc returned), so a kernel that accumulates into them or skips an entry must fail here
        self.grad = np.full(self.q, np.nan)
        self.H = np.full(self.lay.total, np.nan)
        self.scal[:] = np.nan
        if self.etr:
            lib().emu_etr_state(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.grad), p(self.scal), p(self.H), p(self.feas))
        elif self.sps:
            lib().emu_sps_state(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.wone:
            lib().emu_wone_state(self.K, p(self.off), p(self.dims), p(self.Rs), p(self.voff), p(self.vecs), p(self.kidx),
                                 p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.weuc:
            lib().emu_weuc_state(self.K, p(self.off), p(self.dims), p(self.Rs), p(self.voff), p(self.vecs), p(self.kidx),
                                 p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.wpsd:
            lib().emu_wpsd_state(self.K, p(self.off), p(self.dims), p(self.Rs), p(self.voff), p(self.vecs), p(self.kidx),
                                 p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.mep:
            lib().emu_mep_state(self.K, p(self.off), p(self.dims), p(self.d1), p(self.voff), p(self.vecs), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.dual), p(self.grad), p(self.scal), p(self.H),
                                p(self.feas), p(self.dual_feas))
        elif self.dnn:
            lib().emu_dnn_state(self.K, p(self.off), p(self.dims), p(self.sides), p(self.voff), p(self.vecs), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.lmi:
            lib().emu_lmi_state(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.wsos:
            lib().emu_wsos_state(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(self.kidx),
                                 p(self.lay.moff), p(self.point), p(self.grad), p(self.H), p(self.feas))
        elif self.ens:
            lib().emu_ens_state(self.K, p(self.off), p(self.dims), p(self.d1), p(self.voff), p(self.vecs), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.dual), p(self.grad), p(self.scal), p(self.H),
                                p(self.feas), p(self.dual_feas))
        elif self.hpm:
            lib().emu_hpm_state(self.K, p(self.off), p(self.dims), p(self.aoff), p(self.alpha), p(self.kidx),
                                p(self.lay.moff), p(self.point), p(self.dual), p(self.grad), p(self.scal), p(self.H),
                                p(self.feas), p(self.dual_feas))
        else:
            lib().emu_gpow_state(self.K, p(self.off), p(self.dims), p(self.mu), p(self.aoff), p(self.alpha),
                                 p(self.kidx), p(self.lay.moff), p(self.point), p(self.dual), p(self.grad),
                                 p(self.scal), p(self.H), p(self.feas), p(self.dual_feas))
        # hess_fact: the batched factor-and-invert kernel of chol.cu on a copy of the explicit Hessians
        self.U = self.H.copy()
        self.Ui = np.full(self.lay.total, np.nan)
        lib().emu_chol_batched(self.K, p(self.lay.sides), p(self.lay.moff), p(self.kidx), p(self.U), p(self.Ui),
                               p(self.feas))

    def prod(self, arr, mode, in_place=False):
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(self.q, -1, order="F")).copy(order="F")
        out = a if in_place else np.full_like(a, np.nan, order="F")
        hess_dual, inv_dual = {0: (-1, -2), 1: (-2, -1), 4: (0, 1), 5: (1, 0)}[int(mode)]
        L = lib()
        if hess_dual > -2 and self.etr:
            L.emu_etr_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(self.dualf),
                           p(self.scal), p(a), i64(self.q), p(out), i64(self.q), i64(a.shape[1]), i64(0))
        elif hess_dual > -2 and self.mep:
            L.emu_mep_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.d1), p(self.voff), p(self.vecs),
                           p(self.dualf), p(self.scal), p(self.point), p(a), i64(self.q), p(out), i64(self.q),
                           i64(a.shape[1]), i64(0))
        elif hess_dual > -2 and self.dnn:
            L.emu_dnn_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.sides), p(self.voff), p(self.vecs),
                           p(self.dualf), p(self.point), p(a), i64(self.q), p(out), i64(self.q), i64(a.shape[1]), i64(0))
        elif hess_dual > -2 and (self.wsos or self.lmi or self.wpsd or self.weuc or self.wone or self.sps):
            L.emu_gen_hess_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.lay.moff), p(self.dualf), p(self.H),
                                p(a), i64(self.q), p(out), i64(self.q), i64(a.shape[1]), i64(0))
        elif hess_dual > -2 and self.ens:
            L.emu_ens_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.d1), p(self.voff), p(self.vecs),
                           p(self.dualf), p(self.scal), p(self.point), p(a), i64(self.q), p(out), i64(self.q),
                           i64(a.shape[1]), i64(0))
        elif hess_dual > -2 and self.hpm:
            L.emu_hpm_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.aoff), p(self.alpha), p(self.dualf),
                           p(self.scal), p(self.point), p(a), i64(self.q), p(out), i64(self.q), i64(a.shape[1]), i64(0))
        elif hess_dual > -2:
            L.emu_gpow_prod(self.K, hess_dual, p(self.off), p(self.dims), p(self.mu), p(self.aoff), p(self.alpha),
                            p(self.dualf), p(self.scal), p(self.point), p(a), i64(self.q), p(out), i64(self.q),
                            i64(a.shape[1]), i64(0))
        if inv_dual > -2:
            L.emu_gen_invhess_prod(self.K, inv_dual, p(self.off), p(self.dims), p(self.lay.moff), p(self.dualf),
                                   p(self.Ui), p(a), i64(self.q), p(out), i64(self.q), i64(a.shape[1]), i64(0))
        return out[:, 0] if np.ndim(arr) == 1 else out

    def dder3(self, direction):
        d = np.ascontiguousarray(direction, dtype=np.float64)
        out = np.full(self.q, np.nan)
        if self.etr:
            lib().emu_etr_dder3(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(self.scal), p(d), p(out))
        elif self.sps:
            lib().emu_sps_dder3(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(d), p(out))
        elif self.wone:
            lib().emu_wone_dder3(self.K, p(self.off), p(self.dims), p(self.Rs), p(self.voff), p(self.vecs), p(d), p(out))
        elif self.weuc:
            lib().emu_weuc_dder3(self.K, p(self.off), p(self.dims), p(self.Rs), p(self.voff), p(self.vecs), p(d), p(out))
        elif self.wpsd:
            lib().emu_wpsd_dder3(self.K, p(self.off), p(self.dims), p(self.Rs), p(self.voff), p(self.vecs), p(d), p(out))
        elif self.mep:
            lib().emu_mep_dder3(self.K, p(self.off), p(self.dims), p(self.d1), p(self.voff), p(self.vecs), p(self.scal),
                                p(self.point), p(d), p(out))
        elif self.dnn:
            lib().emu_dnn_dder3(self.K, p(self.off), p(self.dims), p(self.sides), p(self.voff), p(self.vecs),
                                p(self.point), p(d), p(out))
        elif self.lmi:
            lib().emu_lmi_dder3(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(d), p(out))
        elif self.wsos:
            lib().emu_wsos_dder3(self.K, p(self.off), p(self.dims), p(self.voff), p(self.vecs), p(d), p(out))
        elif self.ens:
            lib().emu_ens_dder3(self.K, p(self.off), p(self.dims), p(self.d1), p(self.voff), p(self.vecs), p(self.scal),
                                p(self.point), p(d), p(out))
        elif self.hpm:
            lib().emu_hpm_dder3(self.K, p(self.off), p(self.dims), p(self.aoff), p(self.alpha), p(self.scal),
                                p(self.point), p(d), p(out))
        else:
            lib().emu_gpow_dder3(self.K, p(self.off), p(self.dims), p(self.mu), p(self.aoff), p(self.alpha),
                                 p(self.scal), p(self.point), p(d), p(out))
        return out

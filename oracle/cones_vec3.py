"""CPU restatement of EpiPerSquare and HypoPerLog (oracle; test infrastructure).

reference: src/Cones/epipersquare.jl:42-274 (rotated second-order cone (u, v, w): 2uv >= |w|^2,
barrier -log(2uv - |w|^2), nu = 2, self-dual, closed-form sqrt oracles),
src/Cones/hypoperlog.jl:54-287 (hypograph of the perspective of sum-log, (u, v, w):
u <= v sum log(w_i / v), barrier -log(v sum log(w_i/v) - u) - log v - sum log w_i, nu = dim).
"""
import numpy as np
import scipy.linalg as sla

from hypatia_b200.host import models as M
from .cones import Cone, EPS, _as2d, _ret, get_central_ray_hypoperlog


class EpiPerSquare(Cone):
    ctype = M.CONE_EPIPERSQUARE
    nu = 2.0

    def __init__(self, dim, use_dual=False):
        assert not use_dual            # epipersquare.jl:42 (self-dual cone)
        super().__init__(dim)

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[:2] = 1.0
        return arr

    def update_feas(self):
        u, v = self.point[0], self.point[1]
        if u > EPS and v > EPS:
            w = self.point[2:]
            self.dist = u * v - float(w @ w) / 2
            return self.dist > EPS
        return False

    def is_dual_feas(self):
        u, v = self.dual_point[0], self.dual_point[1]
        if u > EPS and v > EPS:
            w = self.dual_point[2:]
            return (u * v - float(w @ w) / 2) > EPS
        return False

    def update_grad(self):
        g = self._grad
        g[:] = self.point / self.dist
        g2 = g[1]
        g[1] = -g[0]
        g[0] = -g2

    def update_hess(self):
        g = self.grad()
        H = np.outer(g, g)
        inv_dist = 1.0 / self.dist
        idx = np.arange(2, self.dim)
        H[idx, idx] += inv_dist
        H[0, 1] -= inv_dist
        H[1, 0] -= inv_dist
        return H

    def update_inv_hess(self):
        Hi = np.outer(self.point, self.point)
        idx = np.arange(2, self.dim)
        Hi[idx, idx] += self.dist
        Hi[0, 1] -= self.dist
        Hi[1, 0] -= self.dist
        return Hi

    def use_sqrt_hess_oracles(self, arr_dim):
        return True

    @staticmethod
    def _swap(a):
        """J a = (-a_2, -a_1, a_w)."""
        out = a.copy()
        out[0] = -a[1]
        out[1] = -a[0]
        return out

    def hess_prod(self, arr):
        assert self.is_feas()
        a, vec = _as2d(arr)
        u, v, w = self.point[0], self.point[1], self.point[2:]
        uj, vj, wj = a[0], a[1], a[2:]
        ga = (w @ wj - v * uj - u * vj) / self.dist
        prod = np.empty_like(a)
        prod[0] = -ga * v - vj
        prod[1] = -ga * u - uj
        prod[2:] = ga[None, :] * w[:, None] + wj
        prod /= self.dist
        return _ret(prod, vec)

    def inv_hess_prod(self, arr):
        assert self.is_feas()
        a, vec = _as2d(arr)
        pa = self.point @ a
        prod = pa[None, :] * self.point[:, None] + self.dist * self._swap(a)
        return _ret(prod, vec)

    def sqrt_hess_prod(self, arr):
        assert self.is_feas()
        a, vec = _as2d(arr)
        rtdist = np.sqrt(self.dist)
        denom = 2 * rtdist + self.point[0] + self.point[1]
        sv = self.point / rtdist
        sv[0] = -self.point[1] / rtdist - 1
        sv[1] = -self.point[0] / rtdist - 1
        dotj = (sv @ a) / denom
        prod = dotj[None, :] * sv[:, None] + self._swap(a) / rtdist
        return _ret(prod, vec)

    def inv_sqrt_hess_prod(self, arr):
        assert self.is_feas()
        a, vec = _as2d(arr)
        rtdist = np.sqrt(self.dist)
        denom = 2 * rtdist + self.point[0] + self.point[1]
        sv = self.point.copy()
        sv[:2] += rtdist
        dotj = (sv @ a) / denom
        prod = dotj[None, :] * sv[:, None] + self._swap(a) * rtdist
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        point = self.point
        u, v, w = point[0], point[1], point[2:]
        u_dir, v_dir, w_dir = direction[0], direction[1], direction[2:]
        jdotpd = u * v_dir + v * u_dir - float(w @ w_dir)
        d3 = self.hess_prod(direction).copy()
        dotdHd = -float(direction @ d3)
        dotpHd = float(point @ d3)
        d3 *= jdotpd
        d3[2:] += dotdHd * w + dotpHd * w_dir
        d3[0] += -dotdHd * v - dotpHd * v_dir
        d3[1] += -dotdHd * u - dotpHd * u_dir
        d3 /= 2 * self.dist
        return d3


class HypoPerLog(Cone):
    ctype = M.CONE_HYPOPERLOG

    def __init__(self, dim, use_dual=False):
        self.use_dual_barrier = use_dual
        super().__init__(dim)

    @property
    def nu(self):
        return float(self.dim)

    def set_initial_point(self, arr):
        u, v, w = get_central_ray_hypoperlog(self.dim - 2)
        arr[0], arr[1] = u, v
        arr[2:] = w
        return arr

    def update_feas(self):
        v, w = self.point[1], self.point[2:]
        if v > EPS and (w > EPS).all():
            u = self.point[0]
            self.phi = float(np.sum(np.log(w / v)))
            self.zeta = v * self.phi - u
            return self.zeta > EPS
        return False

    def is_dual_feas(self):
        u, w = self.dual_point[0], self.dual_point[2:]
        if (w > EPS).all() and u < -EPS:
            v = self.dual_point[1]
            sumlog = float(np.sum(np.log(w / -u)))
            return (v - u * (sumlog + w.size)) > EPS
        return False

    def update_grad(self):
        v, w = self.point[1], self.point[2:]
        d = w.size
        zeta = self.zeta
        g = self._grad
        g[0] = 1.0 / zeta
        g[1] = -(self.phi - d) / zeta - 1.0 / v
        g[2:] = (-1 - v / zeta) / w

    def update_hess(self):
        g = self.grad()
        v, w = self.point[1], self.point[2:]
        d = w.size
        zeta = self.zeta
        sigzi = (self.phi - d) / zeta
        vzi = v / zeta
        wivzi = vzi / w
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zeta ** -2
        H[0, 1] = H[1, 0] = -sigzi / zeta
        H[1, 1] = v ** -2 + sigzi ** 2 + d / zeta / v
        H[0, 2:] = H[2:, 0] = (-vzi / zeta) / w
        H[1, 2:] = H[2:, 1] = (((self.phi - d) * vzi - 1) / zeta) / w
        H[2:, 2:] = np.outer(wivzi, wivzi)
        idx = np.arange(2, self.dim)
        H[idx, idx] -= g[2:] / w
        return H

    def hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        v, w = self.point[1], self.point[2:]
        zeta = self.zeta
        d = w.size
        sigma = self.phi - d
        vzi1 = v / zeta + 1
        p, q = a[0], a[1]
        rwi = a[2:] / w[:, None]
        qzi = q / zeta
        c0 = rwi.sum(axis=0) / zeta
        c1 = (v * c0 - p / zeta + sigma * qzi) / zeta
        c3 = c1 * v - qzi
        prod = np.empty_like(a)
        prod[0] = -c1
        prod[1] = c1 * sigma - c0 + (qzi * d + q / v) / v
        prod[2:] = (c3[None, :] + vzi1 * rwi) / w[:, None]
        return _ret(prod, vec)

    def _ih_consts(self):
        v, w = self.point[1], self.point[2:]
        d = w.size
        zeta, phi = self.zeta, self.phi
        zv = zeta + v
        zzvi = zeta / zv
        c3 = v / (zv + d * v)
        c0 = phi - d * zzvi
        c4 = v * c3 * zv
        c6 = (v * phi) ** 2 + zeta * (zeta + d * v) - d * (zeta + v * phi) ** 2 * c3
        return v, w, d, zeta, phi, zv, zzvi, c3, c0, c4, c6

    def update_inv_hess(self):
        self.grad()
        v, w, d, zeta, phi, zv, zzvi, c3, c0, c4, c6 = self._ih_consts()
        c2 = v * c3
        c1 = v * zzvi + c0 * c2
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = c6
        Hi[0, 1] = Hi[1, 0] = c0 * c4
        Hi[1, 1] = c4
        Hi[0, 2:] = Hi[2:, 0] = c1 * w
        Hi[1, 2:] = Hi[2:, 1] = c2 * w
        Hi[2:, 2:] = zzvi * np.diag(w ** 2) + (c2 / zv) * np.outer(w, w)
        return Hi

    def inv_hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        v, w, d, zeta, phi, zv, zzvi, c3, c0, c4, c6 = self._ih_consts()
        c7 = c4 * c0
        c8 = c7 + v * zeta
        p, q = a[0], a[1]
        rw = a[2:] * w[:, None]
        c1 = rw.sum(axis=0) / zv
        c5 = c0 * p + q + c1
        c2 = v * (zzvi * p + c3 * c5)
        prod = np.empty_like(a)
        prod[0] = c6 * p + c7 * q + c8 * c1
        prod[1] = c4 * c5
        prod[2:] = (c2[None, :] + zzvi * rw) * w[:, None]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        v, w = self.point[1], self.point[2:]
        p, q = direction[0], direction[1]
        zeta = self.zeta
        d = w.size
        sigma = self.phi - d
        viq = q / v
        viq2 = viq ** 2
        vzi = v / zeta
        vzi1 = vzi + 1
        rwi = direction[2:] / w
        c0 = float(rwi.sum())
        c7 = float((rwi ** 2).sum())
        zichi = (-p + sigma * q + c0 * v) / zeta
        c4 = (viq * (-viq * d + 2 * c0) - c7) / zeta / 2
        c1 = (zichi ** 2 - v * c4) / zeta
        c3 = -(zichi + viq) / zeta
        c5 = c3 * q + vzi * viq2
        c6 = -2 * vzi * viq - c3 * v
        c8 = c5 + c1 * v
        d3 = np.empty(self.dim)
        d3[0] = -c1
        d3[1] = c1 * sigma + (viq2 - (d * c5 + c6 * c0 + vzi * c7)) / v - c4
        d3[2:] = (c8 + rwi * (c6 + vzi1 * rwi)) / w
        return d3


class EpiNormInf(Cone):
    """epinorminf.jl:8-406 (real case): (u, w) with u >= |w|_inf, barrier
    -sum_j log(u^2 - w_j^2) + (n - 1) log u, nu = n + 1 = dim; the dual barrier variant
    (use_dual = true) serves the epigraph of the l1 norm."""
    ctype = M.CONE_EPINORMINF

    def __init__(self, dim, use_dual=False):
        self.use_dual_barrier = use_dual
        self.n = dim - 1
        super().__init__(dim)

    @property
    def nu(self):
        return float(self.n + 1)

    def reset_data(self):
        super().reset_data()
        self.hess_aux_updated = self.inv_hess_aux_updated = False

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[0] = np.sqrt(self.nu)
        return arr

    def update_feas(self):
        u, w = self.point[0], self.point[1:]
        return bool(u > EPS and u - np.abs(w).max() > EPS)

    def is_dual_feas(self):
        u = self.dual_point[0]
        return bool(u > EPS and u - np.abs(self.dual_point[1:]).sum() > EPS)

    def update_grad(self):
        u, w = self.point[0], self.point[1:]
        self.den = 0.5 * (u * u - w * w)
        self.uden = u / self.den
        self.wden = w / self.den
        self._grad[0] = (self.n - 1) / u - self.uden.sum()
        self._grad[1:] = self.wden

    def update_hess_aux(self):
        if self.hess_aux_updated:
            return
        self.grad()
        u = self.point[0]
        self.Hure = -self.wden * self.uden
        self.Hrere = self.wden ** 2 + 1.0 / self.den
        self.Huu = float((self.uden ** 2).sum()) - ((self.n - 1) / u + self.uden.sum()) / u
        self.hess_aux_updated = True

    def update_hess(self):
        self.update_hess_aux()
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = self.Huu
        H[0, 1:] = H[1:, 0] = self.Hure
        idx = np.arange(1, self.dim)
        H[idx, idx] = self.Hrere
        return H

    def hess_prod(self, arr):
        self.update_hess_aux()
        a, vec = _as2d(arr)
        prod = np.empty_like(a)
        prod[0] = self.Huu * a[0] + self.Hure @ a[1:]
        prod[1:] = self.Hure[:, None] * a[0][None, :] + self.Hrere[:, None] * a[1:]
        return _ret(prod, vec)

    def update_inv_hess_aux(self):
        if self.inv_hess_aux_updated:
            return
        self.update_hess_aux()
        u, w = self.point[0], self.point[1:]
        u2pw2 = 0.5 * (u * u + w * w)
        self.Hiure = u / u2pw2 * w
        self.schur = (1 - self.n) / (u * u) + float((1.0 / u2pw2).sum())
        self.inv_hess_aux_updated = True

    def update_inv_hess(self):
        self.update_inv_hess_aux()
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = 1.0 / self.schur
        col = self.Hiure
        Hi[1:, 0] = Hi[0, 1:] = col / self.schur
        Hi[1:, 1:] = np.outer(col, col) / self.schur + np.diag(1.0 / self.Hrere)
        return Hi

    def inv_hess_prod(self, arr):
        self.update_inv_hess_aux()
        a, vec = _as2d(arr)
        prod = np.empty_like(a)
        prod[0] = (a[0] + self.Hiure @ a[1:]) / self.schur
        prod[1:] = self.Hiure[:, None] * prod[0][None, :] + a[1:] / self.Hrere[:, None]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        u, w = self.point[0], self.point[1:]
        udir, d = direction[0], direction[1:]
        u3 = 1.5 / u
        udu = udir / u
        z = self.uden
        d3 = np.empty(self.dim)
        d3[0] = -udir * float((z * (u3 - z) * z).sum()) * udir - udu * (self.n - 1) / u * udu
        deni = -4 * self.den
        udeni = 2 * self.uden
        suuw = udir * (-1 + udeni * u)
        wdeni = 2 * self.wden
        uuw = suuw * wdeni
        uimim = 1 + wdeni * w
        uimim2 = -udeni * uimim * d
        d3[0] += float((d * (2 * uuw + uimim2) / deni).sum())
        d3[1:] = (udir * (uuw + 2 * uimim2) + d * wdeni * (2 + uimim) * d) / deni
        return d3


class HypoGeoMean(Cone):
    """hypogeomean.jl:8-264: (u, w), w in R^d_++, u <= geomean(w); barrier
    -log(prod(w_i)^(1/d) - u) - sum log w_i, nu = dim."""
    ctype = M.CONE_HYPOGEOMEAN

    def __init__(self, dim, use_dual=False):
        self.use_dual_barrier = use_dual
        self.d = dim - 1
        self.di = 1.0 / self.d
        super().__init__(dim)

    @property
    def nu(self):
        return float(self.dim)

    def set_initial_point(self, arr):
        d = self.d
        c = np.sqrt(5.0 * d * d + 2 * d + 1)
        arr[0] = -np.sqrt((-c + 3 * d + 1) / (2.0 + 2 * d))
        arr[1:] = (c - d + 1) / np.sqrt((1 + d) * (-2 * c + 6 * d + 2))
        return arr

    def update_feas(self):
        u, w = self.point[0], self.point[1:]
        if (w > EPS).all():
            self.phi = float(np.exp(self.di * np.sum(np.log(w))))
            self.zeta = self.phi - u
            return self.zeta > EPS
        return False

    def is_dual_feas(self):
        u, w = self.dual_point[0], self.dual_point[1:]
        if u < -EPS and (w > EPS).all():
            return bool(w.size * np.exp(self.di * np.sum(np.log(w))) + u > EPS)
        return False

    def update_grad(self):
        w = self.point[1:]
        self.pzd = self.phi / self.zeta * self.di
        self._grad[0] = 1.0 / self.zeta
        self._grad[1:] = (-self.pzd - 1) / w

    def update_hess(self):
        self.grad()
        w, zeta, pzd = self.point[1:], self.zeta, self.pzd
        c4 = pzd - self.di
        c1 = pzd * (1 + c4) + 1
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zeta ** -2
        H[0, 1:] = H[1:, 0] = -(pzd / w) / zeta
        wi = 1.0 / w
        H[1:, 1:] = pzd * c4 * np.outer(wi, wi)
        idx = np.arange(1, self.dim)
        H[idx, idx] = c1 * wi * wi
        return H

    def hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        w, zeta, pzd, di = self.point[1:], self.zeta, self.pzd, self.di
        p = a[0]
        rwi = a[1:] / w[:, None]
        c0 = pzd * rwi.sum(axis=0)
        c1 = c0 - p / zeta
        c2 = pzd * c1 - di * c0
        prod = np.empty_like(a)
        prod[0] = c1 / -zeta
        prod[1:] = (c2[None, :] + (pzd + 1) * rwi) / w[:, None]
        return _ret(prod, vec)

    def update_inv_hess(self):
        self.grad()
        w, zeta, phi, di = self.point[1:], self.zeta, self.phi, self.di
        phidi = phi * di
        c2 = 1.0 / (self.pzd + 1)
        c3 = c2 / zeta * di
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = zeta ** 2 + phidi * phi
        Hi[0, 1:] = Hi[1:, 0] = phidi * w
        Hi[1:, 1:] = c3 * phidi * np.outer(w, w) + np.diag(c2 * w * w)
        return Hi

    def inv_hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        w, zeta, phi, di = self.point[1:], self.zeta, self.phi, self.di
        phidi = phi * di
        c2 = 1.0 / (self.pzd + 1)
        c3 = c2 / zeta * di
        c4 = zeta ** 2 + phidi * phi
        p = a[0]
        rw = a[1:] * w[:, None]
        c5 = rw.sum(axis=0)
        c6 = phidi * (c3 * c5 + p)
        prod = np.empty_like(a)
        prod[0] = phidi * c5 + c4 * p
        prod[1:] = (c6[None, :] + c2 * rw) * w[:, None]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        w, zeta, phi, di, pzd = self.point[1:], self.zeta, self.phi, self.di, self.pzd
        p, r = direction[0], direction[1:]
        rwi = r / w
        c0 = float(rwi.sum()) * di
        c6 = float((rwi ** 2).sum()) * di
        zichi = (p - phi * c0) / zeta
        c1 = zichi ** 2 + phi / zeta * (c6 - c0 ** 2) / 2
        c7 = pzd * (c1 - c6 / 2 + c0 * (zichi + c0 / 2))
        c8 = -pzd * (zichi + c0)
        c9 = pzd + 1
        d3 = np.empty(self.dim)
        d3[0] = c1 / -zeta
        d3[1:] = (c7 + rwi * (c8 + c9 * rwi)) / w
        return d3


class EpiRelEntropy(Cone):
    """epirelentropy.jl:8-410: (u, v, w) with v, w in R^d_++, u >= sum_i w_i log(w_i / v_i); barrier
    -log(u - sum w_i log(w_i / v_i)) - sum log v_i - sum log w_i, nu = dim = 1 + 2d."""
    ctype = M.CONE_EPIRELENTROPY

    _CENTRAL = np.array([      # epirelentropy.jl:398-409
        [0.827838399, 1.290927714, 0.805102005], [0.708612491, 1.256859155, 0.818070438],
        [0.622618845, 1.231401008, 0.829317079], [0.558111266, 1.211710888, 0.838978357],
        [0.508038611, 1.196018952, 0.847300431], [0.468039614, 1.183194753, 0.854521307],
        [0.435316653, 1.172492397, 0.860840992], [0.408009282, 1.163403374, 0.866420017],
        [0.38483862, 1.155570329, 0.871385499], [0.364899122, 1.148735192, 0.875838068]])

    def __init__(self, dim, use_dual=False):
        assert dim >= 3 and dim % 2 == 1
        self.use_dual_barrier = use_dual
        self.d = (dim - 1) // 2
        super().__init__(dim)

    @property
    def nu(self):
        return float(self.dim)

    @classmethod
    def central_ray(cls, d):
        # epirelentropy.jl:377-396
        if d <= 10:
            return cls._CENTRAL[d - 1]
        rt = np.sqrt(d)
        if d <= 20:
            return np.array([1.2023 / rt - 0.015, 0.432 / rt + 1.0125, -0.3057 / rt + 0.972])
        return np.array([1.1513 / rt - 0.0069, 0.4873 / rt + 1.0008, -0.4247 / rt + 0.9961])

    def set_initial_point(self, arr):
        u, v, w = self.central_ray(self.d)
        arr[0] = u
        arr[1:1 + self.d] = v
        arr[1 + self.d:] = w
        return arr

    def _uvw(self, vec):
        return vec[0], vec[1:1 + self.d], vec[1 + self.d:]

    def update_feas(self):
        # epirelentropy.jl:91-108
        u, v, w = self._uvw(self.point)
        if (v > EPS).all() and (w > EPS).all():
            self.lwv = np.log(w / v)
            self.z = float(u - w @ self.lwv)
            return self.z > EPS
        return False

    def is_dual_feas(self):
        # epirelentropy.jl:110-121
        u, v, w = self._uvw(self.dual_point)
        if (v > EPS).all() and u > EPS:
            return bool((u * (1 + np.log(v / u)) + w > EPS).all())
        return False

    def update_grad(self):
        # epirelentropy.jl:123-140
        u, v, w = self._uvw(self.point)
        z = self.z
        self.sigma = w / v / z
        self.tau = (self.lwv + 1) / -z
        self._grad[0] = -1.0 / z
        self._grad[1:1 + self.d] = -self.sigma - 1.0 / v
        self._grad[1 + self.d:] = -self.tau - 1.0 / w

    def update_hess(self):
        # epirelentropy.jl:142-186
        self.grad()
        d, z, sigma, tau = self.d, self.z, self.sigma, self.tau
        u, v, w = self._uvw(self.point)
        H = np.zeros((self.dim, self.dim))
        vi, wi = np.arange(1, 1 + d), np.arange(1 + d, self.dim)
        H[0, 0] = z ** -2
        H[0, vi] = H[vi, 0] = sigma / z
        H[0, wi] = H[wi, 0] = tau / z
        H[np.ix_(vi, vi)] = np.outer(sigma, sigma)
        H[np.ix_(wi, wi)] = np.outer(tau, tau)
        H[np.ix_(vi, wi)] = np.outer(sigma, tau)
        H[vi, vi] = sigma ** 2 + (sigma + 1.0 / v) / v
        H[wi, wi] = tau ** 2 + (1.0 / z + 1.0 / w) / w
        H[vi, wi] -= 1.0 / z / v
        H[np.ix_(wi, vi)] = H[np.ix_(vi, wi)].T
        return H

    def hess_prod(self, arr):
        # epirelentropy.jl:260-293
        self.grad()
        a, vec = _as2d(arr)
        d, z, sigma, tau = self.d, self.z, self.sigma, self.tau
        u, v, w = self._uvw(self.point)
        p, av, aw = a[0], a[1:1 + d], a[1 + d:]
        up = sigma @ av + tau @ aw + p / z
        prod = np.empty_like(a)
        prod[1:1 + d] = sigma[:, None] * up[None, :] + (sigma[:, None] * av + av / v[:, None] - aw / z) / v[:, None]
        prod[1 + d:] = tau[:, None] * up[None, :] + (aw / z + aw / w[:, None]) / w[:, None] - av / v[:, None] / z
        prod[0] = up / z
        return _ret(prod, vec)

    def _inv_hess_aux(self):
        # epirelentropy.jl:188-222
        u, v, w = self._uvw(self.point)
        z, lwv = self.z, self.lwv
        zw = z + w
        z2w = zw + w
        wz2w = w / z2w
        vz2w = v / z2w
        uvv = w * (w * lwv - z)
        uww = w * (z + lwv * zw) * wz2w
        HiuHu = float(np.sum(wz2w * uvv - uww * (lwv + 1)))
        return dict(Hiuu=z * z - HiuHu, Hiuv=vz2w * uvv, Hiuw=uww, Hivw=w * v * wz2w, Hiww=w * zw * wz2w,
                    Hivv=v * zw * vz2w)

    def update_inv_hess(self):
        # epirelentropy.jl:224-258 (sparse arrow + diagonal bands, formed dense here)
        self.grad()
        d, x = self.d, self._inv_hess_aux()
        vi, wi = np.arange(1, 1 + d), np.arange(1 + d, self.dim)
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = x["Hiuu"]
        Hi[0, vi] = Hi[vi, 0] = x["Hiuv"]
        Hi[0, wi] = Hi[wi, 0] = x["Hiuw"]
        Hi[vi, vi] = x["Hivv"]
        Hi[wi, wi] = x["Hiww"]
        Hi[vi, wi] = Hi[wi, vi] = x["Hivw"]
        return Hi

    def inv_hess_prod(self, arr):
        # epirelentropy.jl:295-321
        self.grad()
        a, vec = _as2d(arr)
        d, x = self.d, self._inv_hess_aux()
        p, av, aw = a[0], a[1:1 + d], a[1 + d:]
        prod = np.empty_like(a)
        prod[0] = x["Hiuu"] * p + x["Hiuv"] @ av + x["Hiuw"] @ aw
        prod[1:1 + d] = x["Hiuv"][:, None] * p[None, :] + x["Hivv"][:, None] * av + x["Hivw"][:, None] * aw
        prod[1 + d:] = x["Hiuw"][:, None] * p[None, :] + x["Hivw"][:, None] * av + x["Hiww"][:, None] * aw
        return _ret(prod, vec)

    def dder3(self, direction):
        # epirelentropy.jl:323-364
        self.grad()
        d, z, tau = self.d, self.z, self.tau
        u, v, w = self._uvw(self.point)
        p, dv, dw = direction[0], direction[1:1 + d], direction[1 + d:]
        i2z = 1.0 / (2 * z)
        wdw = dw / w
        vdv = dv / v
        const0 = (p + w @ vdv) / z + tau @ dw
        const1 = const0 ** 2 + float(np.sum(w * vdv ** 2 + dw * (wdw - 2 * vdv))) / (2 * z)
        d3 = np.empty(self.dim)
        d3[0] = const1 / z
        t = const1 + (const0 + vdv) * vdv - i2z * wdw * dw
        t = t * w + (z * vdv - dw) * vdv + (-const0 + i2z * dw) * dw
        d3[1:1 + d] = t / v / z
        d3[1 + d:] = (const1 * tau + ((const0 - w * vdv / z) / z + (1.0 / w + i2z) * wdw) * wdw
                      + (-const0 + dw / z - vdv / 2) / z * vdv)
        return d3


class EpiNormSpectral(Cone):
    """epinormspectral.jl:10-294 (real case): (u, W) with W a d1 x d2 matrix stacked column-wise, d1 <= d2,
    u >= sigma_max(W); barrier -logdet(u I - W W' / u) - log u, nu = d1 + 1; the dual cone is the nuclear-norm
    epigraph.  No closed-form inverse Hessian: inv_hess_prod!, inv_hess and the sqrt oracles are the generic ones of
    Cones.jl:113-118, 189-259 (explicit Hessian + posdef_fact_copy!)."""
    ctype = M.CONE_EPINORMSPECTRAL

    def __init__(self, d1, d2, use_dual=False):
        assert 1 <= d1 <= d2
        self.d1, self.d2 = d1, d2
        self.use_dual_barrier = use_dual
        super().__init__(1 + d1 * d2)

    @property
    def nu(self):
        return float(self.d1 + 1)

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[0] = np.sqrt(self.d1 + 1.0)
        return arr

    def _mat(self, vec):
        return vec.reshape(self.d1, self.d2, order="F")

    def update_feas(self):
        # epinormspectral.jl:107-124
        u = self.point[0]
        if u > EPS:
            self.W = self._mat(self.point[1:]).copy()
            Z = u * u * np.eye(self.d1) - self.W @ self.W.T
            try:
                self.fact_Z = sla.cho_factor(Z, lower=False, check_finite=False)
            except np.linalg.LinAlgError:
                return False
            return True
        return False

    def is_dual_feas(self):
        # epinormspectral.jl:126-133
        u = self.dual_point[0]
        if u > EPS:
            return bool(u - np.sum(np.linalg.svd(self._mat(self.dual_point[1:]), compute_uv=False)) > EPS)
        return False

    def _zsolve(self, X):
        return sla.cho_solve(self.fact_Z, X, check_finite=False)

    def update_grad(self):
        # epinormspectral.jl:135-151
        u = self.point[0]
        self.tau = self._zsolve(self.W)
        self.Zi = self._zsolve(np.eye(self.d1))
        self.Zi = (self.Zi + self.Zi.T) / 2
        self._grad[0] = -2 * u * np.trace(self.Zi) + (self.d1 - 1) / u
        self._grad[1:] = 2 * self.tau.ravel(order="F")
        # update_hess_aux, epinormspectral.jl:153-172
        self.Zitau = self._zsolve(self.tau)
        self.HuW = -4 * u * self.Zitau
        self.trZi2 = float(np.sum(self.Zi ** 2))
        self.Huu = 4 * u * u * self.trZi2 + (self._grad[0] - 2 * (self.d1 - 1) / u) / u
        self.WtauI = np.eye(self.d2) + self.W.T @ self.tau

    def update_hess(self):
        # epinormspectral.jl:174-214: H[(j,i),(l,k)] = 2 (Zi[l,j] WtauI[i,k] + tau[l,i] tau[j,k]), row index j + i d1
        self.grad()
        d1, d2 = self.d1, self.d2
        H = np.empty((self.dim, self.dim))
        T1 = np.einsum("lj,ik->jilk", self.Zi, self.WtauI)
        T2 = np.einsum("li,jk->jilk", self.tau, self.tau)
        HW = 2 * (T1 + T2)                                  # [j, i, l, k]
        H[1:, 1:] = HW.transpose(1, 0, 3, 2).reshape(d1 * d2, d1 * d2)   # (i, j) -> i * d1 + j
        H[0, 1:] = H[1:, 0] = self.HuW.ravel(order="F")
        H[0, 0] = self.Huu
        return H

    def hess_prod(self, arr):
        # epinormspectral.jl:216-246
        self.grad()
        a, vec = _as2d(arr)
        u = self.point[0]
        prod = np.empty_like(a)
        for j in range(a.shape[1]):
            p, R = a[0, j], self._mat(a[1:, j])
            prod[0, j] = self.Huu * p + float(np.sum(self.HuW * R))
            T = R @ self.W.T
            T = T + T.T - 2 * u * p * np.eye(self.d1)
            prod[1:, j] = self._zsolve(2 * (T @ self.tau) + 2 * R).ravel(order="F")
        return _ret(prod, vec)

    def dder3(self, direction):
        # epinormspectral.jl:248-294
        self.grad()
        u, W, tau, Zitau, WtauI, Zi = self.point[0], self.W, self.tau, self.Zitau, self.WtauI, self.Zi
        ud, Wd = direction[0], self._mat(direction[1:])
        B = Wd.T @ tau
        D = self._zsolve(Wd)
        E = D @ WtauI
        C = D @ B.T
        F = D @ W.T
        G = B @ B + Wd.T @ E
        D2 = tau @ G + C @ WtauI + E @ B
        E2 = self._zsolve(E) + Zitau @ B
        F2 = F + tau @ Wd.T
        E3 = -2 * u * (E2 + F2 @ Zitau)
        C2 = self._zsolve(4 * u * u * ud * Zitau - ud * tau)
        E4 = E3 + C2
        d3 = np.empty(self.dim)
        d3[1:] = (-2 * ud * E4 - 2 * D2).ravel(order="F")
        U = self.fact_Z[0]
        trZi3 = float(np.sum(sla.solve_triangular(U, Zi, trans="T", lower=False, check_finite=False) ** 2))
        E5 = E4 + 3 * C2
        d3[0] = -float(np.sum(Wd * E5)) - u * ud * (6 * self.trZi2 - 8 * u * trZi3 * u) * ud \
            - (self.d1 - 1) * (ud / u) ** 2 / u
        return d3


class WSOSInterpNonnegative(Cone):
    """wsosinterpnonnegative.jl:15-200 (real case): interpolant-basis weighted sum-of-squares cone of dimension U,
    parametrised by matrices Ps[k] (U x L_k); the barrier is that of the DUAL cone
    {s : Ps[k]' Diagonal(s) Ps[k] psd for all k}, -sum_k logdet(Ps[k]' Diagonal(s) Ps[k]), so use_dual_barrier =
    !use_dual (wsosinterpnonnegative.jl:59); nu = sum L_k.  hess_prod! is the generic explicit-Hessian product
    (Cones.jl:101-105), inv_hess_prod! the generic factorisation fallback (Cones.jl:113-118), is_dual_feas the generic
    `true` (Cones.jl:64)."""
    ctype = M.CONE_WSOSINTERPNONNEGATIVE

    def __init__(self, U, Ps, use_dual=False):
        self.Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
        assert all(P.shape[0] == U for P in self.Ps)
        self.use_dual_barrier = not use_dual
        super().__init__(U)

    @property
    def nu(self):
        return float(sum(P.shape[1] for P in self.Ps))

    def set_initial_point(self, arr):
        arr[:] = 1.0
        return arr

    def update_feas(self):
        # wsosinterpnonnegative.jl:91-121: Lambda_k = P_k' Diagonal(point) P_k, Cholesky (lower)
        self.LF = []
        for P in self.Ps:
            Lam = P.T @ (self.point[:, None] * P)
            try:
                self.LF.append(np.linalg.cholesky(Lam))
            except np.linalg.LinAlgError:
                return False
        return True

    def update_grad(self):
        # wsosinterpnonnegative.jl:123-138: LFLP_k = L_k^-1 P_k'; grad_j = -sum_k |LFLP_k[:, j]|^2
        self.LFLP = [sla.solve_triangular(L, P.T, lower=True, check_finite=False) for L, P in zip(self.LF, self.Ps)]
        self._grad[:] = -sum(np.sum(F * F, axis=0) for F in self.LFLP)

    def update_hess(self):
        # wsosinterpnonnegative.jl:140-156
        self.grad()
        return sum((F.T @ F) ** 2 for F in self.LFLP)

    def hess_prod(self, arr):
        a, vec = _as2d(arr)
        return _ret(np.asarray(self.hess()) @ a, vec)

    def dder3(self, direction):
        # wsosinterpnonnegative.jl:180-200
        self.grad()
        d3 = np.zeros(self.dim)
        for F in self.LFLP:
            S = (F * direction[None, :]) @ F.T
            T = S @ F
            d3 += np.sum(T * T, axis=0)
        return d3


from hypatia_b200.host.instances import dnn_initial_point  # noqa: E402  (doublynonnegativetri.jl:72-126)


class DoublyNonnegativeTri(Cone):
    """doublynonnegativetri.jl:8-205: symmetric matrices (svec) that are positive semidefinite AND entrywise nonnegative;
    barrier -logdet(W) - sum of log over the off-diagonal svec entries, nu = dim.  inv_hess_prod! is the generic
    factorisation fallback (Cones.jl:113-118), is_dual_feas the generic `true`."""
    ctype = M.CONE_DOUBLYNONNEGATIVETRI

    def __init__(self, dim, use_dual=False):
        self.side = M.svec_side(dim)
        self.use_dual_barrier = use_dual
        self.offdiag = np.array([j * (j + 1) // 2 + i for j in range(self.side) for i in range(j)], dtype=np.int64)
        super().__init__(dim)

    @property
    def nu(self):
        return float(self.dim)

    def set_initial_point(self, arr):
        ond, offd = dnn_initial_point(self.side)
        arr[:] = offd
        arr[[j * (j + 1) // 2 + j for j in range(self.side)]] = ond
        return arr

    def update_feas(self):
        # doublynonnegativetri.jl:128-142
        from . import arrayutil as au
        if (self.point > EPS).all():
            self.mat = au.svec_to_smat(self.point)
            try:
                self.fact = sla.cho_factor(self.mat, lower=False, check_finite=False)
            except np.linalg.LinAlgError:
                return False
            return True
        return False

    def update_grad(self):
        # doublynonnegativetri.jl:144-155
        from . import arrayutil as au
        self.inv_mat = sla.cho_solve(self.fact, np.eye(self.side), check_finite=False)
        self.inv_mat = (self.inv_mat + self.inv_mat.T) / 2
        self._grad[:] = -au.smat_to_svec(self.inv_mat)
        self._grad[self.offdiag] -= 1.0 / self.point[self.offdiag]

    def update_hess(self):
        # doublynonnegativetri.jl:157-171
        from . import arrayutil as au
        self.grad()
        H = au.symm_kron(self.inv_mat)
        H[self.offdiag, self.offdiag] += self.point[self.offdiag] ** -2
        return H

    def hess_prod(self, arr):
        # doublynonnegativetri.jl:173-193
        from . import arrayutil as au
        self.grad()
        a, vec = _as2d(arr)
        prod = np.empty_like(a)
        for j in range(a.shape[1]):
            Mx = au.svec_to_smat(a[:, j])
            prod[:, j] = au.smat_to_svec(self.inv_mat @ Mx @ self.inv_mat)
        so = self.point[self.offdiag]
        prod[self.offdiag] += a[self.offdiag] / (so * so)[:, None]
        return _ret(prod, vec)

    def dder3(self, direction):
        # doublynonnegativetri.jl:195-205
        from . import arrayutil as au
        self.grad()
        D = au.svec_to_smat(direction)
        T1 = self.inv_mat @ D
        d3 = au.smat_to_svec(T1 @ self.inv_mat @ T1.T)
        so = self.point[self.offdiag]
        d3[self.offdiag] += (direction[self.offdiag] / so) ** 2 / so
        return d3


class WSOSInterpPosSemidefTri(Cone):
    """wsosinterppossemideftri.jl:14-321: interpolant-basis weighted sum-of-squares cone of R x R matrix polynomials
    (svec order of the blocks, U coefficients each); the barrier is the dual cone's,
    -sum_k logdet((I_R kron P_k)' D(s) (I_R kron P_k)) with D(s) the R U x R U matrix of diagonal blocks
    Diagonal(smat(s)_pq), so use_dual_barrier = !use_dual; nu = R sum L_k.  hess_prod! / inv_hess_prod! are the generic
    explicit-Hessian oracles (Cones.jl:101-118).  Dense restatement: the reference exploits the block-triangular
    structure of the factors, the mathematics is the same."""
    ctype = M.CONE_WSOSINTERPPOSSEMIDEFTRI

    def __init__(self, R, U, Ps, use_dual=False):
        self.R, self.U = R, U
        self.Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
        assert all(P.shape[0] == U for P in self.Ps)
        self.use_dual_barrier = not use_dual
        self.blocks = [(p, q) for p in range(R) for q in range(p + 1)]        # svec order: (p, q), q <= p
        super().__init__(U * R * (R + 1) // 2)

    @property
    def nu(self):
        return float(self.R * sum(P.shape[1] for P in self.Ps))

    def set_initial_point(self, arr):
        arr[:] = 0.0
        for b, (p, q) in enumerate(self.blocks):
            if p == q:
                arr[b * self.U:(b + 1) * self.U] = 1.0
        return arr

    def _D(self, s):
        R, U = self.R, self.U
        D = np.zeros((R * U, R * U))
        idx = np.arange(U)
        for b, (p, q) in enumerate(self.blocks):
            v = s[b * U:(b + 1) * U] * (1.0 if p == q else 1 / np.sqrt(2.0))
            D[p * U + idx, q * U + idx] = v
            D[q * U + idx, p * U + idx] = v
        return D

    def update_feas(self):
        # wsosinterppossemideftri.jl:108-142
        D = self._D(self.point)
        self.Pb = [np.kron(np.eye(self.R), P) for P in self.Ps]
        self.LF = []
        for Pb in self.Pb:
            try:
                self.LF.append(np.linalg.cholesky(Pb.T @ D @ Pb))
            except np.linalg.LinAlgError:
                return False
        return True

    def _blockdiag(self, G, G2=None):
        # block_diag_prod! (:259-282): diagonal of every (p, q) block of G (= mat1' mat2), sqrt(2) off the diagonal blocks
        U = self.U
        idx = np.arange(U)
        out = np.empty(self.dim)
        for b, (p, q) in enumerate(self.blocks):
            out[b * U:(b + 1) * U] = G[q * U + idx, p * U + idx] * (1.0 if p == q else np.sqrt(2.0))
        return out

    def update_grad(self):
        # wsosinterppossemideftri.jl:144-188
        self.F = [sla.solve_triangular(L, Pb.T, lower=True, check_finite=False) for L, Pb in zip(self.LF, self.Pb)]
        self.G = [F.T @ F for F in self.F]
        self._grad[:] = -sum(self._blockdiag(G) for G in self.G)

    def update_hess(self):
        # wsosinterppossemideftri.jl:190-239
        self.grad()
        U = self.U
        H = np.zeros((self.dim, self.dim))
        rt2 = np.sqrt(2.0)
        for G in self.G:
            B = lambda p, q: G[p * U:(p + 1) * U, q * U:(q + 1) * U]
            for b, (p, q) in enumerate(self.blocks):
                for b2, (p2, q2) in enumerate(self.blocks):
                    scal = rt2 if (p == q) != (p2 == q2) else 1.0
                    blk = B(p, p2) * B(q, q2) * scal
                    if p != q and p2 != q2:
                        blk = blk + B(p, q2) * B(q, p2)
                    H[b * U:(b + 1) * U, b2 * U:(b2 + 1) * U] += blk
        return (H + H.T) / 2

    def hess_prod(self, arr):
        a, vec = _as2d(arr)
        return _ret(np.asarray(self.hess()) @ a, vec)

    def dder3(self, direction):
        # partial_prod! with use_symm_prod = true (:284-321)
        self.grad()
        Dd = self._D(direction)
        out = np.zeros(self.dim)
        for L, Pb, F in zip(self.LF, self.Pb, self.F):
            S = Pb.T @ Dd @ Pb
            S = sla.solve_triangular(L, S, lower=True, check_finite=False)
            S = sla.solve_triangular(L, S.T, lower=True, check_finite=False).T
            T = ((S + S.T) / 2) @ F
            out += self._blockdiag(T.T @ T)
        return out


class WSOSInterpEpiNormEucl(Cone):
    """wsosinterpepinormeucl.jl:14-382: (f_1, .., f_R) polynomials (U interpolant coefficients each) with
    f_1 >= |(f_2 .. f_R)|_2 in the WSOS sense; the barrier is the dual cone's,
    -sum_k [logdet(L11_k) + logdet(L11_k - sum_r L1r_k L11_k^-1 L1r_k)] with L1r_k = P_k' Diagonal(s_r) P_k, nu = 2 sum L_k,
    use_dual_barrier = !use_dual.  Dense restatement: with A_k(s) the R L x R L block-arrow matrix (L11 on the diagonal
    blocks, L1r on the first block row / column), the barrier is -logdet A_k + (R - 2) logdet L11_k, so gradient, Hessian
    and dder3 are the logdet derivatives of two matrices that are LINEAR in s (the reference exploits the arrow structure,
    the mathematics is the same).  hess_prod! / inv_hess_prod! are the generic explicit-Hessian oracles."""
    ctype = M.CONE_WSOSINTERPEPINORMEUCL

    def __init__(self, R, U, Ps, use_dual=False):
        assert R >= 2
        self.R, self.U = R, U
        self.Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
        assert all(P.shape[0] == U for P in self.Ps)
        self.use_dual_barrier = not use_dual
        super().__init__(U * R)

    @property
    def nu(self):
        return float(2 * sum(P.shape[1] for P in self.Ps))

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[:self.U] = 1.0
        return arr

    def _D(self, s):
        R, U = self.R, self.U
        D = np.zeros((R * U, R * U))
        idx = np.arange(U)
        for r in range(R):
            D[r * U + idx, r * U + idx] = s[:U]
        for j in range(1, R):
            D[idx, j * U + idx] = s[j * U:(j + 1) * U]
            D[j * U + idx, idx] = s[j * U:(j + 1) * U]
        return D

    def update_feas(self):
        # wsosinterpepinormeucl.jl:119-167: L11 and the Schur complement positive definite <=> the arrow matrix is
        D = self._D(self.point)
        self.Pb = [np.kron(np.eye(self.R), P) for P in self.Ps]
        self.LA, self.L11 = [], []
        for P, Pb in zip(self.Ps, self.Pb):
            try:
                self.L11.append(np.linalg.cholesky(P.T @ (self.point[:self.U, None] * P)))
                self.LA.append(np.linalg.cholesky(Pb.T @ D @ Pb))
            except np.linalg.LinAlgError:
                return False
        return True

    def _collect(self, G, G11):
        """Derivative of logdet A (through G = (I kron P) X (I kron P)') minus (R - 2) times that of logdet L11."""
        R, U = self.R, self.U
        idx = np.arange(U)
        out = np.empty(self.dim)
        out[:U] = sum(G[r * U + idx, r * U + idx] for r in range(R)) - (R - 2) * G11[idx, idx]
        for j in range(1, R):
            out[j * U:(j + 1) * U] = 2 * G[idx, j * U + idx]
        return out

    def update_grad(self):
        # wsosinterpepinormeucl.jl:169-211
        tri = lambda L, B: sla.solve_triangular(L, B, lower=True, check_finite=False)
        self.F = [tri(L, Pb.T) for L, Pb in zip(self.LA, self.Pb)]
        self.F11 = [tri(L, P.T) for L, P in zip(self.L11, self.Ps)]
        self.G = [F.T @ F for F in self.F]
        self.G11 = [F.T @ F for F in self.F11]
        self._grad[:] = -sum(self._collect(G, G11) for G, G11 in zip(self.G, self.G11))

    def update_hess(self):
        # wsosinterpepinormeucl.jl:213-290: H_ij = tr(A^-1 E_i A^-1 E_j) - (R - 2) tr(L11^-1 E_i L11^-1 E_j)
        self.grad()
        R, U = self.R, self.U
        H = np.zeros((self.dim, self.dim))
        for G, G11 in zip(self.G, self.G11):
            B = lambda x, y: G[x * U:(x + 1) * U, y * U:(y + 1) * U]
            H[:U, :U] += sum(B(r, r2) ** 2 for r in range(R) for r2 in range(R)) - (R - 2) * G11 ** 2
            for j in range(1, R):
                blk = 2 * sum(B(r, 0) * B(r, j) for r in range(R))
                H[:U, j * U:(j + 1) * U] += blk
                H[j * U:(j + 1) * U, :U] += blk.T
                for j2 in range(1, R):
                    H[j * U:(j + 1) * U, j2 * U:(j2 + 1) * U] += 2 * (B(0, 0) * B(j, j2) + B(0, j2) * B(j, 0))
        return (H + H.T) / 2

    def hess_prod(self, arr):
        a, vec = _as2d(arr)
        return _ret(np.asarray(self.hess()) @ a, vec)

    def dder3(self, direction):
        # wsosinterpepinormeucl.jl:292-382
        self.grad()
        tri = lambda L, B: sla.solve_triangular(L, B, lower=True, check_finite=False)
        Dd = self._D(direction)
        out = np.zeros(self.dim)
        for P, Pb, L, L11, F, F11 in zip(self.Ps, self.Pb, self.LA, self.L11, self.F, self.F11):
            S = tri(L, tri(L, Pb.T @ Dd @ Pb).T).T
            T = S @ F
            S11 = tri(L11, tri(L11, P.T @ (direction[:self.U, None] * P)).T).T
            T11 = S11 @ F11
            out += self._collect(T.T @ T, T11.T @ T11)
        return out


class WSOSInterpEpiNormOne(Cone):
    """wsosinterpepinormone.jl:14-493: (f_1, .., f_R) polynomials with f_1 >= sum_r |f_r| in the WSOS sense; dual barrier
    -sum_k [logdet L11_k + sum_{r >= 2} logdet(L11_k - L1r_k L11_k^-1 L1r_k)], nu = R sum L_k, use_dual_barrier = !use_dual.
    Dense restatement: the barrier is sum_{r >= 2} -logdet A2(s_1, s_r) + (R - 2) logdet L11 with A2 the 2 x 2 block-arrow
    matrix of the R = 2 Euclidean-norm cone above, so every oracle is a sum of those of R - 1 pair cones plus the L11
    correction.  The reference has closed forms for hess_prod! / inv_hess_prod! that exploit the arrow-shaped Hessian; here
    the generic explicit-Hessian oracles (Cones.jl:101-118) are used - the same linear maps."""
    ctype = M.CONE_WSOSINTERPEPINORMONE

    def __init__(self, R, U, Ps, use_dual=False):
        assert R >= 2
        self.R, self.U = R, U
        self.Ps = [np.asarray(P, dtype=np.float64) for P in Ps]
        self.use_dual_barrier = not use_dual
        self.pairs = [WSOSInterpEpiNormEucl(2, U, self.Ps) for _ in range(R - 1)]
        for pc in self.pairs:
            pc.setup_data()
        super().__init__(U * R)

    @property
    def nu(self):
        return float(self.R * sum(P.shape[1] for P in self.Ps))

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[:self.U] = 1.0
        return arr

    def _pair_vec(self, vec, r):
        return np.concatenate((vec[:self.U], vec[r * self.U:(r + 1) * self.U]))

    def update_feas(self):
        # wsosinterpepinormone.jl:147-204
        for r, pc in enumerate(self.pairs, start=1):
            pc.reset_data()
            pc.load_point(self._pair_vec(self.point, r))
            if not pc.is_feas():
                return False
        return True

    def update_grad(self):
        # wsosinterpepinormone.jl:206-245
        U, R = self.U, self.R
        g = np.zeros(self.dim)
        for r, pc in enumerate(self.pairs, start=1):
            gp = pc.grad()
            g[:U] += gp[:U]
            g[r * U:(r + 1) * U] = gp[U:]
        self.G11 = self.pairs[0].G11
        g[:U] += (R - 2) * sum(np.diag(G11) for G11 in self.G11)
        self._grad[:] = g

    def update_hess(self):
        self.grad()
        U, R = self.U, self.R
        H = np.zeros((self.dim, self.dim))
        for r, pc in enumerate(self.pairs, start=1):
            Hp = np.asarray(pc.hess())
            sl = slice(r * U, (r + 1) * U)
            H[:U, :U] += Hp[:U, :U]
            H[:U, sl] = Hp[:U, U:]
            H[sl, :U] = Hp[U:, :U]
            H[sl, sl] = Hp[U:, U:]
        H[:U, :U] -= (R - 2) * sum(G11 ** 2 for G11 in self.G11)
        return H

    def hess_prod(self, arr):
        a, vec = _as2d(arr)
        return _ret(np.asarray(self.hess()) @ a, vec)

    def dder3(self, direction):
        # wsosinterpepinormone.jl:406-493
        self.grad()
        U, R = self.U, self.R
        tri = lambda L, B: sla.solve_triangular(L, B, lower=True, check_finite=False)
        out = np.zeros(self.dim)
        for r, pc in enumerate(self.pairs, start=1):
            dp = pc.dder3(self._pair_vec(direction, r))
            out[:U] += dp[:U]
            out[r * U:(r + 1) * U] = dp[U:]
        pc = self.pairs[0]
        for P, L11, F11 in zip(self.Ps, pc.L11, pc.F11):
            S11 = tri(L11, tri(L11, P.T @ (direction[:U, None] * P)).T).T
            T11 = S11 @ F11
            out[:U] -= (R - 2) * np.sum(T11 * T11, axis=0)
        return out


class PosSemidefTriSparse(Cone):
    """possemideftrisparse/{possemideftrisparse,denseimpl}.jl (real case, the reference's dense implementation
    PSDSparseDense): the entries (row_idxs, col_idxs) of the lower triangle of a symmetric side x side matrix (every
    diagonal entry present, off-diagonals scaled by sqrt 2) such that the matrix with zeros elsewhere is psd; barrier
    -logdet, nu = side.  use_dual = true gives the cone of psd-completable partial matrices.  hess_prod! /
    inv_hess_prod! are the generic explicit-Hessian oracles."""
    ctype = M.CONE_POSSEMIDEFTRISPARSE

    def __init__(self, side, row_idxs, col_idxs, use_dual=False):
        self.side = side
        self.rows = np.asarray(row_idxs, dtype=np.int64)      # 0-based, col <= row
        self.cols = np.asarray(col_idxs, dtype=np.int64)
        assert (self.cols <= self.rows).all() and (self.rows < side).all()
        assert sorted(self.rows[self.rows == self.cols]) == list(range(side))
        self.use_dual_barrier = use_dual
        self.scal = np.where(self.rows == self.cols, 1.0, np.sqrt(2.0))
        super().__init__(self.rows.size)

    @property
    def nu(self):
        return float(self.side)

    def set_initial_point(self, arr):
        arr[:] = (self.rows == self.cols).astype(float)
        return arr

    def _smat(self, vec):
        Mx = np.zeros((self.side, self.side))
        Mx[self.rows, self.cols] = vec / self.scal
        Mx[self.cols, self.rows] = vec / self.scal
        return Mx

    def update_feas(self):
        # denseimpl.jl:30-41
        try:
            self.fact = sla.cho_factor(self._smat(self.point), lower=True, check_finite=False)
        except np.linalg.LinAlgError:
            return False
        return True

    def update_grad(self):
        # denseimpl.jl:43-55
        Li = sla.cho_solve(self.fact, np.eye(self.side), check_finite=False)
        self.Li = (Li + Li.T) / 2
        self._grad[:] = -self.Li[self.rows, self.cols] * self.scal

    def update_hess(self):
        # denseimpl.jl:57-83
        self.grad()
        Li, i, j = self.Li, self.rows, self.cols
        H = Li[np.ix_(i, i)] * Li[np.ix_(j, j)] + Li[np.ix_(i, j)] * Li[np.ix_(j, i)]
        dg = (i == j)
        H[np.ix_(dg, dg)] = Li[np.ix_(i[dg], i[dg])] ** 2
        mixed = np.logical_xor.outer(dg, dg)
        Hm = np.sqrt(2.0) * Li[np.ix_(i, i)] * Li[np.ix_(j, j)]
        H[mixed] = Hm[mixed]
        return H

    def hess_prod(self, arr):
        a, vec = _as2d(arr)
        return _ret(np.asarray(self.hess()) @ a, vec)

    def dder3(self, direction):
        # denseimpl.jl:153-167
        self.grad()
        T = self.Li @ self._smat(direction) @ self.Li @ self._smat(direction) @ self.Li
        return T[self.rows, self.cols] * self.scal


class MatrixEpiPerSquare(Cone):
    """matrixepipersquare.jl:10-397 (real case): (svec(U), v, vec(W)) with U symmetric d1 x d1, W d1 x d2 (d1 <= d2),
    2 v U - W W' psd; barrier -logdet(2 v U - W W') + (d1 - 1) log v, nu = d1 + 1.  inv_hess_prod! is the generic
    factorisation fallback (Cones.jl:113-118); the explicit Hessian is assembled column by column from hess_prod!
    (equal to update_hess, matrixepipersquare.jl:187-279, which the identity tests check through hess * inv_hess)."""
    ctype = M.CONE_MATRIXEPIPERSQUARE

    def __init__(self, d1, d2, use_dual=False):
        assert 1 <= d1 <= d2
        self.d1, self.d2 = d1, d2
        self.v_idx = d1 * (d1 + 1) // 2
        self.use_dual_barrier = use_dual
        super().__init__(self.v_idx + 1 + d1 * d2)

    @property
    def nu(self):
        return float(self.d1 + 1)

    def set_initial_point(self, arr):
        # matrixepipersquare.jl:103-116: U = I, v = 1, W = 0
        arr[:] = 0.0
        arr[[j * (j + 1) // 2 + j for j in range(self.d1)]] = 1.0
        arr[self.v_idx] = 1.0
        return arr

    def _split(self, vec):
        from . import arrayutil as au
        return (au.svec_to_smat(vec[:self.v_idx]), vec[self.v_idx],
                vec[self.v_idx + 1:].reshape(self.d1, self.d2, order="F"))

    def update_feas(self):
        # matrixepipersquare.jl:118-135
        U, v, W = self._split(self.point)
        if v > EPS:
            self.U, self.v, self.W = U, float(v), W.copy()
            try:
                self.fact_Z = sla.cho_factor(2 * v * U - W @ W.T, lower=False, check_finite=False)
            except np.linalg.LinAlgError:
                return False
            return True
        return False

    def is_dual_feas(self):
        # matrixepipersquare.jl:137-150
        U, v, W = self._split(self.dual_point)
        if v > EPS:
            try:
                R = np.linalg.cholesky(U).T
            except np.linalg.LinAlgError:
                return False
            LW = sla.solve_triangular(R, W, trans="T", lower=False, check_finite=False)
            return bool(2 * v - float(np.sum(LW * LW)) > EPS)
        return False

    def _zs(self, X):
        return sla.cho_solve(self.fact_Z, X, check_finite=False)

    def update_grad(self):
        # matrixepipersquare.jl:152-170, update_hess_aux :172-185
        from . import arrayutil as au
        U, v, W, d1 = self.U, self.v, self.W, self.d1
        Zi = self._zs(np.eye(d1))
        self.Zi = (Zi + Zi.T) / 2
        self.ZiW = self._zs(W)
        self._grad[:self.v_idx] = -2 * v * au.smat_to_svec(self.Zi)
        self._grad[self.v_idx] = -2 * float(np.sum(self.Zi * U)) + (d1 - 1) / v
        self._grad[self.v_idx + 1:] = 2 * self.ZiW.ravel(order="F")
        ZiUZi = self._zs(self._zs(U).T).T
        self.ZiUZi = (ZiUZi + ZiUZi.T) / 2
        self.Hvv = 4 * float(np.sum(self.ZiUZi * U)) - (d1 - 1) / v / v
        self.ZiUZiW = self.ZiUZi @ W
        self.WtZiW = W.T @ self.ZiW

    def hess_prod(self, arr):
        # matrixepipersquare.jl:281-325
        from . import arrayutil as au
        self.grad()
        a, vec = _as2d(arr)
        U, v, W = self.U, self.v, self.W
        v2 = 2 * v
        prod = np.empty_like(a)
        for i in range(a.shape[1]):
            tU, va, tW = self._split(a[:, i])
            ZiWd = self._zs(tW)
            U3 = ZiWd @ self.ZiW.T
            ZtUZ = self._zs(self._zs(tU).T).T
            U2 = U3 + U3.T - v2 * ZtUZ
            T1 = U2 - 2 * va * self.ZiUZi
            prod[self.v_idx + 1:, i] = (2 * (T1 @ W) + 2 * ZiWd).ravel(order="F")
            T2 = 2 * v2 * self.ZiUZi - 2 * self.Zi
            prod[self.v_idx, i] = float(np.sum(T2 * tU)) - 4 * float(np.sum(U * U3)) + self.Hvv * va
            prod[:self.v_idx, i] = au.smat_to_svec(va * T2 - v2 * U2)
        return _ret(prod, vec)

    def update_hess(self):
        H = self.hess_prod(np.eye(self.dim))
        return (H + H.T) / 2

    def dder3(self, direction):
        # matrixepipersquare.jl:327-397
        from . import arrayutil as au
        self.grad()
        d1, U, v, W = self.d1, self.U, self.v, self.W
        Ud, vd, Wd = self._split(direction)
        v2, vd2 = 2 * v, 2 * vd
        ZiW, ZiUZi = self.ZiW, self.ZiUZi
        ZiU = self._zs(U)
        ZiWd = self._zs(Wd)
        ZiUd = self._zs(Ud)
        ZiUZiUZi = ZiUZi @ ZiU.T
        ZiUZiUZi = (ZiUZiUZi + ZiUZiUZi.T) / 2
        ZiUZi2v = ZiUZi - v2 * ZiUZiUZi
        WdWZi = Wd @ ZiW.T
        WdZiW = Wd.T @ ZiW
        UdZiW = Ud @ ZiW
        ZiWdWZi = self._zs(WdWZi)
        ZiWdWZi2 = ZiWdWZi + ZiWdWZi.T
        ZiUdZiW = self._zs(UdZiW)
        ZiUZiWdWZi = ZiU @ ZiWdWZi
        ZiUZiUdZiW = ZiU @ ZiUdZiW + ZiUd @ self.ZiUZiW
        ZiWdWZiUZi = ZiWdWZi @ ZiU.T
        ZiWdWZiUZi2 = ZiWdWZiUZi + ZiWdWZiUZi.T + ZiUZiWdWZi.T
        ZiUdZi = self._zs(ZiUd.T).T
        ZiUdZi = (ZiUdZi + ZiUdZi.T) / 2
        ZiUZiUdZi = ZiU @ ZiUdZi
        ZiUZiUdZi2 = ZiUZiUdZi + ZiUZiUdZi.T
        ZiUdZiWdWZi = ZiUd @ ZiWdWZi + ZiWdWZi @ ZiUd.T
        WtZiWI = self.WtZiW + np.eye(self.d2)
        ZiWdWtZiWI = ZiWd @ WtZiWI
        vdZiUZiUZiW = vd2 * ZiUZiUZi @ W
        WdWtZiWI = Wd @ WtZiWI
        ZiUZiWdWZiWI = ZiUZi @ WdWtZiWI + ZiWdWZiUZi2 @ W
        vZiUZiUdZi2 = v * ZiUZiUdZi2 - ZiUdZi
        Utemp = vd2 * (-vd2 * ZiUZi2v + ZiWdWZi2 - v2 * (ZiUZiWdWZi + ZiWdWZiUZi2 - 2 * vZiUZiUdZi2)) + \
            v2 * (ZiWdWZi @ WdWZi + WdWZi.T @ ZiWdWZi2 + ZiWdWtZiWI @ ZiWd.T +
                  v2 * (v2 * ZiUd @ ZiUdZi - ZiUdZiWdWZi - ZiUdZiWdWZi.T))
        d3 = np.empty(self.dim)
        d3[:self.v_idx] = au.smat_to_svec((Utemp + Utemp.T) / 2)
        v_Wd_dot = -4 * (v * ZiUZiUdZiW + vdZiUZiUZiW) + ZiUZiWdWZiWI + 2 * ZiUdZiW
        d3[self.v_idx] = vd * (-8 * float(np.sum(ZiUZi2v * Ud)) + vd * (8 * float(np.sum(ZiUZiUZi * U)) -
                                                                        (d1 - 1) / v / v / v)) + \
            4 * v * float(np.sum(vZiUZiUdZi2 * Ud)) + 2 * float(np.sum(v_Wd_dot * Wd))
        Wtemp = 4 * vd * (ZiUdZiW - v2 * ZiUZiUdZiW + ZiUZiWdWZiWI - vdZiUZiUZiW) + \
            4 * v * (ZiUdZiW @ WdZiW + ZiWdWZi @ UdZiW + WdWZi.T @ ZiUdZiW + ZiUdZi @ WdWtZiWI - v2 * ZiUd @ ZiUdZiW) + \
            -2 * (ZiW @ WdZiW @ WdZiW + WdWZi.T @ ZiWdWtZiWI + ZiWdWtZiWI @ WdZiW + ZiWd @ WdZiW.T @ WtZiWI)
        d3[self.v_idx + 1:] = Wtemp.ravel(order="F")
        return d3


class LinMatrixIneq(Cone):
    """linmatrixineq.jl:8-159 (real dense matrices): {w : sum_i w_i A_i psd} for symmetric A_i (side x side, A_1 positive
    definite), barrier -logdet(sum_i w_i A_i), nu = side.  hess_prod! / inv_hess_prod! are the generic explicit-Hessian
    oracles of Cones.jl:101-118; is_dual_feas the generic `true`."""
    ctype = M.CONE_LINMATRIXINEQ

    def __init__(self, As, use_dual=False):
        self.As = [np.asarray(A, dtype=np.float64) for A in As]
        self.side = self.As[0].shape[0]
        assert len(self.As) > 1 and all(A.shape == (self.side, self.side) and np.allclose(A, A.T) for A in self.As)
        assert self.side * (self.side + 1) // 2 >= len(self.As)
        self.use_dual_barrier = use_dual
        super().__init__(len(self.As))

    @property
    def nu(self):
        return float(self.side)

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[0] = 1.0
        return arr

    def update_feas(self):
        # linmatrixineq.jl:87-96
        S = sum(w * A for w, A in zip(self.point, self.As))
        try:
            self.L = np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            return False
        return True

    def update_grad(self):
        # linmatrixineq.jl:98-109: B_i = L^-1 A_i L^-T, grad_i = -tr(B_i)
        L = self.L
        self.B = []
        for A in self.As:
            X = sla.solve_triangular(L, A, lower=True, check_finite=False)
            self.B.append(sla.solve_triangular(L, X.T, lower=True, check_finite=False).T)
        self._grad[:] = [-np.trace(B) for B in self.B]

    def update_hess(self):
        # linmatrixineq.jl:111-123
        self.grad()
        Bm = np.stack([B.ravel() for B in self.B])
        return Bm @ Bm.T

    def hess_prod(self, arr):
        a, vec = _as2d(arr)
        return _ret(np.asarray(self.hess()) @ a, vec)

    def dder3(self, direction):
        # linmatrixineq.jl:147-159
        self.grad()
        D = sum(d * B for d, B in zip(direction, self.B))
        Z = D @ D.T
        return np.array([np.sum(Z * B) for B in self.B])


class GeneralizedPower(Cone):
    """generalizedpower.jl:8-236: (u in R^m_++, w in R^n), prod u_i^(alpha_i) >= |w|_2; barrier
    -log(prod u_i^(2 alpha_i) - |w|^2) - sum (1 - alpha_i) log u_i, nu = m + 1.  No closed-form inverse Hessian:
    inv_hess_prod!, inv_hess and the sqrt oracles are the generic ones of Cones.jl:113-118, 189-259 (explicit
    Hessian + posdef_fact_copy!)."""
    ctype = M.CONE_GENERALIZEDPOWER

    def __init__(self, alpha, n, use_dual=False):
        self.alpha = np.array(alpha, dtype=np.float64)
        self.m = self.alpha.size
        self.n = n
        self.use_dual_barrier = use_dual
        super().__init__(self.m + n)

    @property
    def nu(self):
        return float(self.m + 1)

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[:self.m] = np.sqrt(1 + self.alpha)
        return arr

    def update_feas(self):
        u, w = self.point[:self.m], self.point[self.m:]
        if (u > EPS).all():
            self.z = float(np.exp(2 * np.sum(self.alpha * np.log(u))))
            self.w2 = float(w @ w)
            self.zw = self.z - self.w2
            return self.zw > EPS
        return False

    def is_dual_feas(self):
        u, w = self.dual_point[:self.m], self.dual_point[self.m:]
        if (u > EPS).all():
            p = float(np.exp(2 * np.sum(self.alpha * np.log(u / self.alpha))))
            return (p - float(w @ w)) > EPS
        return False

    def update_grad(self):
        u, w = self.point[:self.m], self.point[self.m:]
        self.zwzwi = (self.z + self.w2) / self.zw
        self._grad[:self.m] = -(self.zwzwi * self.alpha + 1) / u
        self._grad[self.m:] = 2 * w / self.zw

    def update_hess(self):
        g = self.grad()
        m = self.m
        u = self.point[:m]
        aui = 2 * self.alpha / u
        auizzwi = -self.z * aui / self.zw
        zzwim1 = -self.w2 / self.zw
        H = np.zeros((self.dim, self.dim))
        H[:m, :m] = np.outer(aui, auizzwi) * zzwim1
        H[np.arange(m), np.arange(m)] -= g[:m] / u
        H[:m, m:] = np.outer(auizzwi, g[m:])
        H[m:, :m] = H[:m, m:].T
        H[m:, m:] = np.outer(g[m:], g[m:]) + (2 / self.zw) * np.eye(self.n)
        return H

    def hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        m = self.m
        u, w = self.point[:m], self.point[m:]
        prod_u = a[:m] / u[:, None]
        prod_w = 2 * a[m:] / self.zw
        dot1 = -4 * (self.alpha @ prod_u) * self.z / self.zw
        dot2 = (dot1 + 2 * (w @ prod_w)) / self.zw
        dot3 = dot1 - dot2 * self.z
        prod = np.empty_like(a)
        prod[:m] = (prod_u * (1 + self.zwzwi * self.alpha)[:, None] + dot3[None, :] * self.alpha[:, None]) / u[:, None]
        prod[m:] = prod_w + dot2[None, :] * w[:, None]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        m = self.m
        u, w = self.point[:m], self.point[m:]
        u_dir, w_dir = direction[:m], direction[m:]
        alpha, z, zw, w2, zwzwi = self.alpha, self.z, self.zw, self.w2, self.zwzwi
        zzwi = 2 * z / zw
        zwi = 2 / zw
        udu = u_dir / u
        wwd = 2 * float(w @ w_dir)
        c15 = wwd / zw
        audu = float(alpha @ udu)
        sumaudu2 = float(alpha @ (udu ** 2))
        c1 = 2 * zwzwi * audu ** 2 + sumaudu2
        c10 = float(w_dir @ w_dir) + wwd * c15
        c13 = zzwi * (w2 * c1 - 2 * wwd * zwzwi * audu + c10) / zw
        c14 = zzwi * (2 * audu * w2 - wwd) / zw
        d3 = np.empty(self.dim)
        d3[:m] = (c13 * alpha + ((c14 + zwzwi * udu) * alpha + udu) * udu) / u
        c6 = zwi * (z * (4 * audu * c15 - c1) - c10) / zw
        c7 = zwi * (2 * z * audu - wwd) / zw
        d3[m:] = c7 * w_dir + c6 * w
        return d3


def get_central_ray_hypopowermean(alpha):
    """hypopowermean.jl:205-232 (fitted constants of the reference)."""
    alpha = np.asarray(alpha, dtype=np.float64)
    d = alpha.size
    if d == 1:
        w = np.full(1, 1.306563)
    elif d == 2:
        w = 1.0049885 + 0.2986276 * alpha
    elif d <= 5:
        w = 1.0040142949 - 0.0004885108 * d + 0.3016645951 * alpha
    elif d <= 20:
        w = 1.001168 - 4.547017e-05 * d + 3.032880e-01 * alpha
    elif d <= 100:
        w = 1.000069 - 5.469926e-07 * d + 3.074084e-01 * alpha
    else:
        w = 1 + 3.086535e-01 * alpha
    p = np.exp(np.sum(alpha * np.log(w)))
    u = p - p / d * np.sum(alpha / (w * w - 1))
    return u, w


class HypoPowerMean(Cone):
    """hypopowermean.jl:8-232: (u, w in R^d_++), u <= prod w_i^alpha_i; barrier -log(prod w_i^alpha_i - u) - sum log w_i,
    nu = dim.  No closed-form inverse Hessian: the generic oracles of Cones.jl:113-118, 189-259 apply."""
    ctype = M.CONE_HYPOPOWERMEAN

    def __init__(self, alpha, use_dual=False):
        self.alpha = np.array(alpha, dtype=np.float64)
        self.use_dual_barrier = use_dual
        super().__init__(1 + self.alpha.size)

    @property
    def nu(self):
        return float(self.dim)

    def set_initial_point(self, arr):
        d = self.dim - 1
        if np.all(self.alpha == 1.0 / d):
            c = np.sqrt(5.0 * d * d + 2 * d + 1)
            arr[0] = -np.sqrt((-c + 3 * d + 1) / (2.0 + 2 * d))
            arr[1:] = (c - d + 1) / np.sqrt((1 + d) * (-2 * c + 6 * d + 2))
        else:
            arr[0], arr[1:] = get_central_ray_hypopowermean(self.alpha)
        return arr

    def update_feas(self):
        u, w = self.point[0], self.point[1:]
        if (w > EPS).all():
            self.phi = float(np.exp(np.sum(self.alpha * np.log(w))))
            self.zeta = self.phi - u
            return self.zeta > EPS
        return False

    def is_dual_feas(self):
        u, w = self.dual_point[0], self.dual_point[1:]
        if u < -EPS and (w > EPS).all():
            return bool(np.exp(np.sum(self.alpha * np.log(w / self.alpha))) + u > EPS)
        return False

    def update_grad(self):
        w = self.point[1:]
        self._grad[0] = 1.0 / self.zeta
        self._grad[1:] = (-self.phi / self.zeta * self.alpha - 1) / w

    def update_hess(self):
        self.grad()
        w, alpha, zeta = self.point[1:], self.alpha, self.zeta
        zip_ = self.phi / zeta
        awi = alpha / w
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zeta ** -2
        H[0, 1:] = H[1:, 0] = -(zip_ * awi) / zeta
        H[1:, 1:] = zip_ * (zip_ - 1) * np.outer(awi, awi)
        idx = np.arange(1, self.dim)
        H[idx, idx] = (zip_ * awi * (1 + alpha * (zip_ - 1)) + 1.0 / w) / w
        return H

    def hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        w, alpha, zeta = self.point[1:], self.alpha, self.zeta
        zip_ = self.phi / zeta
        p = a[0]
        rwi = a[1:] / w[:, None]
        c0 = alpha @ rwi
        c1 = zip_ * c0 - p / zeta
        c2 = c1 - c0
        prod = np.empty_like(a)
        prod[0] = c1 / -zeta
        prod[1:] = (alpha[:, None] * zip_ * (c2[None, :] + rwi) + rwi) / w[:, None]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        w, alpha, zeta, phi = self.point[1:], self.alpha, self.zeta, self.phi
        p, r = direction[0], direction[1:]
        zip_ = phi / zeta
        rwi = r / w
        c0 = float(rwi @ alpha)
        c6 = float((rwi ** 2) @ alpha)
        zichi = (p - phi * c0) / zeta
        c1 = zichi ** 2 + zip_ * (c6 - c0 ** 2) / 2
        c7 = zip_ * (c1 - c6 / 2 + c0 * (zichi + c0 / 2))
        c8 = -zip_ * (zichi + c0)
        d3 = np.empty(self.dim)
        d3[0] = -c1 / zeta
        d3[1:] = (alpha * (c7 + rwi * (c8 + zip_ * rwi)) + rwi ** 2) / w
        return d3


# ---------------------------------------------------------------------------------------------------------------------
# EpiTrRelEntropyTri: oracle restatement only (no device kernels yet)
# ---------------------------------------------------------------------------------------------------------------------
def _log_divdiff(nodes):
    """Divided difference log[x_0, .., x_k] with confluent nodes handled by the derivative limit: sort the nodes; a run of
    equal nodes gives log^(k)(x) / k!, otherwise the usual recursion (the Δ2 / Δ3 / Δ4 matrices of
    epitrrelentropytri.jl:385-573 are these numbers for k = 1, 2, 3)."""
    xs = sorted(float(x) for x in nodes)

    def dd(i, j):
        k = j - i
        if k == 0:
            return np.log(xs[i])
        if abs(xs[j] - xs[i]) <= 1e-9 * max(abs(xs[i]), abs(xs[j])):
            x = 0.5 * (xs[i] + xs[j])
            # log^(k)(x) / k! = (-1)^(k-1) / (k x^k)
            return (-1.0) ** (k - 1) / (k * x ** k)
        return (dd(i + 1, j) - dd(i, j - 1)) / (xs[j] - xs[i])
    return dd(0, len(xs) - 1)


class _LogFrechet:
    """Frechet derivatives of the matrix logarithm at X = Q diag(lam) Q' (Daleckii-Krein)."""

    def __init__(self, X):
        self.lam, self.Q = np.linalg.eigh(X)
        d = self.lam.size
        lam = self.lam
        self.D1 = np.array([[_log_divdiff((lam[i], lam[j])) for j in range(d)] for i in range(d)])
        self.D2 = np.array([[[_log_divdiff((lam[i], lam[j], lam[k])) for k in range(d)] for j in range(d)]
                            for i in range(d)])
        self._D3 = None
        self.log = (self.Q * np.log(lam)) @ self.Q.T

    @property
    def D3(self):
        if self._D3 is None:
            lam, d = self.lam, self.lam.size
            self._D3 = np.array([[[[_log_divdiff((lam[i], lam[j], lam[k], lam[l])) for l in range(d)] for k in range(d)]
                                  for j in range(d)] for i in range(d)])
        return self._D3

    def _in(self, H):
        return self.Q.T @ H @ self.Q

    def _out(self, Ht):
        return self.Q @ Ht @ self.Q.T

    def d1(self, H):
        return self._out(self.D1 * self._in(H))

    def d2(self, H, K):
        """D^2 log(X)[H, K] (symmetric bilinear)."""
        Ht, Kt = self._in(H), self._in(K)
        out = np.einsum("ikj,ik,kj->ij", self.D2, Ht, Kt) + np.einsum("ikj,ik,kj->ij", self.D2, Kt, Ht)
        return self._out(out)

    def d3(self, A, B, C):
        """D^3 log(X)[A, B, C] (symmetric trilinear): sum over the six orderings of D3[i,k,l,j] a_ik b_kl c_lj."""
        import itertools
        mats = [self._in(A), self._in(B), self._in(C)]
        out = np.zeros_like(mats[0])
        for p in itertools.permutations(range(3)):
            out += np.einsum("iklj,ik,kl,lj->ij", self.D3, mats[p[0]], mats[p[1]], mats[p[2]])
        return self._out(out)


class EpiTrRelEntropyTri(Cone):
    """epitrrelentropytri.jl:8-573: (u, svec(V), svec(W)) with V, W positive definite d x d and
    u >= tr(W log W - W log V); barrier -log(u - tr(W log W - W log V)) - logdet V - logdet W, nu = 2 d + 1.
    (ctype = M.CONE_EPITRRELENTROPYTRI.)
    Restated through the Frechet derivatives of the matrix logarithm (Daleckii-Krein with confluent divided differences):
    with z = u - phi, phi = tr(W log W) - tr(W log V),
        phi_W = log W + I - log V,         phi_V = -Dlog(V)[W],
        phi_WW[H] = Dlog(W)[H],            phi_WV[K] = -Dlog(V)[K],     phi_VV[K] = -D2log(V)[W, K],
    and the third derivatives D2log(W)[.,.], D2log(V)[.,.], D3log(V)[W,.,.]; the barrier derivatives follow from
    F = -log z - logdet V - logdet W.  inv_hess_prod! is the generic factorisation fallback.  Oracle only: the device
    kernels of this cone (two batched Jacobi decompositions + a post kernel, like EpiPerSepSpectral) are not built yet."""

    def __init__(self, dim, use_dual=False):
        assert dim > 2 and dim % 2 == 1
        self.vw = (dim - 1) // 2
        self.d = M.svec_side(self.vw)
        self.use_dual_barrier = use_dual
        super().__init__(dim)

    @property
    def nu(self):
        return float(2 * self.d + 1)

    def set_initial_point(self, arr):
        # epitrrelentropytri.jl:121-135: diagonal V, W from the vector cone's central ray
        arr[:] = 0.0
        u, v, w = EpiRelEntropy.central_ray(self.d)
        arr[0] = u
        dg = np.array([j * (j + 1) // 2 + j for j in range(self.d)])
        arr[1 + dg] = v
        arr[1 + self.vw + dg] = w
        return arr

    def _split(self, vec):
        from . import arrayutil as au
        return vec[0], au.svec_to_smat(vec[1:1 + self.vw]), au.svec_to_smat(vec[1 + self.vw:])

    def _join(self, u, Vm, Wm):
        from . import arrayutil as au
        return np.concatenate(([u], au.smat_to_svec((Vm + Vm.T) / 2), au.smat_to_svec((Wm + Wm.T) / 2)))

    def update_feas(self):
        # epitrrelentropytri.jl:137-166
        u, V, W = self._split(self.point)
        for X in (V, W):
            try:
                np.linalg.cholesky(X)
            except np.linalg.LinAlgError:
                return False
        self.V, self.W = V, W
        self.LV, self.LW = _LogFrechet(V), _LogFrechet(W)
        if self.LV.lam.min() <= 0 or self.LW.lam.min() <= 0:
            return False
        self.z = float(u - np.sum(W * (self.LW.log - self.LV.log)))
        return self.z > 0

    def _dz(self):
        """Gradient of z = u - phi as (1, dz/dV, dz/dW) (matrices)."""
        d = self.d
        return 1.0, self.LV.d1(self.W), -(self.LW.log + np.eye(d) - self.LV.log)

    def update_grad(self):
        # epitrrelentropytri.jl:168-208
        self.Vi, self.Wi = np.linalg.inv(self.V), np.linalg.inv(self.W)
        zu, zV, zW = self._dz()
        self._grad[:] = self._join(-zu / self.z, -zV / self.z - self.Vi, -zW / self.z - self.Wi)

    def _d2z(self, dV, dW):
        """Hessian of z applied to the direction (0, dV, dW): returns (matrix for V, matrix for W)."""
        hV = self.LV.d2(self.W, dV) + self.LV.d1(dW)
        hW = self.LV.d1(dV) - self.LW.d1(dW)
        return hV, hW

    def hess_prod(self, arr):
        # epitrrelentropytri.jl:210-267 (update_hess; hess_prod! is the generic explicit product there)
        self.grad()
        a, vec = _as2d(arr)
        z = self.z
        zu, zV, zW = self._dz()
        prod = np.empty_like(a)
        for j in range(a.shape[1]):
            du, dV, dW = self._split(a[:, j])
            dz = zu * du + float(np.sum(zV * dV)) + float(np.sum(zW * dW))
            hV, hW = self._d2z(dV, dW)
            prod[:, j] = self._join(dz * zu / z ** 2,
                                    dz * zV / z ** 2 - hV / z + self.Vi @ dV @ self.Vi,
                                    dz * zW / z ** 2 - hW / z + self.Wi @ dW @ self.Wi)
        return _ret(prod, vec)

    def update_hess(self):
        H = self.hess_prod(np.eye(self.dim))
        return (H + H.T) / 2

    def dder3(self, direction):
        # epitrrelentropytri.jl:269-383: -1/2 of the third directional derivative of the barrier
        self.grad()
        z = self.z
        zu, zV, zW = self._dz()
        du, dV, dW = self._split(direction)
        dz = zu * du + float(np.sum(zV * dV)) + float(np.sum(zW * dW))
        hV, hW = self._d2z(dV, dW)
        dHd = float(np.sum(hV * dV)) + float(np.sum(hW * dW))
        # third derivative of z along (dV, dW) twice
        tV = self.LV.d3(self.W, dV, dV) + 2 * self.LV.d2(dW, dV)
        tW = self.LV.d2(dV, dV) - self.LW.d2(dW, dW)
        c1 = -2 * dz ** 2 / z ** 3
        c2 = 2 * dz / z ** 2
        c3 = dHd / z ** 2
        t3u = c1 * zu + c3 * zu
        t3V = c1 * zV + c2 * hV + c3 * zV - tV / z - 2 * self.Vi @ dV @ self.Vi @ dV @ self.Vi
        t3W = c1 * zW + c2 * hW + c3 * zW - tW / z - 2 * self.Wi @ dW @ self.Wi @ dW @ self.Wi
        return -0.5 * self._join(t3u, t3V, t3W)

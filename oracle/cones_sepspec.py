"""CPU restatement of EpiPerSepSpectral{MatrixCSqr} (oracle; test infrastructure).

Epigraph of the perspective of a separable spectral function h over the real symmetric PSD cone:
points (u, v, svec W), barrier -log(u - v * sum h(lambda_i(W / v))) - log(v) - logdet(W).

reference: src/Cones/epipersepspectral/epipersepspectral.jl:28-86 (type, nu = 2 + d),
matrixcsqr.jl:75-564 (oracles), sepspectralfun.jl:17-116 (h functions),
arrayutilities.jl:387-424 (eig_dot_kron!), dense.jl:69 (update_eigen! = LAPACK syev, here numpy's
eigh = LAPACK syevd: same spectrum, eigenvectors equal up to sign / rotation inside eigenspaces,
to which every oracle below is invariant).
"""
import numpy as np

from hypatia_b200.host import models as M
from . import arrayutil as au
from . import linalg as la
from .cones import Cone, EPS, _as2d, _ret

H_INV, H_NEGLOG, H_NEGENTROPY, H_POWER12 = 0, 1, 2, 3


class SepSpectralFun:
    """sepspectralfun.jl:17-116: h_val, h_conj_dom_pos, h_conj, h_der1..3, get_initial_point."""

    def __init__(self, kind, p=1.5):
        self.kind = int(kind)
        self.p = float(p)
        if self.kind == H_POWER12:
            assert 1 < self.p <= 2

    def val(self, x):
        k = self.kind
        if k == H_INV:
            return float(np.sum(1.0 / x))
        if k == H_NEGLOG:
            return float(-np.sum(np.log(x)))
        if k == H_NEGENTROPY:
            return float(np.sum(x * np.log(x)))
        return float(np.sum(x ** self.p))

    def conj_dom_pos(self):
        return self.kind in (H_INV, H_NEGLOG)

    def conj(self, x):
        k = self.kind
        if k == H_INV:
            return float(-2 * np.sum(np.sqrt(x)))
        if k == H_NEGLOG:
            return float(-x.size - np.sum(np.log(x)))
        if k == H_NEGENTROPY:
            return float(np.sum(np.exp(-x - 1)))
        p = self.p
        qq = p / (p - 1)
        return float((p - 1) * np.sum(np.where(x >= 0, 0.0, (np.abs(x) / p) ** qq)))

    def der1(self, x):
        k = self.kind
        if k == H_INV:
            return -x ** -2.0
        if k == H_NEGLOG:
            return -1.0 / x
        if k == H_NEGENTROPY:
            return 1 + np.log(x)
        return self.p * x ** (self.p - 1)

    def der2(self, x):
        k = self.kind
        if k == H_INV:
            return 2 * x ** -3.0
        if k == H_NEGLOG:
            return x ** -2.0
        if k == H_NEGENTROPY:
            return 1.0 / x
        p = self.p
        return p * (p - 1) * x ** (p - 2)

    def der3(self, x):
        k = self.kind
        if k == H_INV:
            return -6 * x ** -4.0
        if k == H_NEGLOG:
            return -2 * x ** -3.0
        if k == H_NEGENTROPY:
            return -x ** -2.0
        p = self.p
        return p * (p - 1) * (p - 2) * x ** (p - 3)

    def initial_point(self, d):
        if self.kind in (H_INV, H_POWER12):
            return (2.0 * d, 1.0, 1.0)
        return (1.0, 1.0, 1.0)


class EpiPerSepSpectralMat(Cone):
    """matrixcsqr.jl:75-564 (real symmetric case)."""
    ctype = M.CONE_EPIPERSEPSPECTRAL_MAT

    def __init__(self, dim, hkind=H_NEGLOG, hparam=1.5, use_dual=False):
        self.d = au.svec_side(dim - 2)
        self.h = SepSpectralFun(hkind, hparam)
        self.use_dual_barrier = use_dual
        super().__init__(dim)

    @property
    def nu(self):
        return 2.0 + self.d

    def reset_data(self):
        super().reset_data()
        self.hess_aux_updated = self.inv_hess_aux_updated = self.dder3_aux_updated = False

    def set_initial_point(self, arr):
        # matrixcsqr.jl:75-88
        u, v, w0 = self.h.initial_point(self.d)
        arr[:] = 0.0
        arr[0], arr[1] = u, v
        arr[2:] = au.smat_to_svec(w0 * np.eye(self.d))
        return arr

    def update_feas(self):
        # matrixcsqr.jl:91-115: Cholesky gate, then syev of W / v
        v = self.point[1]
        if v > EPS:
            W = au.svec_to_smat(self.point[2:])
            if la.posdef_fact(W).issuccess():
                lam, X = np.linalg.eigh(W / v)
                self.viw_lam, self.viw_X = lam, X
                if (lam > EPS).all():
                    self.phi = self.h.val(lam)
                    self.zeta = self.point[0] - v * self.phi
                    return self.zeta > EPS
        return False

    def is_dual_feas(self):
        # matrixcsqr.jl:119-138
        u = self.dual_point[0]
        if u < EPS:
            return False
        W = au.svec_to_smat(self.dual_point[2:])
        if self.h.conj_dom_pos():
            if not la.posdef_fact(W).issuccess():
                return False
        lam = np.linalg.eigvalsh(W / u)
        return bool(self.dual_point[1] - u * self.h.conj(lam) > EPS)

    def update_grad(self):
        # matrixcsqr.jl:140-165
        v = self.point[1]
        self.zetai = 1.0 / self.zeta
        lam, X = self.viw_lam, self.viw_X
        self.dh = self.h.der1(lam)
        self.sigma = self.phi - float(lam @ self.dh)
        self.w_lam = v * lam
        self.w_lami = 1.0 / self.w_lam
        g = self._grad
        g[0] = -self.zetai
        g[1] = -1.0 / v + self.zetai * self.sigma
        wd = self.zetai * self.dh - self.w_lami
        g[2:] = au.smat_to_svec((X * wd) @ X.T)

    def update_hess_aux(self):
        # matrixcsqr.jl:167-217
        if self.hess_aux_updated:
            return
        self.grad()
        lam, dh = self.viw_lam, self.dh
        d = self.d
        self.d2h = self.h.der2(lam)
        rteps = np.sqrt(EPS)
        lam_d = lam[:, None] - lam[None, :]          # [i, j] = lam_i - lam_j
        lam_d[np.abs(lam_d) < rteps] = 0.0
        self.lam_d = lam_d
        zero = lam_d == 0.0
        safe = np.where(zero, 1.0, lam_d)
        Dh = np.where(zero, (self.d2h[:, None] + self.d2h[None, :]) / 2,
                      (dh[:, None] - dh[None, :]) / safe)
        Dh[np.diag_indices(d)] = self.d2h
        self.Dh = Dh
        zetaivi = self.zetai / self.point[1]
        self.theta = zetaivi * Dh + np.outer(self.w_lami, self.w_lami)
        self.hess_aux_updated = True

    def _rot(self, a_w):
        """svec columns -> stack of V' M_j V."""
        X = self.viw_X
        return X.T @ au.svecs_to_smats(a_w) @ X

    def _unrot(self, mats):
        X = self.viw_X
        return au.smats_to_svecs(X @ mats @ X.T)

    def _eig_dot_kron(self, inner):
        """arrayutilities.jl:387-424: column (i, j) = svec(X (inner .* (X' E_ij X)) X')."""
        L = au.svec_length(self.d)
        E = au.svecs_to_smats(np.eye(L))
        X = self.viw_X
        return au.smats_to_svecs(X @ (inner[None] * (X.T @ E @ X)) @ X.T)

    def update_hess(self):
        # matrixcsqr.jl:219-271
        self.update_hess_aux()
        v, zetai, sigma = self.point[1], self.zetai, self.sigma
        X, lam, dh, d2h = self.viw_X, self.viw_lam, self.dh, self.d2h
        zetai2 = zetai ** 2
        zetaivi = zetai / v
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zetai2
        H[0, 1] = H[1, 0] = -zetai2 * sigma
        H[1, 1] = v ** -2 + (zetai * sigma) ** 2 + zetaivi * float((lam ** 2) @ d2h)
        wd = -zetai * dh
        Hwu = au.smat_to_svec((X * wd) @ X.T)
        H[0, 2:] = H[2:, 0] = zetai * Hwu
        wd = wd * (-zetai * sigma) - zetaivi * d2h * lam
        H[1, 2:] = H[2:, 1] = au.smat_to_svec((X * wd) @ X.T)
        H[2:, 2:] = self._eig_dot_kron(self.theta) + np.outer(Hwu, Hwu)
        return H

    def hess_prod(self, arr):
        # matrixcsqr.jl:273-319
        self.update_hess_aux()
        a, vec = _as2d(arr)
        v, zetai, sigma = self.point[1], self.zetai, self.sigma
        lam, dh, Dh, w_lami = self.viw_lam, self.dh, self.Dh, self.w_lami
        zetaivi = zetai / v
        d = self.d
        idx = np.arange(d)
        p, q = a[0], a[1]
        r = self._rot(a[2:])
        rdiag = r[:, idx, idx]
        sum1 = rdiag @ dh
        c1 = -zetai * (p - sigma * q - sum1) * zetai
        t = r.copy()
        t[:, idx, idx] -= q[:, None] * lam[None, :]
        w_aux = zetaivi * Dh[None] * t
        c2 = w_aux[:, idx, idx] @ lam
        w_aux = w_aux + np.outer(w_lami, w_lami)[None] * r
        w_aux[:, idx, idx] += c1[:, None] * dh[None, :]
        prod = np.empty_like(a)
        prod[0] = -c1
        prod[1] = c1 * sigma - c2 + q / v / v
        prod[2:] = self._unrot(w_aux)
        return _ret(prod, vec)

    def update_inv_hess_aux(self):
        # matrixcsqr.jl:321-359
        if self.inv_hess_aux_updated:
            return
        self.update_hess_aux()
        v, sigma = self.point[1], self.sigma
        lam, dh = self.viw_lam, self.dh
        zetaivi = self.zetai / v
        diag_theta = np.diag(self.theta)
        wd = zetaivi * self.d2h
        self.alpha = dh / diag_theta
        wd = wd * lam
        self.gamma = wd / diag_theta
        zeta2beta = self.zeta ** 2 + float(dh @ self.alpha)
        c0 = sigma + float(dh @ self.gamma)
        c1 = c0 / zeta2beta
        sum1 = float(((lam + c1 * self.alpha - self.gamma) * wd).sum())
        c3 = v ** -2 + sigma * c1 + sum1
        self.c0 = c0
        self.c4 = 1.0 / (c3 - c0 * c1)
        self.c5 = zeta2beta * c3
        self.inv_hess_aux_updated = True

    def update_inv_hess(self):
        # matrixcsqr.jl:361-400
        self.update_inv_hess_aux()
        X, c4 = self.viw_X, self.c4
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = c4 * self.c5
        Hiuv = Hi[0, 1] = Hi[1, 0] = c4 * self.c0
        Hi[1, 1] = c4
        gamma_vec = au.smat_to_svec((X * self.gamma) @ X.T)
        Hi[1, 2:] = Hi[2:, 1] = c4 * gamma_vec
        HiuW = au.smat_to_svec((X * self.alpha) @ X.T) + Hiuv * gamma_vec
        Hi[0, 2:] = Hi[2:, 0] = HiuW
        Hi[2:, 2:] = self._eig_dot_kron(1.0 / self.theta) + c4 * np.outer(gamma_vec, gamma_vec)
        return Hi

    def inv_hess_prod(self, arr):
        # matrixcsqr.jl:402-447
        self.update_inv_hess_aux()
        a, vec = _as2d(arr)
        d = self.d
        idx = np.arange(d)
        alpha, gamma, c0, c4, c5 = self.alpha, self.gamma, self.c0, self.c4, self.c5
        p, q = a[0], a[1]
        r = self._rot(a[2:])
        rdiag = r[:, idx, idx]
        qgr = q + rdiag @ gamma
        cu = c4 * (c5 * p + c0 * qgr)
        cv = c4 * (c0 * p + qgr)
        prod = np.empty_like(a)
        prod[0] = cu + rdiag @ alpha
        prod[1] = cv
        w_prod = r / self.theta[None]
        w_prod[:, idx, idx] += p[:, None] * alpha[None, :] + cv[:, None] * gamma[None, :]
        prod[2:] = self._unrot(w_prod)
        return _ret(prod, vec)

    def _d2h_slice(self, i, j):
        """Delta2h[:, (i, j)] of update_dder3_aux (matrixcsqr.jl:449-502): second divided
        differences over the index triple sorted ascending (a <= b <= c):
        Delta_ab == 0: (Delta_ac == 0 ? (d3h_a + d3h_b + d3h_c) / 6 : (Dh[a, b] - Dh[b, c]) / Delta_ac)
        else (Dh[a, c] - Dh[b, c]) / Delta_ab."""
        d = self.d
        out = np.empty(d)
        for k in range(d):
            a, b, c = sorted((i, j, k))
            dab = self.lam_d[a, b]
            if dab == 0:
                dac = self.lam_d[a, c]
                if dac == 0:
                    t = (self.d3h[a] + self.d3h[b] + self.d3h[c]) / 6
                else:
                    t = (self.Dh[a, b] - self.Dh[b, c]) / dac
            else:
                t = (self.Dh[a, c] - self.Dh[b, c]) / dab
            out[k] = t
        return out

    def dder3(self, direction):
        # matrixcsqr.jl:504-564
        self.update_hess_aux()
        if not self.dder3_aux_updated:
            self.d3h = self.h.der3(self.viw_lam)
            self.dder3_aux_updated = True
        d = self.d
        v, zetai, sigma = self.point[1], self.zetai, self.sigma
        lam, dh, Dh, w_lami = self.viw_lam, self.dh, self.Dh, self.w_lami
        vi = 1.0 / v
        p, q = direction[0], direction[1]
        r = self._rot(direction[2:].reshape(-1, 1))[0]
        viq = vi * q
        xi = vi * r - viq * np.diag(lam)
        xib = zetai * Dh * xi
        sum1 = float(dh @ np.diag(r))
        zetaichi = zetai * (p - sigma * q - sum1)
        xibxi = float((xib * xi).sum()) / 2
        c1 = -zetai * (zetaichi ** 2 + v * xibxi)
        w_aux = xib * (zetaichi + viq)
        for j in range(d):
            for i in range(j + 1):
                t = zetai * float((xi[:, i] * self._d2h_slice(i, j)) @ xi[:, j])
                w_aux[i, j] -= t
                if i != j:
                    w_aux[j, i] -= t
        c2 = float(lam @ np.diag(w_aux))
        rs = (w_lami[:, None] * r) * np.sqrt(w_lami)[None, :]
        w_aux = w_aux + rs @ rs.T
        w_aux[np.diag_indices(d)] += c1 * dh
        d3 = np.empty(self.dim)
        d3[0] = -c1
        d3[1] = c1 * sigma - c2 + xibxi + viq ** 2 / v
        d3[2:] = self._unrot(w_aux[None])[:, 0]
        return d3


class EpiPerSepSpectralVec(Cone):
    """EpiPerSepSpectral{VectorCSqr} (vectorcsqr.jl:1-357): (u, v, w), w in R^d_++, barrier
    -log(u - v sum h(w_i / v)) - log(v) - sum log(w_i), nu = 2 + d."""
    ctype = M.CONE_EPIPERSEPSPECTRAL_VEC

    def __init__(self, dim, hkind=H_NEGLOG, hparam=1.5, use_dual=False):
        self.d = dim - 2
        self.h = SepSpectralFun(hkind, hparam)
        self.use_dual_barrier = use_dual
        super().__init__(dim)

    @property
    def nu(self):
        return 2.0 + self.d

    def reset_data(self):
        super().reset_data()
        self.hess_aux_updated = self.inv_hess_aux_updated = False

    def set_initial_point(self, arr):
        u, v, w0 = self.h.initial_point(self.d)
        arr[0], arr[1] = u, v
        arr[2:] = w0
        return arr

    def update_feas(self):
        v, w = self.point[1], self.point[2:]
        if v > EPS and (w > EPS).all():
            self.viw = w / v
            self.phi = self.h.val(self.viw)
            self.zeta = self.point[0] - v * self.phi
            return self.zeta > EPS
        return False

    def is_dual_feas(self):
        u = self.dual_point[0]
        if u < EPS:
            return False
        w = self.dual_point[2:]
        if self.h.conj_dom_pos() and (w < EPS).any():
            return False
        return bool(self.dual_point[1] - u * self.h.conj(w / u) > EPS)

    def update_grad(self):
        v, w = self.point[1], self.point[2:]
        self.zetai = 1.0 / self.zeta
        self.dh = self.h.der1(self.viw)
        self.sigma = self.phi - float(self.viw @ self.dh)
        self.wi = 1.0 / w
        g = self._grad
        g[0] = -self.zetai
        g[1] = -1.0 / v + self.zetai * self.sigma
        g[2:] = -self.wi + self.zetai * self.dh

    def update_hess_aux(self):
        if not self.hess_aux_updated:
            self.grad()
            self.d2h = self.h.der2(self.viw)
            self.hess_aux_updated = True

    def update_hess(self):
        self.update_hess_aux()
        v, zetai, sigma = self.point[1], self.zetai, self.sigma
        viw, wi, dh, d2h = self.viw, self.wi, self.dh, self.d2h
        zetaivi = zetai / v
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zetai ** 2
        H[0, 1] = H[1, 0] = -zetai ** 2 * sigma
        term = viw * zetaivi * d2h
        H[1, 1] = v ** -2 + (zetai * sigma) ** 2 + float(viw @ term)
        H[0, 2:] = H[2:, 0] = -zetai ** 2 * dh
        H[1, 2:] = H[2:, 1] = sigma * zetai ** 2 * dh - term
        H[2:, 2:] = zetai ** 2 * np.outer(dh, dh) + np.diag(zetaivi * d2h + wi ** 2)
        return H

    def hess_prod(self, arr):
        self.update_hess_aux()
        a, vec = _as2d(arr)
        v, w = self.point[1], self.point[2:]
        zetai, sigma = self.zetai, self.sigma
        zetaivi = zetai / v
        p, q, r = a[0], a[1], a[2:]
        viq = q / v
        xib = (zetaivi * self.d2h)[:, None] * (r - viq[None, :] * w[:, None])
        c1 = -zetai * (p - sigma * q - self.dh @ r) * zetai
        prod = np.empty_like(a)
        prod[0] = -c1
        prod[1] = c1 * sigma - self.viw @ xib + viq / v
        prod[2:] = c1[None, :] * self.dh[:, None] + xib + (self.wi ** 2)[:, None] * r
        return _ret(prod, vec)

    def update_inv_hess_aux(self):
        if self.inv_hess_aux_updated:
            return
        self.update_hess_aux()
        v, sigma = self.point[1], self.sigma
        zetaivi = self.zetai / v
        w1 = zetaivi * self.d2h
        self.m = 1.0 / (w1 + self.wi ** 2)
        self.alpha = self.m * self.dh
        w1 = w1 * self.viw
        self.gamma = self.m * w1
        zeta2beta = self.zeta ** 2 + float(self.dh @ self.alpha)
        c0 = sigma + float(self.dh @ self.gamma)
        c1 = c0 / zeta2beta
        sum1 = float(((self.viw + c1 * self.alpha - self.gamma) * w1).sum())
        c3 = v ** -2 + sigma * c1 + sum1
        self.c0, self.c4, self.c5 = c0, 1.0 / (c3 - c0 * c1), zeta2beta * c3
        self.inv_hess_aux_updated = True

    def update_inv_hess(self):
        self.update_inv_hess_aux()
        c0, c4, c5 = self.c0, self.c4, self.c5
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = c4 * c5
        Hi[0, 1] = Hi[1, 0] = c4 * c0
        Hi[1, 1] = c4
        Hiv = c4 * self.gamma
        Hi[1, 2:] = Hi[2:, 1] = Hiv
        Hi[0, 2:] = Hi[2:, 0] = self.alpha + c0 * Hiv
        Hi[2:, 2:] = np.outer(Hiv, self.gamma) + np.diag(self.m)
        return Hi

    def inv_hess_prod(self, arr):
        self.update_inv_hess_aux()
        a, vec = _as2d(arr)
        c0, c4, c5 = self.c0, self.c4, self.c5
        p, q, r = a[0], a[1], a[2:]
        qgr = q + self.gamma @ r
        cu = c4 * (c5 * p + c0 * qgr)
        cv = c4 * (c0 * p + qgr)
        prod = np.empty_like(a)
        prod[0] = cu + self.alpha @ r
        prod[1] = cv
        prod[2:] = p[None, :] * self.alpha[:, None] + cv[None, :] * self.gamma[:, None] + self.m[:, None] * r
        return _ret(prod, vec)

    def dder3(self, direction):
        self.update_hess_aux()
        d3h = self.h.der3(self.viw)
        v, w = self.point[1], self.point[2:]
        zetai, sigma = self.zetai, self.sigma
        zetaivi = zetai / v
        p, q, r = direction[0], direction[1], direction[2:]
        viq = q / v
        xi = r - viq * w
        xib = zetaivi * self.d2h * xi
        zetaichi = zetai * (p - sigma * q - float(self.dh @ r))
        xibxi = float(xib @ xi) / 2
        c1 = -zetai * (zetaichi ** 2 + xibxi)
        c2 = -zetai / 2
        xi = xi / v
        w_aux = xib * (zetaichi + viq) + c2 * d3h * xi * xi
        d3 = np.empty(self.dim)
        d3[0] = -c1
        d3[1] = c1 * sigma - float(self.viw @ w_aux) + (xibxi + viq ** 2) / v
        d3[2:] = c1 * self.dh + w_aux + (r * self.wi) ** 2 * self.wi
        return d3

"""CPU restatement of the reference's cone barrier oracles (oracle; test infrastructure).

Per-cone classes follow the reference's lazily-updated state machine (feas -> grad -> hess /
inv_hess / hess_fact) and method names; `OracleConeBlock` adapts a list of them to the batched
batched ConeBlock interface (oracle/layout.py) by looping over cones, as the reference's callers do.

reference: src/Cones/Cones.jl:27-310 (generic API, use_sqrt_hess_oracles, update_hess_fact,
check_numerics, get_proxsqr), nonnegative.jl:42-145, epinormeucl.jl:44-228,
possemideftri.jl:69-207, hypoperlogdettri.jl:80-368 (+ hypoperlog.jl:289-319 central ray),
hyporootdettri.jl:82-324.
"""
import numpy as np
import scipy.linalg as sla

from hypatia_b200.host import models as M
from .layout import OracleConeBlockBase as ConeBlock
from . import arrayutil as au
from . import linalg as la

EPS = np.finfo(np.float64).eps
RT2 = np.sqrt(2.0)


def _as2d(arr):
    a = np.asarray(arr, dtype=np.float64)
    return (a.reshape(-1, 1), True) if a.ndim == 1 else (a, False)


def _ret(prod, was1d):
    return prod[:, 0] if was1d else prod


class Cone:
    use_dual_barrier = False

    def __init__(self, dim):
        self.dim = dim
        self.setup_data()

    # ---- generic state machine (Cones.jl:136-186) ----
    def setup_data(self):
        self.reset_data()
        self.point = np.zeros(self.dim)
        self.dual_point = np.zeros(self.dim)
        self._grad = np.zeros(self.dim)

    def reset_data(self):
        self.feas_updated = self.grad_updated = self.hess_updated = False
        self.inv_hess_updated = self.hess_fact_updated = False

    def load_point(self, point, scal=None):
        self.point[:] = point if scal is None else scal * np.asarray(point)

    def load_dual_point(self, point):
        self.dual_point[:] = point

    def is_feas(self):
        if not self.feas_updated:
            self._is_feas = bool(self.update_feas())
            self.feas_updated = True
        return self._is_feas

    def is_dual_feas(self):
        return True

    def grad(self):
        if not self.grad_updated:
            assert self.is_feas()
            self.update_grad()
            self.grad_updated = True
        return self._grad

    def hess(self):
        if not self.hess_updated:
            self.grad()
            self._hess = self.update_hess()
            self.hess_updated = True
        return self._hess

    def inv_hess(self):
        if not self.inv_hess_updated:
            self.grad()
            self._inv_hess = self.update_inv_hess()
            self.inv_hess_updated = True
        return self._inv_hess

    def use_dder3(self):
        return True

    # ---- generic fallbacks (Cones.jl:101-118, 189-259) ----
    def update_hess_fact(self):
        if self.hess_fact_updated:
            return self.hess_fact.issuccess()
        self.hess_fact = la.posdef_fact_copy(self.hess(), try_shift=False)
        self.hess_fact_updated = True
        return self.hess_fact.issuccess()

    def use_sqrt_hess_oracles(self, arr_dim):
        if not self.hess_fact_updated:
            if arr_dim < self.dim:
                return False
            if not self.update_hess_fact():
                return False
        return isinstance(self.hess_fact, la.Cholesky)

    def sqrt_hess_prod(self, arr):
        assert self.hess_fact_updated
        return self.hess_fact.U @ arr

    def inv_sqrt_hess_prod(self, arr):
        assert self.hess_fact_updated
        return sla.solve_triangular(self.hess_fact.U, arr, trans="T")

    def hess_prod_slow(self, arr):
        return self.hess_prod(arr)

    # generic inverse-Hessian oracles for cones without closed forms (Cones.jl:113-118, 253-259)
    def inv_hess_prod(self, arr):
        self.update_hess_fact()
        a, vec = _as2d(arr)
        return _ret(self.hess_fact.solve(np.array(a, dtype=np.float64, order="F")), vec)

    def update_inv_hess(self):
        self.update_hess_fact()
        return self.hess_fact.solve(np.eye(self.dim))

    def check_numerics(self, gtol=EPS ** 0.25, Htol=None):
        """Cones.jl:273-290"""
        Htol = 10 * np.sqrt(gtol) if Htol is None else Htol
        g = self.grad()
        dim = g.size
        nu = self.nu
        if abs(1 + g @ self.point / nu) > gtol * dim:
            return False
        Hig = self.inv_hess_prod(g)
        if abs(1 - Hig @ g / nu) > Htol * dim:
            return False
        return True

    def get_proxsqr(self, irtmu, use_max_prox, negtol=np.sqrt(EPS)):
        """Cones.jl:294-310"""
        g = self.grad()
        vec1 = irtmu * self.dual_point + g
        vec2 = self.inv_hess_prod(vec1)
        prox_sqr = float(vec2 @ vec1)
        if prox_sqr < -negtol * g.size:
            return np.inf
        return abs(prox_sqr)


# ======================================================================================
class Nonnegative(Cone):
    """nonnegative.jl:42-145"""
    ctype = M.CONE_NONNEGATIVE

    @property
    def nu(self):
        return float(self.dim)

    def set_initial_point(self, arr):
        arr[:] = 1.0
        return arr

    def update_feas(self):
        return bool((self.point > EPS).all())

    def is_dual_feas(self):
        return bool((self.dual_point > EPS).all())

    def update_grad(self):
        self._grad[:] = -1.0 / self.point

    def update_hess(self):
        return np.diag(self.grad() ** 2)

    def update_inv_hess(self):
        return np.diag(self.point ** 2)

    def use_sqrt_hess_oracles(self, arr_dim):
        return True

    def hess_prod(self, arr):
        a, v = _as2d(arr)
        return _ret(a / self.point[:, None] / self.point[:, None], v)

    def inv_hess_prod(self, arr):
        a, v = _as2d(arr)
        return _ret(a * self.point[:, None] * self.point[:, None], v)

    def sqrt_hess_prod(self, arr):
        a, v = _as2d(arr)
        return _ret(a / self.point[:, None], v)

    def inv_sqrt_hess_prod(self, arr):
        a, v = _as2d(arr)
        return _ret(a * self.point[:, None], v)

    def dder3(self, direction):
        return (direction / self.point) ** 2 / self.point

    def get_proxsqr(self, irtmu, use_max_prox):
        vals = (self.point * self.dual_point * irtmu - 1.0) ** 2
        return float(vals.max() if use_max_prox else vals.sum())


# ======================================================================================
class EpiNormEucl(Cone):
    """epinormeucl.jl:44-228"""
    ctype = M.CONE_EPINORMEUCL
    nu = 2.0

    def set_initial_point(self, arr):
        arr[:] = 0.0
        arr[0] = np.sqrt(2.0)
        return arr

    def update_feas(self):
        u = self.point[0]
        if u > EPS:
            w = self.point[1:]
            self.dist = (u * u - float(w @ w)) / 2
            return self.dist > EPS
        return False

    def is_dual_feas(self):
        u = self.dual_point[0]
        if u > EPS:
            w = self.dual_point[1:]
            return (u * u - float(w @ w)) > 2 * EPS
        return False

    def update_grad(self):
        self._grad[:] = self.point / self.dist
        self._grad[0] *= -1

    def update_hess(self):
        g = self.grad()
        H = np.outer(g, g)
        inv_dist = 1.0 / self.dist
        H[np.diag_indices(self.dim)] += inv_dist
        H[0, 0] -= inv_dist + inv_dist
        return H

    def update_inv_hess(self):
        Hi = np.outer(self.point, self.point)
        Hi[np.diag_indices(self.dim)] += self.dist
        Hi[0, 0] -= self.dist + self.dist
        return Hi

    def use_sqrt_hess_oracles(self, arr_dim):
        return True

    def hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        u, w = self.point[0], self.point[1:]
        uj, wj = a[0], a[1:]
        ga = (w @ wj - u * uj) / self.dist
        prod = np.empty_like(a)
        prod[0] = -ga * u - uj
        prod[1:] = ga[None, :] * w[:, None] + wj
        prod /= self.dist
        return _ret(prod, v)

    def inv_hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        pa = self.point @ a
        prod = pa[None, :] * self.point[:, None]
        prod[0] -= self.dist * a[0]
        prod[1:] += self.dist * a[1:]
        return _ret(prod, v)

    def sqrt_hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        u, w = self.point[0], self.point[1:]
        distrt2 = self.dist * RT2
        rtdist = np.sqrt(self.dist)
        urtdist = u + rtdist * RT2
        uj, wj = a[0], a[1:]
        dotwwj = w @ wj
        prod = np.empty_like(a)
        prod[0] = (u * uj - dotwwj) / distrt2
        wmulj = (dotwwj / urtdist - uj) / distrt2
        prod[1:] = w[:, None] * wmulj[None, :] + wj / rtdist
        return _ret(prod, v)

    def inv_sqrt_hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        u, w = self.point[0], self.point[1:]
        rtdist = np.sqrt(self.dist)
        urtdist = u + rtdist * RT2
        uj, wj = a[0], a[1:]
        dotwwj = w @ wj
        prod = np.empty_like(a)
        prod[0] = (u * uj + dotwwj) / RT2
        wmulj = (dotwwj / urtdist + uj) / RT2
        prod[1:] = w[:, None] * wmulj[None, :] + wj * rtdist
        return _ret(prod, v)

    def dder3(self, direction):
        self.grad()
        point = self.point
        u, w = point[0], point[1:]
        u_dir, w_dir = direction[0], direction[1:]
        jdotpd = u * u_dir - float(w @ w_dir)
        d3 = self.hess_prod(direction).copy()
        dotdHd = -float(direction @ d3)
        dotpHd = float(point @ d3)
        d3 *= jdotpd
        d3[1:] += dotdHd * w + dotpHd * w_dir
        d3[0] += -dotdHd * u - dotpHd * u_dir
        d3 /= 2 * self.dist
        return d3


# ======================================================================================
class _ChoFact:
    """Upper Cholesky of a small symmetric matrix with the solves the matrix cones need."""

    def __init__(self, mat):
        self.fact = la.posdef_fact(mat)
        self.ok = self.fact.issuccess()
        self.Uf = self.fact.U if self.ok else None

    def logdet(self):
        return self.fact.logdet()

    def inverse(self):
        return self.fact.inverse()

    @staticmethod
    def _hs(mats):
        """(c, d, d) stack -> d x (c*d) matrix [M_1 M_2 ... M_c]."""
        c, d, _ = mats.shape
        return mats.transpose(1, 0, 2).reshape(d, c * d)

    @staticmethod
    def _unhs(mat, c):
        d = mat.shape[0]
        return mat.reshape(d, c, d).transpose(1, 0, 2)

    def _two_sided(self, mats, left):
        """left(left(M)') for every (symmetric-in) M of the stack: applies the same one-sided
        solve from the left, transposes, and applies it again -> L M L' for the operator L."""
        c = mats.shape[0]
        t = self._unhs(left(self._hs(mats)), c)                 # L M_i
        t = np.ascontiguousarray(np.swapaxes(t, -1, -2))        # M_i' L' = (L M_i)'
        return self._unhs(left(self._hs(t)), c)                 # L M_i' L'

    def congr_inv_sqrt(self, mats):
        """U^-T * M * U^-1 for a stack of symmetric matrices (rdiv! by U then ldiv! by U')."""
        U = self.Uf
        return self._two_sided(mats, lambda X: sla.solve_triangular(U, X, trans="T"))

    def congr_inv(self, mats):
        """S^-1 * M * S^-1 for a stack (rdiv! / ldiv! by the Cholesky factorisation)."""
        return self._two_sided(mats, lambda X: self.fact.solve(X))

    def congr_fwd_inv_sqrt_t(self, mats):
        """U^-1 * M * U^-T for a stack (rdiv! by U' then ldiv! by U)."""
        U = self.Uf
        return self._two_sided(mats, lambda X: sla.solve_triangular(U, X))


class PosSemidefTri(Cone):
    """possemideftri.jl:69-207 (real symmetric case)"""
    ctype = M.CONE_POSSEMIDEFTRI

    def __init__(self, dim):
        self.side = au.svec_side(dim)
        super().__init__(dim)

    @property
    def nu(self):
        return float(self.side)

    def set_initial_point(self, arr):
        arr[:] = au.smat_to_svec(np.eye(self.side))
        return arr

    def update_feas(self):
        self.mat = au.svec_to_smat(self.point)
        self.fact_mat = _ChoFact(self.mat)
        return self.fact_mat.ok

    def is_dual_feas(self):
        return _ChoFact(au.svec_to_smat(self.dual_point)).ok

    def update_grad(self):
        self.inv_mat = self.fact_mat.inverse()
        self._grad[:] = -au.smat_to_svec(self.inv_mat)

    def update_hess(self):
        return au.symm_kron(self.inv_mat)

    def update_inv_hess(self):
        assert self.is_feas()
        return au.symm_kron(self.mat)

    def use_sqrt_hess_oracles(self, arr_dim):
        return True

    def hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        return _ret(au.smats_to_svecs(self.fact_mat.congr_inv(au.svecs_to_smats(a))), v)

    def inv_hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        return _ret(au.smats_to_svecs(self.mat @ au.svecs_to_smats(a) @ self.mat), v)

    def sqrt_hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        return _ret(au.smats_to_svecs(self.fact_mat.congr_inv_sqrt(au.svecs_to_smats(a))), v)

    def inv_sqrt_hess_prod(self, arr):
        assert self.is_feas()
        a, v = _as2d(arr)
        U = self.fact_mat.Uf
        return _ret(au.smats_to_svecs(U @ au.svecs_to_smats(a) @ U.T), v)

    def dder3(self, direction):
        self.grad()
        S = au.svec_to_smat(direction)
        S = self.fact_mat.fact.solve(S)                                     # S^-1 * D
        S = sla.solve_triangular(self.fact_mat.Uf, S.T, trans="T").T        # ... * U^-1
        return au.smat_to_svec(S @ S.T)


# ======================================================================================
_CENTRAL_RAYS_HYPOPERLOG = np.array([
    [-0.827838387, 0.805102007, 1.290927686],
    [-0.689607388, 0.724605082, 1.224617936],
    [-0.584372665, 0.68128058, 1.182421942],
    [-0.503499342, 0.65448622, 1.153053152],
    [-0.440285893, 0.636444224, 1.131466926],
    [-0.389979809, 0.623569352, 1.114979519],
    [-0.349255921, 0.613978276, 1.102013921],
    [-0.315769104, 0.606589839, 1.091577908],
    [-0.287837744, 0.600745284, 1.083013],
    [-0.264242734, 0.596019009, 1.075868782],
])


def get_central_ray_hypoperlog(d):
    """Tabulated / fitted central ray of the hypoperlog barrier (data from
    hypoperlog.jl:289-319; the constants are numerical data of the reference)."""
    if d <= 10:
        return _CENTRAL_RAYS_HYPOPERLOG[d - 1]
    x = 1.0 / d
    if d <= 70:
        return np.array([4.657876 * x ** 2 - 3.116192 * x + 0.000647,
                         0.424682 * x + 0.553392, 0.760412 * x + 1.001795])
    return np.array([-3.011166 * x - 0.000122, 0.395308 * x + 0.553955, 0.837545 * x + 1.000024])


class HypoPerLogdetTri(Cone):
    """hypoperlogdettri.jl:80-368 (real symmetric case): (u, v, svec W), barrier
    -log(v*logdet(W/v) - u) - log(v) - logdet(W)."""
    ctype = M.CONE_HYPOPERLOGDETTRI

    def __init__(self, dim, use_dual=False):
        self.d = au.svec_side(dim - 2)
        self.use_dual_barrier = use_dual
        super().__init__(dim)

    @property
    def nu(self):
        return 2.0 + self.d

    def set_initial_point(self, arr):
        arr[:] = 0.0
        u, v, w = get_central_ray_hypoperlog(self.d)
        arr[0], arr[1] = u, v
        arr[2:] = au.smat_to_svec(w * np.eye(self.d))
        return arr

    def update_feas(self):
        v = self.point[1]
        if v > EPS:
            u = self.point[0]
            self.mat = au.svec_to_smat(self.point[2:])
            self.fact_W = _ChoFact(self.mat)
            if self.fact_W.ok:
                self.phi = self.fact_W.logdet() - self.d * np.log(v)
                self.zeta = v * self.phi - u
                return self.zeta > EPS
        return False

    def is_dual_feas(self):
        u = self.dual_point[0]
        if u < -EPS:
            v = self.dual_point[1]
            f = _ChoFact(au.svec_to_smat(self.dual_point[2:]))
            if f.ok:
                return (v - u * (f.logdet() + self.d * (1 - np.log(-u)))) > EPS
        return False

    def update_grad(self):
        v, zeta = self.point[1], self.zeta
        g = self._grad
        self.zetai = 1.0 / zeta
        g[0] = self.zetai
        g[1] = -1.0 / v - (self.phi - self.d) / zeta
        self.Wi = self.fact_W.inverse()
        self.Wi_vec = au.smat_to_svec(self.Wi)
        g[2:] = (-1 - v / zeta) * self.Wi_vec

    def update_hess(self):
        v, d, zeta, zetai = self.point[1], self.d, self.zeta, self.zetai
        sigma = self.phi - d
        Wi_vec = self.Wi_vec
        zis = sigma / zeta
        vzi = v / zeta
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zetai ** 2
        H[0, 1] = H[1, 0] = -zetai * zis
        H[1, 1] = v ** -2 + zis ** 2 + d / (v * zeta)
        H[0, 2:] = H[2:, 0] = (-vzi / zeta) * Wi_vec
        H[1, 2:] = H[2:, 1] = ((sigma * vzi - 1) / zeta) * Wi_vec
        Wivzi = vzi * Wi_vec
        H[2:, 2:] = (1 + vzi) * au.symm_kron(self.Wi) + np.outer(Wivzi, Wivzi)
        return H

    def hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        v, d, zeta = self.point[1], self.d, self.zeta
        sigma = self.phi - d
        vzi1 = v / zeta + 1
        p, q = a[0], a[1]
        w_aux = self.fact_W.congr_inv_sqrt(au.svecs_to_smats(a[2:]))   # U^-T R U^-1
        qzi = q / zeta
        c0 = np.trace(w_aux, axis1=1, axis2=2) / zeta
        c1 = (v * c0 - p / zeta + sigma * qzi) / zeta
        c3 = c1 * v - qzi
        prod = np.empty_like(a)
        prod[0] = -c1
        prod[1] = c1 * sigma - c0 + (qzi * d + q / v) / v
        w_aux = vzi1 * w_aux
        idx = np.arange(d)
        w_aux[:, idx, idx] += c3[:, None]
        w_aux = self.fact_W.congr_fwd_inv_sqrt_t(w_aux)                # U^-1 (.) U^-T
        prod[2:] = au.smats_to_svecs(w_aux)
        return _ret(prod, vec)

    def _ih_consts(self):
        v, d, zeta, phi = self.point[1], self.d, self.zeta, self.phi
        zv = zeta + v
        zzvi = zeta / zv
        c3 = v / (zv + d * v)
        c0 = phi - d * zzvi
        c4 = v * c3 * zv
        c6 = (v * phi) ** 2 + zeta * (zeta + d * v) - d * (zeta + v * phi) ** 2 * c3
        return v, d, zeta, phi, zv, zzvi, c3, c0, c4, c6

    def update_inv_hess(self):
        v, d, zeta, phi, zv, zzvi, c3, c0, c4, c6 = self._ih_consts()
        w = self.point[2:]
        c2 = v * c3
        c1 = v * zzvi + c0 * c2
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = c6
        Hi[0, 1] = Hi[1, 0] = c0 * c4
        Hi[1, 1] = c4
        Hi[0, 2:] = Hi[2:, 0] = c1 * w
        Hi[1, 2:] = Hi[2:, 1] = c2 * w
        Hi[2:, 2:] = zzvi * au.symm_kron(self.mat) + (c2 / zv) * np.outer(w, w)
        return Hi

    def inv_hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        v, d, zeta, phi, zv, zzvi, c3, c0, c4, c6 = self._ih_consts()
        w = self.point[2:]
        W = self.mat
        c7 = c4 * c0
        c8 = c7 + v * zeta
        p, q, r = a[0], a[1], a[2:]
        c1 = (w @ r) / zv
        c5 = c0 * p + q + c1
        c2 = v * (zzvi * p + c3 * c5)
        prod = np.empty_like(a)
        prod[0] = c6 * p + c7 * q + c8 * c1
        prod[1] = c4 * c5
        WRW = W @ au.svecs_to_smats(r) @ W
        prod[2:] = zzvi * au.smats_to_svecs(WRW) + w[:, None] * c2[None, :]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        v, d, zeta = self.point[1], self.d, self.zeta
        p, q, r = direction[0], direction[1], direction[2:]
        sigma = self.phi - d
        viq = q / v
        viq2 = viq ** 2
        vzi = v / zeta
        vzi1 = vzi + 1
        rwi = self.fact_W.congr_inv_sqrt(au.svec_to_smat(r)[None])[0]
        c0 = np.trace(rwi)
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
        w_aux2 = c6 * np.eye(d) + vzi1 * rwi
        w_aux = rwi @ w_aux2
        w_aux[np.diag_indices(d)] += c8
        w_aux = self.fact_W.congr_fwd_inv_sqrt_t(w_aux[None])[0]
        d3[2:] = au.smat_to_svec(w_aux)
        return d3


# ======================================================================================
class HypoRootdetTri(Cone):
    """hyporootdettri.jl:82-324 (real symmetric case): (u, svec W), barrier
    -log(det(W)^(1/d) - u) - logdet(W)."""
    ctype = M.CONE_HYPOROOTDETTRI

    def __init__(self, dim, use_dual=False):
        self.d = au.svec_side(dim - 1)
        self.di = 1.0 / self.d
        self.use_dual_barrier = use_dual
        super().__init__(dim)

    @property
    def nu(self):
        return 1.0 + self.d

    def set_initial_point(self, arr):
        d = self.d
        arr[:] = 0.0
        c1 = np.sqrt(5.0 * d * d + 2 * d + 1)
        c2 = arr[0] = -np.sqrt((3 * d + 1 - c1) / (2.0 * d + 2))
        c3 = -c2 * (d + 1 + c1) / (2.0 * d)
        arr[1:] = au.smat_to_svec(c3 * np.eye(d))
        return arr

    def update_feas(self):
        self.mat = au.svec_to_smat(self.point[1:])
        self.fact_W = _ChoFact(self.mat)
        if self.fact_W.ok:
            self.phi = np.exp(self.fact_W.logdet() / self.d)
            self.zeta = self.phi - self.point[0]
            return self.zeta > EPS
        return False

    def is_dual_feas(self):
        u = self.dual_point[0]
        if u < -EPS:
            f = _ChoFact(au.svec_to_smat(self.dual_point[1:]))
            if f.ok:
                return (f.logdet() - self.d * np.log(-u / self.d)) > EPS
        return False

    def update_grad(self):
        zeta = self.zeta
        self.pzd = self.phi / zeta * self.di          # phi / zeta / d
        g = self._grad
        g[0] = 1.0 / zeta
        self.Wi = self.fact_W.inverse()
        self.Wi_vec = au.smat_to_svec(self.Wi)
        g[1:] = (-self.pzd - 1) * self.Wi_vec

    def update_hess(self):
        zeta, pzd, Wi_vec = self.zeta, self.pzd, self.Wi_vec
        H = np.zeros((self.dim, self.dim))
        H[0, 0] = zeta ** -2
        H[0, 1:] = H[1:, 0] = (-pzd / zeta) * Wi_vec
        c2 = pzd * (pzd - self.di)
        H[1:, 1:] = (pzd + 1) * au.symm_kron(self.Wi) + c2 * np.outer(Wi_vec, Wi_vec)
        return H

    def hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        di, zeta, pzd, d = self.di, self.zeta, self.pzd, self.d
        p = a[0]
        w_aux = self.fact_W.congr_inv_sqrt(au.svecs_to_smats(a[1:]))
        c0 = pzd * np.trace(w_aux, axis1=1, axis2=2)
        c1 = c0 - p / zeta
        c2 = pzd * c1 - di * c0
        w_aux = (pzd + 1) * w_aux
        idx = np.arange(d)
        w_aux[:, idx, idx] += c2[:, None]
        w_aux = self.fact_W.congr_fwd_inv_sqrt_t(w_aux)
        prod = np.empty_like(a)
        prod[0] = c1 / -zeta
        prod[1:] = au.smats_to_svecs(w_aux)
        return _ret(prod, vec)

    def update_inv_hess(self):
        w = self.point[1:]
        zeta, phi, di = self.zeta, self.phi, self.di
        phidi = phi * di
        c2 = 1.0 / (self.pzd + 1)
        c3 = phidi * c2 / zeta * di
        Hi = np.zeros((self.dim, self.dim))
        Hi[0, 0] = zeta ** 2 + phidi * phi
        Hi[0, 1:] = Hi[1:, 0] = phidi * w
        Hi[1:, 1:] = c2 * au.symm_kron(self.mat) + c3 * np.outer(w, w)
        return Hi

    def inv_hess_prod(self, arr):
        self.grad()
        a, vec = _as2d(arr)
        w = self.point[1:]
        W = self.mat
        zeta, phi, di = self.zeta, self.phi, self.di
        phidi = phi * di
        c2 = 1.0 / (self.pzd + 1)
        c3 = c2 / zeta * di
        c4 = zeta ** 2 + phidi * phi
        p, r = a[0], a[1:]
        c5 = w @ r
        c6 = phidi * (c3 * c5 + p)
        prod = np.empty_like(a)
        prod[0] = phidi * c5 + c4 * p
        WRW = W @ au.svecs_to_smats(r) @ W
        prod[1:] = c2 * au.smats_to_svecs(WRW) + w[:, None] * c6[None, :]
        return _ret(prod, vec)

    def dder3(self, direction):
        self.grad()
        p, r = direction[0], direction[1:]
        zeta, phi, di, pzd, d = self.zeta, self.phi, self.di, self.pzd, self.d
        rwi = self.fact_W.congr_inv_sqrt(au.svec_to_smat(r)[None])[0]
        c0 = np.trace(rwi) * di
        c6 = float((rwi ** 2).sum()) * di
        zichi = (p - phi * c0) / zeta
        c1 = zichi ** 2 + phi / zeta * (c6 - c0 ** 2) / 2
        c7 = pzd * (c1 - c6 / 2 + c0 * (zichi + c0 / 2))
        c8 = -pzd * (zichi + c0)
        c9 = pzd + 1
        d3 = np.empty(self.dim)
        d3[0] = c1 / -zeta
        w_aux2 = c8 * np.eye(d) + c9 * rwi
        w_aux = rwi @ w_aux2
        w_aux[np.diag_indices(d)] += c7
        w_aux = self.fact_W.congr_fwd_inv_sqrt_t(w_aux[None])[0]
        d3[1:] = au.smat_to_svec(w_aux)
        return d3


# ======================================================================================
_CLASSES = {
    M.CONE_NONNEGATIVE: Nonnegative,
    M.CONE_EPINORMEUCL: EpiNormEucl,
    M.CONE_POSSEMIDEFTRI: PosSemidefTri,
    M.CONE_HYPOPERLOGDETTRI: HypoPerLogdetTri,
    M.CONE_HYPOROOTDETTRI: HypoRootdetTri,
}


def make_cone(spec):
    if spec.ctype == M.CONE_EPIPERSEPSPECTRAL_MAT:
        from .cones_sepspec import EpiPerSepSpectralMat
        return EpiPerSepSpectralMat(spec.dim, spec.hkind, spec.hparam, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_EPIPERSEPSPECTRAL_VEC:
        from .cones_sepspec import EpiPerSepSpectralVec
        return EpiPerSepSpectralVec(spec.dim, spec.hkind, spec.hparam, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_WSOSINTERPEPINORMONE:
        from .cones_vec3 import WSOSInterpEpiNormOne
        return WSOSInterpEpiNormOne(spec.hkind, spec.dim // spec.hkind, M.wsos_unpack(spec), use_dual=not spec.use_dual)
    if spec.ctype == M.CONE_WSOSINTERPEPINORMEUCL:
        from .cones_vec3 import WSOSInterpEpiNormEucl
        return WSOSInterpEpiNormEucl(spec.hkind, spec.dim // spec.hkind, M.wsos_unpack(spec), use_dual=not spec.use_dual)
    if spec.ctype == M.CONE_WSOSINTERPPOSSEMIDEFTRI:
        from .cones_vec3 import WSOSInterpPosSemidefTri
        Rr = spec.hkind
        return WSOSInterpPosSemidefTri(Rr, spec.dim // (Rr * (Rr + 1) // 2), M.wsos_unpack(spec), use_dual=not spec.use_dual)
    if spec.ctype == M.CONE_EPITRRELENTROPYTRI:
        from .cones_vec3 import EpiTrRelEntropyTri
        return EpiTrRelEntropyTri(spec.dim, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_POSSEMIDEFTRISPARSE:
        from .cones_vec3 import PosSemidefTriSparse
        a = np.asarray(spec.alpha)
        return PosSemidefTriSparse(int(a[0]), a[1:1 + spec.dim].astype(int), a[1 + spec.dim:].astype(int),
                                   use_dual=spec.use_dual)
    if spec.ctype == M.CONE_MATRIXEPIPERSQUARE:
        from .cones_vec3 import MatrixEpiPerSquare
        d1 = spec.hkind
        return MatrixEpiPerSquare(d1, (spec.dim - d1 * (d1 + 1) // 2 - 1) // d1, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_DOUBLYNONNEGATIVETRI:
        from .cones_vec3 import DoublyNonnegativeTri
        return DoublyNonnegativeTri(spec.dim, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_LINMATRIXINEQ:
        from .cones_vec3 import LinMatrixIneq
        return LinMatrixIneq(M.lmi_unpack(spec), use_dual=spec.use_dual)
    if spec.ctype == M.CONE_WSOSINTERPNONNEGATIVE:
        from .cones_vec3 import WSOSInterpNonnegative
        return WSOSInterpNonnegative(spec.dim, M.wsos_unpack(spec), use_dual=not spec.use_dual)
    if spec.ctype == M.CONE_EPINORMSPECTRAL:
        from .cones_vec3 import EpiNormSpectral
        return EpiNormSpectral(spec.hkind, (spec.dim - 1) // spec.hkind, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_EPIRELENTROPY:
        from .cones_vec3 import EpiRelEntropy
        return EpiRelEntropy(spec.dim, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_HYPOPOWERMEAN:
        from .cones_vec3 import HypoPowerMean
        return HypoPowerMean(spec.alpha, use_dual=spec.use_dual)
    if spec.ctype == M.CONE_GENERALIZEDPOWER:
        from .cones_vec3 import GeneralizedPower
        return GeneralizedPower(spec.alpha, spec.dim - len(spec.alpha), use_dual=spec.use_dual)
    if spec.ctype in (M.CONE_EPIPERSQUARE, M.CONE_HYPOPERLOG, M.CONE_EPINORMINF, M.CONE_HYPOGEOMEAN):
        from . import cones_vec3
        cls = {M.CONE_EPIPERSQUARE: cones_vec3.EpiPerSquare, M.CONE_HYPOPERLOG: cones_vec3.HypoPerLog,
               M.CONE_EPINORMINF: cones_vec3.EpiNormInf, M.CONE_HYPOGEOMEAN: cones_vec3.HypoGeoMean}[spec.ctype]
        return cls(spec.dim, use_dual=spec.use_dual)
    cls = _CLASSES[spec.ctype]
    if spec.ctype in (M.CONE_HYPOPERLOGDETTRI, M.CONE_HYPOROOTDETTRI):
        return cls(spec.dim, use_dual=spec.use_dual)
    return cls(spec.dim)


class OracleConeBlock(ConeBlock):
    """ConeBlock over a list of per-cone CPU oracles (loops `for k in cones` like the reference)."""

    def __init__(self, model):
        super().__init__(model)
        self.cones = [make_cone(s) for s in self.specs]
        self.slices = list(model.cone_idxs)
        self.point = np.zeros(self.q)
        self.dual_point = np.zeros(self.q)

    def load_point(self, primal, dual, scal=1.0):
        self.point[:] = scal * np.asarray(primal)
        self.dual_point[:] = dual
        for ck, sl in zip(self.cones, self.slices):
            ck.load_point(self.point[sl])
            ck.load_dual_point(self.dual_point[sl])
            ck.reset_data()

    def _map(self, fn, arr):
        a = np.asarray(arr, dtype=np.float64)
        out = np.empty_like(a)
        for ck, sl in zip(self.cones, self.slices):
            out[sl] = fn(ck, a[sl])
        return out

    def is_feas(self):
        return np.array([ck.is_feas() for ck in self.cones], dtype=bool)

    def is_dual_feas(self):
        return np.array([ck.is_dual_feas() for ck in self.cones], dtype=bool)

    def grad(self):
        out = np.empty(self.q)
        for ck, sl in zip(self.cones, self.slices):
            out[sl] = ck.grad()
        return out

    def hess_prod(self, arr):
        return self._map(lambda ck, a: ck.hess_prod(a), arr)

    def inv_hess_prod(self, arr):
        return self._map(lambda ck, a: ck.inv_hess_prod(a), arr)

    def block_hess_prod(self, arr):
        return self._map(lambda ck, a: ck.inv_hess_prod(a) if ck.use_dual_barrier
                         else ck.hess_prod(a), arr)

    def sqrt_hess_prod(self, arr):
        return self._map(lambda ck, a: ck.sqrt_hess_prod(a), arr)

    def inv_sqrt_hess_prod(self, arr):
        return self._map(lambda ck, a: ck.inv_sqrt_hess_prod(a), arr)

    def use_dder3(self):
        return np.array([ck.use_dder3() for ck in self.cones], dtype=bool)

    def dder3(self, direction):
        return self._map(lambda ck, a: ck.dder3(a), direction)

    def check_numerics(self, irtmu=None, use_max_prox=None):
        return np.array([ck.check_numerics() for ck in self.cones], dtype=bool)

    def get_proxsqr(self, irtmu, use_max_prox):
        return np.array([ck.get_proxsqr(irtmu, use_max_prox) for ck in self.cones])

    def initial_point(self):
        out = np.zeros(self.q)
        for ck, sl in zip(self.cones, self.slices):
            ck.set_initial_point(out[sl])
        return out

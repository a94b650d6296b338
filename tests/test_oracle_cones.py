"""Pins the CPU oracle's cone restatements with the reference's own implementation-independent
oracle identities (reference: test/cone.jl:23-114, sizes from :321-325, :336-340, :463-467,
:648-655 and the HypoRootdetTri block), at the reference tolerance tol = 1e3*eps."""
import numpy as np
import pytest

from oracle import cones as oc

EPS = np.finfo(np.float64).eps
TOL = 1e3 * EPS


def perturb_scale(rng, point, noise, scale):
    # reference: test/cone.jl:236-248
    if noise:
        point += 2 * noise * rng.random(point.size) - noise
    if scale != 1:
        point *= scale
    return point


def close(a, b, tol=TOL):
    # Julia's isapprox(a, b, atol, rtol): norm(a-b) <= max(atol, rtol*max(norm a, norm b))
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) <= max(tol, tol * max(np.linalg.norm(a), np.linalg.norm(b)))


def run_oracles(cone, init_tol=TOL, init_only=False, noise=0.1, scale=0.1, tol=TOL):
    rng = np.random.default_rng(1)
    dim = cone.dim
    cone.setup_data()
    point = np.zeros(dim)
    cone.set_initial_point(point)
    cone.load_point(point)
    assert cone.is_feas()
    dual_point = -cone.grad().copy()
    cone.load_dual_point(dual_point)
    assert cone.is_dual_feas()
    assert cone.get_proxsqr(1.0, True) <= 1
    assert cone.get_proxsqr(1.0, False) <= dim
    assert close(cone.hess_prod(point), dual_point, tol)
    if np.isfinite(init_tol):
        assert close(point, dual_point, init_tol)
    if init_only:
        return

    perturb_scale(rng, point, noise, scale)
    perturb_scale(rng, dual_point, noise, 1 / scale)
    cone.reset_data()
    cone.load_point(point)
    assert cone.is_feas()
    cone.load_dual_point(dual_point)
    assert cone.is_dual_feas()

    nu = cone.nu
    grad = cone.grad().copy()
    assert abs(point @ grad + nu) <= tol * max(1, nu)
    hess = np.array(cone.hess())
    inv_hess = np.array(cone.inv_hess())
    I = np.eye(dim)
    assert close(hess @ inv_hess, I, tol)
    assert close(hess @ point, -grad, tol)
    assert close(cone.hess_prod(point), -grad, tol)
    assert close(cone.inv_hess_prod(grad), -point, tol)
    assert close(cone.hess_prod(inv_hess), I, tol)
    assert close(cone.inv_hess_prod(hess), I, tol)
    psi = dual_point + grad
    proxsqr = psi @ cone.inv_hess_prod(psi)
    assert abs(cone.get_proxsqr(1.0, False) - proxsqr) <= tol * max(1, abs(proxsqr))

    if cone.use_sqrt_hess_oracles(dim + 1):
        prod_mat2 = cone.sqrt_hess_prod(inv_hess).T.copy()
        assert close(cone.sqrt_hess_prod(prod_mat2), I, tol)
        pm = cone.inv_sqrt_hess_prod(I)
        assert close(pm.T @ pm, inv_hess, tol)

    if cone.use_dder3():
        assert close(-cone.dder3(point), grad, tol)
        direction = perturb_scale(rng, np.zeros(dim), noise, 1.0)
        d3 = cone.dder3(direction)
        ref = direction @ hess @ direction
        assert abs(d3 @ point - ref) <= tol * max(1, abs(ref))


@pytest.mark.parametrize("dim", [1, 2, 6])
def test_nonnegative(dim):
    run_oracles(oc.Nonnegative(dim))


@pytest.mark.parametrize("side", [1, 2, 3, 5])
def test_possemideftri(side):
    run_oracles(oc.PosSemidefTri(side * (side + 1) // 2))


@pytest.mark.parametrize("dim", [2, 3, 4, 6, 25])
def test_epinormeucl(dim):
    run_oracles(oc.EpiNormEucl(dim))


@pytest.mark.parametrize("side", [1, 2, 4])
def test_hypoperlogdettri(side):
    # reference: test/cone.jl:648-655 (init_tol = 1e-4 for this cone)
    run_oracles(oc.HypoPerLogdetTri(2 + side * (side + 1) // 2), init_tol=1e-4)


@pytest.mark.parametrize("side", [8, 12])
def test_hypoperlogdettri_init_only(side):
    run_oracles(oc.HypoPerLogdetTri(2 + side * (side + 1) // 2), init_tol=1e-1, init_only=True)


@pytest.mark.parametrize("side", [1, 2, 4, 5])
def test_hyporootdettri(side):
    run_oracles(oc.HypoRootdetTri(1 + side * (side + 1) // 2))


@pytest.mark.parametrize("side", [3, 6])
def test_generic_sqrt_oracles_logdet(side):
    """Cones without closed-form sqrt oracles use the Cholesky of the explicit Hessian
    (reference: Cones.jl:189-218)."""
    cone = oc.HypoPerLogdetTri(2 + side * (side + 1) // 2)
    rng = np.random.default_rng(3)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.05, 0.5)
    cone.load_point(point)
    assert cone.is_feas()
    assert not cone.use_sqrt_hess_oracles(cone.dim - 1)   # array too small
    assert cone.use_sqrt_hess_oracles(cone.dim)
    A = rng.standard_normal((cone.dim, 4))
    RA = cone.sqrt_hess_prod(A)
    assert close(RA.T @ RA, A.T @ cone.hess_prod(A), 1e-12)


@pytest.mark.parametrize("dim", [3, 4, 6, 25])
def test_epipersquare(dim):
    # reference: test/cone.jl EpiPerSquare block (dims 3, 4, 6)
    from oracle.cones_vec3 import EpiPerSquare
    run_oracles(EpiPerSquare(dim))


@pytest.mark.parametrize("dw", [1, 2, 5])
def test_hypoperlog(dw):
    # reference: test/cone.jl:627-630 (init_tol = 1e-5)
    from oracle.cones_vec3 import HypoPerLog
    run_oracles(HypoPerLog(2 + dw), init_tol=1e-5)


@pytest.mark.parametrize("dw", [15, 40, 100])
def test_hypoperlog_init_only(dw):
    # reference: test/cone.jl:631-633
    from oracle.cones_vec3 import HypoPerLog
    run_oracles(HypoPerLog(2 + dw), init_tol=1e-1, init_only=True)


@pytest.mark.parametrize("dw", [1, 2, 5])
def test_epinorminf(dw):
    # reference: test/cone.jl:443-447
    from oracle.cones_vec3 import EpiNormInf
    run_oracles(EpiNormInf(1 + dw))


@pytest.mark.parametrize("dw", [1, 2, 5])
def test_hypogeomean(dw):
    # reference: test/cone.jl:587-591
    from oracle.cones_vec3 import HypoGeoMean
    run_oracles(HypoGeoMean(1 + dw))


@pytest.mark.parametrize("du,dw", [(2, 1), (3, 2), (4, 1), (2, 4)])
def test_generalizedpower(du, dw):
    # reference: test/cone.jl:541-545 (random powers); inv_hess / inv_hess_prod / sqrt oracles are the generic
    # Hessian-factorisation fallbacks of Cones.jl:113-118, 189-259
    from oracle.cones_vec3 import GeneralizedPower
    a = np.random.default_rng(du * 10 + dw).random(du) + 1e-3
    run_oracles(GeneralizedPower(a / a.sum(), dw))


@pytest.mark.parametrize("dw,init_only", [(1, False), (2, False), (5, False), (15, True), (40, True), (100, True)])
def test_hypopowermean(dw, init_only):
    # reference: test/cone.jl:563-570 (powers rand + 1 normalised, init_tol 1e-2 / 1e-1)
    from oracle.cones_vec3 import HypoPowerMean
    a = np.random.default_rng(dw).random(dw) + 1
    run_oracles(HypoPowerMean(a / a.sum()), init_tol=1e-1 if init_only else 1e-2, init_only=init_only)


@pytest.mark.parametrize("dw,init_only", [(1, False), (2, False), (4, False), (15, True), (40, True), (100, True)])
def test_epirelentropy(dw, init_only):
    # reference: test/cone.jl:707-715
    from oracle.cones_vec3 import EpiRelEntropy
    run_oracles(EpiRelEntropy(1 + 2 * dw), init_tol=1e-1 if init_only else 1e-5, init_only=init_only)


@pytest.mark.parametrize("dr,ds", [(1, 1), (1, 2), (2, 2), (2, 4), (3, 4)])
def test_epinormspectral(dr, ds):
    # reference: test/cone.jl:499-503
    from oracle.cones_vec3 import EpiNormSpectral
    run_oracles(EpiNormSpectral(dr, ds))


def test_epinormspectral_barrier():
    """test_barrier of test/cone.jl:505-513 with central differences: grad, hess_prod and dder3 against derivatives of
    -logdet(u^2 I - W W') + (d1 - 1) log u."""
    from oracle.cones_vec3 import EpiNormSpectral
    dr, ds = 2, 3
    cone = EpiNormSpectral(dr, ds)

    def barrier(s):
        W = s[1:].reshape(dr, ds, order="F")
        return -np.linalg.slogdet(s[0] ** 2 * np.eye(dr) - W @ W.T)[1] + (dr - 1) * np.log(s[0])

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-7)
    assert close(cone.hess() @ direction, fd_hess_dir, 1e-7)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


@pytest.mark.parametrize("n,halfdeg", [(1, 1), (1, 3), (2, 2), (3, 1)])
def test_wsosinterpnonnegative(n, halfdeg):
    # reference: test/cone.jl WSOSInterpNonnegative block (interpolations of free / box domains, init_tol = Inf)
    from oracle.cones_vec3 import WSOSInterpNonnegative
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    run_oracles(WSOSInterpNonnegative(U, Ps), init_tol=np.inf)
    run_oracles(WSOSInterpNonnegative(U, Ps, use_dual=True), init_tol=np.inf)


def test_wsosinterpnonnegative_barrier():
    """grad, hess_prod and dder3 against central differences of -sum_k logdet(P_k' Diagonal(s) P_k)
    (test/cone.jl WSOSInterpNonnegative test_barrier)."""
    from oracle.cones_vec3 import WSOSInterpNonnegative
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box([-1.0, 0.0], [1.0, 2.0], 2)
    cone = WSOSInterpNonnegative(U, Ps)

    def barrier(s):
        return -sum(np.linalg.slogdet(P.T @ (s[:, None] * P))[1] for P in Ps)

    rng = np.random.default_rng(1)
    point = np.ones(U)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(U)])
    assert close(g, fd_grad, 1e-7)
    direction = rng.standard_normal(U)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


@pytest.mark.parametrize("side,init_only", [(1, False), (2, False), (5, False), (10, True), (20, True)])
def test_doublynonnegativetri(side, init_only):
    # reference: test/cone.jl:353-361 (init_tol = sqrt(eps))
    from oracle.cones_vec3 import DoublyNonnegativeTri
    run_oracles(DoublyNonnegativeTri(side * (side + 1) // 2), init_tol=np.sqrt(EPS), init_only=init_only)


def test_doublynonnegativetri_barrier():
    """test/cone.jl:363-371 with central differences: -logdet(W) - sum log of the off-diagonal svec entries."""
    from oracle import arrayutil as au
    from oracle.cones_vec3 import DoublyNonnegativeTri
    side = 3
    cone = DoublyNonnegativeTri(6)
    off = cone.offdiag

    def barrier(s):
        return -np.linalg.slogdet(au.svec_to_smat(s))[1] - np.sum(np.log(s[off]))

    rng = np.random.default_rng(1)
    point = np.zeros(6)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(6)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.3 * rng.standard_normal(6)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    assert close(cone.hess() @ direction, fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


@pytest.mark.parametrize("dr,ds", [(1, 1), (1, 2), (2, 2), (2, 4), (3, 4)])
def test_matrixepipersquare(dr, ds):
    # reference: test/cone.jl:519-523
    from oracle.cones_vec3 import MatrixEpiPerSquare
    run_oracles(MatrixEpiPerSquare(dr, ds))


def test_matrixepipersquare_barrier():
    """test/cone.jl:525-535 with central differences: -logdet(2 v U - W W') + (d1 - 1) log v."""
    from oracle import arrayutil as au
    from oracle.cones_vec3 import MatrixEpiPerSquare
    dr, ds = 2, 2
    cone = MatrixEpiPerSquare(dr, ds)
    du = dr * (dr + 1) // 2

    def barrier(s):
        U, v, W = au.svec_to_smat(s[:du]), s[du], s[du + 1:].reshape(dr, ds, order="F")
        return -np.linalg.slogdet(2 * v * U - W @ W.T)[1] + (dr - 1) * np.log(v)

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.3 * rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


@pytest.mark.parametrize("R,n,halfdeg", [(1, 1, 1), (2, 1, 2), (3, 1, 1), (2, 2, 1), (3, 2, 1)])
def test_wsosinterppossemideftri(R, n, halfdeg):
    # reference: test/cone.jl WSOSInterpPosSemidefTri block (init_tol = Inf)
    from oracle.cones_vec3 import WSOSInterpPosSemidefTri
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    run_oracles(WSOSInterpPosSemidefTri(R, U, Ps), init_tol=np.inf)
    run_oracles(WSOSInterpPosSemidefTri(R, U, Ps, use_dual=True), init_tol=np.inf)


def test_wsosinterppossemideftri_barrier():
    """grad, hess_prod and dder3 against central differences of -sum_k logdet((I kron P_k)' D(s) (I kron P_k))."""
    from oracle.cones_vec3 import WSOSInterpPosSemidefTri
    from wsos_util import interpolate_box
    R = 2
    U, _, Ps = interpolate_box([-1.0], [1.0], 2)
    cone = WSOSInterpPosSemidefTri(R, U, Ps)

    def barrier(s):
        D = cone._D(s)
        return -sum(np.linalg.slogdet(np.kron(np.eye(R), P).T @ D @ np.kron(np.eye(R), P))[1] for P in Ps)

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.3 * rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


@pytest.mark.parametrize("R,n,halfdeg", [(2, 1, 1), (2, 1, 2), (3, 1, 2), (3, 2, 1), (4, 2, 1)])
def test_wsosinterpepinormeucl(R, n, halfdeg):
    # reference: test/cone.jl WSOSInterpEpiNormEucl block (init_tol = Inf)
    from oracle.cones_vec3 import WSOSInterpEpiNormEucl
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    run_oracles(WSOSInterpEpiNormEucl(R, U, Ps), init_tol=np.inf)
    run_oracles(WSOSInterpEpiNormEucl(R, U, Ps, use_dual=True), init_tol=np.inf)


def test_wsosinterpepinormeucl_barrier():
    """grad, hess_prod and dder3 of the dense (arrow-matrix) restatement against central differences of the REFERENCE's
    barrier -sum_k [logdet L11 + logdet(L11 - sum_r L1r L11^-1 L1r)] (wsosinterpepinormeucl.jl:119-167)."""
    from oracle.cones_vec3 import WSOSInterpEpiNormEucl
    from wsos_util import interpolate_box
    R = 3
    U, _, Ps = interpolate_box([-1.0], [1.0], 2)
    cone = WSOSInterpEpiNormEucl(R, U, Ps)

    def barrier(s):
        tot = 0.0
        for P in Ps:
            L11 = P.T @ (s[:U, None] * P)
            mat = L11.copy()
            for r in range(1, R):
                L1r = P.T @ (s[r * U:(r + 1) * U, None] * P)
                mat -= L1r @ np.linalg.solve(L11, L1r)
            tot -= np.linalg.slogdet(L11)[1] + np.linalg.slogdet(mat)[1]
        return tot

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.3 * rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


@pytest.mark.parametrize("R,n,halfdeg", [(2, 1, 1), (2, 1, 2), (3, 1, 2), (3, 2, 1), (4, 2, 1)])
def test_wsosinterpepinormone(R, n, halfdeg):
    # reference: test/cone.jl WSOSInterpEpiNormOne block (init_tol = Inf)
    from oracle.cones_vec3 import WSOSInterpEpiNormOne
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    run_oracles(WSOSInterpEpiNormOne(R, U, Ps), init_tol=np.inf)
    run_oracles(WSOSInterpEpiNormOne(R, U, Ps, use_dual=True), init_tol=np.inf)


def test_wsosinterpepinormone_barrier():
    """grad, hess_prod and dder3 of the pairwise (arrow-matrix) restatement against central differences of the REFERENCE's
    barrier -sum_k [logdet L11 + sum_r logdet(L11 - L1r L11^-1 L1r)] (wsosinterpepinormone.jl:147-204)."""
    from oracle.cones_vec3 import WSOSInterpEpiNormOne
    from wsos_util import interpolate_box
    R = 4
    U, _, Ps = interpolate_box([-1.0], [1.0], 2)
    cone = WSOSInterpEpiNormOne(R, U, Ps)

    def barrier(s):
        tot = 0.0
        for P in Ps:
            L11 = P.T @ (s[:U, None] * P)
            tot -= np.linalg.slogdet(L11)[1]
            for r in range(1, R):
                L1r = P.T @ (s[r * U:(r + 1) * U, None] * P)
                tot -= np.linalg.slogdet(L11 - L1r @ np.linalg.solve(L11, L1r))[1]
        return tot

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.3 * rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


def rand_sppsd_pattern(rng, side):
    """rand_sppsd_pattern of test/cone.jl: a random sparse lower-triangular pattern that contains the diagonal."""
    mask = np.tril(rng.random((side, side)) < 1 / np.sqrt(side)) | np.eye(side, dtype=bool)
    rows, cols = np.nonzero(mask)
    return rows, cols


@pytest.mark.parametrize("side", [1, 2, 10, 25, 40])
def test_possemideftrisparse(side):
    # reference: test/cone.jl:377-381
    from oracle.cones_vec3 import PosSemidefTriSparse
    rows, cols = rand_sppsd_pattern(np.random.default_rng(side), side)
    run_oracles(PosSemidefTriSparse(side, rows, cols))


@pytest.mark.parametrize("dW,init_only", [(1, False), (2, False), (4, False), (6, True), (10, True)])
def test_epitrrelentropytri(dW, init_only):
    # reference: test/cone.jl:731-739 (init_tol 1e-4 / 1e-1).  Oracle restatement only: no device kernels for this cone yet
    from oracle.cones_vec3 import EpiTrRelEntropyTri
    run_oracles(EpiTrRelEntropyTri(1 + 2 * (dW * (dW + 1) // 2)), init_tol=1e-1 if init_only else 1e-4,
                init_only=init_only)


def test_epitrrelentropytri_barrier():
    """grad, hess_prod and dder3 against central differences of
    -log(u - tr(W log W - W log V)) - logdet V - logdet W (test/cone.jl:741-752)."""
    from oracle import arrayutil as au
    from oracle.cones_vec3 import EpiTrRelEntropyTri
    cone = EpiTrRelEntropyTri(1 + 2 * 6)

    def mlog(X):
        lam, Q = np.linalg.eigh(X)
        return (Q * np.log(lam)) @ Q.T, lam

    def barrier(s):
        V, W = au.svec_to_smat(s[1:7]), au.svec_to_smat(s[7:13])
        (lV, lamV), (lW, lamW) = mlog(V), mlog(W)
        return -np.log(s[0] - np.sum(W * (lW - lV))) - np.sum(np.log(lamV)) - np.sum(np.log(lamW))

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.3 * rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


def rand_lmi(rng, side, dim):
    """rand_herms of test/cone.jl (real case): symmetric matrices with a positive definite first one."""
    As = []
    for i in range(dim):
        X = rng.random((side, side))
        As.append(X @ X.T + np.eye(side) if i == 0 else (X + X.T) / 2 - 0.5)
    return As


@pytest.mark.parametrize("side,dim", [(2, 2), (3, 2), (4, 3), (3, 6)])
def test_linmatrixineq(side, dim):
    # reference: test/cone.jl:423-429 (noise 1e-2, init_tol = Inf)
    from oracle.cones_vec3 import LinMatrixIneq
    run_oracles(LinMatrixIneq(rand_lmi(np.random.default_rng(1), side, dim)), init_tol=np.inf, noise=1e-2)


def test_linmatrixineq_barrier():
    """test/cone.jl:431-436 with central differences: -logdet(sum_i s_i A_i)."""
    from oracle.cones_vec3 import LinMatrixIneq
    As = rand_lmi(np.random.default_rng(1), 3, 3)
    cone = LinMatrixIneq(As)

    def barrier(s):
        return -np.linalg.slogdet(sum(w * A for w, A in zip(s, As)))[1]

    rng = np.random.default_rng(2)
    point = np.zeros(3)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.01, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps) for e in np.eye(3)])
    assert close(g, fd_grad, 1e-7)
    direction = 0.1 * rng.standard_normal(3)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-6)
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


SSF = [(0, 0.0), (1, 0.0), (2, 0.0), (3, 1.5), (3, 2.0), (3, 1.1)]   # Inv, NegLog, NegEntropy, Power12(p)


@pytest.mark.parametrize("side", [1, 2, 3, 6])
@pytest.mark.parametrize("hkind,hparam", SSF)
def test_epipersepspectral_matrix(side, hkind, hparam):
    # reference: test/cone.jl:672-676 (d in [1, 2, 3, 6], every separable spectral function, init_tol = Inf)
    from oracle.cones_sepspec import EpiPerSepSpectralMat
    run_oracles(EpiPerSepSpectralMat(2 + side * (side + 1) // 2, hkind, hparam), init_tol=np.inf)


@pytest.mark.parametrize("d", [1, 2, 3, 6])
@pytest.mark.parametrize("hkind,hparam", SSF)
def test_epipersepspectral_vector(d, hkind, hparam):
    # reference: test/cone.jl:672-676 with VectorCSqr
    from oracle.cones_sepspec import EpiPerSepSpectralVec
    run_oracles(EpiPerSepSpectralVec(2 + d, hkind, hparam), init_tol=np.inf)


@pytest.mark.parametrize("hkind,hparam", SSF)
def test_epipersepspectral_matrix_barrier(hkind, hparam):
    """test_barrier of test/cone.jl:117-160, :688-698 with central differences in place of ForwardDiff:
    grad, hess_prod and dder3 against derivatives of the barrier
    -log(u - v * sum h(eig(W) / v)) - log(v) - sum log eig(W)."""
    from oracle import arrayutil as au
    from oracle.cones_sepspec import EpiPerSepSpectralMat, SepSpectralFun
    side = 3
    cone = EpiPerSepSpectralMat(2 + side * (side + 1) // 2, hkind, hparam)
    h = SepSpectralFun(hkind, hparam)

    def barrier(s):
        lam = np.linalg.eigvalsh(au.svec_to_smat(s[2:]))
        return -np.log(s[0] - s[1] * h.val(lam / s[1])) - np.log(s[1]) - np.sum(np.log(lam))

    rng = np.random.default_rng(1)
    point = np.zeros(cone.dim)
    cone.set_initial_point(point)
    perturb_scale(rng, point, 0.1, 1.0)

    def grad_at(s):
        cone.reset_data()
        cone.load_point(s)
        assert cone.is_feas()
        return cone.grad().copy()

    g = grad_at(point)
    eps = 1e-6
    fd_grad = np.array([(barrier(point + eps * e) - barrier(point - eps * e)) / (2 * eps)
                        for e in np.eye(cone.dim)])
    assert close(g, fd_grad, 1e-7)
    direction = rng.standard_normal(cone.dim)
    fd_hess_dir = (grad_at(point + eps * direction) - grad_at(point - eps * direction)) / (2 * eps)
    grad_at(point)
    assert close(cone.hess_prod(direction), fd_hess_dir, 1e-7)
    assert close(cone.hess() @ direction, fd_hess_dir, 1e-7)
    # -2 dder3 = third directional derivative (cone.jl:155): second difference of the gradient
    e2 = 1e-4
    fd_third = (grad_at(point + e2 * direction) - 2 * g + grad_at(point - e2 * direction)) / e2 ** 2
    grad_at(point)
    assert close(-2 * cone.dder3(direction), fd_third, 1e-5)


def test_svec_roundtrip():
    from oracle import arrayutil as au
    rng = np.random.default_rng(0)
    for side in (1, 2, 5, 9):
        Mx = rng.standard_normal((side, side))
        Mx = Mx + Mx.T
        v = au.smat_to_svec(Mx)
        assert v.size == au.svec_length(side)
        assert np.allclose(au.svec_to_smat(v), Mx)
        # svec is an isometry: <A,B>_F = svec(A)'svec(B)
        N = rng.standard_normal((side, side))
        N = N + N.T
        assert np.isclose(v @ au.smat_to_svec(N), np.sum(Mx * N))
        K = au.symm_kron(N)
        assert np.allclose(K @ v, au.smat_to_svec(N @ Mx @ N.T))

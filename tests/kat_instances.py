"""Deterministic known-answer instances restated from the reference's test suite
(reference: test/nativeinstances.jl; line ranges per instance below).  Each entry returns
(model, expected) where `expected` holds the closed-form values the reference asserts at
tol = eps^(1/4) (nativeinstances.jl:29).  Only instances whose cones are on the hot path
(Nonnegative, EpiNormEucl, PosSemidefTri, HypoPerLogdetTri, HypoRootdetTri) are restated."""
import numpy as np

from hypatia_b200.host import models as M

RT2 = np.sqrt(2.0)
RT3 = np.sqrt(3.0)


def _m(c, A, b, G, h, cones):
    c = np.asarray(c, float)
    A = np.zeros((0, c.size)) if A is None else np.asarray(A, float)
    return M.Model(c, A, np.asarray(b if b is not None else [], float), np.asarray(G, float),
                   np.asarray(h, float), cones)


def dimension1():  # nativeinstances.jl:88-108
    return _m([-1, 0], None, None, [[1, 0]], [1], [M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=-1, x=[1, 0])


def nonnegative4():  # :295-310
    G = np.zeros((3, 2))
    G[0, 0], G[0, 1], G[1, 1], G[2, 1] = 1, -1, 1, -1
    return _m([-2, 0], None, None, G, [0, 2, 0], [M.Nonnegative(3)]), \
        dict(status="Optimal", primal_obj=-4, x=[2, 2], s=[0, 0, 2], z=[2, 2, 0])


def possemideftri1():  # :312-325
    return _m([0, -1, 0], [[1, 0, 0], [0, 0, 1]], [0.5, 1], -np.eye(3), np.zeros(3),
              [M.PosSemidefTri(3)]), dict(status="Optimal", primal_obj=-1, x_idx={1: 1.0})


def possemideftri2():  # :327-340
    return _m([0, -1, 0], [[1, 0, 1]], [0], -np.eye(3), np.zeros(3), [M.PosSemidefTri(3)]), \
        dict(status="Optimal", primal_obj=0, x=[0, 0, 0])


def possemideftri8():  # :439-462
    G = np.zeros((15, 1))
    G[[0, 2, 5, 9, 14], 0] = -1
    h = np.zeros(15)
    h[[6, 7, 8, 10, 11, 12]] = RT2 * np.array([1, 1, 0, 1, -1, 1])
    inv6, rt2inv6, invrt6 = 1 / 6, RT2 / 6, 1 / (RT2 * RT3)
    return _m([1], None, None, G, h, [M.PosSemidefTri(15)]), dict(
        status="Optimal", primal_obj=RT3,
        s=[RT3, 0, RT3, 0, 0, RT3, RT2, RT2, 0, RT3, RT2, -RT2, RT2, 0, RT3],
        z=[inv6, -rt2inv6, inv6, rt2inv6, -rt2inv6, inv6, 0, 0, 0, 0, -invrt6, invrt6, -invrt6,
           0, 0.5])


def possemideftri9():  # :464-491
    G = np.zeros((16, 10))
    for j in (1, 3, 6, 7, 9):
        G[0, j] = 0.5
    for (i, j) in ((0, 0), (1, 1), (3, 3), (6, 6), (10, 7), (15, 9)):
        G[i, j] = -1
    for (i, j) in ((2, 2), (4, 4), (5, 5), (14, 8)):
        G[i, j] = -RT2
    h = np.zeros(16)
    h[[7, 8, 9, 11, 12, 13]] = RT2 * np.array([1, 1, 0, 1, -1, 1])
    c = np.zeros(10)
    c[0] = 1
    i2, i3 = 1 / RT2, 1 / RT3
    i6 = i2 * i3
    return _m(c, None, None, G, h, [M.Nonnegative(1), M.PosSemidefTri(15)]), dict(
        status="Optimal", primal_obj=RT2 + RT3,
        s=[0, i2 + i3, 1 - RT2 / RT3, i2 + i3, RT2 * i3, -RT2 * i3, i3, RT2, RT2, 0, RT2, RT2,
           -RT2, RT2, 0, RT3],
        z=[1, 0.5, 0, 0.5, 0, 0, 0.5, -0.5, -0.5, 0, 0.5, -i6, i6, -i6, 0, 0.5])


def epinormeucl1():  # :915-931
    return _m([0, -1, -1], [[10, 0, 0], [0, 10, 0]], [10, 10 / RT2], -np.eye(3), np.zeros(3),
              [M.EpiNormEucl(3)]), dict(status="Optimal", primal_obj=-RT2,
                                        x=[1, 1 / RT2, 1 / RT2], y=[RT2 / 10, 0])


def epinormeucl2():  # :933-946
    return _m([0, -1, -1], [[1, 0, 0]], [0], -np.eye(3), np.zeros(3), [M.EpiNormEucl(3)]), \
        dict(status="Optimal", primal_obj=0, x=[0, 0, 0])


def epinormeucl3():  # :948-961
    return _m([1, 0, 0], [[0, 1, 0]], [1], -np.eye(3), np.zeros(3), [M.EpiNormEucl(3)]), \
        dict(status="Optimal", primal_obj=1, x=[1, 1, 0])


def hyporootdettri4():  # :1657-1675
    G = np.zeros((6, 4))
    G[0, 0] = G[1, 1] = G[3, 3] = -1
    G[2, 2] = -RT2
    G[4, 1] = G[5, 3] = 1
    return _m([-1, 0, 0, 0], None, None, G, [0, 0, 0, 0, 1, 1],
              [M.HypoRootdetTri(4), M.Nonnegative(2)]), dict(
        status="Optimal", primal_obj=-1, x=[1, 1, 0, 1], z=[-1, 0.5, 0, 0.5, 0.5, 0.5])


def hypoperlogdettri4():  # :1886-1906
    A = np.zeros((1, 5))
    A[0, 1] = 1
    G = np.zeros((7, 5))
    G[0, 0] = G[1, 1] = G[2, 2] = G[4, 4] = -1
    G[3, 3] = -RT2
    G[5, 2] = G[6, 4] = 1
    return _m([-1, 0, 0, 0, 0], A, [1], G, [0, 0, 0, 0, 0, 1, 1],
              [M.HypoPerLogdetTri(5), M.Nonnegative(2)]), dict(
        status="Optimal", primal_obj=0, x=[0, 1, 1, 0, 1], y=[-2], z=[-1, -2, 1, 0, 1, 1, 1])


def primalinfeas1():  # :169-180
    return _m([1, 0], [[1, 1]], [-2], -np.eye(2), np.zeros(2), [M.Nonnegative(2)]), \
        dict(status="PrimalInfeasible")


def primalinfeas2():  # :182-195
    G = np.vstack((-np.eye(3), np.diag([1.0, 1.0, -1.0])))
    return _m([1, 1, 1], None, None, G, [0, 0, 0, 1, 1, -2],
              [M.EpiNormEucl(3), M.Nonnegative(3)]), dict(status="PrimalInfeasible")


def dualinfeas_lp():
    """min -x1 : x >= 0 (unbounded).  Not a reference instance (its dualinfeas1-3 use cones
    outside the hot path); exercises the DualInfeasible status branch on a hot-path cone."""
    return _m([-1, 0], None, None, -np.eye(2), np.zeros(2), [M.Nonnegative(2)]), \
        dict(status="DualInfeasible")


ALL = [dimension1, nonnegative4, possemideftri1, possemideftri2, possemideftri8, possemideftri9,
       epinormeucl1, epinormeucl2, epinormeucl3, hyporootdettri4, hypoperlogdettri4,
       primalinfeas1, primalinfeas2, dualinfeas_lp]

TOL = np.finfo(np.float64).eps ** 0.25


def _approx(a, b, tol):
    a, b = np.atleast_1d(np.asarray(a, float)), np.atleast_1d(np.asarray(b, float))
    return np.linalg.norm(a - b) <= max(tol, tol * max(np.linalg.norm(a), np.linalg.norm(b)))


def check_solution(solver, model, expected, tol=TOL):
    """Certificate checks of build_solve_check (nativeinstances.jl:32-86) + pinned values."""
    assert solver.status == expected["status"], (solver.status, expected["status"])
    x, y, z, s = solver.get_x(), solver.get_y(), solver.get_z(), solver.get_s()
    c, A, b, G, h = model.c, model.A, model.b, model.G, model.h
    rt_tol = np.sqrt(tol)
    if solver.status == "Optimal":
        assert _approx(solver.primal_obj, solver.dual_obj, tol)
        assert _approx(c @ x + model.obj_offset, solver.primal_obj, tol)
        assert _approx(-(b @ y) - h @ z + model.obj_offset, solver.dual_obj, tol)
        assert _approx(A @ x, b, tol)
        assert _approx(G @ x + s, h, tol)
        assert _approx(G.T @ z + A.T @ y, -c, tol)
        assert _approx(s @ z, 0.0, rt_tol)
    elif solver.status == "PrimalInfeasible":
        assert _approx(-(b @ y) - h @ z, solver.dual_obj, tol)
        assert _approx(G.T @ z, -A.T @ y, rt_tol)
    elif solver.status == "DualInfeasible":
        assert _approx(c @ x, solver.primal_obj, tol)
        assert _approx(G @ x, -s, rt_tol)
        assert _approx(A @ x, np.zeros(y.size), rt_tol)
    if "primal_obj" in expected:
        assert _approx(solver.primal_obj, expected["primal_obj"], tol)
    for key, val in (("x", x), ("y", y), ("z", z), ("s", s)):
        if key in expected:
            assert _approx(val, expected[key], tol), (key, val, expected[key])
    for i, v in expected.get("x_idx", {}).items():
        assert _approx(x[i], v, tol)

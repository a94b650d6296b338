"""Deterministic known-answer instances restated from the reference's test suite
(reference: test/nativeinstances.jl; line ranges per instance below).  Each entry returns
(model, expected) where `expected` holds the closed-form values the reference asserts at
tol = eps^(1/4) (nativeinstances.jl:29).  Only instances whose cones are on the hot path
(Nonnegative, EpiNormEucl, PosSemidefTri, HypoPerLogdetTri, HypoRootdetTri, EpiPerSquare, HypoPerLog,
EpiNormInf (real), EpiPerSepSpectral{MatrixCSqr}) are restated."""
import numpy as np

from hypatia_b200.host import models as M

RT2 = np.sqrt(2.0)
RT3 = np.sqrt(3.0)


TOL_EPS4 = np.finfo(np.float64).eps ** 0.25      # test_tol of the reference (nativeinstances.jl:29)


def _m(c, A, b, G, h, cones, obj_offset=0.0):
    c = np.asarray(c, float)
    A = np.zeros((0, c.size)) if A is None else np.asarray(A, float)
    return M.Model(c, A, np.asarray(b if b is not None else [], float), np.asarray(G, float),
                   np.asarray(h, float), cones, obj_offset)


def _sparse(rows, cols, vals, m, n):
    G = np.zeros((m, n))
    for i, j, v in zip(rows, cols, vals):
        G[i - 1, j - 1] += v      # 1-based like the reference's sparse(...) calls
    return G


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


def primalinfeas3():  # :196-207
    return _m(np.zeros(3), -np.eye(3), [1, 1, 3], -np.eye(3), np.zeros(3), [M.HypoPerLog(3)]), \
        dict(status="PrimalInfeasible")


def dualinfeas2():  # :223-234
    return _m([-1, 0], None, None, [[-1, 0], [0, 0], [0, -1]], [0, 1, 0], [M.EpiPerSquare(3)]), \
        dict(status="DualInfeasible")


def dualinfeas3():  # :236-247
    return _m([0, 1, 1, 0], None, None, -np.eye(4), np.zeros(4), [M.EpiPerSquare(4)]), \
        dict(status="DualInfeasible")


def epipersquare1():  # :963-977
    return _m([0, 0, -1, -1], [[1, 0, 0, 0], [0, 1, 0, 0]], [0.5, 1], -np.eye(4), np.zeros(4),
              [M.EpiPerSquare(4)]), dict(status="Optimal", primal_obj=-RT2,
                                         x_idx={2: 1 / RT2, 3: 1 / RT2})


def epipersquare2():  # :979-994
    i2 = 1 / RT2
    return _m([0, 0, -1], [[1, 0, 0], [0, 1, 0]], [i2 / 2, i2], -np.eye(3), np.zeros(3),
              [M.EpiPerSquare(3)], obj_offset=-1.0), \
        dict(status="Optimal", primal_obj=-i2 - 1, x_idx={1: i2})


def epipersquare3():  # :996-1009
    return _m([0, 1, -1, -1], [[1, 0, 0, 0]], [0], -np.eye(4), np.zeros(4), [M.EpiPerSquare(4)]), \
        dict(status="Optimal", primal_obj=0, x=[0, 0, 0, 0])


def epipersquare4():  # :1011-1036
    G = _sparse([1, 1, 2, 2, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [1, 7, 2, 3, 4, 2, 3, 5, 4, 7, 6, 5, 6, 7],
                [1, -0.5, 1, 1, 1, -1, -1, -1, -1, -0.5, -1, -1, -1, -1], 11, 7)
    h = np.zeros(11)
    h[1] = 3
    c = np.zeros(7)
    c[0] = -1
    i3 = 1 / 3
    r3 = i3 * RT2
    return _m(c, None, None, G, h, [M.Nonnegative(2), M.EpiPerSquare(3), M.EpiPerSquare(3),
                                    M.EpiPerSquare(3)]), dict(
        status="Optimal", primal_obj=-1, x=[1, 1, 1, 1, RT2, RT2, 2],
        s=[0, 0, 1, 1, RT2, 1, 1, RT2, RT2, RT2, 2],
        z=[1, i3, i3, i3, -r3, i3, i3, -r3, r3, r3, -2 * i3])


def epinorminf1():  # :791-806
    i2 = 1 / RT2
    return _m([0, -1, -1], [[1, 0, 0], [0, 1, 0]], [1, i2], -np.eye(3), np.zeros(3), [M.EpiNormInf(3)]), \
        dict(status="Optimal", primal_obj=-1 - i2, x=[1, i2, 1], y=[1, 1])


def epinorminf2():  # :808-829
    l = 3
    L = 2 * l + 1
    A = np.zeros((2, L))
    A[0, 0] = A[0, L - 1] = A[1, 0] = 1
    A[1, L - 1] = -1
    G = np.vstack((np.zeros((1, L)), np.eye(L), np.zeros((1, L)), 2 * np.eye(L)))
    h = np.zeros(2 * L + 2)
    h[0] = 1
    h[L + 1] = 1
    return _m(np.arange(-l, l + 1, dtype=float), A, [0, 0], G, h,
              [M.EpiNormInf(L + 1, use_dual=True), M.EpiNormInf(L + 1)], obj_offset=1.0), \
        dict(status="Optimal", primal_obj=-l + 2, x_idx={1: 0.5, L - 2: -0.5}, tol_scale=10)


def epinorminf3():  # :831-847 (primal barrier)
    return _m([1, 0, 0, 0, 0, 0], None, None, -np.eye(6), np.zeros(6), [M.EpiNormInf(6)]), \
        dict(status="Optimal", primal_obj=0, x=np.zeros(6))


def epinorminf3_dual():  # :831-847 (dual barrier)
    return _m([1, 0, 0, 0, 0, 0], None, None, -np.eye(6), np.zeros(6), [M.EpiNormInf(6, use_dual=True)]), \
        dict(status="Optimal", primal_obj=0, x=np.zeros(6))


def epinorminf4():  # :849-863
    return _m([0, 1, -1], [[1, 0, 0], [0, 1, 0]], [1, -0.4], -np.eye(3), np.zeros(3),
              [M.EpiNormInf(3, use_dual=True)]), \
        dict(status="Optimal", primal_obj=-1, x=[1, -0.4, 0.6], y=[1, 0])


def dualinfeas1():  # :209-221
    G = np.vstack((-np.eye(3), -np.eye(3)))
    return _m([-1, -1, 0], None, None, G, np.zeros(6), [M.EpiNormInf(3), M.EpiNormInf(3, use_dual=True)]), \
        dict(status="DualInfeasible")


def _generalizedpower1(use_dual):  # :1261-1281
    v = -RT2 if use_dual else -1 / RT2
    return _m([0, 0, 1], [[1, 0, 0], [0, 1, 0]], [0.5, 1], -np.eye(3), np.zeros(3),
              [M.GeneralizedPower([0.5, 0.5], 1, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=v, x=[0.5, 1, v])


def _generalizedpower2(use_dual):  # :1283-1304
    w = 1.0 if use_dual else 0.5
    return _m([0, 0, -1, -1], [[0, 1, 0, 0], [1, 0, 0, 0]], [0.5, 1], -np.eye(4), np.zeros(4),
              [M.GeneralizedPower([0.5, 0.5], 2, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=-2.0 if use_dual else -1.0, x=[1, 0.5, w, w])


def _generalizedpower3(use_dual):  # :1306-1325
    l = 4
    A = np.zeros((2, l + 2))
    A[0, l] = A[1, l + 1] = 1
    return _m(np.concatenate((np.full(l, 10.0), [0, 0])), A, [1, 0], -10 * np.eye(l + 2), np.zeros(l + 2),
              [M.GeneralizedPower(np.full(l, 1 / l), 2, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=10.0 if use_dual else 10.0 * l,
             x_idx={i: (1 / l if use_dual else 1.0) for i in range(l)})


def _generalizedpower4(use_dual):  # :1327-1345
    l = 4
    G = np.vstack((np.zeros((3, l)), -np.eye(l)))
    return _m(np.ones(l), None, None, G, np.zeros(l + 3), [M.GeneralizedPower(np.full(l, 1 / l), 3, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=0, x=np.zeros(l))


def _hypopowermean1(use_dual):  # :1347-1364
    return _m([-1, 0, 0], [[0, 0, 1], [0, 1, 0]], [0.5, 1], -np.eye(3), np.zeros(3),
              [M.HypoPowerMean([0.5, 0.5], use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=0.0 if use_dual else -1 / RT2, x_idx={1: 1.0, 2: 0.5})


def _hypopowermean2(use_dual):  # :1366-1385
    l = 4
    A = np.zeros((1, l + 1))
    A[0, 0] = 1
    return _m(np.concatenate(([0.0], np.ones(l))), A, [-1.0 if use_dual else 1.0], -np.eye(l + 1), np.zeros(l + 1),
              [M.HypoPowerMean(np.full(l, 1 / l), use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=1.0 if use_dual else float(l),
             x_idx={i: (1 / l if use_dual else 1.0) for i in range(1, l + 1)})


def hypopowermean4():  # :1407-1423
    G = np.vstack((-np.eye(4), [[0, 1, 1, 1]]))
    return _m([-1, 0, 0, 0], None, None, G, [0, 0, 0, 0, 3], [M.HypoPowerMean(np.full(3, 1 / 3)), M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=-1, x=[1, 1, 1, 1], s=[1, 1, 1, 1, 0], z=[-1, 1 / 3, 1 / 3, 1 / 3, 1 / 3])


def hypopowermean5():  # :1425-1440
    G = _sparse([1, 2, 3], [1, 2, 2], [-1, -1, 1], 3, 2)
    return _m([-2, 0], None, None, G, [0, 0, 2], [M.HypoPowerMean([1.0]), M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=-4, x=[2, 2], s=[2, 2, 0], z=[-2, 2, 2])


def hypopowermean6():  # :1442-1458
    c = np.zeros(10)
    c[0] = -1
    A = np.hstack((np.zeros((9, 1)), np.eye(9)))
    return _m(c, A, np.ones(9), -np.eye(10), np.zeros(10), [M.HypoPowerMean(np.full(9, 1 / 9))]), \
        dict(status="Optimal", primal_obj=-1, x=np.ones(10), z=np.concatenate(([-1.0], np.full(9, 1 / 9))),
             y=np.full(9, 1 / 9))


def hypogeomean1():  # :1460-1476 (primal barrier)
    return _m([-1, 0, 0], [[0, 0, 1], [0, 1, 0]], [0.5, 1], -np.eye(3), np.zeros(3), [M.HypoGeoMean(3)]), \
        dict(status="Optimal", primal_obj=-1 / RT2, x_idx={1: 1.0, 2: 0.5})


def hypogeomean1_dual():  # :1460-1476 (dual barrier)
    return _m([-1, 0, 0], [[0, 0, 1], [0, 1, 0]], [0.5, 1], -np.eye(3), np.zeros(3),
              [M.HypoGeoMean(3, use_dual=True)]), dict(status="Optimal", primal_obj=0, x_idx={1: 1.0, 2: 0.5})


def hypogeomean2():  # :1478-1496 (primal barrier)
    l = 4
    A = np.zeros((1, l + 1))
    A[0, 0] = 1
    return _m(np.concatenate(([0.0], np.ones(l))), A, [1], -np.eye(l + 1), np.zeros(l + 1), [M.HypoGeoMean(l + 1)]), \
        dict(status="Optimal", primal_obj=l, x_idx={i: 1.0 for i in range(1, l + 1)})


def hypogeomean2_dual():  # :1478-1496 (dual barrier)
    l = 4
    A = np.zeros((1, l + 1))
    A[0, 0] = 1
    return _m(np.concatenate(([0.0], np.ones(l))), A, [-1], -np.eye(l + 1), np.zeros(l + 1),
              [M.HypoGeoMean(l + 1, use_dual=True)]), \
        dict(status="Optimal", primal_obj=1, x_idx={i: 1.0 / l for i in range(1, l + 1)})


def hypogeomean4():  # :1517-1532
    G = np.vstack((-np.eye(4), [[0, 1, 1, 1]]))
    return _m([-1, 0, 0, 0], None, None, G, [0, 0, 0, 0, 3], [M.HypoGeoMean(4), M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=-1, x=[1, 1, 1, 1], s=[1, 1, 1, 1, 0], z=[-1, 1 / 3, 1 / 3, 1 / 3, 1 / 3])


def hypogeomean5():  # :1534-1549
    G = _sparse([1, 2, 3], [1, 2, 2], [-1, -1, 1], 3, 2)
    return _m([-2, 0], None, None, G, [0, 0, 2], [M.HypoGeoMean(2), M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=-4, x=[2, 2], s=[2, 2, 0], z=[-2, 2, 2])


def hypogeomean6():  # :1551-1567
    c = np.zeros(10)
    c[0] = -1
    A = np.hstack((np.zeros((9, 1)), np.eye(9)))
    return _m(c, A, np.ones(9), -np.eye(10), np.zeros(10), [M.HypoGeoMean(10)]), \
        dict(status="Optimal", primal_obj=-1, x=np.ones(10), z=np.concatenate(([-1.0], np.full(9, 1 / 9))),
             y=np.full(9, 1 / 9))


def hypoperlog1():  # :1677-1693
    e = np.exp(0.5)
    return _m([1, 1, 1], [[0, 1, 0], [1, 0, 0]], [2, 1], -np.eye(3), np.zeros(3), [M.HypoPerLog(3)]), \
        dict(status="Optimal", primal_obj=2 * e + 3, x=[1, 2, 2 * e], y=[-(1 + e / 2), -(1 + e)])


def hypoperlog2():  # :1695-1707
    return _m([-1, 0, 0], [[0, 1, 0]], [0], -np.eye(3), np.zeros(3), [M.HypoPerLog(3)]), \
        dict(status="Optimal", primal_obj=0)


def hypoperlog3():  # :1709-1723
    G = _sparse([1, 2, 3, 4], [1, 2, 3, 1], [-1, -1, -1, -1], 4, 3)
    return _m([1, 1, 1], None, None, G, np.zeros(4), [M.HypoPerLog(3), M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=0, x=[0, 0, 0])


def hypoperlog4():  # :1725-1739
    e2 = np.exp(-2.0)
    return _m([0, 0, 1], [[0, 1, 0], [1, 0, 0]], [1, -1], -np.eye(3), np.zeros(3),
              [M.HypoPerLog(3, use_dual=True)]), dict(status="Optimal", primal_obj=e2, x=[-1, 1, e2])


def hypoperlog5():  # :1741-1756
    lq = np.log(0.25)
    G = _sparse([1, 3, 4], [1, 2, 3], [-1, -1, -1], 4, 3)
    return _m([-1, 0, 0], [[0, 1, 1]], [1], G, [0, 1, 0, 0], [M.HypoPerLog(4)]), \
        dict(status="Optimal", primal_obj=-lq, x=[lq, 0.5, 0.5], y=[2])


def hypoperlog6():  # :1758-1772
    G = _sparse([1, 3, 4], [1, 2, 3], [-1, -1, -1], 4, 3)
    return _m([-1, 0, 0], None, None, G, np.zeros(4), [M.HypoPerLog(4)]), \
        dict(status="Optimal", primal_obj=0, x_idx={0: 0.0})


def hypoperlog7():  # :1774-1795
    G = _sparse([1, 2, 2, 3, 4, 5, 6], [2, 1, 3, 4, 4, 3, 2], [1, 1, -1, -1, -1, -1, -1], 6, 4)
    h = np.zeros(6)
    h[0] = 2
    return _m([-2, 0, 0, 0], None, None, G, h, [M.Nonnegative(3), M.HypoPerLog(3)]), dict(
        status="Optimal", primal_obj=-4, x=[2, 2, 2, 0], s=[0, 0, 0, 0, 2, 2], z=[2, 2, 2, -2, -2, 2])


# the separable spectral functions the reference's tests loop over (sep_spectral_funs, test/cone.jl /
# nativeinstances.jl): (HYP_SSF kind, parameter)
SEP_SPECTRAL_FUNS = [(M.SSF_INV, 0.0), (M.SSF_NEGLOG, 0.0), (M.SSF_NEGENTROPY, 0.0), (M.SSF_POWER12, 1.5)]


def _ssf(hkind, hparam):
    from oracle.cones_sepspec import SepSpectralFun
    return SepSpectralFun(hkind, hparam)


def _svec(W):
    from oracle import arrayutil as au
    return au.smat_to_svec(W)


def _spectral_matrix1(d, hkind, hparam):  # :2009-2037 (real case): min u : (u, 1, W) in K => h(eig W)
    rng = np.random.default_rng(1)
    W = rng.random((d, d))
    W = W @ W.T + np.eye(d)
    dim = 2 + M.svec_length(d)
    G = np.zeros((dim, 1))
    G[0, 0] = -1
    h = np.zeros(dim)
    h[1] = 1
    h[2:] = _svec(W)
    return _m([1], None, None, G, h, [M.EpiPerSepSpectralMat(dim, hkind, hparam)]), \
        dict(status="Optimal", primal_obj=_ssf(hkind, hparam).val(np.linalg.eigvalsh(W)))


def _spectral_matrix2(d, hkind, hparam):  # :2039-2070: dual barrier => conjugate of h
    rng = np.random.default_rng(1)
    W = rng.random((d, d))
    f = _ssf(hkind, hparam)
    if f.conj_dom_pos():
        W = W @ W.T + np.eye(d)
    else:
        W = (W + W.T) / 2     # the reference reads the upper triangle of a random square matrix
    dim = 2 + M.svec_length(d)
    G = np.zeros((dim, 1))
    G[1, 0] = -1
    h = np.zeros(dim)
    h[0] = 1
    h[2:] = _svec(W)
    return _m([1], None, None, G, h, [M.EpiPerSepSpectralMat(dim, hkind, hparam, use_dual=True)]), \
        dict(status="Optimal", primal_obj=f.conj(np.linalg.eigvalsh(W)))


def _spectral_matrix3(d, hkind, hparam):  # :2072-2098
    dim = 2 + M.svec_length(d)
    c = np.zeros(dim)
    c[0] = 1
    A = np.zeros((1, dim))
    A[0, 1] = 1
    return _m(c, A, [0], -np.eye(dim), np.zeros(dim), [M.EpiPerSepSpectralMat(dim, hkind, hparam)]), \
        dict(status="Optimal", primal_obj=0, x_idx={0: 0.0, 1: 0.0})


def _spectral_vector1(d, hkind, hparam):  # :1908-1932: min u : (u, 1, w) in K => sum h(w_i)
    w = np.random.default_rng(1).random(d) + 1
    G = np.zeros((2 + d, 1))
    G[0, 0] = -1
    h = np.concatenate(([0.0, 1.0], w))
    return _m([1], None, None, G, h, [M.EpiPerSepSpectralVec(2 + d, hkind, hparam)]), \
        dict(status="Optimal", primal_obj=_ssf(hkind, hparam).val(w))


def _spectral_vector2(d, hkind, hparam):  # :1934-1956 (dual barrier)
    dim = 2 + d
    c = np.zeros(dim)
    c[0] = 1
    A = np.zeros((1, dim))
    A[0, 0] = 1
    return _m(c, A, [0], -np.eye(dim), np.zeros(dim), [M.EpiPerSepSpectralVec(dim, hkind, hparam, use_dual=True)]), \
        dict(status="Optimal", primal_obj=0)


def _spectral_vector3(hkind, hparam):  # :1958-1979
    f = _ssf(hkind, hparam)
    val = 5 * f.val(np.array([2.0, 3.0]) / 5)
    return _m([1], None, None, -np.eye(4)[:, :1], [0, 5, 2, 3], [M.EpiPerSepSpectralVec(4, hkind, hparam)]), \
        dict(status="Optimal", primal_obj=val, s=[val, 5, 2, 3], z_idx={0: 1.0})


def _spectral_vector4(hkind, hparam):  # :1981-2007 (dual barrier: conjugate)
    f = _ssf(hkind, hparam)
    w = np.array([2.0, 3.0]) * (1 if f.conj_dom_pos() else -1)
    G = np.zeros((4, 1))
    G[1, 0] = -1
    val = 5 * f.conj(w / 5)
    return _m([1], None, None, G, [5, 0, w[0], w[1]], [M.EpiPerSepSpectralVec(4, hkind, hparam, use_dual=True)]), \
        dict(status="Optimal", primal_obj=val, s=[5, val, w[0], w[1]], z_idx={1: 1.0})


def _named(fn, name):
    fn.__name__ = name
    return fn


SPECTRAL = []
for _k, (_hk, _hp) in enumerate(SEP_SPECTRAL_FUNS):
    for _d in (1, 3):
        SPECTRAL.append(_named(lambda d=_d, hk=_hk, hp=_hp: _spectral_matrix1(d, hk, hp),
                               f"epipersepspectral_matrix1_d{_d}_h{_hk}"))
        SPECTRAL.append(_named(lambda d=_d, hk=_hk, hp=_hp: _spectral_matrix2(d, hk, hp),
                               f"epipersepspectral_matrix2_d{_d}_h{_hk}"))
    for _d in (2, 4):
        SPECTRAL.append(_named(lambda d=_d, hk=_hk, hp=_hp: _spectral_matrix3(d, hk, hp),
                               f"epipersepspectral_matrix3_d{_d}_h{_hk}"))

SPECTRAL_VEC = []
for _k, (_hk, _hp) in enumerate(SEP_SPECTRAL_FUNS):
    for _d in (1, 3):
        SPECTRAL_VEC.append(_named(lambda d=_d, hk=_hk, hp=_hp: _spectral_vector1(d, hk, hp),
                                   f"epipersepspectral_vector1_d{_d}_h{_hk}"))
    for _d in (2, 4):
        SPECTRAL_VEC.append(_named(lambda d=_d, hk=_hk, hp=_hp: _spectral_vector2(d, hk, hp),
                                   f"epipersepspectral_vector2_d{_d}_h{_hk}"))
    SPECTRAL_VEC.append(_named(lambda hk=_hk, hp=_hp: _spectral_vector3(hk, hp), f"epipersepspectral_vector3_h{_hk}"))
    SPECTRAL_VEC.append(_named(lambda hk=_hk, hp=_hp: _spectral_vector4(hk, hp), f"epipersepspectral_vector4_h{_hk}"))

def _epirelentropy1(d):  # :2099-2119: min u : (u, 1, w) in K => sum w log w
    w = np.random.default_rng(1).random(d) + 1
    dim = 1 + 2 * d
    G = np.zeros((dim, 1))
    G[0, 0] = -1
    h = np.concatenate(([0.0], np.ones(d), w))
    return _m([1], None, None, G, h, [M.EpiRelEntropy(dim)]), \
        dict(status="Optimal", primal_obj=float(np.sum(w * np.log(w))))


def _epirelentropy2(d):  # :2121-2141
    dim = 1 + 2 * d
    G = np.zeros((dim, d))
    G[1 + d + np.arange(d), np.arange(d)] = -1
    h = np.zeros(dim)
    h[1:1 + d] = 1
    return _m(-np.ones(d), None, None, G, h, [M.EpiRelEntropy(dim)]), \
        dict(status="Optimal", primal_obj=-d, x=np.ones(d))


def _epirelentropy3(d):  # :2143-2163
    dim = 1 + 2 * d
    G = np.zeros((dim, d))
    G[1 + np.arange(d), np.arange(d)] = -1
    h = np.zeros(dim)
    h[1 + d:] = 1
    return _m(-np.ones(d), np.ones((1, d)), [dim], G, h, [M.EpiRelEntropy(dim)]), \
        dict(status="Optimal", primal_obj=-dim, x=np.full(d, dim / d))


def epirelentropy4():  # :2165-2181
    G = np.zeros((5, 1))
    G[0, 0] = -1
    entr = 2 * np.log(2.0) + 3 * np.log(3 / 5.0)
    return _m([1], None, None, G, [0, 1, 5, 2, 3], [M.EpiRelEntropy(5)]), \
        dict(status="Optimal", primal_obj=entr, s=[entr, 1, 5, 2, 3],
             z=[1, 2, 3 / 5.0, np.log(0.5) - 1, np.log(5 / 3.0) - 1])


def epirelentropy5():  # :2183-2198
    G = np.vstack((np.zeros((4, 2)), -np.ones((3, 2)), [[-1.0, 0.0]]))
    h = np.zeros(8)
    h[1:4] = 1
    return _m([0, -1], None, None, G, h, [M.EpiRelEntropy(7), M.Nonnegative(1)]), \
        dict(status="Optimal", primal_obj=-1, s=[0, 1, 1, 1, 1, 1, 1, 0],
             z=np.array([1, 1, 1, 1, -1, -1, -1, 3]) / 3.0)


def _svals(vec, d1, d2):
    return np.linalg.svd(vec.reshape(d1, d2, order="F"), compute_uv=False)


def _epinormspectral1(use_dual):  # :1038-1073 (real case)
    rng = np.random.default_rng(1)
    d1, d2 = 3, 4
    dim = d1 * d2
    c = np.concatenate(([1.0], np.zeros(dim)))
    A = np.hstack((np.zeros((dim, 1)), np.eye(dim)))
    b = rng.random(dim)
    h = np.concatenate(([0.0], rng.random(dim)))

    def check(s, z, approx):
        ps, ds = _svals(s[1:], d1, d2), _svals(z[1:], d1, d2)
        if use_dual:
            assert approx(ps.sum(), s[0]) and approx(ds[0], z[0])
        else:
            assert approx(ps[0], s[0]) and approx(ds.sum(), z[0])
    return _m(c, A, b, -np.eye(dim + 1), h, [M.EpiNormSpectral(d1, d2, use_dual=use_dual)]), \
        dict(status="Optimal", check=check)


def _epinormspectral2(use_dual):  # :1075-1103 (real case)
    d1, d2 = 3, 4
    dim = d1 * d2
    mat = np.random.default_rng(1).random((d1, d2))
    G = np.vstack((np.zeros((1, dim)), -np.eye(dim)))
    h = np.concatenate(([1.0], np.zeros(dim)))
    sv = np.linalg.svd(mat, compute_uv=False)
    return _m(-mat.ravel(order="F"), None, None, G, h, [M.EpiNormSpectral(d1, d2, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=-sv[0] if use_dual else -sv.sum())


def _epinormspectral3(d1, d2, use_dual):  # :1105-1125 (real case)
    dim = d1 * d2
    G = np.vstack((np.zeros((1, dim)), -np.eye(dim)))
    return _m(-np.ones(dim), None, None, G, np.zeros(dim + 1), [M.EpiNormSpectral(d1, d2, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=0, x=np.zeros(dim))


def _epinormspectral4(use_dual):  # :1127-1155
    G = np.zeros((7, 1))
    G[0, 0] = -1
    h = np.array([0, 1, 1, 1, -1, 0, 1.0])
    rt2, rt3 = np.sqrt(2.0), np.sqrt(3.0)
    if use_dual:
        exp = dict(status="Optimal", primal_obj=rt2 + rt3, s=[rt2 + rt3, 1, 1, 1, -1, 0, 1],
                   z=[1, -1 / rt2, -1 / rt3, -1 / rt2, 1 / rt3, 0, -1 / rt3])
    else:
        exp = dict(status="Optimal", primal_obj=rt3, s=[rt3, 1, 1, 1, -1, 0, 1],
                   z=[1, 0, -1 / rt3, 0, 1 / rt3, 0, -1 / rt3])
    return _m([1], None, None, G, h, [M.EpiNormSpectral(2, 3, use_dual=use_dual)]), exp


NORMSPEC = [_named(lambda f=_f, ud=_ud: f(ud), _f.__name__[1:] + ("_dual" if _ud else ""))
            for _f in (_epinormspectral1, _epinormspectral2, _epinormspectral4) for _ud in (False, True)] + \
    [_named(lambda a=_a, b=_b, ud=_ud: _epinormspectral3(a, b, ud), f"epinormspectral3_{_a}x{_b}" + ("_dual" if _ud else ""))
     for (_a, _b) in ((1, 1), (1, 3), (2, 2), (3, 4)) for _ud in (False, True)]

def _wsos_box(lo, hi, fn):
    from wsos_util import interpolate_box
    U, pts, Ps = interpolate_box(lo, hi, 2)
    return U, Ps, np.array([fn(*p) for p in pts])


def wsosinterpnonnegative1():  # :2286-2304: min of x^4 + x^2 y^2 + 4 y^2 + 4 over [0, 1]^2
    U, Ps, vals = _wsos_box([0, 0], [1, 1], lambda x, y: x ** 4 + x ** 2 * y ** 2 + 4 * y ** 2 + 4)
    return _m([-1], None, None, np.ones((U, 1)), vals, [M.WSOSInterpNonnegative(U, Ps)]), \
        dict(status="Optimal", primal_obj=-4, x=[4.0])


def wsosinterpnonnegative2():  # :2306-2324: min of (x - 2)^2 + (x y - 3)^2 over [0, 3]^2
    U, Ps, vals = _wsos_box([0, 0], [3, 3], lambda x, y: (x - 2) ** 2 + (x * y - 3) ** 2)
    return _m([-1], None, None, np.ones((U, 1)), vals, [M.WSOSInterpNonnegative(U, Ps)]), \
        dict(status="Optimal", primal_obj=0, x=[0.0])


def wsosinterpnonnegative3():  # :2326-2343: the dual formulation with use_dual = true
    U, Ps, vals = _wsos_box([0, 0], [3, 3], lambda x, y: (x - 2) ** 2 + (x * y - 3) ** 2)
    return _m(vals, np.ones((1, U)), [1], -np.eye(U), np.zeros(U), [M.WSOSInterpNonnegative(U, Ps, use_dual=True)]), \
        dict(status="Optimal", primal_obj=0)


def wsosinterppossemideftri1():  # :2385-2405: convexity parameter of (x + 1)^2 (x - 1)^2 on [-1, 1]: H = 12 x^2 - 4
    from wsos_util import interpolate_box
    U, pts, Ps = interpolate_box([-1.0], [1.0], 1)
    h = 12 * pts[:, 0] ** 2 - 4
    return _m([-1], None, None, np.ones((U, 1)), h, [M.WSOSInterpPosSemidefTri(1, U, Ps)]), \
        dict(status="Optimal", primal_obj=4, x=[-4.0])


def wsosinterppossemideftri2():  # :2407-2428: convexity parameter of x1^4 - 3 x2^2: H = [12 x1^2, 0; 0, -6]
    from wsos_util import interpolate_free
    U, pts, Ps = interpolate_free(2, 1)
    G = np.vstack((np.ones((U, 1)), np.zeros((U, 1)), np.ones((U, 1))))
    h = np.concatenate((12 * pts[:, 0] ** 2, np.zeros(U), -6 * np.ones(U)))     # blocks (1,1), (2,1), (2,2)
    return _m([-1], None, None, G, h, [M.WSOSInterpPosSemidefTri(2, U, Ps)]), \
        dict(status="Optimal", primal_obj=6, x=[-6.0])


def wsosinterpepinormeucl1():  # :2522-2542: min constant t : t^2 >= x^4 on [-1, 1]  =>  t = 1
    from wsos_util import interpolate_box
    U, pts, Ps = interpolate_box([-1.0], [1.0], 1)
    G = np.vstack((-np.eye(U), np.zeros((U, U))))
    h = np.concatenate((np.zeros(U), pts[:, 0] ** 2))
    return _m(np.ones(U), [[1, -1, 0], [1, 0, -1]], [0, 0], G, h, [M.WSOSInterpEpiNormEucl(2, U, Ps)]), \
        dict(status="Optimal", primal_obj=U, x=np.ones(U))


def wsosinterpepinormeucl2():  # :2544-2565: t^2 >= x^4 + (x - 1)^2 on [-1, 1]  =>  t = sqrt 5
    from wsos_util import interpolate_box
    U, pts, Ps = interpolate_box([-1.0], [1.0], 1)
    G = np.vstack((-np.eye(U), np.zeros((U, U)), np.zeros((U, U))))
    h = np.concatenate((np.zeros(U), pts[:, 0] ** 2, pts[:, 0] - 1))
    return _m(np.ones(U), [[1, -1, 0], [1, 0, -1]], [0, 0], G, h, [M.WSOSInterpEpiNormEucl(3, U, Ps)]), \
        dict(status="Optimal", primal_obj=np.sqrt(5.0) * U, x=np.full(U, np.sqrt(5.0)))


def wsosinterpepinormone1():  # :2452-2472: min constant t : t >= |x^2| on [-1, 1]  =>  t = 1
    from wsos_util import interpolate_box
    U, pts, Ps = interpolate_box([-1.0], [1.0], 1)
    G = np.vstack((-np.eye(U), np.zeros((U, U))))
    h = np.concatenate((np.zeros(U), pts[:, 0] ** 2))
    return _m(np.ones(U), [[1, -1, 0], [1, 0, -1]], [0, 0], G, h, [M.WSOSInterpEpiNormOne(2, U, Ps)]), \
        dict(status="Optimal", primal_obj=U, x=np.ones(U))


WSOSONE = [wsosinterpepinormone1]

WSOSEUCL = [wsosinterpepinormeucl1, wsosinterpepinormeucl2]

WSOSPSD = [wsosinterppossemideftri1, wsosinterppossemideftri2]

WSOS = [wsosinterpnonnegative1, wsosinterpnonnegative2, wsosinterpnonnegative3]

# ---- further instances of the reference for the cones of the path: seeded-random data with closed-form optima or
# certificate-only checks (the optimum does not depend on the RNG stream, so NumPy's generator stands in for Julia's) ----
def _randint(rng, lo, hi, *shape):
    return rng.integers(lo, hi + 1, size=shape).astype(float)


def consistent1():  # :110-130: dependent equality rows and dependent columns that preprocessing must remove
    rng = np.random.default_rng(1)
    n, p, q = 30, 15, 30
    c = np.zeros(n)
    A = _randint(rng, -9, 9, p, n)
    G = 10.0 * np.eye(q, n)
    r1, r2 = rng.random(2)
    A[10:15] = r1 * A[0:5] - r2 * A[5:10]
    b = A.sum(axis=1)
    r1, r2 = rng.random(2)
    A[:, 10:15] = r1 * A[:, 0:5] - r2 * A[:, 5:10]
    G[:, 10:15] = r1 * G[:, 0:5] - r2 * G[:, 5:10]
    c[10:15] = r1 * c[0:5] - r2 * c[5:10]
    return _m(c, A, b, G, np.zeros(q), [M.Nonnegative(q)]), dict(status="Optimal", tol_scale=10)


def inconsistent1():  # :132-149
    rng = np.random.default_rng(1)
    n, p, q = 30, 15, 30
    c = _randint(rng, 0, 9, n)
    A = _randint(rng, -9, 9, p, n)
    b = rng.random(p)
    r1, r2 = rng.random(2)
    A[10:15] = r1 * A[0:5] - r2 * A[5:10]
    b[10:15] = 2 * (r1 * b[0:5] - r2 * b[5:10])
    return _m(c, A, b, -np.eye(q, n), np.zeros(q), [M.Nonnegative(q)]), dict(status="PrimalInconsistent")


def inconsistent2():  # :151-168
    rng = np.random.default_rng(1)
    n, p, q = 30, 15, 30
    c = _randint(rng, 0, 9, n)
    A = _randint(rng, -9, 9, p, n)
    G = -np.eye(q, n)
    b = rng.random(p)
    r1, r2 = rng.random(2)
    A[:, 10:15] = r1 * A[:, 0:5] - r2 * A[:, 5:10]
    G[:, 10:15] = r1 * G[:, 0:5] - r2 * G[:, 5:10]
    c[10:15] = 2 * (r1 * c[0:5] - r2 * c[5:10])
    return _m(c, A, b, G, np.zeros(q), [M.Nonnegative(q)]), dict(status="DualInconsistent")


def nonnegative1():  # :249-263
    rng = np.random.default_rng(1)
    n, p, q = 6, 3, 6
    c = _randint(rng, 0, 9, n)
    A = _randint(rng, -9, 9, p, n)
    return _m(c, A, A.sum(axis=1), -np.eye(q, n), np.zeros(q), [M.Nonnegative(q)], obj_offset=1.0), dict(status="Optimal")


def nonnegative2():  # :265-278
    rng = np.random.default_rng(1)
    n, p, q = 5, 2, 10
    c = _randint(rng, 0, 9, n)
    A = _randint(rng, 1, 9, p, n)
    G = rng.random((q, n)) - 2.0 * np.eye(q, n)
    return _m(c, A, A.sum(axis=1), G, G.sum(axis=1), [M.Nonnegative(q)]), dict(status="Optimal", tol_scale=2)


def nonnegative3():  # :280-293
    rng = np.random.default_rng(1)
    n, p, q = 15, 6, 15
    c = _randint(rng, 0, 9, n)
    A = _randint(rng, -9, 9, p, n)
    return _m(c, A, A.sum(axis=1), -np.eye(q), np.zeros(q), [M.Nonnegative(q)]), dict(status="Optimal", tol_scale=2)


def _randsym(rng, s):
    X = rng.random((s, s))
    return np.triu(X) + np.triu(X, 1).T        # Hermitian(rand(s, s), :U)


def possemideftri3():  # :342-360: min x : x I - M psd  =>  lambda_max(M)
    Mx = _randsym(np.random.default_rng(1), 2)
    emax = float(np.linalg.eigvalsh(Mx)[-1])
    return _m([1], None, None, [[-1.0], [0.0], [-1.0]], -_svec(Mx), [M.PosSemidefTri(3)]), \
        dict(status="Optimal", primal_obj=emax, x=[emax])


def possemideftri4():  # :362-380: max <M, X> : tr X = 1, X psd  =>  lambda_max(M)
    s = 3
    Mx = _randsym(np.random.default_rng(1), s)
    dim = M.svec_length(s)
    return _m(-_svec(Mx), _svec(np.eye(s))[None, :], [1], -np.eye(dim), np.zeros(dim), [M.PosSemidefTri(dim)]), \
        dict(status="Optimal", primal_obj=-float(np.linalg.eigvalsh(Mx)[-1]))


def _smat(v):
    from oracle import arrayutil as au
    return au.svec_to_smat(np.asarray(v, float))


def hyporootdettri1():  # :1569-1598 (real case)
    side = 3
    H = np.random.default_rng(1).random((side, side))
    dim = 1 + M.svec_length(side)
    G = np.zeros((dim, 1))
    G[0, 0] = -1
    h = np.concatenate(([0.0], _svec(H @ H.T)))

    def check(s, z, approx):
        assert approx(np.linalg.det(_smat(s[1:])) ** (1 / side), s[0])
        assert approx(np.linalg.det(_smat(z[1:] * side)) ** (1 / side), -z[0])
    return _m([-1], None, None, G, h, [M.HypoRootdetTri(dim)]), dict(status="Optimal", check=check)


def hyporootdettri2():  # :1600-1629 (real case, dual barrier)
    side = 4
    H = np.random.default_rng(1).random((side, side))
    dim = 1 + M.svec_length(side)
    G = np.zeros((dim, 1))
    G[0, 0] = -1
    h = np.concatenate(([0.0], _svec(H @ H.T)))

    def check(s, z, approx):
        assert approx(np.linalg.det(_smat(s[1:] * side)) ** (1 / side), -s[0])
        assert approx(np.linalg.det(_smat(z[1:])) ** (1 / side), z[0])
    return _m([1], None, None, G, h, [M.HypoRootdetTri(dim, use_dual=True)]), dict(status="Optimal", check=check)


def hyporootdettri3():  # :1631-1655: W not full rank => optimum 0 (tol eps^0.15)
    side = 3
    H = 0.2 * np.random.default_rng(1).random((side, side - 1))
    dim = 1 + M.svec_length(side)
    G = np.zeros((dim, 1))
    G[0, 0] = -1
    h = np.concatenate(([0.0], _svec(H @ H.T)))
    return _m([-1], None, None, G, h, [M.HypoRootdetTri(dim)]), \
        dict(status="Optimal", primal_obj=0, x=[0.0], tol_scale=np.finfo(float).eps ** 0.15 / TOL_EPS4)


def hypoperlogdettri1():  # :1797-1827 (real case)
    side = 4
    H = np.random.default_rng(1).random((side, side))
    dim = 2 + M.svec_length(side)
    G = np.zeros((dim, 2))
    G[0, 0] = G[1, 1] = -1
    h = np.concatenate(([0.0, 0.0], _svec(H @ H.T + np.eye(side))))

    def check(s, z, approx):
        assert approx(s[1] * np.linalg.slogdet(_smat(s[2:] / s[1]))[1], s[0])
        assert approx(z[0] * (np.linalg.slogdet(_smat(-z[2:] / z[0]))[1] + side), z[1])
    return _m([-1, 0], [[0.0, 1.0]], [1], G, h, [M.HypoPerLogdetTri(dim)]), \
        dict(status="Optimal", x_idx={1: 1.0}, check=check)


def hypoperlogdettri2():  # :1829-1859 (real case, dual barrier)
    side = 2
    H = np.random.default_rng(1).random((side, side))
    dim = 2 + M.svec_length(side)
    G = np.zeros((dim, 2))
    G[0, 0] = G[1, 1] = -1
    h = np.concatenate(([0.0, 0.0], _svec(H @ H.T)))

    def check(s, z, approx):
        assert approx(s[0] * (np.linalg.slogdet(_smat(-s[2:] / s[0]))[1] + side), s[1])
        assert approx(z[1] * np.linalg.slogdet(_smat(z[2:] / z[1]))[1], z[0])
    return _m([0, 1], [[1.0, 0.0]], [-1], G, h, [M.HypoPerLogdetTri(dim, use_dual=True)]), \
        dict(status="Optimal", x_idx={0: -1.0}, check=check)


def hypoperlogdettri3():  # :1861-1884 (real case): perspective variable forced to 0
    side = 3
    H = np.random.default_rng(1).random((side, side))
    dim = 2 + M.svec_length(side)
    G = np.zeros((dim, 2))
    G[0, 0] = G[1, 1] = -1
    h = np.concatenate(([0.0, 0.0], _svec(H @ H.T)))
    return _m([-1, 0], [[0.0, 1.0]], [0], G, h, [M.HypoPerLogdetTri(dim)]), dict(status="Optimal", x=[0.0, 0.0])


def _hypogeomean3(use_dual):  # :1498-1515
    l = 4
    G = np.vstack((np.zeros((1, l)), -np.eye(l)))
    return _m(np.ones(l), None, None, G, np.zeros(l + 1), [M.HypoGeoMean(l + 1, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=0, x=np.zeros(l))


def _hypopowermean3(use_dual):  # :1387-1405
    l = 4
    G = np.vstack((np.zeros((1, l)), -np.eye(l)))
    return _m(np.ones(l), None, None, G, np.zeros(l + 1), [M.HypoPowerMean(np.full(l, 1 / l), use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=0, x=np.zeros(l))


EXTRA = [consistent1, inconsistent1, inconsistent2, nonnegative1, nonnegative2, nonnegative3, possemideftri3,
         possemideftri4, hyporootdettri1, hyporootdettri2, hyporootdettri3, hypoperlogdettri1, hypoperlogdettri2,
         hypoperlogdettri3] + \
    [_named(lambda f=_f, ud=_ud: f(ud), _f.__name__[1:] + ("_dual" if _ud else ""))
     for _f in (_hypogeomean3, _hypopowermean3) for _ud in (False, True)]

def _matrixepipersquare1(d1, d2, use_dual):  # :1157-1195 (real case): U = I, minimise v
    W = np.random.default_rng(d1 * 10 + d2).random((d1, d2))
    per = d1 * (d1 + 1) // 2
    dim = per + 1 + d1 * d2
    G = np.zeros((dim, 1))
    G[per, 0] = -1
    h = np.zeros(dim)
    h[:per] = _svec(np.eye(d1))
    h[per + 1:] = W.ravel(order="F")
    WWt = W @ W.T
    epi = np.trace(WWt) / 2 if use_dual else np.linalg.eigvalsh(WWt)[-1] / 2
    return _m([1], None, None, G, h, [M.MatrixEpiPerSquare(d1, d2, use_dual=use_dual)]), \
        dict(status="Optimal", primal_obj=epi, s_idx={per: epi}, z_idx={per: 1.0})


def _matrixepipersquare2(use_dual):  # :1197-1231 (real case; tol 100 x)
    rng = np.random.default_rng(1)
    d1, d2 = 2, 3
    per = d1 * (d1 + 1) // 2
    dim = per + 1 + d1 * d2
    G = np.zeros((dim, 1))
    G[per, 0] = -1
    Uh = rng.random((d1, d1))
    U = Uh @ Uh.T
    W = rng.random((d1, d2))
    h = np.zeros(dim)
    h[:per] = _svec(U)
    h[per + 1:] = W.ravel(order="F")

    def check(s, z, approx):
        if use_dual:
            assert approx(2 * s[per], np.trace(W.T @ np.linalg.solve(U, W)))
        else:
            assert approx(np.linalg.eigvalsh(2 * s[per] * U - W @ W.T)[0], 0.0)
    return _m([1], None, None, G, h, [M.MatrixEpiPerSquare(d1, d2, use_dual=use_dual)]), \
        dict(status="Optimal", tol_scale=100, check=check)


def _matrixepipersquare3(use_dual):  # :1233-1259 (real case; tol 5 x, 2 x for the norm)
    rng = np.random.default_rng(1)
    d1, d2 = 3, 4
    per = d1 * (d1 + 1) // 2
    wd = d1 * d2
    dim = per + 1 + wd
    G = np.vstack((np.zeros((per + 1, wd)), -10.0 * np.eye(wd)))
    Uh = rng.random((d1, d1))
    h = np.zeros(dim)
    h[:per] = _svec(Uh @ Uh.T)
    return _m(np.ones(wd), None, None, G, h, [M.MatrixEpiPerSquare(d1, d2, use_dual=use_dual)]), \
        dict(status="Optimal", x=np.zeros(wd), tol_scale=10)


MEPS = [_named(lambda a=_a, b=_b, ud=_ud: _matrixepipersquare1(a, b, ud), f"matrixepipersquare1_{_a}x{_b}" + ("_dual" if _ud else ""))
        for (_a, _b) in ((1, 1), (1, 3), (2, 2), (2, 3)) for _ud in (False, True)] + \
    [_named(lambda f=_f, ud=_ud: f(ud), _f.__name__[1:] + ("_dual" if _ud else ""))
     for _f in (_matrixepipersquare2, _matrixepipersquare3) for _ud in (False, True)]


def _mlog(X):
    lam, Q = np.linalg.eigh(X)
    return (Q * np.log(lam)) @ Q.T


def epitrrelentropytri1():  # :2200-2223: min u : (u, V, W) in K  =>  tr(W (log W - log V))
    rng = np.random.default_rng(1)
    side = 4
    sv = M.svec_length(side)
    dim = 2 * sv + 1
    W = rng.random((side, side))
    W = W @ W.T
    V = rng.random((side, side))
    V = V @ V.T
    G = np.zeros((dim, 1))
    G[0, 0] = -1
    h = np.concatenate(([0.0], _svec(V), _svec(W)))
    return _m([1], None, None, G, h, [M.EpiTrRelEntropyTri(dim)]), \
        dict(status="Optimal", primal_obj=float(np.sum(W * (_mlog(W) - _mlog(V)))))


def epitrrelentropytri3():  # :2247-2265
    side = 3
    sv = M.svec_length(side)
    dim = 2 * sv + 1
    c = np.concatenate((np.zeros(sv + 1), np.ones(sv)))
    A = np.concatenate(([1.0], np.zeros(2 * sv)))[None, :]
    return _m(c, A, [0], -np.eye(dim), np.zeros(dim), [M.EpiTrRelEntropyTri(dim)]), \
        dict(status="Optimal", primal_obj=0, s_idx={0: 0.0})


def epitrrelentropytri4():  # :2267-2284
    side = 3
    sv = M.svec_length(side)
    dim = 2 * sv + 1
    c = np.concatenate(([0.0], np.ones(sv), np.zeros(sv)))
    A = np.concatenate(([1.0], np.zeros(2 * sv)))[None, :]
    return _m(c, A, [0], -np.eye(dim), np.zeros(dim), [M.EpiTrRelEntropyTri(dim)]), \
        dict(status="Optimal", primal_obj=0, s=np.zeros(dim))


TRRELENT = [epitrrelentropytri1, epitrrelentropytri3, epitrrelentropytri4]


def possemideftrisparse1():  # :543-563
    return _m([0, -1, 0], [[1, 0, 0], [0, 0, 1]], [0.5, 1], -np.eye(3), np.zeros(3),
              [M.PosSemidefTriSparse(2, [0, 1, 1], [0, 0, 1])]), dict(status="Optimal", primal_obj=-1, x_idx={1: 1.0})


def possemideftrisparse4():  # :631-661
    rt2, rt3 = np.sqrt(2.0), np.sqrt(3.0)
    G = np.zeros((10, 1))
    G[[0, 1, 2, 5, 9], 0] = -1
    h = np.zeros(10)
    h[[3, 4, 6, 7, 8]] = rt2 * np.array([1, 1, 1, -1, 1.0])
    rows = np.array([1, 2, 3, 4, 4, 4, 5, 5, 5, 5]) - 1
    cols = np.array([1, 2, 3, 1, 2, 4, 1, 2, 3, 5]) - 1
    return _m([1], None, None, G, h, [M.PosSemidefTriSparse(5, rows, cols)]), \
        dict(status="Optimal", primal_obj=rt3, s=[rt3, rt3, rt3, rt2, rt2, rt3, rt2, -rt2, rt2, rt3])


PSDSPARSE = [possemideftrisparse1, possemideftrisparse4]


def doublynonnegativetri1():  # :493-511 (the reference's loop overrides use_dual to false)
    return _m([0, 1, 0], [[1, 0, 0], [0, 0, 1]], [1, 1], -np.eye(3), np.zeros(3), [M.DoublyNonnegativeTri(3)]), \
        dict(status="Optimal", primal_obj=0, x=[1, 0, 1], s=[1, 0, 1])


def doublynonnegativetri2():  # :513-526
    return _m([0, -1, 0], [[1, 0, 0], [0, 0, 1]], [1.0, 1.5], -np.eye(3), [-0.5, 0, -0.5], [M.DoublyNonnegativeTri(3)]), \
        dict(status="Optimal", primal_obj=-1, x_idx={1: 1.0})


DNN = [doublynonnegativetri1, doublynonnegativetri2]


def _linmatrixineq1(side):  # :696-719 (real case): min w_1 : w_1 A_1 - 2 v v' psd  =>  2 / lambda_max(A_1)
    H = np.random.default_rng(side).random((side, side))
    A1 = H @ H.T + 2 * np.eye(side)
    vals, vecs = np.linalg.eigh(A1)
    v = vecs[:, -1]
    G = np.zeros((2, 1))
    G[0, 0] = -1
    return _m([1], None, None, G, [0, 2], [M.LinMatrixIneq([A1, -np.outer(v, v)])]), \
        dict(status="Optimal", primal_obj=2 / vals[-1], s=[2 / vals[-1], 2])


def _linmatrixineq2(dim):  # :721-745 (real case)
    rng = np.random.default_rng(1)
    As = []
    for _ in range(dim):
        H = rng.random((3, 3))
        As.append(H @ H.T)
    As[0] = As[0] + np.eye(3)
    G = np.vstack((np.zeros((1, dim - 1)), -np.eye(dim - 1)))
    h = np.zeros(dim)
    h[0] = 1

    def check(s, z, approx):
        assert float(np.sum(s[1:])) < 0       # x = s[1:] and c = 1: the reference asserts primal_obj < 0
    return _m(np.ones(dim - 1), None, None, G, h, [M.LinMatrixIneq(As)]), dict(status="Optimal", check=check)


def linmatrixineq3():  # :747-788 (dense case): min w_1 : w_1 I - diag(1, -1) psd => 1
    G = np.zeros((2, 1))
    G[0, 0] = -1
    return _m([1], None, None, G, [0, -1], [M.LinMatrixIneq([np.eye(2), np.diag([1.0, -1.0])])]), \
        dict(status="Optimal", primal_obj=1, s=[1, -1])


LMI = [_named(lambda s=_s: _linmatrixineq1(s), f"linmatrixineq1_side{_s}") for _s in (2, 4)] + \
    [_named(lambda d=_d: _linmatrixineq2(d), f"linmatrixineq2_dim{_d}") for _d in (2, 3)] + [linmatrixineq3]
EXTRA = EXTRA + LMI + DNN + MEPS + WSOSPSD + WSOSEUCL + WSOSONE + PSDSPARSE + TRRELENT

RELENT = [_named(lambda d=_d: _epirelentropy1(d), f"epirelentropy1_d{_d}") for _d in (1, 2, 3)] + \
    [_named(lambda d=_d: _epirelentropy2(d), f"epirelentropy2_d{_d}") for _d in (1, 2, 4)] + \
    [_named(lambda d=_d: _epirelentropy3(d), f"epirelentropy3_d{_d}") for _d in (2, 4)] + \
    [epirelentropy4, epirelentropy5]

GPOW = []
for _f in (_generalizedpower1, _generalizedpower2, _generalizedpower3, _generalizedpower4):
    for _ud in (False, True):
        GPOW.append(_named(lambda f=_f, ud=_ud: f(ud), _f.__name__[1:] + ("_dual" if _ud else "")))

HPM = [_named(lambda f=_f, ud=_ud: f(ud), _f.__name__[1:] + ("_dual" if _ud else ""))
       for _f in (_hypopowermean1, _hypopowermean2) for _ud in (False, True)] + \
    [hypopowermean4, hypopowermean5, hypopowermean6]

NEW_CONES = GPOW + HPM + RELENT + NORMSPEC + WSOS + [hypogeomean1, hypogeomean1_dual, hypogeomean2, hypogeomean2_dual, hypogeomean4, hypogeomean5,
             hypogeomean6, epinorminf1, epinorminf2, epinorminf3, epinorminf3_dual, epinorminf4, dualinfeas1,
             primalinfeas3, dualinfeas2, dualinfeas3, epipersquare1, epipersquare2, epipersquare3,
             epipersquare4, hypoperlog1, hypoperlog2, hypoperlog3, hypoperlog4, hypoperlog5, hypoperlog6,
             hypoperlog7]

ALL = [dimension1, nonnegative4, possemideftri1, possemideftri2, possemideftri8, possemideftri9,
       epinormeucl1, epinormeucl2, epinormeucl3, hyporootdettri4, hypoperlogdettri4,
       primalinfeas1, primalinfeas2, dualinfeas_lp]

TOL = TOL_EPS4


def _approx(a, b, tol):
    a, b = np.atleast_1d(np.asarray(a, float)), np.atleast_1d(np.asarray(b, float))
    return np.linalg.norm(a - b) <= max(tol, tol * max(np.linalg.norm(a), np.linalg.norm(b)))


def check_solution(solver, model, expected, tol=TOL):
    """Certificate checks of build_solve_check (nativeinstances.jl:32-86) + pinned values."""
    tol = tol * expected.get("tol_scale", 1)
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
    for i, v in expected.get("z_idx", {}).items():
        assert _approx(z[i], v, tol)
    for i, v in expected.get("s_idx", {}).items():
        assert _approx(s[i], v, tol)
    if "check" in expected:       # instance-specific assertions on (s, z), e.g. singular values
        expected["check"](s, z, lambda a, b: _approx(a, b, tol))

"""The oracle's OWN restatement of the reference's flat point layout and of the per-cone loops the
reference's callers write (oracle; test infrastructure).

Deliberately independent of hypatia_b200.host.point / host.coneblock, so that a layout bug on the
product side cannot hide in the checker: tests hand the product's Points to these solvers and the
two layouts must agree entry by entry (tests/test_oracle_solver.py::test_layouts_agree).

reference: src/Solvers/point.jl:24-54 (vec = [x(n); y(p); z(q); tau; s(q); kap], views, ztsk tail),
src/Solvers/systemsolvers/common.jl:184-208 (the (x, y, z) sub-points of the 3x3 system).
"""
import numpy as np


class OraclePoint:
    def __init__(self, n, p, q):
        self.n, self.p, self.q = int(n), int(p), int(q)
        self.vec = np.zeros(self.n + self.p + 2 * self.q + 2)
        o = 0
        self.x = self.vec[o:o + self.n]
        o += self.n
        self.y = self.vec[o:o + self.p]
        o += self.p
        self.ztsk = self.vec[o:]
        self.z = self.vec[o:o + self.q]
        o += self.q
        self.tau_idx = o
        o += 1
        self.s = self.vec[o:o + self.q]
        o += self.q
        assert o == self.vec.size - 1

    tau = property(lambda self: float(self.vec[self.tau_idx]),
                   lambda self, v: self.vec.__setitem__(self.tau_idx, v))
    kap = property(lambda self: float(self.vec[-1]),
                   lambda self, v: self.vec.__setitem__(self.vec.size - 1, v))

    def primal_dual(self, dual_mask):
        """point.jl:46-51: the primal view of a dual-barrier cone is its z block."""
        if dual_mask is None or not np.any(dual_mask):
            return self.s, self.z
        prim, dual = self.s.copy(), self.z.copy()
        prim[dual_mask] = self.z[dual_mask]
        dual[dual_mask] = self.s[dual_mask]
        return prim, dual


class OracleSubPoint:
    def __init__(self, n, p, q):
        self.n, self.p, self.q = int(n), int(p), int(q)
        self.vec = np.zeros(self.n + self.p + self.q)
        self.x = self.vec[:self.n]
        self.y = self.vec[self.n:self.n + self.p]
        self.z = self.vec[self.n + self.p:]


class OracleConeBlockBase:
    """What the reference's `for k in eachindex(cones)` loops need to know about the cone list."""

    def __init__(self, model):
        self.specs = list(model.cones)
        self.K = len(self.specs)
        self.q = int(model.q)
        self.dims = np.array([int(ck.dim) for ck in self.specs], dtype=np.int64)
        self.offsets = np.zeros(self.K, dtype=np.int64)
        if self.K:
            self.offsets[1:] = np.cumsum(self.dims)[:-1]
        self.nus = np.array([float(ck.nu) for ck in self.specs])
        mask = np.zeros(self.q, dtype=bool)
        for ck, o, d in zip(self.specs, self.offsets, self.dims):
            if ck.use_dual:
                mask[o:o + d] = True
        self.dual_mask = mask if mask.any() else None

    def seg_dot(self, a, b):
        return np.array([float(a[o:o + d] @ b[o:o + d]) for o, d in zip(self.offsets, self.dims)])

    def expand(self, per_cone):
        return np.repeat(np.asarray(per_cone), self.dims)

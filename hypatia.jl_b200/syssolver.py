"""Device system solver behind the reference's SystemSolver API (plugin slot 1).

`QRCholDenseSystemSolver` here is the B200 drop-in for the reference type of the same name
(src/Solvers/systemsolvers/qrchol.jl:104-257): same methods - load, update_lhs, solve_system,
solve_subsystem3, apply_lhs, free_memory - same arguments (`solver`, `sol::Point`, `rhs::Point`),
same non-throwing behaviour on factorisation failure (prints the reference's message and
continues, qrchol.jl:252-254).  All arithmetic happens in libhypatia_b200 through the C ABI.

Multi-GPU: with `torch.distributed` initialised (one process per GPU) the cones - and with them
the row panels of G - are partitioned over ranks by `partition_cones`; each rank uploads only its
panel and the library sums the partial Schur matrices with one NCCL allreduce per iteration.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .cones import DeviceConeBlock


def cone_work(spec, m):
    """Per-iteration cost model of one cone block used for balancing ranks: SYRK flops q_k m^2
    plus the oracle's H^{1/2} G_k cost (2 side^3 per column for matrix cones)."""
    w = float(spec.dim) * m * m
    if spec.side:
        w += 4.0 * spec.side ** 3 * m
    return w


def partition_cones(model, nranks):
    """Contiguous cone ranges [lo, hi) per rank balancing `cone_work` (never splits a cone;
    SURVEY.md 8(e)).  Returns a list of (lo, hi) of length nranks."""
    K = len(model.cones)
    m = max(model.n - model.p, 1)
    w = np.array([cone_work(ck, m) for ck in model.cones])
    total = w.sum()
    bounds = [0]
    acc = 0.0
    k = 0
    for r in range(1, nranks):
        target = total * r / nranks
        while k < K and acc + w[k] / 2 <= target:
            acc += w[k]
            k += 1
        bounds.append(k)
    bounds.append(K)
    return [(bounds[r], bounds[r + 1]) for r in range(nranks)]


def giant_cone(model, frac: float = 0.5) -> bool:
    """True when one cone carries more than `frac` of the Schur-assembly work: whole-cone (row-panel) sharding cannot
    split it, so the assembly is sharded by COLUMNS of G_k instead (hess_prod! is independent per
    column, hypoperlogdettri.jl:196-237)."""
    m = model.n - model.p
    w = [cone_work(c, m) for c in model.cones]
    return bool(w) and max(w) > frac * sum(w)


def column_ranges(nmp: int, nranks: int):
    """Column panel [lo, hi) of the Schur matrix each rank assembles under column sharding (equal widths, rounded up
    to an even number of columns: ncclAllGather needs equal counts) - mirrors hyp_load_model."""
    cw = -(-max(nmp, 1) // nranks)
    cw += cw & 1
    return [(min(nmp, r * cw), min(nmp, (r + 1) * cw)) for r in range(nranks)]


class QRCholDenseSystemSolver:
    NEEDS_QR = True
    FAIL_MESSAGE = "positive definite linear system factorization failed"    # qrchol.jl:252-254

    def _after_load(self):
        pass

    def __init__(self, device: int | None = None, dist_group=None, device_residuals: bool = False,
                 column_sharding: bool | None = None):
        self.device = device
        # None: decide per model (giant_cone); True / False: force
        self.column_sharding = column_sharding
        self.dist_group = dist_group
        # True: the driver's calc_convergence_params takes its residuals from hyp_calc_residuals (two
        # passes over G on the device) instead of the host products the reference does (Solvers.jl:425-483)
        self.device_residuals = device_residuals
        self.ctx = None
        self.cones = None
        self.fact_kind = 0
        self.rank, self.nranks = 0, 1

    # ---- load(syssolver, solver): qrchol.jl:138-179 ----
    def load(self, solver):
        model = solver.model
        self._init_comm()
        if self.ctx is None:
            self.ctx = capi.Context(self.device if self.device is not None else 0)
            if self.nranks > 1:
                self._join_comm()
        self.col_shard = self.nranks > 1 and (giant_cone(model) if self.column_sharding is None
                                              else bool(self.column_sharding))
        if self.col_shard:
            self.ctx.set_column_sharding(True)      # every rank keeps all rows; only the assembly is split
        lo, hi = partition_cones(model, self.nranks)[self.rank] if (self.nranks > 1 and not self.col_shard) \
            else (0, len(model.cones))
        self.cone_range = (lo, hi)
        Q = getattr(solver, "Ap_Q", None)
        R = getattr(solver, "Ap_R", None)
        if model.p == 0 or not self.NEEDS_QR:
            Q = R = None
        elif Q is None:
            raise ValueError("QRCholDenseSystemSolver needs solver.Ap_Q / Ap_R when p > 0")
        self.ctx.load_model(model, cone_lo=lo, cone_hi=hi, Ap_Q=Q, Ap_R=R)
        self._after_load()
        self.cones = DeviceConeBlock(model, ctx=self.ctx)
        self.nmp = model.n - model.p
        return self

    def _init_comm(self):
        try:
            import torch.distributed as dist
        except ImportError:
            return
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(self.dist_group)
            self.nranks = dist.get_world_size(self.dist_group)
            if self.device is None:
                import os
                self.device = int(os.environ.get("LOCAL_RANK", self.rank))

    def _join_comm(self):
        import torch
        import torch.distributed as dist
        uid = [capi.comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=self.dist_group)
        self.ctx.comm_init(self.rank, self.nranks, uid[0])

    # ---- update_lhs(syssolver, solver): qrchol.jl:181-257 ----
    def update_lhs(self, solver):
        self.ctx.set_mu_tau(solver.mu, solver.point.tau)
        rc, kind = self.ctx.update_lhs()
        self.fact_kind = kind
        if rc == 2:
            print(self.FAIL_MESSAGE)
        return self

    # ---- solve_system / solve_subsystem3 / apply_lhs: common.jl:129-151, qrchol.jl:39-85,
    #      common.jl:79-121 ----
    def solve_system(self, solver, sol, rhs):
        self.ctx.solve_system(sol.vec, rhs.vec)
        return sol

    def solve_subsystem3(self, solver, sol, rhs):
        self.ctx.solve_subsystem3(sol.vec, rhs.vec)
        return sol

    def apply_lhs(self, solver, direction, res):
        self.ctx.set_mu_tau(solver.mu, solver.point.tau)
        self.ctx.apply_lhs(res.vec, direction.vec)
        return res

    # ---- residual step next to the path (Solvers.jl:425-483) ----
    def calc_residuals(self, solver):
        """x / y / z residual vectors and the ten norms / inner products of calc_convergence_params,
        with the two passes over G on the device."""
        return self.ctx.calc_residuals(solver.point.vec)

    def lhs_full(self):
        """Symmetric Schur matrix (test helper)."""
        S = self.ctx.get_schur()
        U = np.triu(S)
        return U + np.triu(S, 1).T

    # ---- free_memory(syssolver): Solvers.jl:582-584 ----
    def free_memory(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None


class SymIndefDenseSystemSolver(QRCholDenseSystemSolver):
    """Device drop-in for the reference's SymIndefDenseSystemSolver (symindef.jl:203-271): dense
    (n+p+q)^2 symmetric-indefinite LHS [0 A' G'; A 0 0; G 0 -Hinv], rook Bunch-Kaufman with the
    symm_fact_copy! diagonal-shift retry (dense.jl:170-184).  Same C entry points; the context is
    switched with hyp_set_syssolver(ctx, 1).  Single GPU."""
    NEEDS_QR = False
    FAIL_MESSAGE = "symmetric linear system factorization failed"             # symindef.jl:254-256

    def _after_load(self):
        self.ctx.set_syssolver(1)

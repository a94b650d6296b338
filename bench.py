#!/usr/bin/env python
"""bench.py - IPM iterations/sec of the Hypatia KKT hot path on B200 (BASELINE.json metric).

One "step" = one IPM-iteration unit of the reference's CombinedStepper with zero refinement
rounds and the line search excluded (reference timers time_upsys + time_getdir,
src/Solvers/steppers/combined.jl:64-79):

    load_point of every cone at s/sqrt(mu)        (hyp_cones_load_point)
    update_lhs: H^{1/2}G pre-pass, Schur SYRK, Cholesky, constant-column solve   (hyp_update_lhs)
    4 x { solve_system ; apply_lhs residual }      (hyp_solve_system, hyp_apply_lhs)

on the synthetic dense conic instance BASELINE.json's metric is quoted on (default workload "C3":
n = 10000, p = 0, 2000 x EpiNormEucl(25), q = 50000).  The four right-hand sides are the real
cent / centadj / pred / predadj right-hand sides of the first iterate, built once before timing.

  value : steps/s with every input resident in HBM (device pointers through the C ABI)
  e2e   : the same calls with HOST (pinned) buffers: the H2D copies of the point and the four
          right-hand sides and the D2H copies of the directions / residuals are inside the timing
  roofline : the Schur SYRK kernel (tcgen05 int8 digit slicing); frac = executed int8 TOP/s over
             2 x bf16_tflops_sustained of MEASURED_PEAKS.json; the FP64-equivalent rate is a side field
  cpu_baseline : the CPU oracle restatement (OpenBLAS dsyrk/dpotrf/dgemv, all host cores), rank 0, N = 1
  parity : at EVERY N, rank 0 solves the same four right-hand sides with the CPU oracle
           (dir_vs_oracle) and applies the ORACLE's 6x6 operator to the device directions (kkt_residual)
  full_step : wall time of one complete CombinedStepper.step (refinement rounds + line-search sweeps)
  batched_solves : the same unit with the independent right-hand sides solved by hyp_solve_system_multi
  other_workloads : C2 / C4 / C5a at full size (value, phase_ms, parity), --other to select

Launch: `python bench.py --gpus 1 --steps K --warmup W`, or under torchrun for N > 1 (one rank per
GPU, cones / G row panels sharded over ranks, one NCCL reduction of the Schur matrix per step;
"scaling": "strong" - the instance is fixed).  `--impl reference` times the CPU oracle instead.
"""
from __future__ import annotations

import os
import sys

# torch.distributed.run exports OMP_NUM_THREADS=1 when nproc > 1; rank 0 runs the CPU oracle (parity at every
# N, and the whole `--impl reference` arm), which must see all host cores: undo it BEFORE NumPy / OpenBLAS load
if int(os.environ.get("LOCAL_RANK", "0")) == 0:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        if os.environ.get(_v) == "1":
            del os.environ[_v]

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEFAULT_OTHER = "C2,C4,C5a"
METRIC = "IPM iterations/sec (KKT assemble+factor+solve)"
UNIT = "iter/s"

WORKLOADS = {
    # name: (n, cone builder)
    "C2": dict(n=4000, cones=lambda M: [M.Nonnegative(5000)],
               desc="dense LP after reduction: m=4000, Nonnegative(5000)"),
    "C3": dict(n=10000, cones=lambda M: [M.EpiNormEucl(25) for _ in range(2000)],
               desc="n=10000 p=0, 2000 x EpiNormEucl(25), q=50000"),
    "C3s": dict(n=1000, cones=lambda M: [M.EpiNormEucl(25) for _ in range(200)],
                desc="n=1000 p=0, 200 x EpiNormEucl(25), q=5000 (smoke-size)"),
    "C4": dict(n=20000, cones=lambda M: [M.PosSemidefTri(5050) for _ in range(50)],
               desc="n=20000 p=0, 50 x PosSemidefTri(side 100), q=252500"),
    "C4s": dict(n=600, cones=lambda M: [M.PosSemidefTri(M.svec_length(16)) for _ in range(6)],
                desc="n=600 p=0, 6 x PosSemidefTri(side 16), q=816 (smoke-size)"),
    # natvsext-faithful shape of BASELINE config 5 (SURVEY.md 8(d) "C5a"): the largest `nat` log-det
    # D-optimal-design instance (examples/doptimaldesign/JuMP_benchmark.jl:2-5, native.jl:18-86) after
    # the QR reduction: one HypoPerLogdetTri of side 1000 (dim 500502 > m, so the hess_prod! + GEMM
    # branch qrchol.jl:240-246) + two Nonnegative(2000); dense Gaussian G like the other configs
    "C5a": dict(n=2000, cones=lambda M: [M.Nonnegative(2000), M.Nonnegative(2000),
                                         M.HypoPerLogdetTri(2 + M.svec_length(1000))],
                desc="n=2000 p=0, Nonnegative(2000) x 2 + HypoPerLogdetTri(side 1000), q=504502"),
    "C5as": dict(n=120, cones=lambda M: [M.Nonnegative(100), M.Nonnegative(100),
                                          M.HypoPerLogdetTri(2 + M.svec_length(40))],
                 desc="n=120 p=0, Nonnegative(100) x 2 + HypoPerLogdetTri(side 40), q=1022 (smoke-size)"),
    # many-cone mix of BASELINE config 5 (SURVEY.md 8(d) "C5b"): 40 x HypoPerLogdetTri(side 100) +
    # 10000 x EpiNormEucl(25) + Nonnegative(47920), n = 20000, q = 500000
    "C5b": dict(n=20000, cones=lambda M: [M.HypoPerLogdetTri(2 + M.svec_length(100)) for _ in range(40)] +
                [M.EpiNormEucl(25) for _ in range(10000)] + [M.Nonnegative(47920)],
                desc="n=20000 p=0, 40 x HypoPerLogdetTri(side 100) + 10000 x EpiNormEucl(25) + Nonnegative(47920), q=500000"),
    # widening rows (not BASELINE configs): spectral cones through the batched Jacobi eigensolver
    "S1": dict(n=2000, cones=lambda M: [M.EpiPerSepSpectralMat(2 + M.svec_length(100), M.SSF_NEGENTROPY)
                                        for _ in range(24)],
               desc="n=2000 p=0, 24 x EpiPerSepSpectral{MatrixCSqr}(NegEntropy, side 100), q=121248"),
}

def syrk_kernel_name():
    """Name of the digit-sliced SYRK kernel the library launches under the current environment (ozaki.cu, hyp_ozaki_syrk)."""
    c = os.environ.get("HYP_OZAKI_CLUSTER", "")[:1]
    r128 = os.environ.get("HYP_OZAKI_RADIX") == "128"
    if c in ("0", "1") or r128:
        return {"0": "ozaki_syrk_kernel", "1": "ozaki_syrk_cluster_kernel"}.get(c, "ozaki_syrk_pair_kernel")
    return {"2": "ozaki_syrk_pair_kernel", "3": "ozaki_syrk_quad_kernel", "5": "ozaki_syrk_quad64_kernel"}.get(
        c, "ozaki_syrk_pair64_kernel (64-byte k rows, SWIZZLE_64B)")


# DRAM traffic of the Schur SYRK (dram__bytes_read.sum + dram__bytes_write.sum over the launches of ONE SYRK) from the
# committed `ncu --set full` capture of the same kernel build - a CONSTANT taken from that profile, not measured by
# the run that prints it (ncu cannot run inside a timed bench); the source file is named next to it
NCU_TRAFFIC = {"C3:i8": ((17.548279e9 + 0.805614e9) + (18.609838e9 + 0.805811e9) + (18.045313e9 + 0.805962e9),
                         "constant from profiles/r02_ozaki_pair64_tile_order_ncu.md (ncu metrics pass over the three K-chunk "
                         "launches of one SYRK, ozaki_syrk_pair64_kernel with the default tile order of six resident row pairs; "
                         "only quoted when that kernel and that order run)"),
               "C3:i8:row1": ((38.071887e9 + 0.806117e9) + (38.576262e9 + 0.806978e9) + (38.530823e9 + 0.806765e9),
                              "constant from profiles/r02_ozaki_pair64_tile_order_ncu.md (HYP_OZAKI_ORDER=row: one resident row pair)"),
               "C3": (31.497622e9 + 0.411380e9, "constant from profiles/r01_syrk_schur_ncu_full.txt (one launch)")}


class PanelModel:
    """Model fields the C ABI needs, holding only this rank's row panel of G."""

    def __init__(self, n, cones, c, h):
        self.n, self.p = n, 0
        self.cones = cones
        dims = np.array([ck.dim for ck in cones], dtype=np.int64)
        self.cone_dims = dims
        self.cone_offsets = np.concatenate(([0], np.cumsum(dims)))[:-1].astype(np.int64)
        self.q = int(dims.sum())
        self.cone_idxs = [slice(int(o), int(o + d)) for o, d in zip(self.cone_offsets, dims)]
        self.cone_nus = np.array([ck.nu for ck in cones])
        self.nu = float(self.cone_nus.sum())
        self.c, self.h = c, h
        self.b = np.zeros(0)
        self.A = np.zeros((0, n))
        self.G = None


def gen_panel(n, row_lo, row_hi, seed, block=2048):
    """Rows [row_lo, row_hi) of the N(0,1) matrix G; block-seeded so that any rank generates the
    same global matrix."""
    G = np.empty((row_hi - row_lo, n), order="F")
    b0, b1 = row_lo // block, (row_hi + block - 1) // block
    for b in range(b0, b1):
        rng = np.random.Generator(np.random.PCG64([seed, b]))
        blk = rng.standard_normal((block, n))
        lo, hi = max(row_lo, b * block), min(row_hi, (b + 1) * block)
        G[lo - row_lo:hi - row_lo] = blk[lo - b * block:hi - b * block]
    return G


def gen_panel_device(torch, device, n, row_lo, row_hi, seed, block=4096):
    """The same idea on the device for the multi-GB workloads (C4: 40 GB): rows [row_lo, row_hi) of a
    block-seeded N(0,1) matrix, returned as the (n, rows) row-major tensor = the column-major panel."""
    rows = row_hi - row_lo
    GT = torch.empty((n, rows), dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    b0, b1 = row_lo // block, (row_hi + block - 1) // block
    for b in range(b0, b1):
        gen.manual_seed(seed * 1000003 + b)
        blk = torch.randn((n, block), dtype=torch.float64, device=device, generator=gen)
        lo, hi = max(row_lo, b * block), min(row_hi, (b + 1) * block)
        GT[:, lo - row_lo:hi - row_lo] = blk[:, lo - b * block:hi - b * block]
        del blk
    return GT


def _planted_point(cones, model, seed):
    """Planted interior primal-dual pair: cone central points perturbed like
    test/cone.jl:236-248.  Returns s0, z0 and the svec off-diagonal rows of G that carry sqrt(2)."""
    from hypatia_b200.host import instances as inst
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    q = model.q
    s0, z0 = np.empty(q), np.empty(q)
    rt2_rows = []
    for ck, sl in zip(cones, model.cone_idxs):
        prim = inst.cone_initial_point(ck)
        dual = inst._cone_dual_initial(ck, prim)
        s0[sl] = inst._perturb(rng, ck, prim, 0.1)
        z0[sl] = inst._perturb(rng, ck, dual, 0.1)
        if ck.side:
            rt2_rows.append(sl.start + inst.mat_offset(ck) + np.nonzero(inst._svec_offdiag_mask(ck.side))[0])
    rt2_rows = np.concatenate(rt2_rows) if rt2_rows else np.zeros(0, dtype=np.int64)
    return s0, z0, rt2_rows, rng


def build_instance(workload, rank, nranks, dist=None, device=None, on_device=False, replicate_rows=False):
    """This rank's share of the synthetic instance.  on_device: G is generated on the GPU (torch) and stays
    there (`G_dev`, the (n, rows) tensor); otherwise NumPy on the host (`G_local`).  replicate_rows: every rank
    holds ALL rows (single-giant-cone workloads that shard by COLUMNS of G_k)."""
    from hypatia_b200.host import models as M
    from hypatia_b200.syssolver import partition_cones
    w = WORKLOADS[workload]
    n = w["n"]
    cones = w["cones"](M)
    seed = 1000 + sorted(WORKLOADS).index(workload)
    model = PanelModel(n, cones, None, None)
    q = model.q
    K = len(cones)
    if replicate_rows or nranks == 1:
        lo, hi = 0, K
    else:
        lo, hi = partition_cones(model, nranks)[rank]
    row_lo = int(model.cone_offsets[lo]) if lo < K else q
    row_hi = int(model.cone_offsets[hi]) if hi < K else q
    s0, z0, rt2_rows, rng = _planted_point(cones, model, seed)
    x0 = rng.standard_normal(n)
    rows = rt2_rows[(rt2_rows >= row_lo) & (rt2_rows < row_hi)] - row_lo
    h = np.zeros(q)
    G_local = G_dev = None
    if on_device:
        import torch
        G_dev = gen_panel_device(torch, device, n, row_lo, row_hi, seed)
        if rows.size:
            G_dev[:, torch.from_numpy(rows).to(device)] *= float(np.sqrt(2.0))
        tx = torch.from_numpy(x0).to(device)
        tz = torch.from_numpy(z0[row_lo:row_hi].copy()).to(device)
        h[row_lo:row_hi] = torch.mv(G_dev.t(), tx).cpu().numpy() + s0[row_lo:row_hi]
        c = -torch.mv(G_dev, tz).cpu().numpy()
    else:
        G_local = gen_panel(n, row_lo, row_hi, seed)
        if rows.size:
            G_local[rows] *= np.sqrt(2.0)
        h[row_lo:row_hi] = G_local @ x0 + s0[row_lo:row_hi]
        c = -(G_local.T @ z0[row_lo:row_hi])
    if nranks > 1 and not replicate_rows:
        import torch
        th = torch.from_numpy(h).to(device)
        tc = torch.from_numpy(c).to(device)
        dist.all_reduce(th)
        dist.all_reduce(tc)
        h, c = th.cpu().numpy(), tc.cpu().numpy()
    model.c, model.h = c, h
    mu = (float(z0 @ s0) + 1.0) / (model.nu + 1)
    return dict(model=model, G_local=G_local, G_dev=G_dev, s0=s0, z0=z0, x0=x0, mu=mu, cone_range=(lo, hi),
                rows=(row_lo, row_hi), desc=w["desc"])


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def dgemm_peak_tflops(torch, device):
    """FP64 tensor (DMMA) peak measured live with cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json has
    only HBM and bf16 entries); the roofline of the FP64 DMMA kernels (Cholesky trailing updates)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def use_all_cores():
    """BLAS threads = all host cores, whatever the launcher exported (torchrun: OMP_NUM_THREADS=1)."""
    try:
        import threadpoolctl
        n = os.cpu_count() or 1
        threadpoolctl.threadpool_limits(limits=n)
        return max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
def cpu_unit(workload, steps, warmup, budget_s=25.0, rhs_list=None, check_sols=None):
    """CPU oracle restatement of the same unit on the host cores.  Returns a dict: seconds per unit, sample
    description, cores, BLAS threads, directions of the last unit, and - when `check_sols` (device directions)
    is given - the residual of each one under the ORACLE's 6x6 operator, ||K_oracle d - r|| / ||r||.
    The SYRK is timed on a bounded row sample only when `steps + warmup` full units would exceed `budget_s`
    (its cost is linear in the rows; the factorisation then uses the Schur matrix of one full, untimed
    assembly), everything else always runs in full.  `rhs_list` replaces the oracle's own four right-hand
    sides (the parity check hands in the ones the device solved)."""
    from scipy.linalg import blas as _blas
    from hypatia_b200.host import models as M
    from oracle.cones import OracleConeBlock
    from oracle.bench_unit import iterate_shell
    from oracle.layout import OraclePoint
    cores = os.cpu_count() or 1
    blas_threads = use_all_cores()
    I = build_instance(workload, 0, 1)
    pm = I["model"]
    model = M.Model(pm.c, None, pm.b, I["G_local"], pm.h, pm.cones)
    q, n = model.q, model.n
    # estimate dsyrk speed
    rs = min(q, 4096)
    t0 = time.perf_counter()
    _blas.dsyrk(1.0, model.G[:rs], trans=1, lower=0)
    t_s = time.perf_counter() - t0
    est_full = t_s * q / rs
    frac = 1.0
    if est_full * (steps + warmup) > budget_s:
        frac = max(rs / q, min(1.0, budget_s / (est_full * (steps + warmup))))
    shell = iterate_shell(model, I["s0"], I["z0"], I["x0"], I["mu"], OracleConeBlock,
                          syrk_row_fraction=frac)
    if rhs_list is None:
        rhs_list = shell.rhs_list
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        extra = shell.unit(rhs_list)
        dt = time.perf_counter() - t0 + extra
        if it >= warmup:
            times.append(dt)
    sample = (f"{steps} full unit(s) of {workload}" if frac >= 1.0 else
              f"{steps} unit(s) of {workload}; dsyrk timed on the first {frac:.3f} of the rows and "
              f"scaled by 1/{frac:.3f}, Cholesky and all solves in full")
    kkt = None
    if check_sols is not None:
        d, r = (OraclePoint(n, model.p, q) for _ in range(2))
        kkt = []
        for v, rv in zip(check_sols, rhs_list):
            d.vec[:] = v
            shell.syssolver.apply_lhs(shell, d, r)
            kkt.append(float(np.linalg.norm(r.vec - rv) / max(np.linalg.norm(rv), 1e-300)))
    return dict(sec=float(np.median(times)), sample=sample, cores=cores, blas_threads=blas_threads,
                sols=shell.sols, kkt_oracle=kkt)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # full units (no dsyrk sampling) as long as the whole run stays within ~12 minutes of host time
    r = cpu_unit(args.workload, max(1, args.steps), min(args.warmup, 1), budget_s=720.0)
    v = 1.0 / r["sec"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload]['desc']}"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "blas_threads": r["blas_threads"],
                             "kind": "port", "sample": r["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
class DeviceShell:
    """What the host RHS builders / stepper (hypatia.jl_b200/host/stepper.py) read from `solver`, positioned at
    the planted first iterate, with the device plug-ins behind it."""

    def __init__(self, I, ctx, cones):
        from hypatia_b200.host.point import Point
        from hypatia_b200.syssolver import QRCholDenseSystemSolver
        model = I["model"]
        self.model, self.mu, self.cones = model, I["mu"], cones
        pt = self.point = Point(model)
        pt.x[:] = I["x0"]
        pt.z[:] = I["z0"]
        pt.s[:] = I["s0"]
        pt.tau = pt.kap = 1.0
        self.x_residual = np.zeros(model.n)
        self.y_residual = np.zeros(0)
        self.z_residual = np.zeros(model.q)
        self.tau_residual = float(model.c @ pt.x) + float(model.h @ pt.z) + pt.kap
        sysv = self.syssolver = QRCholDenseSystemSolver()
        sysv.ctx, sysv.cones = ctx, cones
        self.max_ref_steps = 5                        # Solvers.jl:264
        self.res_norm_cutoff = 1e-4 * abs(self.tau_residual)      # Solvers.jl:381-382 at this iterate
        self.time_upsys = self.time_uprhs = self.time_getdir = self.time_search = 0.0
        self.n_solve_system = self.n_apply_lhs = 0
        self.worst_dir_res = 0.0
        self.status = "SolveCalled"


def measure_workload(args, workload, torch, dist, device, rank, world, local_rank, full):
    """Times one workload on this process group.  full: the headline treatment (e2e pass with host buffers,
    clocks, CPU oracle parity, full step, batched solves); otherwise value + phases + an independent KKT
    residual only (other_workloads)."""
    from hypatia_b200 import capi
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.host.point import Point
    from hypatia_b200.host import stepper as st

    big = workload in ("C4", "C5a", "C5b") or (not full)
    from hypatia_b200.host import models as M_
    from hypatia_b200.syssolver import giant_cone
    # one cone carrying most of the assembly work cannot be split by whole-cone sharding: shard by columns instead
    giant = world > 1 and giant_cone(PanelModel(WORKLOADS[workload]["n"], WORKLOADS[workload]["cones"](M_), None, None))
    I = build_instance(workload, rank, world, dist, device, on_device=big, replicate_rows=giant)
    model = I["model"]
    n, q = model.n, model.q
    ctx = capi.Context(local_rank)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    lo, hi = I["cone_range"]
    if giant:
        ctx.set_column_sharding(True)
    ctx.load_model(model, G_local=I["G_dev"] if big else I["G_local"], cone_lo=lo, cone_hi=hi)
    # models with a cone that has no closed-form square root assemble S with the TWO-operand product P'(HG)
    # (qrchol.jl:240-246 branch); since round 2 the library runs that one on the int8 tensor pipe as well
    syrk = args.syrk
    ctx.set_syrk_mode(1 if syrk == "i8" else 0)
    # the library keeps its own copy of the panel: drop ours before update_lhs allocates its work buffers
    # (C4: G 40 GB + H^{1/2}G 40 GB + digit slices 40 GB + Schur / factor 6.4 GB of the 180 GB)
    I["G_local"] = I["G_dev"] = None
    if big:
        torch.cuda.empty_cache()
    cones = DeviceConeBlock(model, ctx=ctx)

    # ---- the four right-hand sides of the first iterate (built once, untimed) ----
    sh = DeviceShell(I, ctx, cones)
    pt = sh.point
    irtmu = 1.0 / np.sqrt(sh.mu)
    cones.load_point(pt.s, pt.z, irtmu)
    ctx.set_mu_tau(sh.mu, pt.tau)
    rc, kind = ctx.update_lhs()
    if rc != 0 or kind != 0:
        raise SystemExit(f"bench.py: Cholesky of the Schur complement failed (rc={rc}, kind={kind})")
    rhs, d = Point(model), Point(model)
    rhs_list = []
    st.update_rhs_cent(sh, rhs)
    rhs_list.append(rhs.vec.copy())
    ctx.solve_system(d.vec, rhs.vec)
    st.update_rhs_centadj(sh, rhs, d)
    rhs_list.append(rhs.vec.copy())
    st.update_rhs_pred(sh, rhs)
    rhs_list.append(rhs.vec.copy())
    ctx.solve_system(d.vec, rhs.vec)
    st.update_rhs_predadj(sh, rhs, d)
    rhs_list.append(rhs.vec.copy())
    dim6 = rhs.vec.size

    # device-resident and pinned-host copies of the inputs / outputs
    dev_in = dict(s=torch.from_numpy(pt.s.copy()).to(device), z=torch.from_numpy(pt.z.copy()).to(device),
                  rhs=[torch.from_numpy(r).to(device) for r in rhs_list])
    dev_out = dict(sol=[torch.empty(dim6, dtype=torch.float64, device=device) for _ in range(4)],
                   res=[torch.empty(dim6, dtype=torch.float64, device=device) for _ in range(4)])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host_in = dict(s=pin(pt.s), z=pin(pt.z), rhs=[pin(r) for r in rhs_list])
    host_out = dict(sol=[torch.empty(dim6, dtype=torch.float64).pin_memory() for _ in range(4)],
                    res=[torch.empty(dim6, dtype=torch.float64).pin_memory() for _ in range(4)])

    def step(inp, out):
        ctx.cones_load_point(inp["s"], inp["z"], irtmu)
        ctx.update_lhs()
        for i in range(4):
            ctx.solve_system(out["sol"][i], inp["rhs"][i])
            ctx.apply_lhs(out["res"][i], out["sol"][i])

    # batched variant: the data flow of combined.jl:67-79 allows {cent, pred} and {centadj, predadj} to share
    # one multi-column sweep each; a stepper-side change, hence reported separately
    dev_in["rhs2"] = [torch.stack([dev_in["rhs"][0], dev_in["rhs"][2]]).contiguous(),
                      torch.stack([dev_in["rhs"][1], dev_in["rhs"][3]]).contiguous()]
    dev_out["sol2"] = [torch.empty((2, dim6), dtype=torch.float64, device=device) for _ in range(2)]
    dev_out["res2"] = [torch.empty((2, dim6), dtype=torch.float64, device=device) for _ in range(2)]

    def step_batched(inp, out):
        ctx.cones_load_point(inp["s"], inp["z"], irtmu)
        ctx.update_lhs()
        for i in range(2):
            ctx.solve_system_multi(out["sol2"][i], inp["rhs2"][i], 2)
            ctx.apply_lhs_multi(out["res2"][i], out["sol2"][i], 2)

    ext = torch.cuda.ExternalStream(ctx.stream(), device=device)

    def timed(fn, inp, out, steps, warmup, sample_clocks):
        for _ in range(warmup):
            fn(inp, out)
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        for _ in range(steps):
            fn(inp, out)
        with torch.cuda.stream(ext):
            e1.record()
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0, clocks

    steps, warmup = (args.steps, args.warmup) if full else (args.other_steps, 1)
    ms, launches, clocks = timed(step, dev_in, dev_out, steps, warmup, full)
    out = {"workload": f"{workload}: {I['desc']}", "value": steps / (ms * 1e-3), "ms_per_step": ms / steps,
           "steps": steps, "gpu_launches": launches, "syrk": syrk, "clocks": clocks,
           "sharding": ("columns of the giant cone (all-gather)" if giant else "cones / row panels") if world > 1 else "none"}
    if full:
        ms_e2e, _, _ = timed(step, host_in, host_out, args.steps, 1, False)
        out["e2e"] = {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                      "h2d_bytes_per_step": 8 * (2 * q + 4 * dim6 + 4 * dim6),     # point + 4 rhs + 4 dirs (apply_lhs input)
                      "d2h_bytes_per_step": 8 * (4 * dim6 + 4 * dim6) + 4}        # 4 dirs + 4 residuals + Cholesky info word
        dev_sols = [t.numpy().copy() for t in host_out["sol"]]
        dev_res = [t.numpy().copy() for t in host_out["res"]]
    else:
        dev_sols = [t.cpu().numpy() for t in dev_out["sol"]]
        dev_res = [t.cpu().numpy() for t in dev_out["res"]]
    if ctx.has_multi():
        try:
            ms_b, launches_b, _ = timed(step_batched, dev_in, dev_out, steps, 1, False)
            solb = [dev_out["sol2"][0][0], dev_out["sol2"][1][0], dev_out["sol2"][0][1], dev_out["sol2"][1][1]]
            dmax = max(float(torch.linalg.norm(solb[i] - dev_out["sol"][i]) /
                             torch.linalg.norm(dev_out["sol"][i])) for i in range(4))
            out["batched_solves"] = {"value": steps / (ms_b * 1e-3), "unit": UNIT, "ms_per_step": ms_b / steps,
                                     "gpu_launches": launches_b, "max_rel_diff_vs_single_column": dmax,
                                     "what": "same unit with {cent, pred} and {centadj, predadj} solved by "
                                             "hyp_solve_system_multi / hyp_apply_lhs_multi (2 columns per sweep); a stepper-side "
                                             "change of the call sequence, so NOT the headline"}
        except Exception as e:
            out["batched_solves"] = {"error": f"{type(e).__name__}: {e}"}

    # per-phase device times (library CUDA-event timers; separate untimed pass)
    ctx.timing_enable(True)
    ctx.timing_reset()
    nprof = 2
    for _ in range(nprof):
        step(dev_in, dev_out)
    ctx.sync()
    out["phase_ms"] = {k: v[0] / nprof for k, v in ctx.timing().items() if v[1]}
    ctx.timing_enable(False)
    out["rows"] = I["rows"]
    out["dims"] = (n, q)

    relerr = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    parity = {"tol": 1e-8, "kkt_residual_device_operator": [relerr(dev_res[i], rhs_list[i]) for i in range(4)]}
    out["parity"] = parity
    out["_dev_sols"], out["_rhs_list"] = dev_sols, rhs_list

    if full:
        # one complete CombinedStepper.step at this iterate: update_lhs, four get_directions WITH iterative refinement
        # (up to 5 rounds each, common.jl:15-76), the line search with its cone-oracle sweeps (search.jl:46-138)
        try:
            out["full_step"] = full_step(sh, I, ctx, cones, irtmu, torch)
        except Exception as e:
            out["full_step"] = {"error": f"{type(e).__name__}: {e}"}
    ctx.close()
    if big:
        torch.cuda.empty_cache()
    if not full:
        # independent KKT operator: the two G products by torch (cuBLAS dgemv on the regenerated panel), the cone
        # Hessian products by the CPU oracle's cones - nothing of libhypatia_b200 in it
        try:
            I2 = build_instance(workload, rank, world, dist, device, on_device=True, replicate_rows=giant)
            parity["kkt_residual_independent"] = independent_kkt(torch, dist, device, I2, sh, dev_sols, rhs_list, giant)
            parity["kkt_operator"] = "torch dgemv on the generated G panel + CPU-oracle cone Hessians (no libhypatia_b200 code)"
            del I2
            torch.cuda.empty_cache()
        except Exception as e:
            parity["kkt_residual_independent"] = f"failed: {type(e).__name__}: {e}"
    return out


def independent_kkt(torch, dist, device, I, sh, dev_sols, rhs_list, giant):
    """||K d - r|| / ||r|| of the device directions with an operator that shares no code with the library."""
    from oracle.cones import OracleConeBlock
    from oracle.layout import OraclePoint
    model = I["model"]
    n, q = model.n, model.q
    row_lo, row_hi = I["rows"]
    GT = I["G_dev"]                                # (n, rows) on the device
    ora = OracleConeBlock(model)
    pt = sh.point
    ora.load_point(pt.s, pt.z, 1.0 / np.sqrt(sh.mu))
    world = dist.get_world_size() if dist is not None else 1
    out = []
    d, r = OraclePoint(n, 0, q), OraclePoint(n, 0, q)
    for v, rv in zip(dev_sols, rhs_list):
        d.vec[:] = v
        tz = torch.from_numpy(d.z[row_lo:row_hi].copy()).to(device)
        tx = torch.from_numpy(d.x.copy()).to(device)
        gtz = torch.mv(GT, tz)
        gx = torch.zeros(q, dtype=torch.float64, device=device)
        gx[row_lo:row_hi] = torch.mv(GT.t(), tx)
        if world > 1 and not giant:
            dist.all_reduce(gtz)
            dist.all_reduce(gx)
        gtz, gx = gtz.cpu().numpy(), gx.cpu().numpy()
        r.x[:] = gtz + model.c * d.tau                                   # common.jl:91-94
        r.z[:] = model.h * d.tau - d.s - gx                              # common.jl:100-103
        r.tau = -float(model.c @ d.x) - float(model.h @ d.z) - d.kap
        r.s[:] = ora.hess_prod(d.s) + d.z                                # common.jl:109-115 (primal-barrier cones)
        r.kap = sh.mu / pt.tau * d.tau / pt.tau + d.kap
        out.append(float(np.linalg.norm(r.vec - rv) / max(np.linalg.norm(rv), 1e-300)))
    return out


def full_step(sh, I, ctx, cones, irtmu, torch):
    from hypatia_b200.host import stepper as st
    stp = st.CombinedStepper().load(sh)
    pt = sh.point
    p0 = pt.vec.copy()
    times, info = [], None
    for it in range(3):
        pt.vec[:] = p0
        sh.n_solve_system = sh.n_apply_lhs = 0
        sh.worst_dir_res = 0.0
        stp.searcher.n_oracle_sweeps = 0
        ctx.sync()
        t0 = time.perf_counter()
        cones.load_point(pt.s, pt.z, irtmu)
        ok = stp.step(sh)
        ctx.sync()
        times.append(time.perf_counter() - t0)
        info = {"ok": bool(ok), "alpha": float(stp.prev_alpha), "n_solve_system": sh.n_solve_system,
                "n_apply_lhs": sh.n_apply_lhs, "n_oracle_sweeps": stp.searcher.n_oracle_sweeps,
                "worst_dir_res": float(sh.worst_dir_res)}
    pt.vec[:] = p0
    info.update({"ms": float(np.median(times[1:])) * 1e3, "value": 1.0 / float(np.median(times[1:])), "unit": UNIT,
                 "timing": "host wall clock around CombinedStepper.step (host control flow, device plug-ins), median of 2 after 1 warm-up",
                 "what": "load_point + update_lhs + 4 x get_directions incl. iterative refinement (max_ref_steps 5, cutoff "
                         "1e-4 x residual norm) + search_alpha with its cone-oracle sweeps + point update"})
    return info


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - libhypatia_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)

    main = measure_workload(args, args.workload, torch, dist, device, rank, world, local_rank, True)
    # the headline line is complete (roofline, CPU oracle parity) BEFORE the extra workloads start
    line = build_line(args, torch, device, world, main) if rank == 0 else None
    others = {}
    names = [x for x in args.other.split(",") if x and x != "none"]
    # BASELINE config 5's many-cone mix needs 240 GB of panels (G, H G, the mixed left operand): from 4 GPUs on
    if world >= 4 and names and "C5b" not in names and args.other == DEFAULT_OTHER:
        names.append("C5b")

    # The extra workloads must never cost the headline line: if one of them hangs (a rank lost in a collective inside
    # a C call - a Python signal handler would never run), a watchdog THREAD makes every rank leave after
    # `--other-timeout` seconds and rank 0 print what it has.
    lock = threading.Lock()
    state = {"printed": False}

    def emit():
        with lock:
            if rank == 0 and not state["printed"]:
                state["printed"] = True
                line["other_workloads"] = others
                print(json.dumps(line), flush=True)

    def on_timeout():
        others["_watchdog"] = f"other_workloads abandoned after {args.other_timeout} s"
        emit()
        os._exit(0)

    dog = None
    if names:
        dog = threading.Timer(float(args.other_timeout), on_timeout)
        dog.daemon = True
        dog.start()
    for w in names:
        try:
            r = measure_workload(args, w, torch, dist, device, rank, world, local_rank, False)
            r.pop("_dev_sols", None), r.pop("_rhs_list", None), r.pop("clocks", None)
            others[w] = r
        except Exception as e:
            others[w] = {"error": f"{type(e).__name__}: {e}"}
    if dog is not None:
        dog.cancel()
    emit()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def build_line(args, torch, device, world, main):

    ms_step = main["ms_per_step"]
    phases = main["phase_ms"]
    n, q = main["dims"]
    m = n
    qloc = main["rows"][1] - main["rows"][0]
    syrk_ms = phases.get("schur_syrk", float("nan"))
    syrk_flops = float(qloc) * m * (m + 1)
    fp64_eq = syrk_flops / (syrk_ms * 1e-3) / 1e12 if syrk_ms == syrk_ms and syrk_ms > 0 else None
    peaks = measured_peaks()
    fp64_peak = dgemm_peak_tflops(torch, device)
    if main["syrk"] == "i8":
        # exact int8 digit-pair products per FP64 product: 28 (seven balanced radix-256 digits, s + t <= 6; default)
        # or 36 (eight radix-128 digits, s + t <= 7: HYP_OZAKI_RADIX=128 or the single-CTA / cluster kernels)
        r128 = os.environ.get("HYP_OZAKI_RADIX") == "128" or os.environ.get("HYP_OZAKI_CLUSTER", "2")[:1] in ("0", "1")
        npairs = 28 if (not r128 or os.environ.get("HYP_OZAKI_SLICES") == "7") else 36
        nt = (m + 127) // 128
        int8_ops = 2.0 * npairs * float(qloc) * 128 * 128 * (nt * (nt + 1) // 2)
        int8_tops = int8_ops / (syrk_ms * 1e-3) / 1e12 if fp64_eq else None
        # the kernel is timed inside a long step: the sustained bf16 figure is the denominator (x 2 for int8)
        bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
        int8_peak = 2.0 * bf16 if bf16 else None
        order = os.environ.get("HYP_OZAKI_ORDER", "row6")
        tkey = {"row6": ":i8", "row": ":i8:row1", "row1": ":i8:row1"}.get(order)
        traffic = NCU_TRAFFIC.get(args.workload + tkey) if (world == 1 and tkey and "pair64" in syrk_kernel_name()) else None
        roofline = {"bound": "tensor",
                    "kernel": "%s (Schur SYRK: FP64-accurate digit slicing, %d exact int8 digit-pair products, tcgen05 kind::i8 "
                              "cta_group::2 M=256, TMEM accumulators, 3-D TMA) + slicing kernels" % (syrk_kernel_name(), npairs),
                    "achieved": int8_tops, "peak": int8_peak, "unit": "TOP/s (int8 MACs x 2 executed on tcgen05 kind::i8)",
                    "frac": (int8_tops / int8_peak) if int8_tops and int8_peak else None,
                    "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (assumes int8 dense = 2 x bf16 on B200; "
                                   "no int8 peak is measured by the driver)",
                    "fp64_equivalent_tflops": fp64_eq, "fp64_dgemm_peak_tflops": fp64_peak,
                    "fp64_equivalent_over_dgemm": (fp64_eq / fp64_peak) if fp64_eq else None,
                    "fp64_note": "algorithmic FP64 flops q*m*(m+1) per SYRK over its time, next to cuBLAS DGEMM 8192^3 measured "
                                 "live in this run (the FP64 DMMA roofline the digit-sliced kernel replaces)",
                    "traffic": traffic[0] if traffic else None,
                    "traffic_source": traffic[1] if traffic else None,
                    "algorithmic_flops_per_launch": syrk_flops, "executed_int8_ops_per_launch": int8_ops,
                    "avg_launch_ms": syrk_ms,
                    "step_share": syrk_ms / ms_step if syrk_ms == syrk_ms else None,
                    "phase_ms": phases}
    else:
        traffic = NCU_TRAFFIC.get(args.workload) if world == 1 else None
        roofline = {"bound": "tensor", "kernel": "atb_upper_kernel (Schur SYRK, TMA + FP64 DMMA)",
                    "achieved": fp64_eq, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (fp64_eq / fp64_peak) if fp64_eq else None,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (FP64 tensor pipe; "
                                   "MEASURED_PEAKS.json has no FP64 entry)",
                    "traffic": traffic[0] if traffic else None,
                    "traffic_source": traffic[1] if traffic else None,
                    "algorithmic_flops_per_launch": syrk_flops, "avg_launch_ms": syrk_ms,
                    "step_share": syrk_ms / ms_step if syrk_ms == syrk_ms else None,
                    "phase_ms": phases}
    potrf_ms = phases.get("potrf")
    if potrf_ms:
        roofline["potrf"] = {"kernel": "hyp_potrf_upper (blocked Cholesky: panel kernel + FP64 DMMA block rows + tcgen05 digit-sliced depth-512 "
                                       "trailing updates; HYP_POTRF=dag: task-graph FP64-DMMA kernel)",
                             "algorithmic_flops": m ** 3 / 3.0, "ms": potrf_ms,
                             "fp64_tflops": m ** 3 / 3.0 / (potrf_ms * 1e-3) / 1e12,
                             "frac_of_dgemm": m ** 3 / 3.0 / (potrf_ms * 1e-3) / 1e12 / fp64_peak}
    g_bytes = 8.0 * qloc * n
    gemv_ms = phases.get("gemv")
    if gemv_ms:
        # passes over G per step: const column 2 + 4 x (solve_subsystem3 2 + apply_lhs 2) = 18; with one rank
        # apply_lhs reads G once for both of its products (gemv_nt_kernel): 14
        npass = 14 if (world == 1 and not os.environ.get("HYP_NO_FUSED_GEMV")) else 18
        roofline["hbm_phase"] = {"kernel": "gemv_t / gemv_n / gemv_nt passes over G",
                                 "algorithmic_bytes_per_step": npass * g_bytes,
                                 "achieved_gbs": npass * g_bytes / (gemv_ms * 1e-3) / 1e9,
                                 "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "%d passes per step (reference count 22; the s-lift reuses G*x; apply_lhs "
                                         "reads G once for G'z and G x when the panel is not sharded)" % npass}
    trsv_ms = phases.get("trsv")
    if trsv_ms:
        roofline["trsv"] = {"algorithmic_bytes_per_step": 5 * 8.0 * m * m, "achieved_gbs": 5 * 8.0 * m * m / (trsv_ms * 1e-3) / 1e9,
                            "peak_gbs": peaks.get("hbm_gbs"), "note": "5 potrs per step, two triangular sweeps each (4 m^2 B per sweep)"}

    # parity at the benchmarked size, at EVERY N: rank 0 runs the CPU oracle on the full instance
    parity = main["parity"]
    parity["dir_vs_oracle"] = parity["kkt_residual"] = None
    cpu = None
    if not args.no_cpu_baseline:
        try:
            r = cpu_unit(args.workload, 1, 0, rhs_list=main["_rhs_list"], check_sols=main["_dev_sols"])
            relerr = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
            parity["dir_vs_oracle"] = [relerr(main["_dev_sols"][i], r["sols"][i]) for i in range(4)]
            parity["kkt_residual"] = r["kkt_oracle"]
            parity["kkt_operator"] = "oracle apply_lhs (NumPy, CPU) applied to the directions returned over the C ABI"
            if world == 1:
                cpu = {"value": 1.0 / r["sec"], "unit": UNIT, "cores": r["cores"], "blas_threads": r["blas_threads"],
                       "kind": "port", "sample": r["sample"]}
        except Exception as e:      # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"failed: {type(e).__name__}: {e}"}
    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": main["workload"],
                       "unit_of_work": "load_point + update_lhs + 4 x (solve_system + apply_lhs)",
                       "parallelism": f"cone/row-panel sharding over {world} rank(s)",
                       "schur_syrk": "tcgen05 int8 digit slicing (FP64-accurate)" if main["syrk"] == "i8" else "FP64 DMMA",
                       "l2": "inputs larger than L2 (G panel %.1f GB, Schur %.2f GB)" % (g_bytes / 1e9, 8e-9 * m * m)},
            "clocks": main["clocks"], "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"], "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "full_step": main.get("full_step"), "batched_solves": main.get("batched_solves"),
            "other_workloads": {}}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--other", default=os.environ.get("HYP_BENCH_OTHER", DEFAULT_OTHER),
                    help="comma-separated extra workloads reported under other_workloads ('none' to skip)")
    ap.add_argument("--other-steps", type=int, default=2)
    ap.add_argument("--other-timeout", type=float, default=900.0,
                    help="seconds after which the extra workloads are abandoned and the line is printed without them")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--syrk", default=os.environ.get("HYP_SCHUR_SYRK", "i8"), choices=["dmma", "i8"],
                    help="Schur SYRK kernel: FP64 DMMA or FP64-accurate digit slicing on the int8 tcgen05 pipe")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

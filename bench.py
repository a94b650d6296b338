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
  roofline : the Schur SYRK kernel (TMA + DMMA), algorithmic flops q*m*(m+1) per launch
  cpu_baseline : the CPU oracle restatement (OpenBLAS dsyrk/dpotrf/dgemv, all host cores), rank 0

Launch: `python bench.py --gpus 1 --steps K --warmup W`, or under torchrun for N > 1 (one rank per
GPU, cones / G row panels sharded over ranks, one NCCL allreduce of the Schur matrix per step;
"scaling": "strong" - the instance is fixed).  `--impl reference` times the CPU oracle instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

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
    # natvsext-faithful shape of BASELINE config 5 (SURVEY.md 8(d) "C5a"): the largest `nat` log-det
    # D-optimal-design instance (examples/doptimaldesign/JuMP_benchmark.jl:2-5, native.jl:18-86) after
    # the QR reduction: one HypoPerLogdetTri of side 1000 (dim 500502 > m, so the hess_prod! + GEMM
    # branch qrchol.jl:240-246) + two Nonnegative(2000); dense Gaussian G like the other configs
    "C5a": dict(n=2000, cones=lambda M: [M.Nonnegative(2000), M.Nonnegative(2000),
                                         M.HypoPerLogdetTri(2 + M.svec_length(1000))],
                desc="n=2000 p=0, Nonnegative(2000) x 2 + HypoPerLogdetTri(side 1000), q=504502"),
    # widening rows (not BASELINE configs): spectral cones through the batched Jacobi eigensolver
    "S1": dict(n=2000, cones=lambda M: [M.EpiPerSepSpectralMat(2 + M.svec_length(100), M.SSF_NEGENTROPY)
                                        for _ in range(24)],
               desc="n=2000 p=0, 24 x EpiPerSepSpectral{MatrixCSqr}(NegEntropy, side 100), q=121248"),
}


# DRAM traffic of the Schur SYRK launch from the committed `ncu --set full` capture (per launch)
NCU_TRAFFIC = {"C3": 31.497622e9 + 0.411380e9,                       # FP64 DMMA kernel, one launch
               # tcgen05 CTA-pair kernel, radix-256 digits, row-major tile order: the three launches (K chunks) of
               # one SYRK in profiles/r01_ozaki_pair_row_order_ncu.txt
               "C3:i8": (42.40e9 + 0.809e9) + (44.78e9 + 0.808e9) + (44.62e9 + 0.808e9)}


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


def build_instance(workload, rank, nranks, dist=None, device=None):
    from hypatia_b200.host import instances as inst
    from hypatia_b200.host import models as M
    from hypatia_b200.syssolver import partition_cones
    w = WORKLOADS[workload]
    n = w["n"]
    cones = w["cones"](M)
    seed = 1000 + sorted(WORKLOADS).index(workload)
    model = PanelModel(n, cones, None, None)
    q = model.q
    ranges = partition_cones(model, nranks) if nranks > 1 else [(0, len(cones))]
    lo, hi = ranges[rank]
    K = len(cones)
    row_lo = int(model.cone_offsets[lo]) if lo < K else q
    row_hi = int(model.cone_offsets[hi]) if hi < K else q
    G_local = gen_panel(n, row_lo, row_hi, seed)
    # planted interior primal-dual pair: cone central points perturbed like
    # test/cone.jl:236-248
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    s0, z0 = np.empty(q), np.empty(q)
    for ck, sl in zip(cones, model.cone_idxs):
        prim = inst.cone_initial_point(ck)
        dual = inst._cone_dual_initial(ck, prim)
        s0[sl] = inst._perturb(rng, ck, prim, 0.1)
        z0[sl] = inst._perturb(rng, ck, dual, 0.1)
        if ck.side:
            rows = sl.start + inst.mat_offset(ck) + np.nonzero(inst._svec_offdiag_mask(ck.side))[0]
            rows = rows[(rows >= row_lo) & (rows < row_hi)] - row_lo
            G_local[rows] *= np.sqrt(2.0)
    x0 = rng.standard_normal(n)
    h = np.zeros(q)
    h[row_lo:row_hi] = G_local @ x0 + s0[row_lo:row_hi]
    c = -(G_local.T @ z0[row_lo:row_hi])
    if nranks > 1:
        import torch
        th = torch.from_numpy(h).to(device)
        tc = torch.from_numpy(c).to(device)
        dist.all_reduce(th)
        dist.all_reduce(tc)
        h, c = th.cpu().numpy(), tc.cpu().numpy()
    model.c, model.h = c, h
    mu = (float(z0 @ s0) + 1.0) / (model.nu + 1)
    return dict(model=model, G_local=G_local, s0=s0, z0=z0, x0=x0, mu=mu, cone_range=(lo, hi),
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
    only HBM and bf16 entries; the Schur SYRK runs on the FP64 tensor pipe)."""
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


# --------------------------------------------------------------------------------------------
def cpu_unit(workload, steps, warmup, budget_s=25.0, rhs_list=None):
    """CPU oracle restatement of the same unit on the host cores.  Returns (seconds per unit,
    sample description, cores, threadpool info, directions of the last unit).  The SYRK is timed on a
    bounded row sample when a full unit would exceed the budget (its cost is linear in the rows; the
    factorisation then uses the Schur matrix of one full, untimed assembly), everything else runs in
    full.  `rhs_list` replaces the oracle's own four right-hand sides (bench.py's parity check hands in
    the ones the device solved)."""
    import threadpoolctl
    from scipy.linalg import blas as _blas
    from hypatia_b200.host import models as M
    from hypatia_b200.host.point import Point
    from oracle import syssolvers as osys
    from oracle.cones import OracleConeBlock
    from oracle.bench_unit import iterate_shell
    cores = os.cpu_count() or 1
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
    return float(np.median(times)), sample, cores, threadpoolctl.threadpool_info(), shell.sols


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sec, sample, cores, _, _ = cpu_unit(args.workload, max(1, args.steps), min(args.warmup, 1))
    v = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload]['desc']}"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
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
    from hypatia_b200 import capi
    from hypatia_b200.cones import DeviceConeBlock
    from hypatia_b200.host.point import Point
    from hypatia_b200.host import stepper as st

    I = build_instance(args.workload, rank, world, dist, device)
    model = I["model"]
    n, q = model.n, model.q
    ctx = capi.Context(local_rank)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    lo, hi = I["cone_range"]
    ctx.load_model(model, G_local=I["G_local"], cone_lo=lo, cone_hi=hi)
    # models with a cone that has no closed-form square root assemble S with the two-operand FP64 DMMA
    # product (qrchol.jl:240-246 branch); the one-operand tcgen05 SYRK needs every cone in sqrt form
    if any(ck.ctype not in (0, 1, 2, 6) for ck in model.cones):
        args.syrk = "dmma"
    ctx.set_syrk_mode(1 if args.syrk == "i8" else 0)
    I["G_local"] = None
    cones = DeviceConeBlock(model, ctx=ctx)

    # ---- the four right-hand sides of the first iterate (built once, untimed) ----
    class Shell:
        pass
    sh = Shell()
    sh.model, sh.mu, sh.cones = model, I["mu"], cones
    pt = sh.point = Point(model)
    pt.x[:] = I["x0"]
    pt.z[:] = I["z0"]
    pt.s[:] = I["s0"]
    pt.tau = pt.kap = 1.0
    sh.x_residual = np.zeros(n)
    sh.y_residual = np.zeros(0)
    sh.z_residual = np.zeros(q)
    sh.tau_residual = float(model.c @ pt.x) + float(model.h @ pt.z) + pt.kap
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

    ext = torch.cuda.ExternalStream(ctx.stream(), device=device)

    def timed(inp, out, steps, warmup, sample_clocks):
        for _ in range(warmup):
            step(inp, out)
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
            step(inp, out)
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

    ms, launches, clocks = timed(dev_in, dev_out, args.steps, args.warmup, True)
    value = args.steps / (ms * 1e-3)
    ms_e2e, _, _ = timed(host_in, host_out, args.steps, 1, False)
    e2e_value = args.steps / (ms_e2e * 1e-3)
    h2d = 8 * (2 * q + 4 * dim6 + 4 * dim6)       # point + 4 rhs + 4 dirs (apply_lhs input)
    d2h = 8 * (4 * dim6 + 4 * dim6) + 4           # 4 dirs + 4 residuals + Cholesky info word

    # per-phase device times (library CUDA-event timers; separate untimed pass)
    ctx.timing_enable(True)
    ctx.timing_reset()
    nprof = 2
    for _ in range(nprof):
        step(dev_in, dev_out)
    ctx.sync()
    phases = {k: v[0] / nprof for k, v in ctx.timing().items() if v[1]}
    ctx.timing_enable(False)
    m = n
    qloc = I["rows"][1] - I["rows"][0]
    syrk_ms = phases.get("schur_syrk", float("nan"))
    syrk_flops = float(qloc) * m * (m + 1)
    achieved = syrk_flops / (syrk_ms * 1e-3) / 1e12 if syrk_ms == syrk_ms and syrk_ms > 0 else None

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    fp64_peak = dgemm_peak_tflops(torch, device)
    if args.syrk == "i8":
        # exact int8 digit-pair products per FP64 product: 28 (seven balanced radix-256 digits, s + t <= 6; default)
        # or 36 (eight radix-128 digits, s + t <= 7: HYP_OZAKI_RADIX=128 or the single-CTA / cluster kernels)
        r128 = os.environ.get("HYP_OZAKI_RADIX") == "128" or os.environ.get("HYP_OZAKI_CLUSTER", "2")[:1] in ("0", "1")
        npairs = 28 if (not r128 or os.environ.get("HYP_OZAKI_SLICES") == "7") else 36
        nt = (m + 127) // 128
        int8_ops = 2.0 * npairs * float(qloc) * 128 * 128 * (nt * (nt + 1) // 2)
        int8_tops = int8_ops / (syrk_ms * 1e-3) / 1e12 if achieved else None
        int8_peak = 2.0 * peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None
        roofline = {"bound": "tensor",
                    "kernel": "ozaki_syrk_pair_kernel (Schur SYRK: FP64-accurate digit slicing, %d exact int8 digit-pair "
                              "products, tcgen05 kind::i8 cta_group::2 M=256, TMEM accumulators, 3-D TMA) + slicing kernels" % npairs,
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s (algorithmic FP64)",
                    "frac": (achieved / fp64_peak) if achieved else None,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run = the FP64 tensor (DMMA) roofline; "
                                   "frac > 1 because the contraction runs as exact int8 products on tcgen05",
                    "int8_top_s_executed": int8_tops,
                    "int8_peak_top_s": int8_peak,
                    "int8_peak_source": "2 x bf16_tflops of MEASURED_PEAKS.json (int8 dense = 2 x bf16 on B200)",
                    "int8_frac": (int8_tops / int8_peak) if int8_tops and int8_peak else None,
                    "traffic": NCU_TRAFFIC.get(args.workload + ":i8") if world == 1 else None,
                    "traffic_unit": "bytes per SYRK (dram__bytes_read.sum + dram__bytes_write.sum over its launches, "
                                    "profiles/r01_ozaki_pair_row_order_ncu.txt)",
                    "algorithmic_flops_per_launch": syrk_flops, "avg_launch_ms": syrk_ms,
                    "step_share": syrk_ms / (ms / args.steps) if syrk_ms == syrk_ms else None,
                    "phase_ms": phases}
    else:
        roofline = {"bound": "tensor", "kernel": "atb_upper_kernel (Schur SYRK, TMA + FP64 DMMA)",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (achieved / fp64_peak) if achieved else None,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (FP64 tensor pipe; "
                                   "MEASURED_PEAKS.json has no FP64 entry)",
                    "peak_bf16_measured": peaks.get("bf16_tflops"),
                    "frac_of_bf16_peak": (achieved / peaks["bf16_tflops"]) if achieved and peaks.get("bf16_tflops") else None,
                    "traffic": NCU_TRAFFIC.get(args.workload) if world == 1 else None,
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, "
                                    "profiles/r01_syrk_schur_ncu_full.txt)",
                    "algorithmic_flops_per_launch": syrk_flops, "avg_launch_ms": syrk_ms,
                    "step_share": syrk_ms / (ms / args.steps) if syrk_ms == syrk_ms else None,
                    "phase_ms": phases}
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
    # parity at the benchmarked size (SURVEY.md 8(d) "parity reported with every timing"): the directions
    # and residuals that came back over the C ABI in the end-to-end pass.  kkt_residual is the size-independent
    # property ||K d - r|| / ||r|| with the device operator; dir_vs_oracle compares each direction with the CPU
    # oracle's solve of the SAME right-hand side (filled in by the cpu_baseline leg below, N = 1 only).
    relerr = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    dev_sols = [t.numpy().copy() for t in host_out["sol"]]
    parity = {"tol": 1e-8,
              "kkt_residual": [relerr(host_out["res"][i].numpy(), rhs_list[i]) for i in range(4)],
              "dir_vs_oracle": None}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            sec, sample, cores, _, ora_sols = cpu_unit(args.workload, 1, 0, rhs_list=rhs_list)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            parity["dir_vs_oracle"] = [relerr(dev_sols[i], ora_sols[i]) for i in range(4)]
        except Exception as e:      # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"failed: {type(e).__name__}: {e}"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {I['desc']}",
                       "unit_of_work": "load_point + update_lhs + 4 x (solve_system + apply_lhs)",
                       "parallelism": f"cone/row-panel sharding over {world} rank(s)",
                       "schur_syrk": "tcgen05 int8 digit slicing (FP64-accurate)" if args.syrk == "i8" else "FP64 DMMA",
                       "l2": "inputs larger than L2 (G panel %.1f GB, Schur %.2f GB)" % (g_bytes / 1e9, 8e-9 * m * m)},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
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

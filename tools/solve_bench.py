"""Full solve to convergence ("for C5 num_iters / solve_time of a full solve to the natvsext
tolerances", benchmarks/natvsext/run.jl:34-47: tol_feas = tol_rel_opt = 1e-7, tol_abs_opt = tol_infeas = 1e-10,
iter_limit 250) of a scaled C5b mix (HypoPerLogdetTri + EpiNormEucl + Nonnegative) with the host driver.

  --impl device : device system solver + device cone oracles through the C ABI (this repo's product path)
  --impl oracle : the CPU oracle plug-ins (NumPy / OpenBLAS), the stand-in for the reference's CPU path

Prints one JSON line.  Not the headline metric (bench.py is); kept under profiles/ as evidence that the
whole loop - preprocessing, stepper, line search, refinement - runs on the device plug-ins at scale.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from hypatia_b200.host import instances as inst  # noqa: E402
from hypatia_b200.host.solver import Solver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="device", choices=["device", "oracle"])
    ap.add_argument("--config", default="C5b")
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--device-residuals", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    I = inst.config(args.config, args.scale)
    model = I.model
    if args.impl == "device":
        from hypatia_b200.cones import DeviceConeBlock as ConeF
        from hypatia_b200.syssolver import QRCholDenseSystemSolver
        sysv = QRCholDenseSystemSolver(device_residuals=args.device_residuals)
    else:
        from oracle.cones import OracleConeBlock as ConeF
        from oracle.syssolvers import QRCholDenseSystemSolver
        sysv = QRCholDenseSystemSolver()
    s = Solver(model, sysv, ConeF, verbose=args.verbose, tol_feas=1e-7, tol_rel_opt=1e-7, tol_abs_opt=1e-10,
               tol_infeas=1e-10, iter_limit=250)
    t0 = time.perf_counter()
    s.solve()
    dt = time.perf_counter() - t0
    x = s.get_x()
    line = {"what": "full solve to convergence", "impl": args.impl, "config": f"{args.config} x {args.scale}",
            "n": model.n, "p": model.p, "q": model.q, "cones": len(model.cones), "status": s.status,
            "num_iters": s.num_iters, "solve_time_s": dt, "iters_per_s": s.num_iters / dt,
            "primal_obj": s.primal_obj, "dual_obj": s.dual_obj, "gap": s.gap,
            "x_feas": s.x_feas, "z_feas": s.z_feas,
            "time_upsys_s": s.time_upsys, "time_getdir_s": s.time_getdir, "time_search_s": s.time_search,
            "time_uprhs_s": s.time_uprhs, "n_solve_system": s.n_solve_system, "n_apply_lhs": s.n_apply_lhs,
            "cores": os.cpu_count(), "x_norm": float(np.linalg.norm(x))}
    if args.impl == "device" and getattr(s.syssolver, "ctx", None) is not None:
        line["gpu_launches"] = s.syssolver.ctx.launch_count()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

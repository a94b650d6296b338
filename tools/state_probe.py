"""Times hyp_cones_load_point (cone state update) for the side-1000 log-det cone of C5a, call by call. GPU only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypatia_b200.host import instances as inst, models as M  # noqa: E402
from hypatia_b200.cones import DeviceConeBlock  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
I = inst.synthetic("x", 4, 0, [M.Nonnegative(2000), M.HypoPerLogdetTri(2 + M.svec_length(side))], seed=3)
dev = DeviceConeBlock(I.model)
prim, dual = I.point.primal_dual(None)
for rep in range(5):
    dev.ctx.sync()
    t0 = time.perf_counter()
    dev.load_point(prim, dual, 1 / np.sqrt(I.mu))
    dev.ctx.sync()
    t1 = time.perf_counter()
    dev.ctx.timing_enable(True); dev.ctx.timing_reset()
    dev.load_point(prim, dual, 1 / np.sqrt(I.mu))
    dev.ctx.sync()
    tm = {k: round(v[0], 3) for k, v in dev.ctx.timing().items() if v[1]}
    dev.ctx.timing_enable(False)
    print(f"side {side} rep {rep}: load_point wall {1e3 * (t1 - t0):.2f} ms; timers {tm}", flush=True)

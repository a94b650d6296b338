"""Where does hyp_potrf_upper spend its time?  Times the full factorisation, the latency chain alone
(HYP_POTRF_MODE=1: panel kernels + in-block updates) and the bulk GEMMs alone (HYP_POTRF_MODE=2) at the
Schur sizes of the BASELINE configs, with the library's own CUDA-event timer.  GPU only."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypatia_b200 import capi  # noqa: E402


def main():
    ctx = capi.Context(0)
    dev = torch.device("cuda", 0)
    out = {}
    for m in [int(a) for a in sys.argv[1:]] or [4000, 10000, 20000]:
        X = torch.randn(m + 64, m, dtype=torch.float64, device=dev)
        A = X.t() @ X + 0.5 * torch.eye(m, dtype=torch.float64, device=dev)
        del X
        res = {}
        impl = os.environ.get("HYP_POTRF", "i8")
        modes = {"stream": ("0", "1", "2"), "i8": ("0", "1", "2", "3")}.get(impl, ("0",))
        names = {"0": "full_ms", "1": "chain_only_ms", "2": "bulk_only_ms"} if impl == "stream" else \
            {"0": "full_ms", "1": "no_sliced_updates_ms", "2": "chain_stream_only_ms", "3": "bulk_stream_only_ms"}
        for mode in modes:
            os.environ["HYP_POTRF_MODE"] = mode
            ts = []
            for rep in range(4):
                F = A.clone()
                info = C.c_int(-1)
                ctx.timing_enable(True)
                ctx.timing_reset()
                ctx.check(ctx.lib.hyp_test_potrf(ctx.h, capi.ptr(F), m, m, C.byref(info)), "potrf")
                ctx.sync()
                ts.append(ctx.timing()["potrf"][0])
                ctx.timing_enable(False)
                if mode == "0" and rep in (0, 3):
                    # rep 0 = first call (tile / pair lists are built, with their stream synchronisations), rep 3 = steady state
                    # the library is column-major: its upper factor U is the LOWER triangle of the row-major view
                    Lw = torch.tril(F)
                    r = torch.linalg.norm(Lw @ Lw.t() - A) / torch.linalg.norm(A)
                    res["residual" if rep == 0 else "residual_steady_state"] = float(r)
                    res["info"] = info.value
                    del Lw
                del F
            res[names[mode]] = float(np.median(ts[1:]))
        res["tflops"] = m ** 3 / 3 / (res["full_ms"] * 1e-3) / 1e12
        out[m] = res
        del A
        torch.cuda.empty_cache()
    os.environ.pop("HYP_POTRF_MODE", None)
    print(json.dumps({"potrf_probe": out, "impl": os.environ.get("HYP_POTRF", "i8")}))


if __name__ == "__main__":
    main()

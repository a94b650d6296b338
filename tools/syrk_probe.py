"""What bounds the digit-sliced Schur SYRK?  Times ozaki_syrk_pair_kernel on the C3 shape (K = 50000, 10000 columns)
in three modes: full, HYP_OZAKI_PROBE=1 (no TMA loads: the MMA issuer runs alone on stale shared memory) and
HYP_OZAKI_PROBE=2 (no MMAs: TMA loads + epilogue only).  Timing only - the probe results are garbage.  GPU only."""
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
    K, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50000, 10000)
    ctx = capi.Context(0)
    dev = torch.device("cuda", 0)
    A = torch.randn(n, K, dtype=torch.float64, device=dev)          # column-major K x n
    Cm = torch.empty(n, n, dtype=torch.float64, device=dev)
    out = {}
    for mode, name in (("0", "full"), ("1", "mma_only_no_tma"), ("2", "tma_and_epilogue_only_no_mma")):
        os.environ["HYP_OZAKI_PROBE"] = mode
        ts = []
        for rep in range(3):
            ctx.timing_enable(True)
            ctx.timing_reset()
            ctx.check(ctx.lib.hyp_test_ozaki_syrk(ctx.h, capi.ptr(A), K, K, n, capi.ptr(Cm), n), "ozaki_syrk")
            ctx.sync()
            ts.append(ctx.timing()["schur_syrk"][0])
            ctx.timing_enable(False)
        out[name + "_ms"] = float(np.median(ts[1:]))
    os.environ.pop("HYP_OZAKI_PROBE", None)
    nt = (n + 127) // 128
    ops = 2.0 * 28 * K * 128 * 128 * (nt * (nt + 1) // 2)
    out["int8_tops_full"] = ops / (out["full_ms"] * 1e-3) / 1e12
    out["int8_tops_mma_only"] = ops / (out["mma_only_no_tma_ms"] * 1e-3) / 1e12
    print(json.dumps({"syrk_probe": out, "K": K, "ncols": n}))


if __name__ == "__main__":
    main()

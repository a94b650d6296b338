"""Issue rate of tcgen05.mma kind::i8 from shared memory vs layout / N / cta_group (hyp_test_mma_rate).  GPU only."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypatia_b200 import capi  # noqa: E402
ctx = capi.Context(0)
rows = []
for style in (0, 1):
  for cg2 in (1, 0):
    for N in (128, 256):
        for swz in (0, 1, 2):
            for ctas in (2, 148):
                out = np.zeros(2)
                rc = ctx.lib.hyp_test_mma_rate(ctx.h, swz | (style << 4), N, cg2, 20000, ctas, capi.ptr(out))
                if rc != 0:
                    print("error:", ctx.lib.hyp_last_error(ctx.h).decode(), flush=True)
                M = 256 if cg2 else 128
                macs = M * N * 32
                sms = ctas
                tops = 2.0 * macs * 20000 * (ctas // 2 if cg2 else ctas) / (out[1] * 1e-3) / 1e12 if rc == 0 and out[1] > 0 else None
                rows.append({"issue": "one thread" if style == 0 else "warp-uniform + elect", "cta_group": 2 if cg2 else 1, "M": M, "N": N, "swizzle_bytes": 32 << swz, "ctas": ctas, "rc": rc,
                             "cycles_per_mma": float(out[0]), "floor_cycles": M * N / (256 * (2 if cg2 else 1)) if True else None,
                             "launch_ms": float(out[1]), "chip_tops": tops})
                print(json.dumps(rows[-1]), flush=True)
print(json.dumps({"mma_probe": rows}))

"""Phase breakdown (clock64) of the one-CTA factor-and-invert kernel of a 128 x 128 block.  GPU only."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hypatia_b200 import capi  # noqa: E402
ctx = capi.Context(0)
rng = np.random.default_rng(0)
B = rng.standard_normal((200, 128))
A = np.asfortranarray(B.T @ B + 0.5 * np.eye(128))
clk = np.zeros(16)
ctx.check(ctx.lib.hyp_test_panel_clocks(ctx.h, capi.ptr(A), 128, 128, capi.ptr(clk)), "panel clocks")
names = ["start", "loaded"] + [f"b{b}:{p}" for b in range(4) for p in ("diag", "row", "trail")] + ["inverse", "stored"]
names[12] = "b3:xcol_thread0"
prev = 0.0
out = {}
for n, c in zip(names, clk):
    if c > 0 or n == "start":
        out[n] = {"at": c, "delta": c - prev}
        prev = c
print(json.dumps({"panel_cycles": out}))

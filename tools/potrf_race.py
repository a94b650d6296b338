"""Steady-state determinism check of hyp_potrf_upper: factor the same matrix many times and compare every result bit
for bit with the first one; report the 128-tiles that differ.  The kernels are deterministic, so ANY difference is a data
race between the two streams of the blocked Cholesky (chol.cu).  GPU only."""
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
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    ctx = capi.Context(0)
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    X = torch.randn(m + 64, m, dtype=torch.float64, device=dev)
    A = X.t() @ X + 0.5 * torch.eye(m, dtype=torch.float64, device=dev)
    del X
    ref = None
    bad = []
    for rep in range(reps):
        F = A.clone()
        torch.cuda.synchronize()
        info = C.c_int(-1)
        ctx.check(ctx.lib.hyp_test_potrf(ctx.h, capi.ptr(F), m, m, C.byref(info)), "potrf")
        ctx.sync()
        U = torch.tril(F)                     # column-major upper factor = lower triangle of the row-major view
        if ref is None:
            ref = U.clone()
            Lw = U
            res0 = float(torch.linalg.norm(Lw @ Lw.t() - A) / torch.linalg.norm(A))
            continue
        if not torch.equal(U, ref):
            d = (U != ref).nonzero()
            # row-major (i, j) = column-major (j, i): tile row of the factor = j // 128, tile column = i // 128
            tiles = sorted({(int(j) // 128, int(i) // 128) for i, j in d[:: max(1, len(d) // 2000)].tolist()})
            relerr = float(torch.linalg.norm(U - ref) / torch.linalg.norm(ref))
            bad.append({"rep": rep, "n_diff": int(len(d)), "rel": relerr, "first_tiles": tiles[:12],
                        "min_tile_row": min(t[0] for t in tiles), "min_tile_col": min(t[1] for t in tiles)})
    print(json.dumps({"m": m, "reps": reps, "residual_first": res0, "n_bad": len(bad), "bad": bad[:6],
                      "env": {k: v for k, v in os.environ.items() if k.startswith("HYP_")}}))


if __name__ == "__main__":
    main()

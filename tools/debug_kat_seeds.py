"""Robustness probe: epipersepspectral_matrix1 (d = 3, every h) over seeds, device plug-ins, with the
device residual step (hyp_calc_residuals) and with the host residuals."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kat_instances as kat
from hypatia_b200.host import models as M
from hypatia_b200.host.solver import Solver
from hypatia_b200.cones import DeviceConeBlock
from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
from collections import Counter

res = {}
for mode in ("dev", "host"):
    for hk, hp in kat.SEP_SPECTRAL_FUNS:
        cnt = Counter()
        for seed in range(1, 9):
            rng = np.random.default_rng(seed)
            d = 3
            W = rng.random((d, d)); W = W @ W.T + np.eye(d)
            dim = 2 + M.svec_length(d)
            G = np.zeros((dim, 1)); G[0, 0] = -1
            h = np.zeros(dim); h[1] = 1; h[2:] = kat._svec(W)
            model = kat._m([1], None, None, G, h, [M.EpiPerSepSpectralMat(dim, hk, hp)])
            sysv = DevQRChol()
            if mode == "host":
                sysv.calc_residuals = None
            s = Solver(model, sysv, DeviceConeBlock)
            s.solve()
            cnt[(s.status, )] += 1
            s.syssolver.free_memory() if s.syssolver.ctx is not None else None
        res[(mode, hk)] = dict(cnt)
        print(mode, "h", hk, dict(cnt), flush=True)

"""Debug helper: solve one KAT instance with the device plug-ins, printing the iteration log."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kat_instances as kat
from hypatia_b200.host.solver import Solver
from hypatia_b200.cones import DeviceConeBlock
from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol

name = sys.argv[1]
host_resid = len(sys.argv) > 2 and sys.argv[2] == "host"
build = [f for f in kat.SPECTRAL + kat.NEW_CONES + kat.ALL if f.__name__ == name][0]
model, expected = build()
sysv = DevQRChol()
if host_resid:
    sysv.calc_residuals = None
s = Solver(model, sysv, DeviceConeBlock, verbose=True)
orig = s.calc_convergence_params
def wrapped():
    r = orig()
    print("   improv %.6e feas %.6e %.6e %.6e %.6e" % (r, s.x_feas, s.y_feas, s.z_feas, s.tau_feas))
    return r
s.calc_convergence_params = wrapped
s.solve()
print(s.status, s.num_iters, s.primal_obj, expected.get("primal_obj"))

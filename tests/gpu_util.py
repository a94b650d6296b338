"""Helpers shared by the -m gpu tests: everything calls the CUDA path through the C ABI."""
import numpy as np
import pytest


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def ctx():
    from hypatia_b200 import capi
    return capi.Context(0)


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def iterate_solver(instance, syssolver, cones=None, Ap=None):
    """Solver shell positioned at the planted iterate (mirrors tests/test_oracle_solver.py)."""
    from hypatia_b200.host.solver import Solver
    from oracle.cones import OracleConeBlock
    model = instance.model
    s = Solver(model, syssolver, OracleConeBlock)
    s.model = model
    s.point = instance.point
    s.mu = instance.mu
    s.Ap_Q, s.Ap_R = (None, np.zeros((0, 0))) if Ap is None else Ap
    s.syssolver.load(s)
    s.cones = s.syssolver.cones if getattr(s.syssolver, "cones", None) is not None \
        else OracleConeBlock(model)
    primal, dual = s.point.primal_dual(s.cones.dual_mask)
    s.cones.load_point(primal, dual, 1 / np.sqrt(s.mu))
    s.syssolver.update_lhs(s)
    return s

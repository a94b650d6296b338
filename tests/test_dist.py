"""world_size = 2 tests of the sharded path: host-side logic on CPU (gloo), device path on 2 GPUs
(nccl, skipped when fewer than two GPUs are visible)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(impl, port, nproc=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py"), "--impl", impl]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK" in r.stdout
    return r.stdout


def test_sharded_schur_sum_gloo_world2():
    _run("oracle", 29531)


def test_column_sharded_schur_allgather_gloo_world2():
    """single-giant-cone path: column panels of S all-gathered, checked against the unsharded oracle"""
    _run("oracle_cols", 29533)


def test_partition_cones_balances_and_covers():
    import numpy as np
    from hypatia_b200.host import instances as inst
    from hypatia_b200.syssolver import cone_work, partition_cones
    I = inst.config("C3", 0.02)
    for nr in (1, 2, 3, 8):
        rg = partition_cones(I.model, nr)
        assert rg[0][0] == 0 and rg[-1][1] == len(I.model.cones)
        assert all(a[1] == b[0] for a, b in zip(rg, rg[1:]))
        w = [sum(cone_work(c, I.model.n) for c in I.model.cones[lo:hi]) for lo, hi in rg]
        assert max(w) <= 1.3 * (sum(w) / nr) + 1


@pytest.mark.gpu
def test_sharded_device_path_nccl_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("device", 29532)


@pytest.mark.gpu
def test_column_sharded_device_path_nccl_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("device_cols", 29534)

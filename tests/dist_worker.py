"""Worker of the world_size > 1 tests, launched with `python -m torch.distributed.run`.

--impl oracle (gloo, CPU): checks the host-side sharding logic - contiguous cone partition, per-rank
    partial Schur matrices summed with one all_reduce, partial G'z sums - against the unsharded
    oracle.
--impl device (nccl, one GPU per rank): the device system solver with cones / G row panels sharded
    over ranks; directions must match the unsharded CPU oracle at 1e-8.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_instance():
    from hypatia_b200.host import instances as inst
    from hypatia_b200.host import models as M
    cones = [M.EpiNormEucl(25) for _ in range(40)] + [M.Nonnegative(120)] + \
        [M.PosSemidefTri(M.svec_length(9)) for _ in range(5)] + [M.HypoPerLogdetTri(2 + M.svec_length(6))] + \
        [M.EpiNormEucl(7) for _ in range(11)]
    return inst.synthetic("dist", 260, 0, cones, seed=2024)


def build_giant_instance():
    """One dominant log-det cone + small ones: whole-cone sharding cannot split it."""
    from hypatia_b200.host import instances as inst
    from hypatia_b200.host import models as M
    cones = [M.Nonnegative(30), M.HypoPerLogdetTri(2 + M.svec_length(36)), M.EpiNormEucl(9), M.Nonnegative(12)]
    return inst.synthetic("giant", 150, 0, cones, seed=2025)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="oracle")
    args = ap.parse_args()
    if args.impl.endswith("_cols"):
        return main_cols(args.impl)
    import torch
    import torch.distributed as dist
    from gpu_util import iterate_solver, rel
    from hypatia_b200.host.point import Point
    from hypatia_b200.syssolver import partition_cones
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    I = build_instance()
    model = I.model
    ora = iterate_solver(I, OraQRChol())
    rng = np.random.default_rng(3)
    rhs = Point(model)
    rhs.vec[:] = rng.standard_normal(rhs.vec.size)
    so = Point(model)
    ora.syssolver.solve_system(ora, so, rhs)
    ranges = partition_cones(model, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(model.cones)
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))

    if args.impl == "oracle":
        dist.init_process_group("gloo")
        lo, hi = ranges[rank]
        cones = ora.cones
        S = np.zeros((model.n, model.n))
        gz = np.zeros(model.n)
        for k in range(lo, hi):
            sl = model.cone_idxs[k]
            ck = cones.cones[k]
            Gk = model.G[sl]
            if ck.use_sqrt_hess_oracles(model.n):
                HG = ck.sqrt_hess_prod(Gk)
                S += HG.T @ HG
            else:
                S += Gk.T @ ck.hess_prod(Gk)
            gz += Gk.T @ rhs.z[sl]
        tS, tg = torch.from_numpy(S), torch.from_numpy(gz)
        dist.all_reduce(tS)
        dist.all_reduce(tg)
        err_S = rel(np.triu(tS.numpy()), np.triu(ora.syssolver.lhs_full()))
        err_g = rel(tg.numpy(), model.G.T @ rhs.z)
        assert err_S <= 1e-12 and err_g <= 1e-12, (err_S, err_g)
        if rank == 0:
            print(f"DIST_OK oracle world={world} schur_err={err_S:.2e} gz_err={err_g:.2e}")
    else:
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
        dev = iterate_solver(I, DevQRChol(device=local_rank))
        assert dev.syssolver.nranks == world
        sd = Point(model)
        dev.syssolver.solve_system(dev, sd, rhs)
        err = rel(sd.vec, so.vec)
        rd, ro = Point(model), Point(model)
        dev.syssolver.apply_lhs(dev, so, rd)
        ora.syssolver.apply_lhs(ora, so, ro)
        err_r = rel(rd.vec, ro.vec)
        err_S = rel(np.triu(dev.syssolver.lhs_full()), np.triu(ora.syssolver.lhs_full()))
        prox_d = dev.cones.get_proxsqr(0.9, True)
        prox_o = ora.cones.get_proxsqr(0.9, True)
        assert err <= 1e-8 and err_r <= 1e-10 and err_S <= 1e-12, (err, err_r, err_S)
        assert np.allclose(prox_d, prox_o, rtol=1e-8, atol=1e-12)
        dev.syssolver.free_memory()
        if rank == 0:
            print(f"DIST_OK device world={world} dir_err={err:.2e} lhs_err={err_r:.2e} schur_err={err_S:.2e}")
    dist.barrier()
    dist.destroy_process_group()


def main_cols(impl):
    """Column sharding of the Schur assembly: rank r builds S[:, J_r] = G' (H G)[:, J_r], the panels are all-gathered."""
    import torch
    import torch.distributed as dist
    from gpu_util import iterate_solver, rel
    from hypatia_b200.host.point import Point
    from hypatia_b200.syssolver import column_ranges, giant_cone
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    I = build_giant_instance()
    model = I.model
    assert giant_cone(model)
    ora = iterate_solver(I, OraQRChol())
    S_ref = np.triu(ora.syssolver.lhs_full())
    rhs, so = Point(model), Point(model)
    rhs.vec[:] = np.random.default_rng(5).standard_normal(rhs.vec.size)
    ora.syssolver.solve_system(ora, so, rhs)
    if impl == "oracle_cols":
        dist.init_process_group("gloo")
        rg = column_ranges(model.n, world)
        assert rg[0][0] == 0 and rg[-1][1] == model.n and all(a[1] == b[0] for a, b in zip(rg, rg[1:]))
        lo, hi = rg[rank]
        cw = rg[0][1] - rg[0][0]
        panel = np.zeros((cw, model.n))                         # row c of `panel` = column lo + c of S
        HGJ = ora.cones.hess_prod(model.G[:, lo:hi])            # hess_prod! on my columns of G, all cones
        panel[:hi - lo] = (model.G.T @ HGJ).T
        out = [torch.zeros(cw, model.n, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(panel))
        S = torch.cat(out)[:model.n].numpy().T
        err_S = rel(np.triu(S), S_ref)
        assert err_S <= 1e-12, err_S
        if rank == 0:
            print(f"DIST_OK oracle_cols world={world} schur_err={err_S:.2e}")
    else:
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
        dev = iterate_solver(I, DevQRChol(device=local_rank))
        assert dev.syssolver.nranks == world and dev.syssolver.col_shard
        sd = Point(model)
        dev.syssolver.solve_system(dev, sd, rhs)
        err = rel(sd.vec, so.vec)
        rd, ro = Point(model), Point(model)
        dev.syssolver.apply_lhs(dev, so, rd)
        ora.syssolver.apply_lhs(ora, so, ro)
        err_r = rel(rd.vec, ro.vec)
        err_S = rel(np.triu(dev.syssolver.lhs_full()), S_ref)
        assert err <= 1e-8 and err_r <= 1e-10 and err_S <= 1e-12, (err, err_r, err_S)
        assert np.allclose(dev.cones.get_proxsqr(0.9, True), ora.cones.get_proxsqr(0.9, True), rtol=1e-8, atol=1e-12)
        dev.syssolver.free_memory()
        if rank == 0:
            print(f"DIST_OK device_cols world={world} dir_err={err:.2e} lhs_err={err_r:.2e} schur_err={err_S:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

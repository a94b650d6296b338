"""-m gpu parity of the device system solver (hyp_update_lhs / hyp_solve_system / hyp_apply_lhs
through the C ABI) against the CPU oracle restatement of QRCholDenseSystemSolver on the same
seeded instances (reduced-size versions of BASELINE.json's configs), with the tolerance the north
star states: ||d_dir|| / ||dir|| <= 1e-8 per direction.  Schur matrices are compared at
||dS||_F / ||S||_F <= 1e-12."""
import numpy as np
import pytest
import scipy.linalg as sla

from gpu_util import iterate_solver, rel
from hypatia_b200.host import instances as inst
from hypatia_b200.host import models as M
from hypatia_b200.host.point import Point, SubPoint

pytestmark = pytest.mark.gpu


def _wsos(n, halfdeg, use_dual=False):
    from wsos_util import interpolate_box
    U, _, Ps = interpolate_box(-np.ones(n), np.ones(n), halfdeg)
    return M.WSOSInterpNonnegative(U, Ps, use_dual=use_dual)
DIR_TOL = 1e-8


def _pair(I, Ap=None):
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    dev = iterate_solver(I, DevQRChol(), Ap=Ap)
    ora = iterate_solver(I, OraQRChol(), Ap=Ap)
    return dev, ora


def _check_system(I, Ap=None, schur_tol=1e-12):
    dev, ora = _pair(I, Ap)
    try:
        if I.model.n - I.model.p > 0:
            Sd, So = dev.syssolver.lhs_full(), ora.syssolver.lhs_full()
            assert rel(Sd, So) <= schur_tol
        assert dev.syssolver.fact_kind == ora.syssolver.fact_kind == 0
        rng = np.random.default_rng(7)
        for trial in range(3):
            rhs = Point(I.model)
            rhs.vec[:] = rng.standard_normal(rhs.vec.size)
            sd, so = Point(I.model), Point(I.model)
            dev.syssolver.solve_system(dev, sd, rhs)
            ora.syssolver.solve_system(ora, so, rhs)
            assert rel(sd.vec, so.vec) <= DIR_TOL, f"direction parity {rel(sd.vec, so.vec):.2e}"
            rd, ro = Point(I.model), Point(I.model)
            dev.syssolver.apply_lhs(dev, so, rd)
            ora.syssolver.apply_lhs(ora, so, ro)
            assert rel(rd.vec, ro.vec) <= 1e-11
            # the device solve satisfies the 6x6 system (checked with the oracle's operator)
            ora.syssolver.apply_lhs(ora, sd, ro)
            assert rel(ro.vec, rhs.vec) <= 1e-7
        # 3x3 subsystem entry point
        r3 = SubPoint(I.model.n, I.model.p, I.model.q)
        r3.vec[:] = rng.standard_normal(r3.vec.size)
        s3d, s3o = SubPoint(I.model.n, I.model.p, I.model.q), SubPoint(I.model.n, I.model.p, I.model.q)
        dev.syssolver.solve_subsystem3(dev, s3d, r3)
        ora.syssolver.solve_subsystem3(ora, s3o, r3)
        assert rel(s3d.vec, s3o.vec) <= DIR_TOL
    finally:
        dev.syssolver.free_memory()


@pytest.mark.parametrize("name,scale", [("C2", 0.05), ("C3", 0.02), ("C3", 0.1)])
@pytest.mark.parametrize("syrk", ["i8", "dmma"])
def test_vector_cone_configs(name, scale, syrk, monkeypatch):
    # both Schur SYRK kernels: FP64-accurate digit slicing on tcgen05 (default) and FP64 DMMA
    monkeypatch.setenv("HYP_SCHUR_SYRK", syrk)
    _check_system(inst.config(name, scale))


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_baseline_configs_at_full_size(name):
    """BASELINE.json's configs[0..1] at their FULL sizes (C2: m = 4000, q = 5000; C3: n = 10000,
    q = 50000, 2000 x EpiNormEucl(25) - the benchmarked workload): Schur matrix, directions, the
    operator and the 3x3 subsystem against the CPU oracle (one dsyrk + dpotrf of the full size, about
    10 s of host time for C3), and the KKT round trip apply_lhs(solve_system(r)) = r."""
    _check_system(inst.config(name, 1.0))


@pytest.mark.parametrize("name,scale", [("C2", 1.0), ("C3", 0.5)])
def test_steady_state_iterations_keep_direction_parity(name, scale):
    """The first update_lhs of a process builds tile / pair lists (with stream synchronisations); later ones run the two
    streams of the blocked Cholesky and every cached path without them.  A race between the streams showed up ONLY in
    later iterations (profiles/r02_potrf_race.md: found by bench.py's parity block, all single-shot tests green), so
    this test repeats update_lhs at the same iterate and compares the directions of the LAST factorisation with the
    oracle - and with the directions of the first one, bit for bit (all kernels are deterministic)."""
    I = inst.config(name, scale)
    dev, ora = _pair(I)
    try:
        rng = np.random.default_rng(17)
        rhs = Point(I.model)
        rhs.vec[:] = rng.standard_normal(rhs.vec.size)
        so = Point(I.model)
        ora.syssolver.solve_system(ora, so, rhs)
        first = None
        for it in range(6):
            if it:
                dev.syssolver.update_lhs(dev)
            sd = Point(I.model)
            dev.syssolver.solve_system(dev, sd, rhs)
            assert dev.syssolver.fact_kind == 0
            assert rel(sd.vec, so.vec) <= DIR_TOL, f"iteration {it}: direction parity {rel(sd.vec, so.vec):.2e}"
            if first is None:
                first = sd.vec.copy()
            else:
                assert np.array_equal(sd.vec, first), f"iteration {it} differs from the first one"
    finally:
        dev.syssolver.free_memory()


def _device_workload(name):
    """bench.py's device-generated instance of a multi-GB BASELINE config + a loaded context at its first iterate."""
    import torch
    import bench
    from hypatia_b200 import capi
    from hypatia_b200.cones import DeviceConeBlock
    dev = torch.device("cuda", 0)
    I = bench.build_instance(name, 0, 1, None, dev, on_device=True)
    model = I["model"]
    ctx = capi.Context(0)
    ctx.load_model(model, G_local=I["G_dev"])
    return torch, bench, dev, I, model, ctx, DeviceConeBlock(model, ctx=ctx)


@pytest.mark.parametrize("name,ncheck", [("C4", 192), ("C5a", 96)])
def test_baseline_configs_4_and_5_at_full_size(name, ncheck):
    """BASELINE.json configs[3] (C4: n = 20000, 50 x PosSemidefTri(side 100), q = 252500 - a 40 GB G) and the
    natvsext-shaped configs[4] (C5a: n = 2000, Nonnegative x 2 + ONE HypoPerLogdetTri of side 1000, q = 504502) at
    their FULL sizes on one GPU.  G is generated on the device (bench.py's block-seeded generator).  Checked:
      * a `ncheck` x `ncheck` sub-block of the Schur matrix (columns from both ends of the range) against the CPU
        oracle's G_J' H G_J built from the same columns (qrchol.jl:219-246) at 1e-12;
      * the directions of two right-hand sides through an INDEPENDENT 6x6 operator (torch dgemv on the panel +
        the CPU oracle's cone Hessians, no library code): ||K d - r|| / ||r|| <= 1e-8;
      * fact_kind = 0 (Cholesky succeeds, as for the oracle on the reduced-size versions of these configs)."""
    from oracle.cones import OracleConeBlock
    from hypatia_b200.host.point import Point
    torch, bench, dev, I, model, ctx, cones = _device_workload(name)
    try:
        n, q = model.n, model.q
        J = np.concatenate([np.arange(ncheck // 2), np.arange(n - ncheck // 2, n)])
        GJ = I["G_dev"][torch.from_numpy(J).to(dev)].t().contiguous().cpu().numpy()      # q x ncheck
        I["G_dev"] = None
        torch.cuda.empty_cache()
        pt_s, pt_z, mu = I["s0"], I["z0"], I["mu"]
        irtmu = 1.0 / np.sqrt(mu)
        cones.load_point(pt_s, pt_z, irtmu)
        ctx.set_mu_tau(mu, 1.0)
        if any(ck.ctype not in (0, 1, 2, 6) for ck in model.cones):
            ctx.set_syrk_mode(0)
        rc, kind = ctx.update_lhs()
        assert rc == 0 and kind == 0
        ora = OracleConeBlock(model)
        ora.load_point(pt_s, pt_z, irtmu)
        S_ora = GJ.T @ ora.hess_prod(GJ)
        S = ctx.get_schur()
        S = np.triu(S) + np.triu(S, 1).T
        S_dev = S[np.ix_(J, J)]
        del S
        assert rel(S_dev, S_ora) <= 1e-12, f"Schur sub-block parity {rel(S_dev, S_ora):.2e}"
        rng = np.random.default_rng(17)
        rhs_list, sols = [], []
        for _ in range(2):
            rhs, sol = Point(model), Point(model)
            rhs.vec[:] = rng.standard_normal(rhs.vec.size)
            ctx.solve_system(sol.vec, rhs.vec)
            assert np.isfinite(sol.vec).all()
            rhs_list.append(rhs.vec.copy())
            sols.append(sol.vec.copy())
    finally:
        ctx.close()
        torch.cuda.empty_cache()

    class Sh:
        pass
    sh = Sh()
    sh.mu = mu
    sh.point = Point(model)
    sh.point.s[:], sh.point.z[:] = pt_s, pt_z
    sh.point.tau = sh.point.kap = 1.0
    I2 = bench.build_instance(name, 0, 1, None, dev, on_device=True)
    res = bench.independent_kkt(torch, None, dev, I2, sh, sols, rhs_list, False)
    del I2
    torch.cuda.empty_cache()
    assert max(res) <= DIR_TOL, f"independent KKT residuals {res}"


@pytest.mark.parametrize("syrk", ["i8", "dmma"])
def test_schur_matrix_accuracy_vs_extended_precision(syrk, monkeypatch):
    """The assembled Schur matrix against a long-double reference of G'HG: both kernels must be at
    FP64 level (the sliced tcgen05 product is exact up to the final FP64 rounding of each entry)."""
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    from oracle.cones import OracleConeBlock
    monkeypatch.setenv("HYP_SCHUR_SYRK", syrk)
    I = inst.config("C3", 0.05)
    dev = iterate_solver(I, DevQRChol())
    try:
        ora = OracleConeBlock(I.model)
        ora.load_point(I.point.s, I.point.z, 1 / np.sqrt(I.mu))
        HG = ora.sqrt_hess_prod(I.model.G).astype(np.longdouble)
        ref = (HG.T @ HG)
        bound = (np.abs(HG).T @ np.abs(HG)).astype(np.float64)
        S = dev.syssolver.ctx.get_schur()
        iu = np.triu_indices(I.model.n)
        err = np.abs(S[iu] - ref[iu].astype(np.float64)) / bound[iu]
        assert err.max() <= (16 if syrk == "i8" else 4 * np.sqrt(I.model.q)) * np.finfo(np.float64).eps
    finally:
        dev.syssolver.free_memory()


def test_matrix_cone_configs():
    # reduced C4 (PSD blocks, q > n so that the Schur complement is positive definite) and C5b
    cones = [M.PosSemidefTri(M.svec_length(12)) for _ in range(6)]
    _check_system(inst.synthetic("C4r", 300, 0, cones, seed=1004))
    _check_system(inst.config("C5b", 0.004))
    cones = [M.PosSemidefTri(M.svec_length(130)), M.HypoPerLogdetTri(2 + M.svec_length(140)),
             M.Nonnegative(10)]
    _check_system(inst.synthetic("bigside", 500, 0, cones, seed=77))


def test_cholesky_failure_takes_the_bunch_kaufman_fallback():
    """q < n: the Schur complement G'HG is singular, Cholesky fails and the device takes the
    posdef_fact_copy! chain (dense.jl:194-215).  Directions of a numerically singular system are
    rounding noise on both sides; what must agree with the oracle is the outcome of the chain, and
    the solves must come back finite."""
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    cones = [M.PosSemidefTri(M.svec_length(10)) for _ in range(2)]
    I = inst.synthetic("rankdef", 200, 0, cones, seed=5)
    dev = iterate_solver(I, DevQRChol())
    ora = iterate_solver(I, OraQRChol())
    try:
        assert ora.syssolver.fact_kind in (1, 2)
        assert dev.syssolver.fact_kind == ora.syssolver.fact_kind
        rhs, sd = Point(I.model), Point(I.model)
        rhs.vec[:] = np.random.default_rng(3).standard_normal(rhs.vec.size)
        dev.syssolver.solve_system(dev, sd, rhs)
        assert np.isfinite(sd.vec).all()
    finally:
        dev.syssolver.free_memory()


def test_shifted_bunch_kaufman_end_of_the_chain_matches_oracle():
    """Two all-zero columns of G give the Schur complement two exactly-zero rows / columns: Cholesky
    fails, dsytrf_rook reports an exactly singular D (info > 0), and the chain ends in increase_diag!
    (dense.jl:106-113: A_jj <- (1 + 1e-5) max(A_jj, 1000 eps)) + Bunch-Kaufman = fact_kind 2.  The shifted
    system is nonsingular and decoupled in those two unknowns, so the directions ARE comparable: same
    fact_kind and direction parity at the north-star tolerance."""
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    from oracle.syssolvers import QRCholDenseSystemSolver as OraQRChol
    cones = [M.EpiNormEucl(5) for _ in range(30)] + [M.Nonnegative(40)]
    I = inst.synthetic("zerocols", 60, 0, cones, seed=9)
    m = I.model
    zero = [7, 41]
    m.G[:, zero] = 0.0
    m.h[:] = m.G @ I.point.x + I.point.s
    m.c[:] = -(m.G.T @ I.point.z)
    dev = iterate_solver(I, DevQRChol())
    ora = iterate_solver(I, OraQRChol())
    try:
        assert ora.syssolver.fact_kind == 2
        assert dev.syssolver.fact_kind == 2
        rng = np.random.default_rng(4)
        for _ in range(2):
            rhs, sd, so = Point(m), Point(m), Point(m)
            rhs.vec[:] = rng.standard_normal(rhs.vec.size)
            dev.syssolver.solve_system(dev, sd, rhs)
            ora.syssolver.solve_system(ora, so, rhs)
            assert rel(sd.vec, so.vec) <= DIR_TOL, f"direction parity {rel(sd.vec, so.vec):.2e}"
    finally:
        dev.syssolver.free_memory()


@pytest.mark.parametrize("p", [0, 3, 40])
def test_mixed_cones_with_equalities(p):
    cones = [M.Nonnegative(5), M.EpiNormEucl(4), M.PosSemidefTri(6), M.HypoPerLogdetTri(8),
             M.HypoRootdetTri(7), M.EpiNormEucl(3), M.HypoPerLogdetTri(5, use_dual=True)]
    I = inst.synthetic("mix", 12 + p, p, cones, seed=11)
    Ap = None
    if p:
        Qf, Rf = sla.qr(I.model.A.T, mode="full")
        Ap = (Qf, np.triu(Rf[:p, :p]))
    _check_system(I, Ap)


@pytest.mark.parametrize("p", [0, 4])
def test_spectral_and_perspective_cones(p):
    """EpiPerSepSpectral{MatrixCSqr} (hess_prod! + GEMM branch of the Schur assembly), EpiPerSquare
    (sqrt branch) and HypoPerLog (primal and dual barrier) next to the other cone types."""
    cones = [M.EpiPerSepSpectralMat(2 + M.svec_length(6), M.SSF_NEGENTROPY), M.EpiPerSquare(7),
             M.HypoPerLog(6), M.Nonnegative(4), M.EpiPerSepSpectralMat(2 + M.svec_length(3), M.SSF_INV),
             M.EpiPerSepSpectralMat(2 + M.svec_length(4), M.SSF_POWER12, 1.5, use_dual=True),
             M.HypoPerLog(5, use_dual=True), M.EpiNormEucl(5), M.EpiPerSquare(3),
             M.EpiPerSepSpectralMat(2 + M.svec_length(5), M.SSF_NEGLOG), M.EpiNormInf(6),
             M.EpiNormInf(4, use_dual=True), M.EpiPerSepSpectralVec(7, M.SSF_POWER12, 1.5),
             M.EpiPerSepSpectralVec(5, M.SSF_NEGLOG, use_dual=True), M.HypoGeoMean(6),
             M.HypoGeoMean(4, use_dual=True), M.GeneralizedPower([0.25, 0.75], 1),
             M.GeneralizedPower([0.2, 0.3, 0.5], 3, use_dual=True), M.HypoPowerMean([0.3, 0.7]),
             M.HypoPowerMean([0.2, 0.2, 0.6], use_dual=True), M.EpiRelEntropy(9),
             M.EpiRelEntropy(5, use_dual=True), M.EpiNormSpectral(2, 4), M.EpiNormSpectral(3, 3, use_dual=True),
             _wsos(2, 2), _wsos(1, 3, use_dual=True)]
    I = inst.synthetic("specmix", 20 + p, p, cones, seed=21)
    Ap = None
    if p:
        Qf, Rf = sla.qr(I.model.A.T, mode="full")
        Ap = (Qf, np.triu(Rf[:p, :p]))
    _check_system(I, Ap)


def test_sqrt_only_model_with_epipersquare_uses_the_syrk_path():
    # all cones have closed-form square roots: the one-operand SYRK (tcgen05 digit slicing) assembles S
    cones = [M.EpiPerSquare(25) for _ in range(40)] + [M.Nonnegative(30), M.EpiNormEucl(12)]
    _check_system(inst.synthetic("rsoc", 150, 0, cones, seed=22))


@pytest.mark.parametrize("p", [0, 5, 150])
def test_vector_cones_with_equalities(p):
    cones = [M.Nonnegative(50), M.EpiNormEucl(25), M.EpiNormEucl(25), M.Nonnegative(1), M.EpiNormEucl(300)]
    I = inst.synthetic("vmix", 200 + p, p, cones, seed=12)
    Ap = None
    if p:
        Qf, Rf = sla.qr(I.model.A.T, mode="full")
        Ap = (Qf, np.triu(Rf[:p, :p]))
    _check_system(I, Ap)


def test_empty_and_tiny_models():
    # n = p (no Schur block), and a 1x1 model
    I = inst.synthetic("tiny", 1, 0, [M.Nonnegative(1)], seed=3)
    _check_system(I)
    I = inst.synthetic("np", 3, 3, [M.Nonnegative(4), M.EpiNormEucl(3)], seed=4)
    Qf, Rf = sla.qr(I.model.A.T, mode="full")
    _check_system(I, (Qf, np.triu(Rf[:3, :3])))


@pytest.mark.parametrize("p", [0, 3])
def test_symindef_dense_device_vs_oracle(p):
    """SymIndefDenseSystemSolver on the device (hyp_set_syssolver(ctx, 1): explicit inv_hess blocks +
    rook Bunch-Kaufman) against the oracle's SymIndefDense and QRCholDense."""
    from hypatia_b200.syssolver import SymIndefDenseSystemSolver as DevSym
    from oracle.syssolvers import SymIndefDenseSystemSolver as OraSym
    cones = [M.Nonnegative(5), M.EpiNormEucl(4), M.PosSemidefTri(6), M.HypoPerLogdetTri(8),
             M.HypoRootdetTri(7), M.EpiNormEucl(3), M.HypoPerLogdetTri(5, use_dual=True)]
    I = inst.synthetic("mixsym", 12 + p, p, cones, seed=13)
    dev = iterate_solver(I, DevSym())
    ora = iterate_solver(I, OraSym())
    try:
        assert dev.syssolver.fact_kind == 1
        rng = np.random.default_rng(9)
        for _ in range(3):
            rhs = Point(I.model)
            rhs.vec[:] = rng.standard_normal(rhs.vec.size)
            sd, so = Point(I.model), Point(I.model)
            dev.syssolver.solve_system(dev, sd, rhs)
            ora.syssolver.solve_system(ora, so, rhs)
            assert rel(sd.vec, so.vec) <= DIR_TOL
            ro = Point(I.model)
            ora.syssolver.apply_lhs(ora, sd, ro)
            assert rel(ro.vec, rhs.vec) <= 1e-8
    finally:
        dev.syssolver.free_memory()


def test_explicit_hess_blocks_match_oracle():
    cones = [M.Nonnegative(3), M.EpiNormEucl(5), M.PosSemidefTri(6), M.HypoPerLogdetTri(8),
             M.HypoRootdetTri(7), M.EpiPerSepSpectralMat(2 + M.svec_length(3), M.SSF_NEGENTROPY),
             M.EpiPerSquare(5), M.HypoPerLog(6), M.EpiNormInf(5), M.GeneralizedPower([0.4, 0.6], 2)]
    I = inst.synthetic("blocks", 4, 0, cones, seed=14)
    from hypatia_b200.cones import DeviceConeBlock
    from oracle.cones import OracleConeBlock
    dev, ora = DeviceConeBlock(I.model), OracleConeBlock(I.model)
    prim, dual = I.point.primal_dual(None)
    dev.load_point(prim, dual, 0.7)
    ora.load_point(prim, dual, 0.7)
    for Hd, Hid, ck in zip(dev.hess(), dev.inv_hess(), ora.cones):
        Ho, Hio = np.asarray(ck.hess()), np.asarray(ck.inv_hess())
        if Ho.ndim == 1:                       # the oracle keeps the orthant's Hessian as a diagonal
            Ho, Hio = np.diag(Ho), np.diag(Hio)
        assert rel(Hd, Ho) <= 1e-10 and rel(Hid, Hio) <= 1e-10
        assert rel(Hd @ Hid, np.eye(Hd.shape[0])) <= 1e-8        # cone.jl:73-75
    dev.free()


@pytest.mark.parametrize("p", [0, 6])
def test_calc_residuals_matches_host_formulas(p):
    """hyp_calc_residuals (the residual step of calc_convergence_params, Solvers.jl:425-483) against the
    NumPy restatement in host/solver.py on the same point."""
    from hypatia_b200.syssolver import QRCholDenseSystemSolver as DevQRChol
    cones = [M.Nonnegative(40), M.EpiNormEucl(25), M.PosSemidefTri(15), M.HypoPerLog(7)]
    I = inst.synthetic("resid", 30 + p, p, cones, seed=31)
    Ap = None
    if p:
        Qf, Rf = sla.qr(I.model.A.T, mode="full")
        Ap = (Qf, np.triu(Rf[:p, :p]))
    dev = iterate_solver(I, DevQRChol(), Ap=Ap)
    try:
        m, pt = I.model, I.point
        rng = np.random.default_rng(5)
        pt.y[:] = rng.standard_normal(m.p)
        pt.tau = 0.7
        xr, yr, zr, st = dev.syssolver.calc_residuals(dev)
        gx = m.G.T @ pt.z + (m.A.T @ pt.y if p else 0)
        ax = m.A @ pt.x if p else np.zeros(0)
        gs = m.G @ pt.x + pt.s
        inf = lambda v: np.linalg.norm(v, np.inf) if v.size else 0.0
        assert rel(xr, -(gx + m.c * pt.tau)) <= 1e-12
        assert rel(zr, gs - m.h * pt.tau) <= 1e-12
        if p:
            assert rel(yr, ax - m.b * pt.tau) <= 1e-12
        ref = [inf(gx), inf(gx + m.c * pt.tau), inf(ax), inf(ax - m.b * pt.tau), inf(gs), inf(gs - m.h * pt.tau),
               m.c @ pt.x, m.b @ pt.y if p else 0.0, m.h @ pt.z, pt.z @ pt.s]
        assert np.allclose(st, ref, rtol=1e-11, atol=1e-11)
    finally:
        dev.syssolver.free_memory()


@pytest.mark.parametrize("name,scale,ncols", [("C3", 0.1, 2), ("C3", 0.1, 3), ("C3", 0.02, 2), ("C2", 1.0, 2)])
def test_multi_column_solves_match_single_column_calls(name, scale, ncols):
    """hyp_solve_system_multi / hyp_apply_lhs_multi (the {cent, pred} and {centadj, predadj} pairs of
    steppers/combined.jl:67-79 solved together: shared passes over G, shared triangular sweeps) against ncols calls of
    hyp_solve_system / hyp_apply_lhs, on device and on host buffers, and against the oracle."""
    import torch
    I = inst.config(name, scale)
    dev, ora = _pair(I)
    try:
        ctx = dev.syssolver.ctx
        dim6 = I.model.n + I.model.p + 2 * I.model.q + 2
        rng = np.random.default_rng(21)
        R = rng.standard_normal((ncols, dim6))
        single = np.empty_like(R)
        res_single = np.empty_like(R)
        for j in range(ncols):
            ctx.solve_system(single[j], R[j])
            ctx.apply_lhs(res_single[j], single[j])
        # host buffers
        multi = np.empty_like(R)
        ctx.solve_system_multi(multi, R, ncols)
        assert rel(multi, single) <= 1e-13
        # device buffers (the path that shares the passes)
        dR = torch.from_numpy(R).cuda()
        dS = torch.empty_like(dR)
        dA = torch.empty_like(dR)
        ctx.solve_system_multi(dS, dR, ncols)
        ctx.apply_lhs_multi(dA, dS, ncols)
        ctx.sync()
        assert rel(dS.cpu().numpy(), single) <= 1e-13
        assert rel(dA.cpu().numpy(), res_single) <= 1e-12
        for j in range(ncols):
            rhs, so = Point(I.model), Point(I.model)
            rhs.vec[:] = R[j]
            ora.syssolver.solve_system(ora, so, rhs)
            assert rel(dS[j].cpu().numpy(), so.vec) <= DIR_TOL
    finally:
        dev.syssolver.free_memory()

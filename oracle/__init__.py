"""CPU oracle for the Hypatia KKT / cone-oracle hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy + SciPy's bundled OpenBLAS/LAPACK, i.e. the same
LAPACK routines Julia's stdlib calls) of the reference algorithms on the hot path:

  oracle/linalg.py      <- src/linearalgebra/dense.jl            (potrf -> BK-rook -> shifted BK chain)
  oracle/arrayutil.py   <- src/Cones/arrayutilities.jl           (svec/smat, symm_kron)
  oracle/cones.py       <- src/Cones/{Cones,nonnegative,epinormeucl,possemideftri,
                                      hypoperlogdettri,hyporootdettri}.jl
  oracle/syssolvers.py  <- src/Solvers/systemsolvers/{common,qrchol,symindef,naive}.jl
  oracle/layout.py      <- src/Solvers/point.jl, common.jl:184-208 (the oracle's own Point / SubPoint / cone loops)
  (instance generators live with the host driver: hypatia.jl_b200/host/instances.py <- examples/linearopt/native.jl
   + the synthetic configs of SURVEY.md 8(d); they only build inputs, both sides read the same Model)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
may import it, and only as the checker / the reported CPU baseline.  Nothing under
hypatia.jl_b200/ imports it; the product path fails loudly when the CUDA library is
missing instead of falling back to this code.

Parity pinning.  The reference is pure Julia and Julia is not installed in this image
(nor can it be: no network), so the reference itself cannot be executed here and it ships
no golden vectors for this path.  The oracle is pinned by (1) the implementation-independent
oracle identities of test/cone.jl:23-114 at the reference's own tolerance 1e3*eps,
(2) the deterministic known-answer instances of test/nativeinstances.jl (closed-form optima),
(3) the reference's own cross-solver equivalence check (QRCholDense == SymIndefDense ==
NaiveDense, test/runnativetests.jl:101-118), all in tests/test_oracle_*.py.  At the
BASELINE tolerance (||d_dir||/||dir|| <= 1e-8) no reference test pins directions, so
direction-level parity is pinned by this restatement only.
"""

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r10_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r10_pytest.log )
tail -6 gpurun_out/r10_pytest.log
( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench_fused.json 2> gpurun_out/r10_bench_fused.err ; echo "fused rc=$?" )
( HYP_NO_FUSED_GEMV=1 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench_nofuse.json 2> gpurun_out/r10_bench_nofuse.err ; echo "nofuse rc=$?" )
python - <<'PY'
import json
for f in ("gpurun_out/r10_bench_fused.json","gpurun_out/r10_bench_nofuse.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"],3), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"], d["roofline"]["hbm_phase"]["achieved_gbs"])
    except Exception as e: print(f, "ERR", e)
PY

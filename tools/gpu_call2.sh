set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout=240 > gpurun_out/r2_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log )
tail -15 gpurun_out/r2_pytest.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_s8.json 2> gpurun_out/r2_bench_c3_s8.err ; echo "bench8 rc=$?" )
( HYP_OZAKI_SLICES=7 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_s7.json 2> gpurun_out/r2_bench_c3_s7.err ; echo "bench7 rc=$?" )
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_c3_s8.json","gpurun_out/r2_bench_c3_s7.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"],3), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
HYP_OZAKI_SLICES=7 timeout 300 python -m pytest tests/test_gpu_system.py tests/test_gpu_ozaki.py -m gpu -q -p no:cacheprovider --timeout=240 2>&1 | tail -15

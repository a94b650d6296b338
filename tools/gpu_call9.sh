set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -s KILL 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_system.py -m gpu -q -p no:cacheprovider > gpurun_out/r9_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r9_pytest.log )
tail -6 gpurun_out/r9_pytest.log
( HYP_OZAKI_RADIX=128 timeout -s KILL 300 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -p no:cacheprovider -k "syrk_matches" > gpurun_out/r9_pytest_r128.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r9_pytest_r128.log )
tail -4 gpurun_out/r9_pytest_r128.log
( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench_r256.json 2> gpurun_out/r9_bench_r256.err ; echo "r256 rc=$?" )
( HYP_OZAKI_RADIX=128 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench_r128.json 2> gpurun_out/r9_bench_r128.err ; echo "r128 rc=$?" )
python - <<'PY'
import json
for f in ("gpurun_out/r9_bench_r256.json","gpurun_out/r9_bench_r128.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"],3), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY

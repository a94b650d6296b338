set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( HYP_OZAKI_RADIX=256 timeout -s KILL 200 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -p no:cacheprovider > gpurun_out/r7_r256_ozaki.log 2>&1 ; echo "r256 ozaki rc=$?" >> gpurun_out/r7_r256_ozaki.log )
tail -12 gpurun_out/r7_r256_ozaki.log
if grep -q "passed" gpurun_out/r7_r256_ozaki.log && ! grep -q "failed" gpurun_out/r7_r256_ozaki.log; then
  ( HYP_OZAKI_RADIX=256 timeout -s KILL 300 python -m pytest tests/test_gpu_system.py tests/test_gpu_solve.py -m gpu -q -p no:cacheprovider > gpurun_out/r7_r256_system.log 2>&1 ; echo "r256 system rc=$?" >> gpurun_out/r7_r256_system.log )
  tail -5 gpurun_out/r7_r256_system.log
  ( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_r128.json 2> gpurun_out/r7_bench_r128.err ; echo "r128 rc=$?" )
  ( HYP_OZAKI_RADIX=256 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_r256.json 2> gpurun_out/r7_bench_r256.err ; echo "r256 rc=$?" )
  ( HYP_OZAKI_RADIX=256 HYP_OZAKI_CLUSTER=3 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench_r256_quad.json 2> gpurun_out/r7_bench_r256_quad.err ; echo "r256 quad rc=$?" )
  python - <<'PY'
import json
for f in ("gpurun_out/r7_bench_r128.json","gpurun_out/r7_bench_r256.json","gpurun_out/r7_bench_r256_quad.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"],3), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
fi

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout=240 > gpurun_out/r4_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r4_pytest.log )
tail -8 gpurun_out/r4_pytest.log
( HYP_MAT_SMALL_MAXCOLS=0 timeout 600 python -m pytest tests/test_gpu_cones.py tests/test_gpu_system.py -m gpu -q -p no:cacheprovider --timeout=240 > gpurun_out/r4_pytest_tensorpath.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r4_pytest_tensorpath.log )
tail -4 gpurun_out/r4_pytest_tensorpath.log
( timeout 300 python bench.py --workload S1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r4_bench_s1.json 2> gpurun_out/r4_bench_s1.err ; echo "s1 rc=$?" )
( timeout 600 python tools/solve_bench.py --impl device --scale 0.1 > gpurun_out/r4_solve_device.json 2> gpurun_out/r4_solve_device.err ; echo "solve rc=$?" )
cut -c1-700 gpurun_out/r4_solve_device.json; tail -3 gpurun_out/r4_solve_device.err
( timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r4_launches_c3.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r4_launches_c3.log 2>&1 ; echo "ncu rc=$?" )
python - <<'PY'
import json
for f in ("gpurun_out/r4_bench_s1.json",):
    try:
        d=json.load(open(f)); print(f, round(d["value"],3), round(d["ms_per_step"],2), d["gpu_launches"], {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/verify_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/verify_pytest.log )
tail -6 gpurun_out/verify_pytest.log
( HYP_MAT_SMALL_MAXCOLS=0 timeout -s KILL 600 python -m pytest tests/test_gpu_cones.py tests/test_gpu_solve.py -m gpu -q -p no:cacheprovider > gpurun_out/verify_pytest_tensorpath.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/verify_pytest_tensorpath.log )
tail -3 gpurun_out/verify_pytest_tensorpath.log
( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err ; echo "bench rc=$?" )
( timeout -s KILL 600 python tools/solve_bench.py --impl device --scale 0.1 > gpurun_out/verify_solve_device.json 2> gpurun_out/verify_solve_device.err ; echo "solve rc=$?" )
python - <<'PY'
import json
d=json.load(open("gpurun_out/verify_bench.json")); print(round(d["value"],3), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"], d["roofline"]["hbm_phase"]["achieved_gbs"])
s=json.load(open("gpurun_out/verify_solve_device.json")); print({k:s[k] for k in ("status","num_iters","solve_time_s","time_upsys_s","time_getdir_s","time_search_s","time_uprhs_s")})
PY

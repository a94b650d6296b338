# Full -m gpu suite + one default bench.py run (with the CPU-oracle leg, which also fills parity.dir_vs_oracle).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=25 --durations=8 -p no:cacheprovider > gpurun_out/verify_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/verify_pytest.log )
tail -16 gpurun_out/verify_pytest.log
( timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/verify_bench_full.json 2> gpurun_out/verify_bench_full.err ; echo "bench rc=$?" )
python - <<'PY'
import json
d=json.load(open("gpurun_out/verify_bench_full.json")); print(round(d["value"],3), round(d["e2e"]["value"],3), d["parity"], d["cpu_baseline"], d["clocks"])
PY

# Final verification of a session: full -m gpu suite, then the C3 bench with the row-major SYRK tile order (with the
# CPU-oracle leg: parity.dir_vs_oracle validates that order at full size) and with the default order on the same box,
# then the DRAM traffic of the row-major order (ncu metrics pass; its timings are not bench values).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -s KILL 300 python -m pytest tests -m gpu -q --maxfail=25 --durations=5 -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/final_pytest.log )
tail -12 gpurun_out/final_pytest.log
( HYP_OZAKI_ORDER=row timeout -s KILL 200 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_row.json 2> gpurun_out/final_bench_row.err ; echo "bench row rc=$?" )
( timeout -s KILL 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_col.json 2> gpurun_out/final_bench_col.err ; echo "bench col rc=$?" )
python - <<'PY'
import json
for f in ("row", "col"):
    try:
        d = json.load(open(f"gpurun_out/final_bench_{f}.json"))
        print(f, round(d["value"], 3), round(d["e2e"]["value"], 3), {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["parity"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
( HYP_OZAKI_ORDER=row timeout -s KILL 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ozaki_syrk -c 3 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/final_ncu_row.txt 2>&1 ; echo "ncu rc=$?" )
grep -E "ozaki_syrk|dram__bytes|gpu__time|lts__t" gpurun_out/final_ncu_row.txt | head -20

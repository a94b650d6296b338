set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/r8_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r8_pytest.log )
tail -6 gpurun_out/r8_pytest.log
( timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r8_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r8_smoke.log ); tail -3 gpurun_out/r8_smoke.log
( timeout -s KILL 400 python bench.py > gpurun_out/r8_bench_c3.json 2> gpurun_out/r8_bench_c3.err ; echo "bench rc=$?" )
( timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r8_bench_ref.json 2> gpurun_out/r8_bench_ref.err ; echo "ref rc=$?" )
cut -c1-300 gpurun_out/r8_bench_ref.json
( timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active --clock-control none -k regex:ozaki_syrk -c 4 --csv --log-file gpurun_out/r8_ozaki_pair_r256_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r8_ncu.log 2>&1; echo "ncu rc=$?" )
python - <<'PY'
import json
d=json.load(open("gpurun_out/r8_bench_c3.json")); print(round(d["value"],3), round(d["ms_per_step"],2), d["e2e"], {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"], d["cpu_baseline"]); print({k:d["roofline"][k] for k in ("achieved","peak","frac","int8_top_s_executed","int8_frac")})
PY

#!/bin/bash
# validation of the current build: full -m gpu suite, default bench line (CPU leg, other workloads), ncu --set full of the
# Schur SYRK (64-byte-row pair kernel), launch list of the bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02v_pytest_gpu.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_syrk_pair64 -c 3 -o gpurun_out/r02v_syrk_pair64 \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --other none > gpurun_out/r02v_ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02v_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --other none > gpurun_out/r02v_launch_bench.log 2>&1; echo "launches rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['phase_ms'], d['roofline']['frac'], d['parity']['dir_vs_oracle'], d['cpu_baseline'])
print('batched', d['batched_solves']['ms_per_step'], 'full_step', d['full_step'])
for w,v in d['other_workloads'].items():
    print(w, v.get('ms_per_step') if isinstance(v,dict) else v, v.get('phase_ms') if isinstance(v,dict) else '', v.get('error') if isinstance(v,dict) else '')
PY
ls -la gpurun_out/r02v_syrk_pair64.ncu-rep

#!/bin/bash
# validation of the final build of round 2 (row6 tile order, byte-parallel slicing, pre-pass, per-cone API): full -m gpu suite, default bench line, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02zz_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02zz_pytest_gpu.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02zz_bench_n1.json 2> gpurun_out/r02zz_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02zz_bench_ref.json 2> gpurun_out/r02zz_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02zz_bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zz_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['phase_ms'], d['roofline']['frac'], d['parity']['dir_vs_oracle'], d['parity']['kkt_residual'], d['cpu_baseline'], d['clocks'])
print('batched', d['batched_solves']['ms_per_step'], d['batched_solves']['max_rel_diff_vs_single_column'], 'full_step', d['full_step'])
for w,v in d['other_workloads'].items():
    print(w, v.get('ms_per_step') if isinstance(v,dict) else v, v.get('phase_ms') if isinstance(v,dict) else '', v.get('parity') if isinstance(v,dict) else '', v.get('error') if isinstance(v,dict) else '')
PY

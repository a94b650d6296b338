#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/panel_probe.py > gpurun_out/r02m_panel_probe.json 2>&1; cat gpurun_out/r02m_panel_probe.json
HYP_POTRF_DEBUG=1 timeout 300 python tools/potrf_probe.py 1000 4000 10000 20000 > gpurun_out/r02m_potrf_dag.json 2> gpurun_out/r02m_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02m_potrf_dag.json; for m in 4000 10000; do grep "m=$m\]" gpurun_out/r02m_potrf_dag.err | tail -2; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02m_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --other C2,C5a > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['phase_ms'], d['parity']['dir_vs_oracle'], d['roofline']['frac'])
print('batched', d['batched_solves']['ms_per_step'], 'full_step', d['full_step'])
for w,v in d['other_workloads'].items(): print(w, v.get('ms_per_step'), v.get('phase_ms'), v.get('error'))
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/r02g_pytest_potrf.log 2>&1; echo "pytest potrf rc=$?"
tail -3 gpurun_out/r02g_pytest_potrf.log
timeout 120 python tools/panel_probe.py > gpurun_out/r02g_panel_probe.json 2>&1; cat gpurun_out/r02g_panel_probe.json
HYP_POTRF_DEBUG=1 timeout 300 python tools/potrf_probe.py 1000 4000 10000 20000 > gpurun_out/r02g_potrf_dag.json 2> gpurun_out/r02g_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02g_potrf_dag.json; for m in 4000 10000 20000; do grep "m=$m\]" gpurun_out/r02g_potrf_dag.err | tail -2; done
timeout 600 python -m pytest tests/test_gpu_system.py -x -q -k "multi_column" > gpurun_out/r02g_pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -12 gpurun_out/r02g_pytest_multi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02g_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --other C2 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['dir_vs_oracle'])
print('batched', d['batched_solves'])
for w,v in d['other_workloads'].items(): print(w, v.get('ms_per_step'), v.get('phase_ms'), v.get('batched_solves'), v.get('error'))
PY

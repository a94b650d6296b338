#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/r02d_pytest_potrf.log 2>&1; echo "pytest potrf rc=$?"
tail -3 gpurun_out/r02d_pytest_potrf.log
HYP_POTRF_DEBUG=1 timeout 300 python tools/potrf_probe.py 1000 4000 10000 20000 > gpurun_out/r02d_potrf_dag.json 2> gpurun_out/r02d_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02d_potrf_dag.json; grep "m=10000\|m=4000\|m=20000" gpurun_out/r02d_potrf_dag.err | tail -6
for n in 4 16; do HYP_POTRF_CHAIN_CTAS=$n timeout 300 python tools/potrf_probe.py 4000 10000 > gpurun_out/r02d_potrf_dag_chain$n.json 2>/dev/null; cat gpurun_out/r02d_potrf_dag_chain$n.json; done
timeout 300 python tools/syrk_probe.py > gpurun_out/r02d_syrk_probe.json 2> gpurun_out/r02d_syrk_probe.err; cat gpurun_out/r02d_syrk_probe.json; tail -3 gpurun_out/r02d_syrk_probe.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02d_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --other C2,C5a > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'])
for w,v in d['other_workloads'].items(): print(w, v.get('ms_per_step'), v.get('phase_ms'), v.get('parity'), v.get('error'))
PY

#!/bin/bash
# quad64 kernel (two CTA pairs share their A tiles by multicast, 64-byte rows) vs pair64
mkdir -p gpurun_out
HYP_OZAKI_CLUSTER=5 timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/r02r_pytest_quad64.log 2>&1; echo "pytest quad64 rc=$?"; tail -3 gpurun_out/r02r_pytest_quad64.log
HYP_OZAKI_CLUSTER=5 timeout 200 python tools/syrk_probe.py > gpurun_out/r02r_syrk_probe_quad64.json 2>gpurun_out/r02r_syrk_probe.err; cat gpurun_out/r02r_syrk_probe_quad64.json; tail -2 gpurun_out/r02r_syrk_probe.err
timeout 200 python tools/syrk_probe.py > gpurun_out/r02r_syrk_probe_pair64.json 2>>gpurun_out/r02r_syrk_probe.err; cat gpurun_out/r02r_syrk_probe_pair64.json
HYP_OZAKI_CLUSTER=5 timeout 600 python -m pytest tests/test_gpu_system.py -x -q > gpurun_out/r02r_pytest_system_quad64.log 2>&1; echo "pytest system quad64 rc=$?"; tail -3 gpurun_out/r02r_pytest_system_quad64.log
HYP_OZAKI_CLUSTER=5 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02r_bench_quad64.json 2> gpurun_out/r02r_bench_quad64.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('bench_quad64',):
    try:
        d=json.loads(open(f'gpurun_out/r02r_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

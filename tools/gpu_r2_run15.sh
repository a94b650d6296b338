#!/bin/bash
# pair64 kernel (64-byte k rows) vs pair kernel (32-byte rows): parity, probes, bench; fused pre-pass parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_system.py -x -q > gpurun_out/r02o_pytest_pair64.log 2>&1; echo "pytest pair64 rc=$?"; tail -5 gpurun_out/r02o_pytest_pair64.log
timeout 300 python tools/syrk_probe.py > gpurun_out/r02o_syrk_probe_pair64.json 2>gpurun_out/r02o_syrk_probe.err; cat gpurun_out/r02o_syrk_probe_pair64.json
HYP_OZAKI_CLUSTER=2 timeout 300 python tools/syrk_probe.py > gpurun_out/r02o_syrk_probe_pair32.json 2>>gpurun_out/r02o_syrk_probe.err; cat gpurun_out/r02o_syrk_probe_pair32.json
tail -3 gpurun_out/r02o_syrk_probe.err
timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02o_bench_pair64.json 2> gpurun_out/r02o_bench_pair64.err; echo "bench pair64 rc=$?"
HYP_OZAKI_CLUSTER=2 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02o_bench_pair32.json 2> gpurun_out/r02o_bench_pair32.err; echo "bench pair32 rc=$?"
HYP_NO_FUSED_PREPASS=1 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02o_bench_pair64_nofuse.json 2> gpurun_out/r02o_bench_pair64_nofuse.err; echo "bench nofuse rc=$?"
python - <<'PY'
import json
for f in ('pair64','pair32','pair64_nofuse'):
    try:
        d=json.loads(open(f'gpurun_out/r02o_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

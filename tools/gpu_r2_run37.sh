#!/bin/bash
# two-operand digit-sliced Schur product for mixed / log-det models (K2 branch on tcgen05), SOC pre-pass with hoisted scalars
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_system.py tests/test_gpu_cones.py -x -q > gpurun_out/r02zk_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zk_pytest.log
timeout 600 python bench.py --workload C5a --steps 3 --warmup 2 --other none --no-cpu-baseline > gpurun_out/r02zk_bench_c5a_i8.json 2> gpurun_out/r02zk_bench_c5a_i8.err; echo "c5a i8 rc=$?"
HYP_K2_DMMA=1 timeout 600 python bench.py --workload C5a --steps 3 --warmup 2 --other none --no-cpu-baseline > gpurun_out/r02zk_bench_c5a_dmma.json 2> gpurun_out/r02zk_bench_c5a_dmma.err; echo "c5a dmma rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --other none > gpurun_out/r02zk_bench_c3.json 2> gpurun_out/r02zk_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
for f in ('bench_c5a_i8','bench_c5a_dmma','bench_c3'):
    try:
        d=json.loads(open(f'gpurun_out/r02zk_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'])
    except Exception as e: print(f, 'failed', e)
PY

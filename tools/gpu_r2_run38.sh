#!/bin/bash
# two-operand digit-sliced Schur product on C5a (1 GPU) vs the FP64 DMMA product; C3 after reverting the pre-pass change
mkdir -p gpurun_out
timeout 600 python bench.py --workload C5a --steps 3 --warmup 2 --other none --no-cpu-baseline > gpurun_out/r02zl_bench_c5a_i8.json 2> gpurun_out/r02zl_bench_c5a_i8.err; echo "c5a i8 rc=$?"
HYP_K2_DMMA=1 timeout 600 python bench.py --workload C5a --steps 3 --warmup 2 --other none --no-cpu-baseline > gpurun_out/r02zl_bench_c5a_dmma.json 2> gpurun_out/r02zl_bench_c5a_dmma.err; echo "c5a dmma rc=$?"
timeout 600 python -m pytest tests/test_gpu_system.py -x -q -k "4_and_5 or mixed or C5" > gpurun_out/r02zl_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zl_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02zl_bench_c3.json 2> gpurun_out/r02zl_bench_c3.err; echo "c3 rc=$?"
python - <<'PY'
import json
for f in ('bench_c5a_i8','bench_c5a_dmma','bench_c3'):
    try:
        d=json.loads(open(f'gpurun_out/r02zl_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'].get('kkt_residual_device_operator'))
    except Exception as e: print(f, 'failed', e)
PY

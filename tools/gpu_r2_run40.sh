#!/bin/bash
# congruences of large matrix cones on tcgen05 (hyp_ozaki_gemm_tn): parity with the threshold lowered to side 128, C5a at full size
mkdir -p gpurun_out
HYP_CONG_I8_MIN=128 timeout 900 python -m pytest tests/test_gpu_cones.py tests/test_gpu_system.py -x -q > gpurun_out/r02zm_pytest_i8min128.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zm_pytest_i8min128.log
timeout 600 python -m pytest tests/test_gpu_system.py -x -q -k "4_and_5" > gpurun_out/r02zm_pytest_c5a.log 2>&1; echo "pytest c5a rc=$?"; tail -3 gpurun_out/r02zm_pytest_c5a.log
timeout 600 python bench.py --workload C5a --steps 3 --warmup 2 --other none --no-cpu-baseline > gpurun_out/r02zm_bench_c5a_i8.json 2> gpurun_out/r02zm_bench_c5a_i8.err; echo "c5a i8 rc=$?"
HYP_CONG_DMMA=1 timeout 600 python bench.py --workload C5a --steps 3 --warmup 2 --other none --no-cpu-baseline > gpurun_out/r02zm_bench_c5a_congdmma.json 2> gpurun_out/r02zm_bench_c5a_congdmma.err; echo "c5a cong dmma rc=$?"
python - <<'PY'
import json
for f in ('bench_c5a_i8','bench_c5a_congdmma'):
    try:
        d=json.loads(open(f'gpurun_out/r02zm_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'].get('kkt_residual_device_operator'))
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/r02zm_bench_c5a_i8.err

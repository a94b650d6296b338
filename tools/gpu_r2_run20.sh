#!/bin/bash
# segmented TRSV, narrow tiles for the far block rows, outer block 1024 option
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02t_pytest.log
HYP_POTRF_OB=1024 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k potrf > gpurun_out/r02t_pytest_ob1024.log 2>&1; echo "pytest ob1024 rc=$?"; tail -2 gpurun_out/r02t_pytest_ob1024.log
timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02t_potrf_i8.json 2> gpurun_out/r02t_potrf_i8.err; echo "probe i8 rc=$?"; cat gpurun_out/r02t_potrf_i8.json; tail -3 gpurun_out/r02t_potrf_i8.err
HYP_POTRF_OB=1024 timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02t_potrf_i8_ob1024.json 2> gpurun_out/r02t_potrf_i8_ob1024.err; cat gpurun_out/r02t_potrf_i8_ob1024.json
timeout 600 python -m pytest tests/test_gpu_system.py -x -q > gpurun_out/r02t_pytest_system.log 2>&1; echo "pytest system rc=$?"; tail -2 gpurun_out/r02t_pytest_system.log
timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"
HYP_TRSV_SEG=0 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02t_bench_oldtrsv.json 2> gpurun_out/r02t_bench_oldtrsv.err; echo "bench old trsv rc=$?"
HYP_POTRF_OB=1024 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02t_bench_ob1024.json 2> gpurun_out/r02t_bench_ob1024.err; echo "bench ob1024 rc=$?"
python - <<'PY'
import json
for f in ('bench','bench_oldtrsv','bench_ob1024'):
    try:
        d=json.loads(open(f'gpurun_out/r02t_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

#!/bin/bash
# blocked Cholesky with tcgen05 (digit-sliced) trailing updates vs the task-graph DMMA kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/r02p_pytest_linalg.log 2>&1; echo "pytest linalg rc=$?"; tail -5 gpurun_out/r02p_pytest_linalg.log
timeout 300 python tools/potrf_probe.py 2000 4000 10000 20000 > gpurun_out/r02p_potrf_i8.json 2> gpurun_out/r02p_potrf_i8.err; echo "probe i8 rc=$?"; cat gpurun_out/r02p_potrf_i8.json; tail -3 gpurun_out/r02p_potrf_i8.err
HYP_POTRF=dag timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02p_potrf_dag.json 2> gpurun_out/r02p_potrf_dag.err; echo "probe dag rc=$?"; cat gpurun_out/r02p_potrf_dag.json
timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/r02p_pytest_ozaki_split3.log 2>&1; echo "pytest ozaki split3 rc=$?"; tail -2 gpurun_out/r02p_pytest_ozaki_split3.log
HYP_OZAKI_SPLIT=4 timeout 300 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/r02p_pytest_ozaki_split4.log 2>&1; echo "pytest ozaki split4 rc=$?"; tail -2 gpurun_out/r02p_pytest_ozaki_split4.log
timeout 300 python tools/syrk_probe.py > gpurun_out/r02p_syrk_probe_split3.json 2>gpurun_out/r02p_syrk_probe.err; cat gpurun_out/r02p_syrk_probe_split3.json
HYP_OZAKI_SPLIT=4 timeout 300 python tools/syrk_probe.py > gpurun_out/r02p_syrk_probe_split4.json 2>>gpurun_out/r02p_syrk_probe.err; cat gpurun_out/r02p_syrk_probe_split4.json
timeout 600 python -m pytest tests/test_gpu_system.py -x -q > gpurun_out/r02p_pytest_system.log 2>&1; echo "pytest system rc=$?"; tail -5 gpurun_out/r02p_pytest_system.log
timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02p_bench_i8.json 2> gpurun_out/r02p_bench_i8.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('i8',):
    try:
        d=json.loads(open(f'gpurun_out/r02p_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

#!/bin/bash
# narrow (128 x 32) DMMA tiles on the Cholesky chain stream + software-pipelined SYRK epilogue
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02s_pytest.log
timeout 200 python tools/syrk_probe.py > gpurun_out/r02s_syrk_probe.json 2>gpurun_out/r02s_syrk_probe.err; cat gpurun_out/r02s_syrk_probe.json
timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02s_potrf_i8.json 2> gpurun_out/r02s_potrf_i8.err; echo "probe i8 rc=$?"; cat gpurun_out/r02s_potrf_i8.json; tail -3 gpurun_out/r02s_potrf_i8.err
HYP_POTRF_TILES=big timeout 300 python tools/potrf_probe.py 10000 > gpurun_out/r02s_potrf_i8_bigtiles.json 2> gpurun_out/r02s_potrf_i8_bigtiles.err; cat gpurun_out/r02s_potrf_i8_bigtiles.json
timeout 600 python -m pytest tests/test_gpu_system.py -x -q > gpurun_out/r02s_pytest_system.log 2>&1; echo "pytest system rc=$?"; tail -2 gpurun_out/r02s_pytest_system.log
timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('bench',):
    try:
        d=json.loads(open(f'gpurun_out/r02s_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

#!/bin/bash
# epilogue fix (batched C loads + L2 prefetch): Schur SYRK and the tcgen05 Cholesky again, with the Cholesky's phase probes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02q_pytest.log
timeout 300 python tools/syrk_probe.py > gpurun_out/r02q_syrk_probe.json 2>gpurun_out/r02q_syrk_probe.err; cat gpurun_out/r02q_syrk_probe.json
timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02q_potrf_i8.json 2> gpurun_out/r02q_potrf_i8.err; echo "probe i8 rc=$?"; cat gpurun_out/r02q_potrf_i8.json; tail -3 gpurun_out/r02q_potrf_i8.err
timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"
HYP_POTRF=dag timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02q_bench_dag.json 2> gpurun_out/r02q_bench_dag.err; echo "bench dag rc=$?"
python - <<'PY'
import json
for f in ('bench','bench_dag'):
    try:
        d=json.loads(open(f'gpurun_out/r02q_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

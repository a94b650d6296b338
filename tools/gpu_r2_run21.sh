#!/bin/bash
# TRSV variants (L2 prefetch of the next tile; segmented with early partial sums), column maxima in the pre-pass, adaptive OB
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ozaki.py -x -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02u_pytest.log
HYP_TRSV_SEG=8 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf or potrs" > gpurun_out/r02u_pytest_seg8.log 2>&1; echo "pytest seg8 rc=$?"; tail -2 gpurun_out/r02u_pytest_seg8.log
timeout 600 python -m pytest tests/test_gpu_system.py -x -q > gpurun_out/r02u_pytest_system.log 2>&1; echo "pytest system rc=$?"; tail -2 gpurun_out/r02u_pytest_system.log
for v in 0 8 16 4; do
HYP_TRSV_SEG=$v timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02u_bench_seg$v.json 2> gpurun_out/r02u_bench_seg$v.err; echo "bench seg$v rc=$?"
done
HYP_NO_PREPASS_COLMAX=1 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02u_bench_nocolmax.json 2> gpurun_out/r02u_bench_nocolmax.err; echo "bench nocolmax rc=$?"
python - <<'PY'
import json
for f in ('bench_seg0','bench_seg8','bench_seg16','bench_seg4','bench_nocolmax'):
    try:
        d=json.loads(open(f'gpurun_out/r02u_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks']['sm_mhz'], d['parity'].get('kkt_residual'))
    except Exception as e: print(f, 'failed', e)
PY

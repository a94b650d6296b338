#!/bin/bash
# TRSV chain-step tweaks (own right-hand-side block prefetched, fence only in the storing threads), smoke()
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02zo_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02zo_smoke.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_system.py -x -q > gpurun_out/r02zo_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02zo_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --other none > gpurun_out/r02zo_bench.json 2> gpurun_out/r02zo_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zo_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['dir_vs_oracle'], d['batched_solves']['ms_per_step'])
PY

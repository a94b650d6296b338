#!/bin/bash
# SOC pre-pass: per-cone constants hoisted into registers, loops unrolled by 8, column maxima collected in the compute loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cones.py tests/test_gpu_system.py tests/test_gpu_ozaki.py -q -m gpu -x --timeout 300 > gpurun_out/r02zw_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zw_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --other none > gpurun_out/r02zw_bench.json 2> gpurun_out/r02zw_bench.err; echo "bench rc=$?"
HYP_FUSED_PREPASS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --other none > gpurun_out/r02zw_bench_fused.json 2> gpurun_out/r02zw_bench_fused.err; echo "bench fused rc=$?"
python - <<'PY'
import json
for f in ("r02zw_bench", "r02zw_bench_fused"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY

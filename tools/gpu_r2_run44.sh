#!/bin/bash
# byte-parallel radix-256 slicing + batched loads of the SOC pre-pass: parity tests, then C3 bench (default and fused pre-pass)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_system.py tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/r02zs_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zs_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --other none > gpurun_out/r02zs_bench.json 2> gpurun_out/r02zs_bench.err; echo "bench rc=$?"
HYP_FUSED_PREPASS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --other none > gpurun_out/r02zs_bench_fused.json 2> gpurun_out/r02zs_bench_fused.err; echo "bench fused rc=$?"
python - <<'PY'
import json
for f in ("r02zs_bench", "r02zs_bench_fused"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"], d["parity"]["dir_vs_oracle"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY

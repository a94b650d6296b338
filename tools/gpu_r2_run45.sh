#!/bin/bash
# packet variant of the triangular solves: kernel + system parity tests (per-test timeout: a lost packet would spin), then
# C3 bench A/B on the same box (HYP_TRSV_PKT=0 = the flag protocol)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_system.py tests/test_gpu_solve.py -q -m gpu -x --timeout 180 > gpurun_out/r02zt_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zt_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --other C2 > gpurun_out/r02zt_bench_pkt.json 2> gpurun_out/r02zt_bench_pkt.err; echo "bench rc=$?"
HYP_TRSV_PKT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --other C2 > gpurun_out/r02zt_bench_flag.json 2> gpurun_out/r02zt_bench_flag.err; echo "bench flag rc=$?"
python - <<'PY'
import json
for f in ("r02zt_bench_pkt", "r02zt_bench_flag"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"], d["batched_solves"]["ms_per_step"], d["clocks"]["sm_mhz"])
        c2 = d["other_workloads"]["C2"]
        print("  C2", c2["ms_per_step"], c2["phase_ms"], c2["parity"])
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# tile order of the Schur SYRK: N row pairs resident (HYP_OZAKI_ORDER=row<N>), same box back to back
mkdir -p gpurun_out
for ord in row row4 row6 row8 row; do
  HYP_OZAKI_ORDER=$ord timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --other none > gpurun_out/r02zx_bench_$ord.json 2> gpurun_out/r02zx_bench_$ord.err; echo "$ord rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02zx_bench_$ord.json").read().strip().splitlines()[-1])
print("$ord", round(d["ms_per_step"], 2), round(d["roofline"]["phase_ms"]["schur_syrk"], 2), d["parity"]["kkt_residual_device_operator"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
PY
done

#!/bin/bash
# localise the steady-state parity failure of the tcgen05 Cholesky with narrow tiles (C2: fast CPU leg)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload C2 --steps 3 --warmup 2 --other none > gpurun_out/r02y_$name.json 2> gpurun_out/r02y_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02y_$name.json').read().strip().splitlines()[-1])
    print('$name', d['ms_per_step'], d['parity']['dir_vs_oracle'], d['roofline']['phase_ms'].get('potrf'))
except Exception as e: print('$name', 'failed', e)
PY
}
run default A=1
run chainonly HYP_POTRF_TILES=chain
run bulkonly HYP_POTRF_TILES=bulk
run sync1 HYP_POTRF_SYNC=1
run sync2 HYP_POTRF_SYNC=2
run sync4 HYP_POTRF_SYNC=4
run sync8 HYP_POTRF_SYNC=8
run sync15 HYP_POTRF_SYNC=15
run nomode HYP_POTRF_MODE=0

#!/bin/bash
# which dependency is missing?  steady-state residual of the tcgen05 Cholesky with device-wide syncs at different points
mkdir -p gpurun_out
for v in 0 1 2 4 8; do
HYP_POTRF_MODE=0 HYP_POTRF_SYNC=$v timeout 200 python tools/potrf_probe.py 5000 10000 > gpurun_out/r02x_probe_sync$v.json 2> gpurun_out/r02x_probe_sync$v.err; echo "sync=$v"; python - <<PY
import json
d=json.loads(open('gpurun_out/r02x_probe_sync$v.json').read().strip().splitlines()[-1])['potrf_probe']
for m,r in d.items(): print(m, r.get('residual'), r.get('residual_steady_state'), r.get('full_ms'))
PY
done
HYP_POTRF_TILES=big timeout 200 python tools/potrf_probe.py 5000 10000 > gpurun_out/r02x_probe_big.json 2> gpurun_out/r02x_probe_big.err; cat gpurun_out/r02x_probe_big.json

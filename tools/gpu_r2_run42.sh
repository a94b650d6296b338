#!/bin/bash
# determinism stress of the shipped Cholesky configuration at the sizes not covered before (outer block 1024 from m = 16384 on, small m)
mkdir -p gpurun_out
for m in 1100 2000 7000 16384 20000; do
reps=40; if [ $m -ge 16000 ]; then reps=25; fi
timeout 400 python tools/potrf_race.py $m $reps >> gpurun_out/r02zp_race.jsonl 2>> gpurun_out/r02zp_race.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02zp_race.jsonl'):
    d=json.loads(l)
    print(d['m'], d['reps'], d['env'], d['n_bad'], d['residual_first'], [ (b['rep'], b['first_tiles'][:3], round(b['rel'],5)) for b in d['bad'][:3]])
PY

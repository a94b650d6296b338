#!/bin/bash
# control (old event position) first: is this box one that shows the race?  then the new event position
mkdir -p gpurun_out
for cfg in "HYP_POTRF_NEAR_EVENT=early HYP_POTRF_NARROW_MASK=8" "HYP_POTRF_NARROW_MASK=8" "HYP_POTRF_NEAR_EVENT=early" "A=1"; do
for m in 5000 10000; do
env $cfg timeout 250 python tools/potrf_race.py $m 60 >> gpurun_out/r02ze_race.jsonl 2>> gpurun_out/r02ze_race.err
done; done
python - <<'PY'
import json
for l in open('gpurun_out/r02ze_race.jsonl'):
    d=json.loads(l)
    print(d['m'], d['env'], d['n_bad'], [ (b['rep'], b['first_tiles'][:3], round(b['rel'],5)) for b in d['bad'][:3]])
PY

#!/bin/bash
mkdir -p gpurun_out
for mode in 4 5; do
HYP_POTRF_NARROW_MASK=8 HYP_POTRF_MODE=$mode timeout 250 python tools/potrf_race.py 5000 60 >> gpurun_out/r02zd_race.jsonl 2>> gpurun_out/r02zd_race.err
tail -1 gpurun_out/r02zd_race.jsonl | cut -c1-300
done

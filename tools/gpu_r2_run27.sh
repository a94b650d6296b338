#!/bin/bash
# which narrow launch races?  determinism stress per launch class
mkdir -p gpurun_out
for mask in 1 2 4 8 3 12; do
for m in 10000; do
HYP_POTRF_NARROW_MASK=$mask timeout 200 python tools/potrf_race.py $m 60 >> gpurun_out/r02za_race.jsonl 2>> gpurun_out/r02za_race.err
tail -1 gpurun_out/r02za_race.jsonl | cut -c1-400
done; done

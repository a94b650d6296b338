#!/bin/bash
# determinism stress after routing the panel kernel's and the DMMA epilogue's reads of C through L2
mkdir -p gpurun_out
for cfg in "HYP_POTRF_NARROW_MASK=8" "A=1" "HYP_POTRF_NARROW_MASK=2"; do
for m in 10000 5000; do
env $cfg timeout 250 python tools/potrf_race.py $m 80 >> gpurun_out/r02zb_race.jsonl 2>> gpurun_out/r02zb_race.err
tail -1 gpurun_out/r02zb_race.jsonl | cut -c1-300
done; done

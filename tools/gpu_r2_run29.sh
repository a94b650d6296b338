#!/bin/bash
# near-UPD narrow (mask 8) determinism with parts of the bulk stream switched off (results are garbage but must repeat)
mkdir -p gpurun_out
for mode in 2 1 0; do
HYP_POTRF_NARROW_MASK=8 HYP_POTRF_MODE=$mode timeout 250 python tools/potrf_race.py 5000 60 >> gpurun_out/r02zc_race.jsonl 2>> gpurun_out/r02zc_race.err
tail -1 gpurun_out/r02zc_race.jsonl | cut -c1-300
done
HYP_POTRF_NARROW_MASK=8 HYP_POTRF_SYNC=2 timeout 250 python tools/potrf_race.py 5000 60 >> gpurun_out/r02zc_race.jsonl 2>> gpurun_out/r02zc_race.err
tail -1 gpurun_out/r02zc_race.jsonl | cut -c1-300
HYP_POTRF_NARROW_MASK=8 HYP_POTRF_SYNC=4 timeout 250 python tools/potrf_race.py 5000 60 >> gpurun_out/r02zc_race.jsonl 2>> gpurun_out/r02zc_race.err
tail -1 gpurun_out/r02zc_race.jsonl | cut -c1-300

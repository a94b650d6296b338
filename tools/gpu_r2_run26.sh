#!/bin/bash
# steady-state determinism of the blocked Cholesky: any bitwise difference between repeated factorisations is a race
mkdir -p gpurun_out
for cfg in "A=1" "HYP_POTRF_TILES=chain" "HYP_POTRF_TILES=bulk" "HYP_POTRF_TILES=big" "HYP_POTRF=dag"; do
for m in 5000 10000; do
env $cfg timeout 200 python tools/potrf_race.py $m 40 >> gpurun_out/r02z_race.jsonl 2>> gpurun_out/r02z_race.err
tail -1 gpurun_out/r02z_race.jsonl | cut -c1-600
done; done

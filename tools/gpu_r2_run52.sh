#!/bin/bash
# launch list of the final build (per-launch durations; shares, not absolutes)
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02zz_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --other none > gpurun_out/r02zz_launches_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02zz_launches.csv

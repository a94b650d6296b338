#!/bin/bash
# DRAM traffic of the Schur SYRK per tile order (ncu metrics pass over the three K-chunk launches of one SYRK)
mkdir -p gpurun_out
for ord in row row6; do
  HYP_OZAKI_ORDER=$ord timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --clock-control none -k regex:ozaki_syrk_pair64 -s 0 -c 3 --csv --log-file gpurun_out/r02zy_ncu_$ord.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --other none > /dev/null 2>&1
  echo "$ord rc=$?"
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02zy_ncu_$ord.csv")) if len(r) > 10 and r[0].isdigit()]
tot = {}
for r in rows:
    name, val = r[-3], float(r[-1].replace(",", ""))
    unit = r[-2]
    tot.setdefault((name, unit), []).append(val)
print("$ord", {k: [round(x, 3) for x in v] for k, v in tot.items()})
PY
done

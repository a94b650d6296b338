#!/bin/bash
# final Cholesky configuration (narrow tiles for the triangular-solve-type launches only, late near event): long stress,
# probe, parity benches
mkdir -p gpurun_out
for m in 3000 5000 10000; do
timeout 400 python tools/potrf_race.py $m 120 >> gpurun_out/r02zf_race.jsonl 2>> gpurun_out/r02zf_race.err
done
HYP_POTRF_NEAR_EVENT=early HYP_POTRF_NARROW_MASK=8 timeout 250 python tools/potrf_race.py 5000 40 >> gpurun_out/r02zf_race.jsonl 2>> gpurun_out/r02zf_race.err
python - <<'PY'
import json
for l in open('gpurun_out/r02zf_race.jsonl'):
    d=json.loads(l)
    print(d['m'], d['reps'], d['env'], d['n_bad'], [ (b['rep'], b['first_tiles'][:3], round(b['rel'],5)) for b in d['bad'][:3]])
PY
timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02zf_potrf.json 2> gpurun_out/r02zf_potrf.err; cat gpurun_out/r02zf_potrf.json
timeout 300 python bench.py --workload C2 --steps 3 --warmup 2 --other none > gpurun_out/r02zf_bench_c2.json 2> gpurun_out/r02zf_bench_c2.err
timeout 900 python bench.py --steps 5 --warmup 3 --other none > gpurun_out/r02zf_bench_c3.json 2> gpurun_out/r02zf_bench_c3.err
python - <<'PY'
import json
for f in ('bench_c2','bench_c3'):
    try:
        d=json.loads(open(f'gpurun_out/r02zf_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['parity']['dir_vs_oracle'], d['roofline']['phase_ms'])
    except Exception as e: print(f, 'failed', e)
PY

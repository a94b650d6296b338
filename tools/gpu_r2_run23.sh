#!/bin/bash
# bisect the parity failure of run 22 on C2 (fast CPU leg)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload C2 --steps 3 --warmup 2 --other none > gpurun_out/r02w_$name.json 2> gpurun_out/r02w_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02w_$name.json').read().strip().splitlines()[-1])
    print('$name', d['ms_per_step'], d['parity']['dir_vs_oracle'], d['roofline']['phase_ms'].get('potrf'))
except Exception as e: print('$name', 'failed', e)
PY
}
run default A=1
run dag HYP_POTRF=dag
run bigtiles HYP_POTRF_TILES=big
run nocolmax HYP_NO_PREPASS_COLMAX=1
run pair32 HYP_OZAKI_CLUSTER=2
run split4 HYP_OZAKI_SPLIT=4
run dmma HYP_SCHUR_SYRK=dmma
run dag_pair32 HYP_POTRF=dag HYP_OZAKI_CLUSTER=2

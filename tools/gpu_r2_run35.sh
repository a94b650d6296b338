#!/bin/bash
# 4 GPUs: bench N = 4 with the extra workloads (C2, C4 row-sharded = BASELINE config 4 at its GPU count, C5a column-sharded, C5b)
mkdir -p gpurun_out
nvidia-smi -L | head -5
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02zi_bench_n4.json 2> gpurun_out/r02zi_bench_n4.err; echo "bench n4 rc=$?"
tail -3 gpurun_out/r02zi_bench_n4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02zi_bench_n4.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'])
    for w,v in d['other_workloads'].items():
        if isinstance(v, dict): print(w, v.get('ms_per_step'), v.get('sharding'), v.get('phase_ms'), v.get('parity'), v.get('error'))
        else: print(w, v)
except Exception as e: print("parse failed", e)
PY

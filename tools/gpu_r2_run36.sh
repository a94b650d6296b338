#!/bin/bash
# 8 GPUs: bench N = 8 (C3 headline + C5b = BASELINE config 5 at its GPU count)
mkdir -p gpurun_out
nvidia-smi -L | head -9
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 5 --warmup 3 --other C5b,C4 > gpurun_out/r02zj_bench_n8.json 2> gpurun_out/r02zj_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r02zj_bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02zj_bench_n8.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'])
    for w,v in d['other_workloads'].items():
        if isinstance(v, dict): print(w, v.get('ms_per_step'), v.get('sharding'), v.get('phase_ms'), v.get('parity'), v.get('error'))
        else: print(w, v)
except Exception as e: print("parse failed", e)
PY

#!/bin/bash
# 2-GPU session: panel phase probe, NCCL tests (row sharding with packed upper-triangle allreduce, column sharding), bench N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 120 python tools/panel_probe.py > gpurun_out/r02e_panel_probe.json 2>&1; cat gpurun_out/r02e_panel_probe.json
timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu > gpurun_out/r02e_pytest_dist.log 2>&1; echo "pytest dist rc=$?"
tail -15 gpurun_out/r02e_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/r02e_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02e_bench_n2.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity'])
    for w,v in d['other_workloads'].items(): print(w, v.get('ms_per_step'), v.get('sharding'), v.get('phase_ms'), v.get('parity'), v.get('error'))
except Exception as e: print("parse failed", e)
PY

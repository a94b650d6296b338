#!/bin/bash
# 2 GPUs: NCCL parity tests (row sharding, column sharding) + bench N = 2 with the final build
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 700 python -m pytest tests/test_dist.py -m gpu -q -p no:cacheprovider > gpurun_out/r02zh_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -4 gpurun_out/r02zh_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --other none > gpurun_out/r02zh_bench_n2.json 2> gpurun_out/r02zh_bench_n2.err; echo "bench2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02zh_bench_ref_n2.json 2> gpurun_out/r02zh_bench_ref_n2.err; echo "ref2 rc=$?"; tail -c 400 gpurun_out/r02zh_bench_ref_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zh_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['roofline']['phase_ms'], d['parity'])
PY

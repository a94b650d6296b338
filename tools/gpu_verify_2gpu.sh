set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
( timeout 600 python -m pytest tests/test_dist.py -m gpu -q -p no:cacheprovider --timeout=500 > gpurun_out/verify2_pytest_dist.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/verify2_pytest_dist.log )
tail -6 gpurun_out/verify2_pytest_dist.log
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/verify2_bench_c3_2gpu.json 2> gpurun_out/verify2_bench_c3_2gpu.err ; echo "bench2 rc=$?" )
tail -3 gpurun_out/verify2_bench_c3_2gpu.err; cut -c1-400 gpurun_out/verify2_bench_c3_2gpu.json

#!/bin/bash
# round-2 GPU session 2: task-graph Cholesky - correctness first (small sizes under timeout), then timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/r02b_pytest_potrf.log 2>&1; echo "pytest potrf rc=$?"
tail -15 gpurun_out/r02b_pytest_potrf.log
timeout 300 python tools/potrf_probe.py 300 1000 2500 4000 10000 20000 > gpurun_out/r02b_potrf_dag.json 2> gpurun_out/r02b_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02b_potrf_dag.json; tail -5 gpurun_out/r02b_potrf_dag.err
HYP_POTRF=stream timeout 300 python tools/potrf_probe.py 4000 10000 20000 > gpurun_out/r02b_potrf_stream.json 2> gpurun_out/r02b_potrf_stream.err
cat gpurun_out/r02b_potrf_stream.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --other C2 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
head -c 1500 gpurun_out/r02b_bench_n1.json

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/r02c_pytest_potrf.log 2>&1; echo "pytest potrf rc=$?"
tail -3 gpurun_out/r02c_pytest_potrf.log
HYP_POTRF_DEBUG=1 timeout 300 python tools/potrf_probe.py 1000 4000 10000 20000 > gpurun_out/r02c_potrf_dag.json 2> gpurun_out/r02c_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02c_potrf_dag.json; grep "m=10000\|m=4000\|m=20000" gpurun_out/r02c_potrf_dag.err | tail -12
for n in 4 16; do HYP_POTRF_CHAIN_CTAS=$n timeout 300 python tools/potrf_probe.py 4000 10000 > gpurun_out/r02c_potrf_dag_chain$n.json 2>/dev/null; cat gpurun_out/r02c_potrf_dag_chain$n.json; done
timeout 200 python tools/state_probe.py 1000 > gpurun_out/r02c_state_probe.log 2>&1; cat gpurun_out/r02c_state_probe.log
HYP_POTRF=stream timeout 200 python tools/state_probe.py 1000 > gpurun_out/r02c_state_probe_stream.log 2>&1; cat gpurun_out/r02c_state_probe_stream.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02c_pytest_gpu.log

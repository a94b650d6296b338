#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/panel_probe.py > gpurun_out/r02l_panel_probe.json 2>&1; cat gpurun_out/r02l_panel_probe.json
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/r02l_pytest_potrf.log 2>&1; echo "pytest potrf rc=$?"
HYP_POTRF_DEBUG=1 timeout 300 python tools/potrf_probe.py 1000 4000 10000 20000 > gpurun_out/r02l_potrf_dag.json 2> gpurun_out/r02l_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02l_potrf_dag.json; for m in 4000 10000; do grep "m=$m\]" gpurun_out/r02l_potrf_dag.err | tail -2; done
for n in 4 16; do HYP_POTRF_CHAIN_CTAS=$n timeout 300 python tools/potrf_probe.py 4000 10000 > gpurun_out/r02l_potrf_dag_chain$n.json 2>/dev/null; cat gpurun_out/r02l_potrf_dag_chain$n.json; done

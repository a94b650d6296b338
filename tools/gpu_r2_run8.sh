#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "potrf" > gpurun_out/r02h_pytest_potrf.log 2>&1; echo "pytest potrf rc=$?"
timeout 120 python tools/panel_probe.py > gpurun_out/r02h_panel_probe.json 2>&1; cat gpurun_out/r02h_panel_probe.json
HYP_POTRF_DEBUG=1 timeout 300 python tools/potrf_probe.py 1000 4000 10000 20000 > gpurun_out/r02h_potrf_dag.json 2> gpurun_out/r02h_potrf_dag.err; echo "probe rc=$?"
cat gpurun_out/r02h_potrf_dag.json; for m in 4000 10000; do grep "m=$m\]" gpurun_out/r02h_potrf_dag.err | tail -2; done
timeout 300 python tools/mma_probe.py > gpurun_out/r02h_mma_probe.json 2> gpurun_out/r02h_mma_probe.err; echo "mma probe rc=$?"; tail -3 gpurun_out/r02h_mma_probe.err
grep -v '"mma_probe"' gpurun_out/r02h_mma_probe.json | cut -c1-300

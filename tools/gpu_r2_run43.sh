#!/bin/bash
# per-cone single-block entry points (hyp_cone_*): parity against the oracle's per-cone objects
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cone_single.py -q -m gpu > gpurun_out/r02zr_pytest_cone_single.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02zr_pytest_cone_single.log

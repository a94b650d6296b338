#!/bin/bash
# full solve to convergence at the natvsext tolerances with the device plug-ins: scaled C5b mix (x 0.2 and x 0.3) and a C3-shaped SOC model (x 0.3)
mkdir -p gpurun_out
timeout 600 python tools/solve_bench.py --impl device --config C5b --scale 0.2 > gpurun_out/r02_fullsolve_c5b_x0.2_device.json 2> gpurun_out/r02_fullsolve_c5b_x0.2_device.err; echo "rc=$?"; cut -c1-900 gpurun_out/r02_fullsolve_c5b_x0.2_device.json; tail -2 gpurun_out/r02_fullsolve_c5b_x0.2_device.err
timeout 900 python tools/solve_bench.py --impl device --config C5b --scale 0.3 > gpurun_out/r02_fullsolve_c5b_x0.3_device.json 2> gpurun_out/r02_fullsolve_c5b_x0.3_device.err; echo "rc=$?"; cut -c1-900 gpurun_out/r02_fullsolve_c5b_x0.3_device.json; tail -2 gpurun_out/r02_fullsolve_c5b_x0.3_device.err
timeout 900 python tools/solve_bench.py --impl device --config C3 --scale 0.3 > gpurun_out/r02_fullsolve_c3_x0.3_device.json 2> gpurun_out/r02_fullsolve_c3_x0.3_device.err; echo "rc=$?"; cut -c1-900 gpurun_out/r02_fullsolve_c3_x0.3_device.json; tail -2 gpurun_out/r02_fullsolve_c3_x0.3_device.err

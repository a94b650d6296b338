#!/bin/bash
# round-2 GPU session 1: potrf probe, -m gpu suite, default bench line, ncu --set full of the Schur SYRK, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_gpu.txt
free -g >> gpurun_out/r02_gpu.txt; nproc >> gpurun_out/r02_gpu.txt
python tools/potrf_probe.py 4000 10000 > gpurun_out/r02_potrf_probe.json 2> gpurun_out/r02_potrf_probe.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_syrk_pair -c 3 -o gpurun_out/r02_syrk_pair \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --other none > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --other none > gpurun_out/r02_launch_bench.log 2>&1; echo "launches rc=$?"
cat gpurun_out/r02_potrf_probe.json
head -c 3000 gpurun_out/r02_bench_n1.json

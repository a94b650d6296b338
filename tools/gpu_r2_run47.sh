#!/bin/bash
# ncu --set full of the second-order-cone pre-pass and the slicing kernel (one launch each, after warm-up)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'soc_prod_chunk_kernel|slice256_kernel' -s 2 -c 2 \
  -o gpurun_out/r02zv_prepass -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --other none > gpurun_out/r02zv_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02zv_ncu.log; ls -la gpurun_out/r02zv_prepass.ncu-rep

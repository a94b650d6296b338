#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -c 3 -o gpurun_out/r02k_panel python tools/panel_probe.py > gpurun_out/r02k_ncu_panel.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02k_ncu_panel.log

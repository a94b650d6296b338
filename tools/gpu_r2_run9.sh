#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/panel_probe.py > gpurun_out/r02i_panel_probe.json 2>&1; cat gpurun_out/r02i_panel_probe.json
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/r02i_pytest_ozaki.log 2>&1; echo "pytest ozaki rc=$?"; tail -3 gpurun_out/r02i_pytest_ozaki.log
timeout 300 python tools/syrk_probe.py > gpurun_out/r02i_syrk_probe_new.json 2>gpurun_out/r02i_syrk_probe.err; cat gpurun_out/r02i_syrk_probe_new.json
HYP_OZAKI_OLD_ISSUE=1 timeout 300 python tools/syrk_probe.py > gpurun_out/r02i_syrk_probe_old.json 2>>gpurun_out/r02i_syrk_probe.err; cat gpurun_out/r02i_syrk_probe_old.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02i_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --other C2 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['parity']['dir_vs_oracle'], d['roofline']['frac'])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/panel_probe.py > gpurun_out/r02j_panel_probe.json 2>&1; cat gpurun_out/r02j_panel_probe.json
HYP_PANEL_FLAGS=1 timeout 120 python tools/panel_probe.py > gpurun_out/r02j_panel_probe_noearly.json 2>&1; cat gpurun_out/r02j_panel_probe_noearly.json
timeout 300 python tools/syrk_probe.py > gpurun_out/r02j_syrk_probe_pair.json 2>gpurun_out/r02j_syrk_probe.err; cat gpurun_out/r02j_syrk_probe_pair.json
HYP_OZAKI_CLUSTER=3 timeout 300 python tools/syrk_probe.py > gpurun_out/r02j_syrk_probe_quad.json 2>>gpurun_out/r02j_syrk_probe.err; cat gpurun_out/r02j_syrk_probe_quad.json
HYP_OZAKI_CLUSTER=3 timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_system.py -x -q > gpurun_out/r02j_pytest_quad.log 2>&1; echo "pytest quad rc=$?"; tail -3 gpurun_out/r02j_pytest_quad.log
HYP_OZAKI_CLUSTER=3 timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02j_bench_quad.json 2> gpurun_out/r02j_bench_quad.err; echo "bench quad rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --other none --no-cpu-baseline > gpurun_out/r02j_bench_pair.json 2> gpurun_out/r02j_bench_pair.err; echo "bench pair rc=$?"
python - <<'PY'
import json
for f in ('quad','pair'):
    d=json.loads(open(f'gpurun_out/r02j_bench_{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['roofline']['phase_ms'], d['clocks'])
PY

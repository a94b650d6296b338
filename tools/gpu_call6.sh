set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( HYP_OZAKI_CLUSTER=3 timeout -s KILL 150 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -p no:cacheprovider -k "syrk_matches" > gpurun_out/r6_quad_ozaki.log 2>&1 ; echo "quad ozaki rc=$?" >> gpurun_out/r6_quad_ozaki.log )
tail -12 gpurun_out/r6_quad_ozaki.log
nvidia-smi --query-gpu=name,utilization.gpu,memory.used --format=csv
if grep -q "passed" gpurun_out/r6_quad_ozaki.log && ! grep -q "failed" gpurun_out/r6_quad_ozaki.log; then
  ( HYP_OZAKI_CLUSTER=3 timeout -s KILL 300 python -m pytest tests/test_gpu_system.py -m gpu -q -p no:cacheprovider -k "vector_cone or schur_matrix or epipersquare" > gpurun_out/r6_quad_system.log 2>&1 ; echo "quad system rc=$?" >> gpurun_out/r6_quad_system.log )
  tail -5 gpurun_out/r6_quad_system.log
  ( timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_pair.json 2> gpurun_out/r6_bench_pair.err ; echo "pair rc=$?" )
  ( HYP_OZAKI_CLUSTER=3 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench_quad.json 2> gpurun_out/r6_bench_quad.err ; echo "quad rc=$?" )
  python - <<'PY'
import json
for f in ("gpurun_out/r6_bench_pair.json","gpurun_out/r6_bench_quad.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"],3), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms"].items()}, d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
fi

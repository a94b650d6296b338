set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout=240 > gpurun_out/r1_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log )
tail -40 gpurun_out/r1_pytest.log
( timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench_c3.json 2> gpurun_out/r1_bench_c3.err ; echo "bench rc=$?" )
cat gpurun_out/r1_bench_c3.json | cut -c1-600
( timeout 300 python bench.py --workload S1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench_s1.json 2> gpurun_out/r1_bench_s1.err ; echo "s1 rc=$?" )
cat gpurun_out/r1_bench_s1.json | cut -c1-300; tail -5 gpurun_out/r1_bench_s1.err
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:'syevj|spec_mid|spec_post' -c 4 -o gpurun_out/r1_spec_prof python bench.py --workload S1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1_ncu_spec.log 2>&1 ; echo "ncu rc=$?" )
ls -la gpurun_out

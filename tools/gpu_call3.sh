set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --timeout=240 > gpurun_out/r3_pytest.log 2>&1 ; echo "pytest rc=$?" >> gpurun_out/r3_pytest.log )
tail -25 gpurun_out/r3_pytest.log
timeout 300 python tools/debug_kat_seeds.py > gpurun_out/r3_seeds.log 2>&1; cat gpurun_out/r3_seeds.log | tail -12
( timeout 600 python bench.py --workload C5a --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_bench_c5a.json 2> gpurun_out/r3_bench_c5a.err ; echo "c5a rc=$?" )
tail -5 gpurun_out/r3_bench_c5a.err; cut -c1-400 gpurun_out/r3_bench_c5a.json

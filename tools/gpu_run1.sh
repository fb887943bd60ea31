set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; tail -c 600 gpurun_out/r02_bench_v3.json
ZF_LEGACY_KERNEL=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_legacy.json 2> gpurun_out/r02_bench_legacy.err
timeout 300 python bench.py --workload c1_16bit_44k1_60s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c1.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_r02a python bench.py --steps 1 --warmup 3 --profile > gpurun_out/r02_ncu.log 2>&1
tail -3 gpurun_out/r02_ncu.log

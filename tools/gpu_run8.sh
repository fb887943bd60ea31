set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 900 python bench.py --workload c5_24bit_96k_10h_shard8 --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5.json 2>gpurun_out/${TAG}_bench_c5.err
tail -c 1200 gpurun_out/${TAG}_bench_c5.json; tail -3 gpurun_out/${TAG}_bench_c5.err
for i in 1 2 3; do timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 run', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done

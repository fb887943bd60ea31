set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
for i in 1 2; do timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 run', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
timeout 300 python bench.py --workload c3_32bit_192k_600s --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 run', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"

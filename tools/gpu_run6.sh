set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 300 python bench.py --workload c3_32bit_192k_600s --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_c3.json').read().strip().splitlines()[-1])
print('c3', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d.get('parity'))
PY
timeout 600 python -m pytest tests -m gpu -x -q -k "32 or wide" 2>&1 | tail -2

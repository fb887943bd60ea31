set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload c3_32bit_192k_600s --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2>gpurun_out/${TAG}_bench_c3.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2>/dev/null
python - <<PY
import json
for f in ['bench_c3','bench_c2']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['kernel'], d['e2e']['value'], d.get('parity'))
    except Exception as e: print(f, e)
PY
tail -3 gpurun_out/${TAG}_bench_c3.err

# usage: bash tools/gpu_run4.sh TAG   -- parity tests, bench (c2, c1, c3), launch list, full ncu capture of the top kernel
set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload c1_16bit_44k1_60s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c1.json 2>&1
timeout 300 python bench.py --workload c3_32bit_192k_600s --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3.json 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1
python - <<PY
import json
for f in ['bench','bench_c1','bench_c3','bench_ref']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('kernel_ms'), d['e2e']['value'], d.get('parity'))
    except Exception as e: print(f, e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --profile > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log

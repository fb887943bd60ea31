set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
for sg in 0 110; do
ZF_V3_STAGGER=$sg ZF_V3_CTAS=3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_sg$sg.json 2> gpurun_out/${TAG}_bench_sg$sg.err
done
python - <<PY
import json
for sg in [0,110]:
    try:
        d=json.loads(open('gpurun_out/${TAG}_bench_sg%d.json'%sg).read().strip().splitlines()[-1])
        print('stagger',sg, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['parity'])
    except Exception as e: print(sg, e)
PY
ZF_V3_CTAS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_${TAG}a python bench.py --steps 1 --warmup 3 --profile > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log

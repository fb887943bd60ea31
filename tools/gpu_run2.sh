set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest.log
tail -3 gpurun_out/r10_pytest.log
ZF_V3_CTAS=4 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r10_bench_c4.json 2> gpurun_out/r10_bench_c4.err
ZF_V3_CTAS=3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r10_bench_c3.json 2> gpurun_out/r10_bench_c3.err
ZF_V3_CTAS=4 timeout 300 python bench.py --workload c1_16bit_44k1_60s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r10_bench_c1.json 2>&1
python - <<'PY'
import json
for f in ['r10_bench_c4','r10_bench_c3','r10_bench_c1']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['parity'])
    except Exception as e: print(f, e)
PY
ZF_V3_CTAS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1 -o gpurun_out/prof_r10a python bench.py --steps 1 --warmup 3 --profile > gpurun_out/r10_ncu.log 2>&1
tail -2 gpurun_out/r10_ncu.log

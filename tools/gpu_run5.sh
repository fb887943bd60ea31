set -x
cd $GRAFT_REPO_ROOT
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -1 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --workload c1_16bit_44k1_60s --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c1.json 2>&1
python - <<PY
import json
for f in ['bench','bench_c1']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('kernel_ms'), d['e2e'], d.get('parity'))
    except Exception as e: print(f, e)
PY
python - <<PY
# raw PCIe rates on this box, for the e2e ceiling
import torch, time
n=345600000
h=torch.empty(n,dtype=torch.uint8,pin_memory=True); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for _ in range(2): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); print('H2D GB/s', 5*n/(time.perf_counter()-t)/1e9)
t=time.perf_counter()
for _ in range(5): h.copy_(d,non_blocking=True)
torch.cuda.synchronize(); print('D2H GB/s', 5*n/(time.perf_counter()-t)/1e9)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream(); h2=torch.empty(n,dtype=torch.uint8,pin_memory=True); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); print('bidirectional GB/s each', 5*n/(time.perf_counter()-t)/1e9)
PY

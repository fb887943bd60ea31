"""Host <-> device copy ceiling of the end-to-end step at N ranks (one process per GPU, torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py
    python tools/pcie_probe_multi.py                      # N = 1

Every rank moves what one end-to-end step of BASELINE config 2 moves (345.6 MB up, 213.7 MB down) between page-locked host
memory and its GPU, with no kernels: upload alone, download alone, and both at once on two streams -- one cudaMemcpyAsync
per direction.  Rank 0 prints one JSON line: GB/s per rank and in aggregate, max over ranks of the time.  This is the
ceiling bench.py's `e2e` is measured against (`e2e.frac_of_copy_ceiling`); it is a property of the box's PCIe / host-memory
fabric, not of the encoder.
"""
import json
import os
import time

import torch

UP, DOWN = 345_600_000, 213_700_000
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
h_up = torch.empty(UP, dtype=torch.uint8, pin_memory=True)
h_down = torch.empty(DOWN, dtype=torch.uint8, pin_memory=True)
d_up = torch.empty(UP, dtype=torch.uint8, device=dev)
d_down = torch.empty(DOWN, dtype=torch.uint8, device=dev)
s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(up, down, iters=10):
    def once():
        if up:
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s_down):
                h_down.copy_(d_down, non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()
    once()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(iters):
        once()
    dt = (time.perf_counter() - t0) / iters
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {"n_gpus": world, "cpus": os.cpu_count()}
for name, up, down in (("up_only", True, False), ("down_only", False, True), ("both", True, True)):
    dt = run(up, down)
    b = (UP if up else 0) + (DOWN if down else 0)
    res[name] = {"ms": round(dt * 1e3, 3), "gbs_per_rank": round(b / dt / 1e9, 1), "gbs_aggregate": round(b * world / dt / 1e9, 1)}
if rank == 0:
    print(json.dumps(res), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()

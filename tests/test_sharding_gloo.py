"""N > 1 path on CPU: world_size-2 gloo processes shard a stream by contiguous frame ranges exactly as
bench.py / the multi-GPU driver do, each rank encodes its range (with the CPU oracle here --
there is no GPU in this container), and the ordered concatenation must equal the single-process stream,
with the order-dependent min/max frame-size replay done on the gathered sizes.  No data-path collective
exists in the product; gloo is used only to gather the results for checking."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_samples, bits, rate, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import bench
    import oracle_lib
    import zigflac_b200 as zf
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f0, nframes, s0, nsamples = bench.shard_of(total_samples, world, rank)
    pcm = zf.synth_pcm(nsamples, rate, bits, first_sample=s0, threads=1)
    out, sizes = oracle_lib.encode_pcm(pcm, nsamples, oracle_lib.config(2, bits), rate, first_frame_number=f0)
    meta = torch.tensor([f0, nframes, nsamples, out.size], dtype=torch.int64)
    metas = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(metas, meta)
    cap = int(max(m[3] for m in metas))
    cap_f = int(max(m[1] for m in metas))
    pad = torch.zeros(cap, dtype=torch.uint8)
    pad[: out.size] = torch.from_numpy(out)
    pads = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(pads, pad)
    sz = torch.zeros(cap_f, dtype=torch.int64)
    sz[: sizes.size] = torch.from_numpy(sizes.astype(np.int64))
    szs = [torch.zeros(cap_f, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(szs, sz)
    if rank == 0:
        stream = b"".join(pads[r][: int(metas[r][3])].numpy().tobytes() for r in range(world))
        all_sizes = [int(v) for r in range(world) for v in szs[r][: int(metas[r][1])]]
        covered = [(int(m[0]), int(m[1]), int(m[2])) for m in metas]
        q.put((stream, all_sizes, covered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_samples", [4096 * 9 + 100, 4096 * 8])
def test_two_rank_frame_sharding_equals_single_stream(oracle, zf, total_samples):
    import torch.multiprocessing as mp
    bits, rate, world = 16, 44100, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_samples, bits, rate, q)) for r in range(world)]
    for p in procs:
        p.start()
    stream, sizes, covered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile the stream: contiguous frame ranges, only the globally last frame may be short
    assert covered[0][0] == 0 and covered[1][0] == covered[0][1]
    assert covered[0][2] == covered[0][1] * 4096 and covered[0][2] + covered[1][2] == total_samples
    pcm = zf.synth_pcm(total_samples, rate, bits)
    ref, ref_sizes = oracle.encode_pcm(pcm, total_samples, oracle.config(2, bits), rate)
    assert stream == ref.tobytes()
    assert sizes == [int(s) for s in ref_sizes]
    assert oracle.replay_frame_sizes(sizes) == oracle.replay_frame_sizes(ref_sizes)

"""Decode timing probe (development tool): encode a BASELINE config on the GPU, then time the decoder device-resident
(FLAC in HBM -> PCM in HBM) and from / to page-locked host buffers.
usage: python tools/decode_probe.py [c1|c2|c3] [reps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import zigflac_b200 as zf  # noqa: E402
import oracle_lib  # noqa: E402

CFG = {"c1": (16, 44100, 60), "c2": (24, 96000, 600), "c3": (32, 192000, 600)}
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bits, rate, seconds = CFG[name]
n = rate * seconds
pcm = zf.synth_pcm(n, rate, bits)
rc, flac = zf.wav_to_flac(oracle_lib.make_wav(pcm, 2, bits, rate))
assert rc == 0
flac_a = np.frombuffer(flac, dtype=np.uint8)
d_flac = torch.from_numpy(flac_a.copy()).cuda()
d_pcm = torch.zeros(pcm.size, dtype=torch.uint8, device="cuda")
with zf.Decoder() as dec, zf.HostBuffer(flac_a.size) as hin, zf.HostBuffer(pcm.size) as hout:
    hin.array[:] = flac_a
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nb, info = dec.decode_device(d_flac.data_ptr(), d_flac.numel(), d_pcm.data_ptr(), d_pcm.numel())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        got, info2 = dec.decode(hin.array, out=hout.array)
        t2 = time.perf_counter()
        print("%s rep %d: device-resident %.3f ms wall (kernels %.3f ms, %d launches) = %.1f GSamples/s;  host pinned %.3f ms = %.1f GSamples/s"
              % (name, r, (t1 - t0) * 1e3, info["kernel_ms"], info["launches"], 2 * n / (t1 - t0) / 1e9, (t2 - t1) * 1e3,
                 2 * n / (t2 - t1) / 1e9), flush=True)
    assert nb == pcm.size and bytes(d_pcm.cpu().numpy()) == pcm.tobytes()
    assert got.tobytes() == pcm.tobytes()
print("ok")

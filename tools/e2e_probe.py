"""Development probe: end-to-end (pinned host -> FLAC in pinned host) time of config 2 for several batch sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zigflac_b200 as zf
bits, rate, n = 24, 96000, 57600000
h_pcm = torch.empty(n * 6, dtype=torch.uint8, pin_memory=True)
zf.synth_pcm(n, rate, bits, out=h_pcm.numpy())
for per in (256, 512, 1024, 2048, 4096, 8192):
    enc = zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=per)
    cap = enc.max_batch_bytes((n + 4095) // 4096)
    h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        enc.encode_pcm(h_pcm.numpy(), n, 0, out=h_out.numpy())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        got, sizes = enc.encode_pcm(h_pcm.numpy(), n, 0, out=h_out.numpy())
    dt = (time.perf_counter() - t0) / 5
    print("frames/batch %5d: %.3f ms  %.1f MSamples/s" % (per, dt * 1e3, 2 * n / dt / 1e6))
    enc.close()
    del h_out

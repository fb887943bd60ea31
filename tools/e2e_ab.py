"""Development probe: A/B of the batch plan (ZF_NO_TAPER) on one box, interleaved repetitions."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zigflac_b200 as zf
bits, rate, n = 24, 96000, 57600000
h_pcm = torch.empty(n * 6, dtype=torch.uint8, pin_memory=True)
zf.synth_pcm(n, rate, bits, out=h_pcm.numpy())
encs = {}
for name, env in (("taper", "0"), ("plain", "1")):
    os.environ["ZF_NO_TAPER"] = env
    encs[name] = zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=2048)
cap = encs["taper"].max_batch_bytes((n + 4095) // 4096)
h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
ref = None
for rnd in range(4):
    for name, enc in encs.items():
        enc.encode_pcm(h_pcm.numpy(), n, 0, out=h_out.numpy())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            got, sizes = enc.encode_pcm(h_pcm.numpy(), n, 0, out=h_out.numpy())
        dt = (time.perf_counter() - t0) / 5
        sig = (int(got) if not hasattr(got, "__len__") else len(got), int(sizes.sum()), h_out[:64].numpy().tobytes())
        if ref is None:
            ref = sig
        assert sig == ref
        print("round %d %s: %.3f ms  %.2f G samples/s" % (rnd, name, dt * 1e3, 2 * n / dt / 1e9))

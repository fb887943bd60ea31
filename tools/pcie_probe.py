"""Development probe: what the PCIe link alone allows for config 2 (345.6 MB up, ~177 MB down), against the e2e time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zigflac_b200 as zf

bits, rate, n = 24, 96000, 57600000
h_pcm = torch.empty(n * 6, dtype=torch.uint8, pin_memory=True)
zf.synth_pcm(n, rate, bits, out=h_pcm.numpy())
d_pcm = torch.empty(n * 6, dtype=torch.uint8, device="cuda")
n_out = 177_000_000
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
d_out = torch.zeros(n_out, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def up():
    with torch.cuda.stream(s_up):
        d_pcm.copy_(h_pcm, non_blocking=True)


def down():
    with torch.cuda.stream(s_dn):
        h_out.copy_(d_out, non_blocking=True)


def both():
    up()
    down()


def up_chunks(k=7):
    step = (n * 6 + k - 1) // k
    with torch.cuda.stream(s_up):
        for i in range(k):
            d_pcm[i * step:(i + 1) * step].copy_(h_pcm[i * step:(i + 1) * step], non_blocking=True)


t_up, t_dn, t_both, t_chunks = timed(up), timed(down), timed(both), timed(up_chunks)
print("H2D 345.6 MB alone: %.3f ms (%.1f GB/s)" % (t_up, n * 6 / t_up / 1e6))
print("D2H 177 MB alone:   %.3f ms (%.1f GB/s)" % (t_dn, n_out / t_dn / 1e6))
print("both concurrently:  %.3f ms" % t_both)
print("H2D in 7 chunks:    %.3f ms" % t_chunks)
for per in (1024, 2048, 4096):
    enc = zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=per)
    cap = enc.max_batch_bytes((n + 4095) // 4096)
    ho = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    t = timed(lambda: enc.encode_pcm(h_pcm.numpy(), n, 0, out=ho.numpy()))
    print("e2e, %d frames/batch: %.3f ms  (%.2f G samples/s)" % (per, t, 2 * n / t / 1e6))
    enc.close()
    del ho

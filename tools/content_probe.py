"""Development probe: kernel time per 1000 frames for content classes that could hit a slow path (B200, 24/16-bit).
Every result is also checked against the oracle on its first 64 frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
os.environ["ZF_NO_TAPER"] = "1"  # one batch = one launch: the kernel time is that of a single kernel
import zigflac_b200 as zf
import oracle_lib as oracle

oracle.lib()
FR = 8880  # 20 frames per CTA
n = FR * 4096


def classes(bits, rng):
    F = 1 << (bits - 1)
    t = np.arange(n)
    music = (0.2 * F * np.sin(t * 0.03) + 0.1 * F * np.sin(t * 0.171) + rng.normal(0, F / 300, n)).astype(np.int64)
    other = (0.15 * F * np.sin(t * 0.021 + 1) + rng.normal(0, F / 300, n)).astype(np.int64)
    syn = np.frombuffer(zf.synth_pcm(n, 96000, bits), dtype=np.uint8)
    if bits == 24:
        b = syn.reshape(-1, 3).astype(np.int64)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
    else:
        v = syn.view("<i2").astype(np.int64)
    yield "bench synthetic stream", v[0::2].copy(), v[1::2].copy()
    yield "music-like (tones + noise)", music, other
    yield "dual mono (L = R)", music, music
    yield "silence", np.zeros(n, np.int64), np.zeros(n, np.int64)
    yield "full-scale noise (VERBATIM)", rng.integers(-F, F, n), rng.integers(-F, F, n)
    yield "8 wasted bits", (music >> 8) << 8, (other >> 8) << 8
    yield "quiet (|x| < 4, escapes)", rng.integers(-3, 4, n), rng.integers(-3, 4, n)
    sp = np.where(rng.random(n) < 0.002, rng.integers(-F, F, n), rng.integers(-2, 3, n))
    yield "sparse spikes (long unary runs)", sp, sp[::-1].copy()
    yield "left only (R = 0)", music, np.zeros(n, np.int64)


for bits in (24, 16):
    rng = np.random.default_rng(bits)
    enc = zf.Encoder(zf.Config.default(2, bits), 48000, max_frames_per_batch=FR)
    cfg = oracle.config(2, bits)
    F = 1 << (bits - 1)
    for name, L, R in classes(bits, rng):
        x = np.stack([np.clip(L, -F, F - 1), np.clip(R, -F, F - 1)], axis=1).reshape(-1)
        pcm = oracle.pcm_bytes_from_int(x, bits)
        for _ in range(3):
            got, sizes = enc.encode_pcm(pcm, n, 0)
        ms = enc.last_batch_stats()[0]
        k = 64
        ref, ref_sizes = oracle.encode_pcm(pcm[: k * 4096 * 2 * (bits // 8)], k * 4096, cfg, 48000, 0)
        ok = np.array_equal(ref_sizes, sizes[:k]) and ref.tobytes() == got[: ref.size].tobytes()
        print("%2d-bit %-34s kernel %.3f ms  %.1f us/1000 frames  ratio %.3f  parity %s" %
              (bits, name, ms, ms * 1e3 / (FR / 1000), got.size / pcm.size, "ok" if ok else "MISMATCH"), flush=True)
    enc.close()

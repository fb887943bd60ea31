"""Input classes that each pin a distinct reference branch.  Shared by the
CPU oracle tests, the kernel-logic emulation tests and the GPU parity tests."""
import numpy as np


def _clip(v, bits):
    f = 1 << (bits - 1)
    return np.clip(np.asarray(v, dtype=np.int64), -f, f - 1)


def stereo_classes(bits, n=2 * 4096 + 777, seed=7):
    """Yields (name, L, R) int64 arrays."""
    rng = np.random.default_rng(seed + bits)
    F = 1 << (bits - 1)
    t = np.arange(n)
    sine = (0.4 * F * np.sin(2 * np.pi * 440 * t / 44100)).astype(np.int64)
    noise = rng.integers(-F, F, n)
    small = rng.integers(-3, 4, n)
    yield "silence", np.zeros(n, np.int64), np.zeros(n, np.int64)
    yield "dc", np.full(n, 1234), np.full(n, -77)
    yield "dc_trailing_zero_bits", np.full(n, 1232), np.full(n, -80)
    yield "wasted_bits", sine & ~0xF, (sine // 3) & ~0x3
    yield "white_noise_full_scale", noise, rng.integers(-F, F, n)
    yield "dither", small, rng.integers(-1, 2, n)
    z = sine.copy()
    z[1000:3000] = 0
    z[5000:5020] = 0
    yield "zero_runs", z, z // 2
    yield "left_equals_right", sine, sine.copy()
    yield "right_is_minus_left", sine, -sine
    yield "full_scale_dc", np.full(n, F - 1), np.full(n, -F)
    sq = np.where((t // 50) % 2 == 0, F - 1, -F)
    yield "full_scale_square", sq, -sq - 1
    yield "tone_plus_noise", sine + small, _clip(sine // 2 + noise // 1024, bits)
    for k in (1, 5, 8, bits):
        base = rng.integers(-F // 4, F // 4, n)
        d = ((rng.integers(-F // 8, F // 8, n) >> k) << k) if k < bits else np.zeros(n, np.int64)
        yield f"side_waste_{k}", base + d, base
    walk = _clip(np.cumsum(rng.integers(-F // 64, F // 64 + 1, n)), bits)
    yield "random_walk", walk, _clip(walk + rng.integers(-F // 256, F // 256 + 1, n), bits)
    sparse = np.where(rng.random(n) < 0.01, rng.integers(-F, F, n), 0)
    yield "sparse_spikes", sparse, np.where(rng.random(n) < 0.02, rng.integers(-8, 8, n), 0)


SHORT_LENGTHS = (1, 2, 3, 4, 5, 6, 7, 15, 16, 17, 31, 33, 100, 255, 256, 257, 576, 1000, 1024, 2048, 4080, 4095)
FRAME_NUMBERS = (126, 127, 2046, 2047, 65534, 65535, (1 << 21) - 2, (1 << 26) - 2, (1 << 26) + 5)
SAMPLE_RATES = (8000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 176400, 192000, 11025, 100000, 200, 12345)


def interleave(channels):
    return np.stack([np.asarray(c, dtype=np.int64) for c in channels], axis=1).reshape(-1)

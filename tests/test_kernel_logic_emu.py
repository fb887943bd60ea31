"""CPU test of the encode KERNELS' integer logic: zig-flac_b200/csrc/zf_kernel*.cuh compiled with
-DZF_HOST_EMU (tests/kernel_emu: CUDA threads as fibers) and compared byte for byte with the oracle.

This is a development/test harness for a box without a GPU.  It is not a product path and not a
fallback: the test library is separate from libzigflac_b200.so, which fails without a CUDA device.
The parity tests proper are tests/test_gpu_parity.py (-m gpu), through the C ABI on a B200.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import signals

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "kernel_emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libzf_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "zig-flac_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_main.cpp", "cuda_emu.h")] + [
        os.path.join(CSRC, f) for f in ("zf_kernel.cuh", "zf_kernel_indep.cuh", "zf_kernel_v3.cuh", "zf_kernel_lpc.cuh", "zf_dev.h")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", EMU_DIR, "-o", EMU_SO,
                        os.path.join(EMU_DIR, "emu_main.cpp")], check=True)
    lib = C.CDLL(EMU_SO)
    lib.emu_encode.restype = C.c_longlong
    lib.emu_encode.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_ulonglong,
                               C.c_uint, C.c_uint, C.c_void_p, C.c_ulonglong, C.c_void_p, C.POINTER(C.c_uint32)]
    lib.emu_best_param.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_ulonglong)]
    return lib


def _emu_encode(emu, pcm, n, bits, channels=2, dec=1, block=4096, rate=44100, first=0, mro=8, mrp=30):
    pcm = np.ascontiguousarray(pcm)
    frames = (n + block - 1) // block
    cap = frames * (channels + 1) * block * 5 + 4096
    out = np.zeros(cap, dtype=np.uint8)
    sizes = np.zeros(frames + 1, dtype=np.uint32)
    nf = C.c_uint32()
    tot = emu.emu_encode(pcm.ctypes.data, n, bits // 8, channels, dec, block, rate, first, mro, mrp, out.ctypes.data, cap,
                         sizes.ctypes.data, C.byref(nf))
    assert tot >= 0, tot
    return out[:tot], sizes[:nf.value]


def _check(emu, oracle, pcm, n, bits, **kw):
    cfg = oracle.config(kw.get("channels", 2), bits, block_size=kw.get("block", 4096),
                        stereo_decorrelation=kw.get("dec", 1), max_rice_order=kw.get("mro", 8),
                        max_rice_param=kw.get("mrp", 30))
    ref, rs = oracle.encode_pcm(pcm, n, cfg, kw.get("rate", 44100), kw.get("first", 0))
    got, gs = _emu_encode(emu, pcm, n, bits, **kw)
    assert np.array_equal(rs, gs)
    assert ref.tobytes() == got.tobytes()


def test_closed_form_parameter_search_equals_brute_force(emu, oracle):
    """best_param() picks the same (choice, cost) as the reference's loop (rice.zig:359-379)."""
    rng = np.random.default_rng(1)
    f = oracle.lib().zo_flac_calc_part_size
    for _ in range(20000):
        n = int(rng.choice([0, 1, 2, 3, 12, 13, 15, 16, 32, 255, 256, 1000, 4092, 4096]))
        S = int(rng.integers(0, 1 << int(rng.integers(1, 50))))
        if rng.random() < 0.25:
            S = int(rng.integers(0, 4 * n + 3))
        B, P = int(rng.integers(0, 33)), int(rng.integers(1, 31))
        best, ch = ((5 + B * n) if B <= 31 else (1 << 64) - 1), 0x80 | B
        for p in range(P):
            c = f(n, p, S)
            if c < best:
                best, ch = c, p
        c_out, k_out = C.c_uint(), C.c_ulonglong()
        emu.emu_best_param(S, B, n, P, C.byref(c_out), C.byref(k_out))
        assert (c_out.value, k_out.value) == (ch, best), (S, B, n, P)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_stereo_kernel_logic_on_input_classes(emu, oracle, bits):
    for name, L, R in signals.stereo_classes(bits, n=4096 + 333):
        pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
        _check(emu, oracle, pcm, L.size, bits)


def test_kernel_logic_short_frames_and_variants(emu, oracle):
    rng = np.random.default_rng(9)
    for bits in (16, 32):
        F = 1 << (bits - 1)
        for m in (1, 4, 5, 16, 17, 255, 256, 1000, 2048, 4080):
            t = np.arange(m)
            L = (0.3 * F * np.sin(t * 0.05)).astype(np.int64) + rng.integers(-3, 4, m)
            pcm = oracle.pcm_bytes_from_int(signals.interleave([L, L // 2]), bits)
            _check(emu, oracle, pcm, m, bits)
        n = 4096 + 100
        L = rng.integers(-F // 8, F // 8, n)
        pcm = oracle.pcm_bytes_from_int(signals.interleave([L, L + rng.integers(-50, 50, n)]), bits)
        for fn in (127, 2047, 65535, (1 << 26) + 5):
            _check(emu, oracle, pcm, n, bits, first=fn)
        for kw in ({"mro": 0}, {"mro": 5}, {"mrp": 1}, {"mrp": 15}, {"block": 576}, {"block": 1000}, {"rate": 11025},
                   {"dec": 0}):
            _check(emu, oracle, pcm, n, bits, **kw)


@pytest.mark.parametrize("channels", [1, 3, 8])
def test_independent_channel_kernel_logic(emu, oracle, channels):
    rng = np.random.default_rng(channels)
    for bits in (16, 24, 32):
        F = 1 << (bits - 1)
        n = 4096 + 77
        planes = [rng.integers(-F // 4, F // 4, n) if c % 2 else np.cumsum(rng.integers(-9, 10, n)) for c in range(channels)]
        pcm = oracle.pcm_bytes_from_int(signals.interleave(planes), bits)
        _check(emu, oracle, pcm, n, bits, channels=channels, rate=48000)


def test_v3_parameter_search_equals_brute_force(emu, oracle):
    """v3::best_param_nw (32-bit costs, widths <= 31, P in {14, 30}) against the reference's loop (rice.zig:359-379)."""
    emu.emu_best_param_nw.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_ulonglong)]
    rng = np.random.default_rng(3)
    f = oracle.lib().zo_flac_calc_part_size
    for _ in range(30000):
        n = int(rng.choice([12, 13, 14, 15, 16, 32, 64, 128, 256, 512, 1024, 2048, 4092, 4096]))
        S = int(rng.integers(0, 1 << int(rng.integers(1, 41))))
        if rng.random() < 0.3:
            S = int(rng.integers(0, 6 * n + 3))
        B, P = int(rng.integers(0, 27)), int(rng.choice([14, 30]))
        best, ch = 5 + B * n, 0x80 | B
        for p in range(P):
            c = f(n, p, S)
            if c < best:
                best, ch = c, p
        c_out, k_out = C.c_uint(), C.c_ulonglong()
        emu.emu_best_param_nw(S, B, n, P, C.byref(c_out), C.byref(k_out))
        assert (c_out.value, k_out.value) == (ch, best), (S, B, n, P)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_v3_kernel_multi_frame_streams(emu, oracle, bits):
    """The lean 256-thread kernel (zf_kernel_v3.cuh) on multi-frame streams: persistent loop, look-back offsets of
    odd-sized frames, every stereo mode, escape partitions, wasted bits, frame numbers with long UTF-8 codes."""
    import zigflac_b200 as zf
    n = 5 * 4096
    pcm = zf.synth_pcm(n, 44100 if bits == 16 else 96000, bits)
    emu.emu_v3_frames.restype = C.c_ulonglong
    before = emu.emu_v3_frames()
    _check(emu, oracle, pcm, n, bits, rate=44100 if bits == 16 else 96000)
    _check(emu, oracle, pcm, n, bits, rate=96000, first=(1 << 21) - 2)
    rng = np.random.default_rng(bits)
    F = 1 << (bits - 1)
    t = np.arange(n)
    env = (0.02 + 0.98 * np.sin(2 * np.pi * t / 3000.0) ** 8)
    L = (0.45 * F * env * np.sin(2 * np.pi * 997 * t / 48000)).astype(np.int64) + rng.integers(-2, 3, n)
    R = (L * 0.9).astype(np.int64) + np.where(rng.random(n) < 0.003, rng.integers(-F // 2, F // 2, n), 0)
    R = np.clip(R, -F, F - 1)
    R[4096:4096 + 2048] &= ~0xff
    L[4096:4096 + 2048] &= ~0xf
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
    _check(emu, oracle, pcm, n, bits, rate=48000)
    for name, a, b in signals.stereo_classes(bits, n=2 * 4096):
        _check(emu, oracle, oracle.pcm_bytes_from_int(signals.interleave([a, b]), bits), a.size, bits)
    assert emu.emu_v3_frames() - before >= 15 + 2 * 18  # the streams above really went through the v3 kernel


def test_v3_32bit_narrow_and_wide_paths(emu, oracle):
    """32-bit PCM runs in 32-bit registers while every difference (and the side channel) fits, with exact overflow
    detection; a frame that overflows is redone by the 64-bit chains.  Both paths, and the boundary between them,
    against the oracle (fixed.zig:85-167 is i64 throughout; encoder.zig:330-339,398-438 for the 33-bit side)."""
    emu.emu_v3_wide_frames.restype = C.c_ulonglong
    emu.emu_v3_narrow_frames.restype = C.c_ulonglong
    bits, F, n = 32, 1 << 31, 4096
    rng = np.random.default_rng(32)
    t = np.arange(n)

    def run(L, R, expect):
        L = np.clip(np.asarray(L, dtype=np.int64), -F, F - 1)
        R = np.clip(np.asarray(R, dtype=np.int64), -F, F - 1)
        w0, n0 = emu.emu_v3_wide_frames(), emu.emu_v3_narrow_frames()
        _check(emu, oracle, oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits), n, bits, rate=96000)
        got = "wide" if emu.emu_v3_wide_frames() > w0 else "narrow"
        assert emu.emu_v3_wide_frames() - w0 + emu.emu_v3_narrow_frames() - n0 == 1
        assert got == expect, (got, expect)

    loud = (0.98 * F * np.sin(2 * np.pi * 60 * t / 96000)).astype(np.int64)
    run(loud + rng.integers(-1000, 1000, n), (0.7 * loud).astype(np.int64), "narrow")      # near full scale, smooth
    run(loud, -loud, "wide")                                                                # |L - R| needs 33 bits
    run(np.full(n, F - 1), np.full(n, -F), "wide")                                          # constant side of 2^32 - 1
    run(rng.integers(-F, F, n), rng.integers(-F, F, n), "wide")                             # full-scale noise
    run(rng.integers(-F // 64, F // 64, n), rng.integers(-F // 64, F // 64, n), "narrow")   # side < 2^26: |delta^4| < 2^30
    step = np.where(t < 2000, -F + 5, F - 9)                                                # one jump of almost 2^32
    run(step, step // 3, "wide")
    jump = np.where(t < 2000, -(F // 2) - 3, F // 2 - 1)                                    # delta^1 = 2^31 - 2 fits, delta^2 does not
    run(jump, np.zeros(n, np.int64), "wide")
    small_jump = np.where(t < 2000, -(F // 4), F // 4)                                      # 2^30: only delta^4 = 3 * 2^30 overflows
    run(small_jump, np.zeros(n, np.int64), "wide")
    tiny_jump = np.where(t < 2000, -(F // 8), F // 8)                                       # delta^4 = 3 * 2^29 still fits
    run(tiny_jump, tiny_jump // 2, "narrow")
    j24 = (rng.integers(-(1 << 23), 1 << 23, n) << 8)                                       # loud 24-bit data in 32 bits
    run(j24, j24 // 2 // 256 * 256, "wide")
    q24 = ((0.4 * (1 << 23) * np.sin(t * 0.01)).astype(np.int64) + rng.integers(-50, 50, n)) << 8   # quiet: 8 wasted bits
    run(q24, q24 // 2 // 256 * 256, "narrow")
    run(np.full(n, -F), np.full(n, -F), "narrow")                                           # CONSTANT at INT32_MIN, side 0
    alt = np.where(t % 2 == 0, F // 2 - 1, -(F // 2))                                       # delta^1 = +-(2^31 - 1): fits; delta^2 not
    run(alt, alt, "wide")
    # thread 0's first samples: a full-scale first sample is not a difference and must not force the slow path
    first = np.zeros(n, np.int64)
    first[0] = F - 1
    first[1:] = (F - 1) - np.minimum(np.arange(1, n) * 1000, F // 3)
    run(first, first // 2, "narrow")


def test_8bit_samples_through_the_16bit_kernels(emu, oracle):
    """bit_depth 8: signed one-byte samples travel in 16-bit containers through the general kernels with depth 8
    (frame header code 2, frame_writer.zig:221-233; Rice limits from bps <= 9).  Stereo and mono against the oracle,
    which reads the raw unsigned bytes the way the reference's reader does."""
    import zigflac_b200 as zf
    rng = np.random.default_rng(88)
    emu.emu_set_bit_depth.argtypes = [C.c_uint]
    try:
        for channels in (2, 1):
            n = 2 * 4096 + 777
            t = np.arange(n)
            cols = [np.clip(128 + 100 * np.sin(t * 0.01 * (c + 1)) + rng.integers(-2, 3, n), 0, 255).astype(np.uint8)
                    for c in range(channels)]
            raw = np.stack(cols, axis=1).reshape(-1).copy()
            ref, rs = oracle.encode_pcm(raw, n, oracle.config(channels, 8), 22050, 0)
            signed = zf.Wav8Reader(channels).convert(raw)
            wide = signed.astype("<i2").view(np.uint8)
            emu.emu_set_bit_depth(8)
            got, gs = _emu_encode(emu, wide, n, 16, channels=channels, rate=22050)
            assert np.array_equal(rs, gs)
            assert ref.tobytes() == got.tobytes()
    finally:
        emu.emu_set_bit_depth(0)


@pytest.mark.parametrize("bits", [16, 24])
def test_lpc_kernel_logic(emu, oracle, bits):
    """LPC extension (zf_kernel_lpc.cuh; the reference has none, encoder.zig:626-640): the kernel's integer window /
    autocorrelation, double-precision Levinson-Durbin, order choice, quantisation, residual and the choice against the
    FIXED candidate, bit for bit against the CPU statement of the same specification (oracle/zigflac_lpc.h), and the
    stream through the independent decoder."""
    import zigflac_b200 as zf
    emu.emu_set_lpc_order.argtypes = [C.c_uint]
    rate = 44100 if bits == 16 else 96000
    rng = np.random.default_rng(bits)
    F = 1 << (bits - 1)
    streams = []
    n = 2 * 4096 + 1500
    streams.append(("synthetic", zf.synth_pcm(n, rate, bits), n))
    t = np.arange(n)
    L = (0.3 * F * np.sin(t * 0.3) + 0.2 * F * np.sin(t * 0.71 + 1) + rng.normal(0, F / 500, n)).astype(np.int64)
    R = (0.25 * F * np.sin(t * 0.3 + 0.4) + rng.normal(0, F / 300, n)).astype(np.int64)
    streams.append(("tones", oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits), n))
    for name, a, b in signals.stereo_classes(bits, n=4096 + 900):
        streams.append((name, oracle.pcm_bytes_from_int(signals.interleave([a, b]), bits), a.size))
    m = 100  # short frames: below 64 samples LPC is off, above it runs on the short geometry
    for k in (5, 63, 64, 100, 1000):
        x = (0.2 * F * np.sin(np.arange(k) * 0.2)).astype(np.int64) + rng.integers(-9, 10, k)
        streams.append((f"short{k}", oracle.pcm_bytes_from_int(signals.interleave([x, x // 3]), bits), k))
    n_lpc = 0
    try:
        emu.emu_set_lpc_order(12)
        for name, pcm, ns in streams:
            ref, rs = oracle.encode_pcm(pcm, ns, oracle.config(2, bits, lpc_order=12), rate, 0)
            got, gs = _emu_encode(emu, pcm, ns, bits, rate=rate)
            assert np.array_equal(rs, gs), name
            assert ref.tobytes() == got.tobytes(), name
            d = oracle.decode(oracle.wrap_frames(got, 2, bits, rate))
            assert d["rc"] == 0, name
            n_lpc += sum(1 for f in d["frames"] for s in f.sub[:2] if s.type == 3)
            fixed, _ = oracle.encode_pcm(pcm, ns, oracle.config(2, bits), rate, 0)
            assert got.size <= fixed.size + 8 * len(d["frames"]), name  # never worse than FIXED beyond rounding per frame
    finally:
        emu.emu_set_lpc_order(0)
    assert n_lpc > 10  # LPC subframes really were chosen and written


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_v3_kernel_any_rice_limit_and_sample_rate(emu, oracle, bits):
    """The lean kernel covers every max_rice_param (1 = parameter 0 only, rice.zig:369) and sample rates without a header
    code of their own, whose trailer carries the block size upstream (frame_writer.zig:258-262)."""
    import zigflac_b200 as zf
    emu.emu_v3_frames.restype = C.c_ulonglong
    n = 2 * 4096
    pcm = zf.synth_pcm(n, 96000, bits, first_sample=300000)
    before = emu.emu_v3_frames()
    for mrp in (1, 2, 5, 14, 15, 29, 30):
        _check(emu, oracle, pcm, n, bits, rate=96000, mrp=mrp)
    for mro in (0, 1, 3, 5, 7):
        _check(emu, oracle, pcm, n, bits, rate=96000, mro=mro)
    for rate in (200, 255, 256, 11025, 12345, 65535, 65536, 100000, 1048575):
        _check(emu, oracle, pcm, n, bits, rate=rate, first=(1 << 21) + 3)
        _check(emu, oracle, pcm, n, bits, rate=rate, first=5)
    assert emu.emu_v3_frames() - before == 2 * ((7 if bits < 32 else 2) + 5 + 18)  # 32-bit with a low parameter limit: general kernel


def test_random_configurations_emulated(emu, oracle):
    """A short randomised sweep over block sizes / tails / Rice limits through the emulated kernels (the long one runs on
    the GPU: tests/test_gpu_parity.py::test_random_configurations)."""
    rng = np.random.default_rng(77)
    for case in range(40):
        bits = int(rng.choice([16, 24, 32]))
        channels = int(rng.choice([1, 2, 2, 3]))
        block = int(rng.choice([4096, 2048, 1024, 576, 192, 1000, 4000, 16, 64, 4080, 255 * 8]))
        mro = int(rng.choice([8, 8, 0, 3, 6]))
        mrp = int(rng.choice([30, 14, 1, 7]))
        n = block + int(rng.integers(1, block + 1))
        F = 1 << (bits - 1)
        planes = [np.clip((rng.normal(0, F * 10.0 ** rng.uniform(-4, -0.5), n)).astype(np.int64), -F, F - 1) for _ in range(channels)]
        pcm = oracle.pcm_bytes_from_int(signals.interleave(planes), bits)
        _check(emu, oracle, pcm, n, bits, channels=channels, block=block, mro=mro, mrp=mrp, rate=48000)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_exact_rice_search_kernel_logic(emu, oracle, bits):
    """Extension (zf_config.exact_rice): Rice parameters by exact code length -- per-partition bit counters, additive over
    the partition tree, E_p = cnt_p + 2 E_{p+1} -- against the oracle's brute-force statement of the same rule; streams
    decode losslessly and are never larger than the estimate's."""
    import zigflac_b200 as zf
    emu.emu_set_exact_rice.argtypes = [C.c_uint]
    rng = np.random.default_rng(100 + bits)
    F = 1 << (bits - 1)
    cases = [("synthetic", zf.synth_pcm(4096 + 1500, 96000, bits), 4096 + 1500, {})]
    for name, a, b in signals.stereo_classes(bits, n=4096 + 700):
        cases.append((name, oracle.pcm_bytes_from_int(signals.interleave([a, b]), bits), a.size, {}))
    for block, mro, mrp in ((576, 8, 30), (1000, 3, 14), (4080, 8, 30), (2048, 8, 5), (16, 8, 30), (4096, 6, 30)):
        n = block + block // 3 + 1
        x = (0.2 * F * np.sin(np.arange(n) * 0.05)).astype(np.int64) + rng.integers(-F // 1000 - 2, F // 1000 + 3, n)
        cases.append((f"block{block}", oracle.pcm_bytes_from_int(signals.interleave([x, x // 2 + rng.integers(-3, 4, n)]), bits), n,
                      dict(block=block, mro=mro, mrp=mrp)))
    try:
        emu.emu_set_exact_rice(1)
        for name, pcm, n, kw in cases:
            cfg = oracle.config(2, bits, block_size=kw.get("block", 4096), max_rice_order=kw.get("mro", 8),
                                max_rice_param=kw.get("mrp", 30), exact_rice=1)
            ref, rs = oracle.encode_pcm(pcm, n, cfg, 48000, 0)
            got, gs = _emu_encode(emu, pcm, n, bits, rate=48000, **kw)
            assert np.array_equal(rs, gs), name
            assert ref.tobytes() == got.tobytes(), name
            assert oracle.decode(oracle.wrap_frames(got, 2, bits, 48000, block_size=kw.get("block", 4096)))["rc"] == 0, name
            cfg.exact_rice = 0
            est, _ = oracle.encode_pcm(pcm, n, cfg, 48000, 0)
            assert got.size <= est.size, name
    finally:
        emu.emu_set_exact_rice(0)

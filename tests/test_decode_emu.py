"""CPU test of the DECODER's device functions: zig-flac_b200/csrc/zf_kernel_decode.cuh compiled for the host
(tests/kernel_emu/emu_decode.cpp runs them thread by thread in the order zf_decode.cu launches the kernels) against
streams of the oracle encoder, with the independent decoder oracle/flac_decode.c as the checker.

A development/test harness for a box without a GPU -- not a product path and not a fallback: libzigflac_b200.so's
decode entries fail without a CUDA device.  The parity tests proper are tests/test_gpu_decode.py (-m gpu).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import signals

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "kernel_emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libzf_emu_decode.so")
CSRC = os.path.join(os.path.dirname(HERE), "zig-flac_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_decode.cpp", "cuda_emu.h")] + [
        os.path.join(CSRC, f) for f in ("zf_kernel_decode.cuh", "zf_decode_host.h", "zf_dev.h")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fwrapv", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", EMU_DIR, "-o", EMU_SO,
                        os.path.join(EMU_DIR, "emu_decode.cpp")], check=True)
    lib = C.CDLL(EMU_SO)
    lib.emu_decode_flac.restype = C.c_longlong
    lib.emu_decode_flac.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    return lib


def emu_decode(emu, flac, cap=None):
    a = np.frombuffer(bytes(flac), dtype=np.uint8)
    cap = cap if cap is not None else max(1 << 20, a.size * 40)
    out = np.zeros(cap, dtype=np.uint8)
    info = (C.c_uint32 * 4)()
    bad = (C.c_uint32 * 2)()
    n = emu.emu_decode_flac(a.ctypes.data, a.size, out.ctypes.data, cap, info, bad)
    return n, out[:max(n, 0)], tuple(info), tuple(bad)


def _roundtrip(emu, oracle, pcm_int, channels, bits, rate=44100, **cfgkw):
    pcm = oracle.pcm_bytes_from_int(pcm_int, bits)
    n = len(pcm_int) // channels
    cfg = oracle.config(channels, bits, **cfgkw)
    frames, sizes = oracle.encode_pcm(pcm, n, cfg, rate)
    flac = oracle.wrap_frames(frames, channels, bits, rate, cfg.block_size, n)
    got_n, got, info, bad = emu_decode(emu, flac)
    assert got_n == pcm.size, (got_n, bad)
    assert got.tobytes() == pcm.tobytes()
    assert info == (channels, bits, rate, len(sizes))
    ref = oracle.decode(flac)
    assert ref["rc"] == 0
    assert oracle.pcm_bytes_from_int(ref["pcm"], bits).tobytes() == got.tobytes()
    return flac


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_decoder_logic_stereo_classes(emu, oracle, bits):
    """CONSTANT / VERBATIM / FIXED subframes, wasted bits, escapes, every stereo assignment, 33-bit side channel."""
    for name, L, R in signals.stereo_classes(bits, n=4096 + 777):
        _roundtrip(emu, oracle, signals.interleave([L, R]), 2, bits)


def test_decoder_logic_block_sizes_channels_and_lpc(emu, oracle):
    rng = np.random.default_rng(5)
    t = np.arange(3000)
    for channels in (1, 2, 3, 8):
        chans = [(3000 * np.sin(2 * np.pi * (200 + 90 * c) * t / 44100)).astype(np.int64) + rng.integers(-20, 21, t.size) for c in range(channels)]
        for block in (4096, 1024, 576, 192, 17):
            _roundtrip(emu, oracle, signals.interleave(chans), channels, 16, block_size=block)
    F = 1 << 23
    t = np.arange(4096 + 1500)
    L = (0.3 * F * np.sin(t * 0.3) + 0.2 * F * np.sin(t * 0.71 + 1) + rng.normal(0, F / 500, t.size)).astype(np.int64)
    R = (0.25 * F * np.sin(t * 0.3 + 0.4) + rng.normal(0, F / 300, t.size)).astype(np.int64)
    for order in (1, 4, 8, 12):
        flac = _roundtrip(emu, oracle, signals.interleave([L, R]), 2, 24, lpc_order=order)
        types = {s.type for f in oracle.decode(flac)["frames"] for s in f.sub[:2]}
        assert order < 4 or 3 in types  # LPC subframes were in the stream
    _roundtrip(emu, oracle, signals.interleave([L >> 8, R >> 8]), 2, 16, exact_rice=1)
    _roundtrip(emu, oracle, signals.interleave([L, R]), 2, 24, stereo_decorrelation=0)
    _roundtrip(emu, oracle, signals.interleave([L, R]), 2, 24, max_rice_param=3, max_rice_order=2)  # long unary runs
    for n in signals.SHORT_LENGTHS:
        _roundtrip(emu, oracle, signals.interleave([L[:n], R[:n]]), 2, 24)


def header_image_stream(oracle, with_expected_number):
    """16-bit stereo noise (VERBATIM subframes: the left channel's samples are byte-aligned big-endian words) with the
    bytes of valid frame headers planted in the samples of frame 0: one carrying a frame number that is not the next
    one, and optionally one carrying exactly the next one (the hardest case for a decoder that finds frames by scanning)."""
    rng = np.random.default_rng(11)
    n = 3 * 4096 + 100
    L = rng.integers(-32768, 32768, n)
    R = rng.integers(-32768, 32768, n)
    L[::5] = -8  # the bytes ff f8

    def plant(at, number):
        h = oracle.frame_header(number, 16, 1, 4096, 48000)
        assert len(h) == 6
        for k in range(3):
            L[at + k] = int.from_bytes(h[2 * k:2 * k + 2], "big", signed=True)

    plant(1001, 7)
    if with_expected_number:
        plant(2001, 1)
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), 16)
    rc, flac = oracle.wav_to_flac(oracle.make_wav(pcm, 2, 16, 48000))
    assert rc == 0
    assert flac.count(b"\xff\xf8") > 500
    assert flac.count(oracle.frame_header(7, 16, 1, 4096, 48000)) == 1
    assert flac.count(oracle.frame_header(1, 16, 1, 4096, 48000)) == (2 if with_expected_number else 1)
    assert oracle.decode(flac)["rc"] == 0
    return pcm, flac


def test_decoder_logic_whole_file_and_sync_patterns_in_data(emu, oracle):
    """A whole file with metadata blocks in front and frame data full of sync patterns and of complete, valid frame
    header images: the chain by frame number discards them, and the one image that carries the expected number is
    dropped after the frame in front of it fails its length / CRC check."""
    for hard in (False, True):
        pcm, flac = header_image_stream(oracle, hard)
        got_n, got, info, bad = emu_decode(emu, flac)
        assert got_n == pcm.size and got.tobytes() == pcm.tobytes(), (hard, got_n, bad)
        assert info[3] == 4
    # 8-bit: signed samples out (what went in passed through the reference reader's quirk, so the independent decoder
    # says what the stream holds)
    x = (np.arange(5000) % 200 - 100).astype(np.int64)
    p8 = oracle.pcm_bytes_from_int(signals.interleave([x, -x]), 8)
    frames, _ = oracle.encode_pcm(p8, 5000, oracle.config(2, 8), 8000)
    flac = oracle.wrap_frames(frames, 2, 8, 8000, 4096, 5000)
    got_n, got, info, bad = emu_decode(emu, flac)
    ref = oracle.decode(flac)
    assert ref["rc"] == 0 and got_n == p8.size
    assert got.tobytes() == oracle.pcm_bytes_from_int(ref["pcm"], 8).tobytes()


def test_decoder_logic_rejects_damage(emu, oracle):
    t = np.arange(2 * 4096)
    L = (9000 * np.sin(2 * np.pi * 300 * t / 44100)).astype(np.int64)
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, L // 3]), 16)
    rc, flac = oracle.wav_to_flac(oracle.make_wav(pcm, 2, 16, 44100))
    assert rc == 0
    assert emu_decode(emu, flac)[0] == pcm.size
    assert emu_decode(emu, b"RIFF" + flac[4:])[0] == -32
    assert emu_decode(emu, flac[:30])[0] in (-32, -33)
    assert emu_decode(emu, flac, cap=100)[0] == -6
    rng = np.random.default_rng(3)
    body = len(flac) - 73
    for _ in range(40):  # a flipped bit anywhere in the frames: CRC-16 / header / length error or a sample-count mismatch, never success
        pos = 73 + int(rng.integers(0, body))
        bad = bytearray(flac)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        n, _, _, why = emu_decode(emu, bytes(bad))
        assert n in (-34, -35, -33), (pos, n, why)
        ref = oracle.decode(bytes(bad))
        assert ref["rc"] != 0
    n, _, _, why = emu_decode(emu, flac[:-5])  # truncated last frame
    assert n == -34


def test_decoder_logic_fuzz(emu, oracle):
    """Mutated and truncated streams: the decoder must come back (no endless parse), must never report success with
    samples other than the independent decoder's, and mostly must report the damage."""
    rng = np.random.default_rng(99)
    t = np.arange(3 * 1024 + 200)
    L = (5000 * np.sin(t * 0.05)).astype(np.int64) + rng.integers(-30, 31, t.size)
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, L // 2 + 7]), 16)
    frames, _ = oracle.encode_pcm(pcm, t.size, oracle.config(2, 16, block_size=1024), 44100)
    flac = oracle.wrap_frames(frames, 2, 16, 44100, 1024, t.size)
    assert emu_decode(emu, flac)[0] == pcm.size
    reported = 0
    for trial in range(300):
        bad = bytearray(flac)
        kind = trial % 3
        if kind == 0:  # a burst of random bytes
            at = int(rng.integers(42, len(bad) - 8))
            for k in range(int(rng.integers(1, 9))):
                bad[at + k] = int(rng.integers(0, 256))
        elif kind == 1:  # zeros (endless unary runs) or ones
            at = int(rng.integers(42, len(bad) - 64))
            fill = 0 if trial % 2 else 0xFF
            for k in range(int(rng.integers(8, 64))):
                bad[at + k] = fill
        else:  # truncation
            bad = bad[:int(rng.integers(43, len(bad)))]
        n, got, info, why = emu_decode(emu, bytes(bad))
        ref = oracle.decode(bytes(bad))
        if n >= 0 and bytes(bad) != flac:
            assert ref["rc"] == 0 and oracle.pcm_bytes_from_int(ref["pcm"], 16).tobytes() == got.tobytes(), (trial, n)
        else:
            reported += 1
    assert reported > 250


def crafted_streams():
    """(name, stream, expected interleaved samples, channels, bits): streams built by tests/flac_craft.py from the format
    itself, with features no encoder of this repository emits."""
    import flac_craft as fc
    rng = np.random.default_rng(2024)
    out = []

    def integrate(res, warm, order):  # samples of a FIXED subframe from its residual
        x = [int(v) for v in warm[:order]]
        coef = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}[order]
        for r in res:
            x.append(int(r) + sum(c * x[-1 - j] for j, c in enumerate(coef)))
        return x

    # 1. 16-bit stereo, 4096: LPC order 20 (precision 12, shift 9) | FIXED order 3 with partition order 10, 5-bit
    #    parameters, escaped partitions (one of width 0) in between
    n = 4096
    t = np.arange(n)
    a = (6000 * np.sin(t * 0.03) + 2000 * np.sin(t * 0.41)).astype(np.int64) + rng.integers(-40, 41, n)
    coefs = [int(v) for v in rng.integers(-300, 301, 20)]
    params_a = [13] * 4
    b = (3000 * np.sin(t * 0.011)).astype(np.int64) + rng.integers(-3, 4, n)
    b[1000:1200] = 77  # a stretch of zero residuals
    d = b.copy()
    for _ in range(3):
        d = np.concatenate([d[:1], d[1:] - d[:-1]])
    res_b = d[3:]
    psize = n >> 10
    params_b = []
    for p in range(1 << 10):
        lo, hi = max(p * psize - 3, 0), (p + 1) * psize - 3
        part = res_b[lo:hi]
        width = max([int(v).bit_length() + 1 for v in part if v != 0] + [0])
        if width == 0 and p % 2:
            params_b.append(("esc", 0))
        elif p % 50 == 7:
            params_b.append(("esc", width))
        else:
            params_b.append(int(rng.integers(0, 4)) if p % 3 else 17)
    f = fc.frame(0, n, 9, 1, 4, [lambda w: fc.subframe_lpc(w, a, 16, 20, coefs, 12, 9, 2, 0, params_a),
                                 lambda w: fc.subframe_fixed(w, b, 16, 3, 10, 1, params_b)])
    out.append(("lpc20_po10", fc.stream([f], 2, 16, 44100, n, n, n), signals.interleave([a, b]), 2, 16))

    # 2. explicit block-size fields: three frames of 1000 (16-bit field), a last one of 200 (8-bit field); mid/side with
    #    wasted bits on the LPC mid channel; metadata blocks full of sync patterns in front
    frames, exp_l, exp_r = [], [], []
    for k, bs in enumerate((1000, 1000, 1000, 200)):
        tt = np.arange(bs) + 1000 * k
        left = (9000 * np.sin(tt * 0.02)).astype(np.int64)
        right = left // 2 + rng.integers(-9, 10, bs)
        left = (left >> 3) << 3
        right = (right >> 3) << 3  # mid and side are multiples of 4 / 8
        mid, side = (left + right) >> 1, left - right
        c5 = [int(v) for v in rng.integers(-200, 201, 5)]
        frames.append(fc.frame(7 + k, bs, 10, 10, 4,
                               [lambda w, mid=mid, c5=c5: fc.subframe_lpc(w, mid, 16, 5, c5, 10, 7, 0, 0, [12], wasted=2),
                                lambda w, side=side: fc.subframe_fixed(w, side, 17, 1, 0, 0, [11], wasted=3)],
                               explicit_block=16 if bs == 1000 else 8))
        exp_l.append(left)
        exp_r.append(right)
    meta = [(1, bytes([0xFF, 0xF8] * 40)), (2, b"test" + bytes([0xFF, 0xF8, 0xC9, 0x18, 0x00, 0xC2] * 9))]
    out.append(("explicit_blocks_ms", fc.stream(frames, 2, 16, 48000, 1000, 1000, 3200, meta),
                signals.interleave([np.concatenate(exp_l), np.concatenate(exp_r)]), 2, 16))

    # 3. 24-bit mono, 4096, FIXED order 1 with partition order 12: 4096 partitions of ONE sample, the first one empty
    res = rng.integers(-3000, 3001, n - 1)
    m = integrate(res, [123456], 1)
    params = [int(v) for v in rng.integers(8, 13, 1 << 12)]
    f = fc.frame(0, n, 11, 0, 6, [lambda w: fc.subframe_fixed(w, m, 24, 1, 12, 0, params)])
    out.append(("po12_single_sample_partitions", fc.stream([f], 1, 24, 96000, n, n, n), np.array(m, dtype=np.int64), 1, 24))

    # 4. LPC order 32, precision 15, shift 14, three channels (one CONSTANT, one VERBATIM)
    n4 = 576
    x = (20000 * np.sin(np.arange(n4) * 0.1)).astype(np.int64)
    c32 = [int(v) for v in rng.integers(-16000, 16001, 32)]
    v = rng.integers(-32768, 32768, n4)

    def const_sub(w):
        w.put(0, 8)
        w.put_signed(-777, 16)

    def verb_sub(w):
        w.put(1 << 1, 8)
        for s in v:
            w.put_signed(int(s), 16)

    f = fc.frame(3, n4, 9, 2, 4, [lambda w: fc.subframe_lpc(w, x, 16, 32, c32, 15, 14, 1, 1, [25, 24]), const_sub, verb_sub])
    out.append(("lpc32_three_channels", fc.stream([f], 3, 16, 44100, n4, n4, n4),
                signals.interleave([x, np.full(n4, -777), v]), 3, 16))

    # 5. 32-bit stereo left/side: a 33-bit side channel through LPC order 2 with parameters 28 / 29, long unary runs
    #    (parameter 0, quotients far beyond 64), an escape of width 31, frame numbers with long codes, a sample-rate trailer,
    #    the depth taken from STREAMINFO (code 0).  One stream per starting frame number.
    n5 = 256
    big = 1 << 31
    i5 = np.arange(n5)
    for k, number in enumerate((0x7F, 0x80, 0x7FF, 0x800)):
        nz = rng.integers(0, 50, n5)
        right = np.where(i5 % 2 == 0, -big + nz, big - 1 - nz).astype(object)          # within a hair of full scale
        if k % 2 == 0:
            left = np.where(i5 % 2 == 0, (1 << 30) - 1, -(1 << 30)).astype(object)     # escape of width 31, order 0
            lsub = lambda w, left=left: fc.subframe_fixed(w, left, 32, 0, 0, 1, [("esc", 31)])
        else:
            q = [int(v) for v in rng.integers(0, 3, n5)]
            q[5], q[100], q[101] = 40, 150, 33                                         # unary runs of 40, 150 and 33 zeros
            left = np.array(integrate([(v >> 1) ^ -(v & 1) for v in q[1:]], [5], 1), dtype=object)
            lsub = lambda w, left=left: fc.subframe_fixed(w, left, 32, 1, 0, 1, [0])   # parameter 0: the code is the run
        side = left - right                                                            # about +-1.5 * 2^31: 33 bits
        ssub = lambda w, side=side: fc.subframe_lpc(w, side, 33, 2, [-2, -1], 4, 0, 1, 1, [28, 29])
        f5 = fc.frame(number, n5, 13, 8, 0, [lsub, ssub], rate_trailer=bytes([0xAC, 0x44]))
        out.append((f"wide_side_{k}", fc.stream([f5], 2, 32, 44100, n5, n5, n5), signals.interleave([left, right]), 2, 32))

    # 6. 8-bit mono with the largest partition order a block of 192 allows (6: partitions of three samples), order 2
    n6 = 192
    x6 = (100 * np.sin(np.arange(n6) * 0.2)).astype(np.int64)
    f = fc.frame(2, n6, 4, 0, 1, [lambda w: fc.subframe_fixed(w, x6, 8, 2, 6, 0, [int(v) for v in rng.integers(0, 6, 64)])])
    out.append(("8bit_192", fc.stream([f], 1, 8, 8000, n6, n6, n6), x6, 1, 8))
    return out


def test_decoder_logic_crafted_streams(emu, oracle):
    """Features beyond what the repository's encoders emit, from a third statement of the format (tests/flac_craft.py):
    the device decoder's logic and the independent CPU decoder must both return the crafted samples."""
    for name, flac, expect, channels, bits in crafted_streams():
        pcm = oracle.pcm_bytes_from_int(expect, bits)
        ref = oracle.decode(flac)
        assert ref["rc"] == 0, (name, ref["rc"])
        assert oracle.pcm_bytes_from_int(ref["pcm"], bits).tobytes() == pcm.tobytes(), name
        n, got, info, bad = emu_decode(emu, flac)
        assert n == pcm.size and got.tobytes() == pcm.tobytes(), (name, n, bad)


def test_decoder_logic_memory_safety(oracle, tmp_path):
    """The decoder's device functions under AddressSanitizer (compute-sanitizer is not available on the GPU pool): the
    crafted streams, and some four hundred damaged copies of them, with every buffer of exactly the size the product
    allocates -- any access outside stops the program."""
    exe = os.path.join(EMU_DIR, "_build", "emu_decode_asan")
    srcs = [os.path.join(EMU_DIR, "emu_decode.cpp"), os.path.join(CSRC, "zf_kernel_decode.cuh"), os.path.join(CSRC, "zf_decode_host.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fwrapv", "-fsanitize=address,undefined", "-fno-sanitize=shift,signed-integer-overflow",
                            "-fno-omit-frame-pointer", "-DZF_EMU_DECODE_MAIN", "-Wno-unknown-pragmas", "-I", EMU_DIR, "-o", exe,
                            os.path.join(EMU_DIR, "emu_decode.cpp")], capture_output=True, text=True)
        if r.returncode != 0 and "asan" in (r.stderr or "").lower():
            pytest.skip("no AddressSanitizer runtime in this toolchain")
        assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(7)
    files = []
    for name, flac, expect, channels, bits in crafted_streams():
        p = tmp_path / f"{name}.flac"
        p.write_bytes(flac)
        files.append(str(p))
        for k in range(45):
            bad = bytearray(flac)
            kind = k % 4
            if kind == 0:
                at = int(rng.integers(4, len(bad)))
                bad[at] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                at = int(rng.integers(42, max(43, len(bad) - 40)))
                bad[at:at + 32] = bytes(32) if k % 8 < 4 else bytes([255]) * 32
            elif kind == 2:
                bad = bad[:int(rng.integers(8, len(bad)))]
            else:
                at = int(rng.integers(42, max(43, len(bad) - 8)))
                bad[at:at + 6] = bytes(int(v) for v in rng.integers(0, 256, 6))
            q = tmp_path / f"{name}_{k}.flac"
            q.write_bytes(bytes(bad))
            files.append(str(q))
    r = subprocess.run([exe] + files, capture_output=True, text=True, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    assert r.returncode in (0, 1), (r.returncode, r.stderr[-2000:])
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(files)
    good = {l.split()[0]: int(l.split()[1]) for l in lines}
    for name, flac, expect, channels, bits in crafted_streams():
        assert good[str(tmp_path / f"{name}.flac")] == len(expect) * (bits // 8), name

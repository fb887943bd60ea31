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

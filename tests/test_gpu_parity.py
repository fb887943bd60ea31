"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs -- byte equality of every frame and every frame size (bit-exact bar: all integer work).

Every input class of SURVEY.md section 4 is covered, plus configuration variants and the whole-file
driver; streams are also decoded by the independent decoder (lossless, CRC-8/CRC-16/MD5 valid).
"""
import numpy as np
import pytest

import signals

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def encoders(zf):
    cache = {}

    def get(channels, bits, sample_rate=44100, **kw):
        key = (channels, bits, sample_rate, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = zf.Encoder(zf.Config(channels, bits, **kw), sample_rate, max_frames_per_batch=64)
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


def _compare(oracle, enc, pcm, n, channels, bits, sample_rate, first=0, **kw):
    cfg = oracle.config(channels, bits, **{k: int(v) for k, v in kw.items()})
    ref, ref_sizes = oracle.encode_pcm(pcm, n, cfg, sample_rate, first)
    got, got_sizes = enc.encode_pcm(pcm, n, first)
    assert np.array_equal(ref_sizes, got_sizes), (ref_sizes[:8], got_sizes[:8])
    assert ref.tobytes() == got.tobytes()
    return got


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_stereo_input_classes(zf, oracle, encoders, bits):
    enc = encoders(2, bits)
    for name, L, R in signals.stereo_classes(bits):
        pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
        got = _compare(oracle, enc, pcm, L.size, 2, bits, 44100)
        d = oracle.decode(oracle.wrap_frames(got, 2, bits, 44100))
        assert d["rc"] == 0, (name, d["rc"])
        assert np.array_equal(d["pcm"], signals.interleave([L, R])), name


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_short_last_frames(zf, oracle, encoders, bits):
    enc = encoders(2, bits)
    rng = np.random.default_rng(bits)
    F = 1 << (bits - 1)
    for m in signals.SHORT_LENGTHS:
        t = np.arange(m)
        L = (0.3 * F * np.sin(t * 0.05)).astype(np.int64) + rng.integers(-3, 4, m)
        R = L // 2
        pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
        _compare(oracle, enc, pcm, m, 2, bits, 44100)
        # and behind a full frame, so the short frame goes through the second launch
        L2 = np.concatenate([L, L])[: 4096 + m] if 2 * m >= 4096 + m else np.concatenate([np.resize(L, 4096), L])
        R2 = L2 // 3
        pcm2 = oracle.pcm_bytes_from_int(signals.interleave([L2, R2]), bits)
        _compare(oracle, enc, pcm2, L2.size, 2, bits, 44100)


def test_long_short_frame_behind_one_full_frame_repeated(zf, oracle, encoders):
    # the short-frame launch (one CTA, up to 4095 samples) outlasts the one-frame persistent kernel that is launched as its
    # programmatic dependent; the append step must still see the finished short frame, every time
    enc = encoders(2, 24)
    rng = np.random.default_rng(99)
    n = 4096 + 4095
    x = rng.integers(-(1 << 22), 1 << 22, size=(n, 2), dtype=np.int64)
    pcm = oracle.pcm_bytes_from_int(x.reshape(-1), 24)
    cfg = oracle.config(2, 24)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, cfg, 44100, 0)
    for _ in range(40):
        got, got_sizes = enc.encode_pcm(pcm, n, 0)
        assert np.array_equal(ref_sizes, got_sizes)
        assert ref.tobytes() == got.tobytes()


@pytest.mark.parametrize("bits", [16, 24])
def test_frame_numbers_and_sample_rates(zf, oracle, encoders, bits):
    rng = np.random.default_rng(5)
    F = 1 << (bits - 1)
    n = 4096 + 100
    L = rng.integers(-F // 8, F // 8, n)
    R = L + rng.integers(-50, 50, n)
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
    enc = encoders(2, bits)
    for fn in signals.FRAME_NUMBERS:
        _compare(oracle, enc, pcm, n, 2, bits, 44100, first=fn)
    for sr in signals.SAMPLE_RATES:
        _compare(oracle, encoders(2, bits, sample_rate=sr), pcm, n, 2, bits, sr)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_config_variants(zf, oracle, encoders, bits):
    rng = np.random.default_rng(11)
    F = 1 << (bits - 1)
    n = 2 * 4096 + 333
    t = np.arange(n)
    L = (0.3 * F * np.sin(t * 0.02)).astype(np.int64) + rng.integers(-F // 512, F // 512 + 1, n)
    R = (0.2 * F * np.sin(t * 0.031)).astype(np.int64) + rng.integers(-4, 5, n)
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
    for mro in (0, 3, 7):
        _compare(oracle, encoders(2, bits, max_rice_order=mro), pcm, n, 2, bits, 44100, max_rice_order=mro)
    for mrp in (1, 5, 14, 15, 29):
        _compare(oracle, encoders(2, bits, max_rice_param=mrp), pcm, n, 2, bits, 44100, max_rice_param=mrp)
    for block in (16, 192, 576, 1000, 1024, 2048, 4000):
        _compare(oracle, encoders(2, bits, block_size=block), pcm, n, 2, bits, 44100, block_size=block)
    _compare(oracle, encoders(2, bits, stereo_decorrelation=False), pcm, n, 2, bits, 44100, stereo_decorrelation=0)


@pytest.mark.parametrize("bits", [16, 24, 32])
@pytest.mark.parametrize("channels", [1, 3, 6, 8])
def test_independent_channels(zf, oracle, encoders, bits, channels):
    rng = np.random.default_rng(channels * 100 + bits)
    F = 1 << (bits - 1)
    for n in (4096 * 2 + 100, 4096, 17, 3):
        t = np.arange(n)
        planes = []
        for c in range(channels):
            k = c % 4
            if k == 0:
                v = (0.3 * F * np.sin(t * 0.01 * (c + 1))).astype(np.int64) + rng.integers(-5, 6, n)
            elif k == 1:
                v = rng.integers(-F, F, n)
            elif k == 2:
                v = np.full(n, (c * 1000) % F)
            else:
                v = (rng.integers(-F // 16, F // 16, n) >> 4) << 4
            planes.append(v)
        pcm = oracle.pcm_bytes_from_int(signals.interleave(planes), bits)
        _compare(oracle, encoders(channels, bits, sample_rate=48000), pcm, n, channels, bits, 48000)


def test_write_frame_matches_batched_path(zf, oracle, encoders):
    """Encoder.writeFrame (planar int32, one frame) == the K = 1 case of the batched entry."""
    rng = np.random.default_rng(2)
    n = 4096
    L = rng.integers(-2000, 2000, n).astype(np.int32)
    R = (L // 2 + rng.integers(-10, 10, n)).astype(np.int32)
    enc = encoders(2, 16)
    import io
    w = io.BytesIO()
    size = enc.write_frame(w, 7, [L, R])
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), 16)
    ref, sizes = oracle.encode_pcm(pcm, n, oracle.config(2, 16), 44100, 7)
    assert size == sizes[0] and w.getvalue() == ref.tobytes()


@pytest.mark.parametrize("bits,rate", [(16, 44100), (24, 96000), (32, 192000)])
def test_synthetic_stream_multi_batch(zf, oracle, bits, rate):
    """BASELINE signal; more frames than one batch so the three-stage host pipeline and the look-back both run."""
    n = 4096 * 150 + 2048
    pcm = zf.synth_pcm(n, rate, bits)
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=64) as enc:
        got, sizes = enc.encode_pcm(pcm, n)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, threads=8)
    assert np.array_equal(sizes, ref_sizes)
    assert got.tobytes() == ref.tobytes()


def test_whole_file_driver(zf, oracle):
    """`flac in.wav out.flac`: byte-identical file, valid MD5 / STREAMINFO (wav2flac.zig:10-97)."""
    for bits, rate, n in [(16, 44100, 44100 * 3 + 17), (24, 96000, 96000 + 5)]:
        pcm = zf.synth_pcm(n, rate, bits)
        wav = oracle.make_wav(pcm, 2, bits, rate, extensible=(bits == 24))
        rc_ref, ref = oracle.wav_to_flac(wav)
        rc, got = zf.wav_to_flac(wav)
        assert rc == 0 and rc_ref == 0
        assert got == ref
        d = oracle.decode(got)
        assert d["rc"] == 0 and d["md5_ok"] == 1


@pytest.mark.parametrize("bits,rate", [(16, 44100), (24, 96000)])
def test_long_single_batch_all_ctas(zf, oracle, bits, rate):
    """One launch with many more frames than resident CTAs (3 per SM): the persistent loop, the deferred epilogue and
    its cp.async look-back run with every CTA in flight; frames must come out in order, byte for byte."""
    n = 4096 * 3000 + 1234
    pcm = zf.synth_pcm(n, rate, bits)
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=3001) as enc:
        got, sizes = enc.encode_pcm(pcm, n, 77)
        got2, sizes2 = enc.encode_pcm(pcm, n, 77)  # and again on the warm handle: no state may leak between launches
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, 77, threads=8)
    assert np.array_equal(sizes, ref_sizes) and np.array_equal(sizes2, ref_sizes)
    assert got.tobytes() == ref.tobytes()
    assert got2.tobytes() == ref.tobytes()


def test_pageable_and_pinned_host_buffers_agree(zf, oracle):
    """zf_encode_pcm stages pageable caller memory through pinned buffers; the three-stage pipeline
    (upload | encode | download) must give the same bytes either way and for any batch size."""
    bits, rate = 24, 96000
    n = 4096 * 37 + 100
    pcm = zf.synth_pcm(n, rate, bits)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, threads=8)
    for per in (1, 2, 5, 16, 64):
        with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=per) as enc:
            got, sizes = enc.encode_pcm(pcm, n)
        assert np.array_equal(sizes, ref_sizes), per
        assert got.tobytes() == ref.tobytes(), per


def test_host_alloc_buffers(zf, oracle):
    # PCM read into, and FLAC returned in, memory from zf_host_alloc: same bytes as through pageable memory
    n = 50 * 4096 + 99
    pcm = zf.synth_pcm(n, 48000, 24)
    enc = zf.Encoder(zf.Config.default(2, 24), 48000, max_frames_per_batch=16)
    try:
        ref, ref_sizes = enc.encode_pcm(pcm, n, 0)
        with zf.HostBuffer(pcm.size) as hin, zf.HostBuffer(enc.max_batch_bytes(51)) as hout:
            hin.array[:] = pcm
            got, got_sizes = enc.encode_pcm(hin.array, n, 0, out=hout.array)
            assert np.array_equal(ref_sizes, got_sizes)
            assert ref.tobytes() == got.tobytes()
        cfg = oracle.config(2, 24)
        o, o_sizes = oracle.encode_pcm(pcm, n, cfg, 48000, 0)
        assert o.tobytes() == ref.tobytes() and np.array_equal(o_sizes, ref_sizes)
    finally:
        enc.close()


def test_full_size_config2_properties(zf, oracle):
    """BASELINE config 2 at full size (24-bit / 96 kHz / 600 s: 14 063 frames, 345.6 MB of PCM), checked through
    size-independent properties: the frame sizes add up to the stream, two frame-range shards concatenate to the
    single-shot stream (what the multi-GPU driver relies on), the independent decoder accepts every frame (CRC-8 and
    CRC-16 valid) and returns the PCM (lossless), and a slice of the stream is byte-identical to the oracle's."""
    bits, rate, seconds = 24, 96000, 600
    n = rate * seconds
    pcm = zf.synth_pcm(n, rate, bits)
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=2048) as enc:
        got, sizes = enc.encode_pcm(pcm, n)
        assert sizes.size == (n + 4095) // 4096 and int(sizes.astype(np.int64).sum()) == got.size
        half = (sizes.size // 2) * 4096
        a, sa = enc.encode_pcm(pcm[: half * 6], half, 0)
        b, sb = enc.encode_pcm(pcm[half * 6:], n - half, sizes.size // 2)
    assert np.array_equal(np.concatenate([sa, sb]), sizes)
    assert np.concatenate([a, b]).tobytes() == got.tobytes()
    # oracle on the first and the last 100 frames
    k = 100 * 4096
    ref, rs = oracle.encode_pcm(pcm[: k * 6], k, oracle.config(2, bits), rate, 0, threads=8)
    assert ref.tobytes() == got[: ref.size].tobytes()
    f0 = sizes.size - 100
    ref, rs = oracle.encode_pcm(pcm[f0 * 4096 * 6:], n - f0 * 4096, oracle.config(2, bits), rate, f0, threads=8)
    assert ref.tobytes() == got[got.size - ref.size:].tobytes()
    # independent decoder: every frame's CRCs, lossless samples
    d = oracle.decode(oracle.wrap_frames(got, 2, bits, rate, total_samples=n), max_frames=1 << 14)
    assert d["rc"] == 0 and d["n_frames"] == sizes.size
    raw = np.frombuffer(pcm, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    expect = (raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16))
    expect = (expect ^ 0x800000) - 0x800000
    assert np.array_equal(d["pcm"], expect)


def test_error_behaviour(zf, oracle):
    """Writer.Error.WriteFailed <-> ZF_ERR_OUT_TOO_SMALL, argument checks, and the handle stays usable afterwards."""
    bits, rate = 16, 44100
    n = 4096 * 6 + 10
    pcm = zf.synth_pcm(n, rate, bits)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate)
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=4) as enc:
        small = np.empty(1000, dtype=np.uint8)
        with pytest.raises(zf.FlacGpuError) as ei:
            enc.encode_pcm(pcm, n, out=small)
        assert ei.value.status == zf.ZF_ERR_OUT_TOO_SMALL
        got, sizes = enc.encode_pcm(pcm, n)  # the same handle still works
        assert np.array_equal(sizes, ref_sizes) and got.tobytes() == ref.tobytes()
        empty, es = enc.encode_pcm(pcm[:0], 0)
        assert empty.size == 0 and es.size == 0
    for bad in (dict(bit_depth=20), dict(block_size=8192), dict(max_rice_order=9), dict(max_rice_param=31)):
        kw = dict(bit_depth=16)
        kw.update(bad)
        depth = kw.pop("bit_depth")
        with pytest.raises(zf.FlacGpuError):
            zf.Encoder(zf.Config(2, depth, **kw), rate)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_patchwork_stream_many_frames_per_cta(zf, oracle, bits):
    """Frames of wildly different sizes follow one another in every CTA (silence, constants, tones, noise at random levels,
    full scale, wasted bits): the bit buffer is not cleared between frames, so stale words of a long frame lie behind a
    short one."""
    rng = np.random.default_rng(1000 + bits)
    F = 1 << (bits - 1)
    frames = 1800  # about four frames per CTA of the persistent kernel
    seg = []
    total = 0
    while total < frames * 4096 + 1234:
        m = int(rng.integers(300, 3 * 4096))
        kind = int(rng.integers(0, 8))
        amp = float(F) * 10.0 ** float(rng.uniform(-4.5, 0))
        t = np.arange(m)
        if kind == 0:
            L = np.zeros(m, dtype=np.int64); R = np.zeros(m, dtype=np.int64)
        elif kind == 1:
            L = np.full(m, int(rng.integers(-F, F)), dtype=np.int64); R = np.full(m, int(rng.integers(-F, F)), dtype=np.int64)
        elif kind == 2:
            L = (amp * 0.9 * np.sin(t * rng.uniform(0.001, 0.5))).astype(np.int64); R = (L * rng.uniform(-1, 1)).astype(np.int64)
        elif kind == 3:
            L = rng.integers(-F, F, m, dtype=np.int64); R = rng.integers(-F, F, m, dtype=np.int64)  # full-scale noise
        elif kind == 4:
            L = rng.normal(0, amp / 4, m).astype(np.int64); R = L + rng.integers(-2, 3, m)
        elif kind == 5:
            sh = int(rng.integers(1, bits - 2))
            L = (rng.integers(-F, F, m, dtype=np.int64) >> sh) << sh; R = (rng.integers(-F, F, m, dtype=np.int64) >> sh) << sh
        elif kind == 6:
            L = np.cumsum(rng.integers(-3, 4, m)).astype(np.int64); R = -L
        else:
            L = np.where(rng.random(m) < 0.01, rng.integers(-F, F, m, dtype=np.int64), 0); R = L[::-1].copy()
        seg.append((np.clip(L, -F, F - 1), np.clip(R, -F, F - 1)))
        total += m
    L = np.concatenate([a for a, _ in seg]); R = np.concatenate([b for _, b in seg])
    n = L.size
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
    cfg = oracle.config(2, bits)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, cfg, 48000, 0)
    enc = zf.Encoder(zf.Config.default(2, bits), 48000, max_frames_per_batch=4096)
    try:
        for _ in range(2):
            got, got_sizes = enc.encode_pcm(pcm, n, 0)
            assert np.array_equal(ref_sizes, got_sizes)
            assert ref.tobytes() == got.tobytes()
    finally:
        enc.close()


def _random_stream(rng, n, channels, bits):
    F = 1 << (bits - 1)
    planes = []
    for c in range(channels):
        kind = int(rng.integers(0, 7))
        t = np.arange(n)
        amp = float(F) * 10.0 ** float(rng.uniform(-4, 0))
        if kind == 0:
            v = np.zeros(n)
        elif kind == 1:
            v = np.full(n, float(rng.integers(-F, F)))
        elif kind == 2:
            v = amp * 0.9 * np.sin(t * rng.uniform(0.001, 1.5) + rng.uniform(0, 6))
        elif kind == 3:
            v = rng.integers(-F, F, n).astype(np.float64)
        elif kind == 4:
            v = rng.normal(0, amp / 4, n)
        elif kind == 5:
            sh = int(rng.integers(1, bits - 1))
            v = ((rng.integers(-F, F, n) >> sh) << sh).astype(np.float64)
        else:
            v = np.cumsum(rng.integers(-5, 6, n)).astype(np.float64)
        planes.append(np.clip(v.astype(np.int64), -F, F - 1))
    if channels == 2 and rng.random() < 0.3:
        planes[1] = np.clip(planes[0] + rng.integers(-3, 4, n), -F, F - 1)  # correlated pair: the side channel wins
    return planes


def test_random_configurations(zf, oracle):
    """Randomised sweep over what Encoder.Config can express (encoder.zig:609-656): block size, channels, depth, Rice
    limits, decorrelation, sample rate, stream length (any tail), first frame number -- every stream byte-identical to
    the oracle, and decoded by the device decoder back to the PCM that went in.  Seeded: a failure prints its case."""
    rng = np.random.default_rng(20250)
    dec_handle = zf.Decoder()
    for case in range(160):
        bits = int(rng.choice([16, 24, 32]))
        channels = int(rng.choice([1, 2, 2, 2, 3, 6]))
        block = int(rng.choice([4096, 4096, 4096, 2048, 1024, 576, 192, 1000, 4000, 16, 64]))
        mro = int(rng.choice([8, 8, 8, 0, 3, 6]))
        mrp = int(rng.choice([30, 30, 14, 1, 7, 20]))
        dec = bool(rng.random() < 0.85)
        rate = int(rng.choice([44100, 48000, 96000, 192000, 8000, 12345, 200, 100000]))
        frames = int(rng.integers(1, 5))
        n = block * (frames - 1) + int(rng.integers(1, block + 1))
        first = int(rng.choice([0, 0, 127, 2047, 65535, (1 << 21) - 1]))
        planes = _random_stream(rng, n, channels, bits)
        pcm = oracle.pcm_bytes_from_int(signals.interleave(planes), bits)
        what = dict(case=case, bits=bits, channels=channels, block=block, mro=mro, mrp=mrp, dec=dec, rate=rate, n=n, first=first)
        cfg = oracle.config(channels, bits, block_size=block, stereo_decorrelation=int(dec), max_rice_order=mro, max_rice_param=mrp)
        ref, rs = oracle.encode_pcm(pcm, n, cfg, rate, first)
        with zf.Encoder(zf.Config(channels, bits, block_size=block, stereo_decorrelation=dec, max_rice_order=mro,
                                  max_rice_param=mrp), rate, max_frames_per_batch=3) as enc:
            got, gs = enc.encode_pcm(pcm, n, first)
        assert np.array_equal(rs, gs), what
        assert ref.tobytes() == got.tobytes(), what
        if rate >= 256:  # below 256 Hz upstream's header writer ORs block-size bits into the frame-number byte:
                         # byte-identical to the reference, but not a FLAC stream any decoder could walk
            back, info = dec_handle.decode(oracle.wrap_frames(got, channels, bits, rate, block, n))
            assert back.tobytes() == pcm.tobytes() and info["n_frames"] == gs.size, what
    dec_handle.close()

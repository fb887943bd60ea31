"""GPU parity at BASELINE.json's full sizes and through every C-ABI entry that carries frames.

* every frame of config 1 (16-bit / 44.1 kHz / 60 s, 646 frames, last 4080 samples), config 2 (24-bit / 96 kHz / 600 s,
  14 063 frames), config 3 (32-bit / 192 kHz / 600 s, 28 125 frames) and the last of eight shards of config 5
  (24-bit / 96 kHz / 10 h: frames 738 283 .. 843 749) is compared with the oracle, slice by slice (1 000 frames)
  so that a failure names the first differing frame -- mirrors the loop of wav2flac.zig:66-97 over encoder.zig:234;
* zf_encode_submit / zf_encode_collect, zf_encode_device, the multi-device whole-file driver
  (zf_encode_wav_memory with a device list) and the `flac` CLI (cli.zig:17-20, wav2flac.zig:24-27 exit codes).
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = 4096


def _first_diff(a, b):
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return int(d[0]) if d.size else n


def _compare_stream(got, sizes, ref, ref_sizes, what):
    """Frame sizes first, then sha256 per 1 000-frame slice; reports the first bad frame."""
    assert sizes.size == ref_sizes.size, (what, sizes.size, ref_sizes.size)
    if not np.array_equal(sizes, ref_sizes):
        f = int(np.nonzero(sizes != ref_sizes)[0][0])
        pytest.fail(f"{what}: frame {f} has {int(sizes[f])} bytes, oracle {int(ref_sizes[f])}")
    assert got.size == ref.size == int(ref_sizes.astype(np.int64).sum()), what
    offs = np.concatenate([[0], np.cumsum(ref_sizes.astype(np.int64))])
    for f0 in range(0, ref_sizes.size, 1000):
        f1 = min(f0 + 1000, ref_sizes.size)
        a, b = got[offs[f0]:offs[f1]], ref[offs[f0]:offs[f1]]
        if hashlib.sha256(a).digest() != hashlib.sha256(b).digest():
            pos = int(offs[f0]) + _first_diff(a, b)
            f = int(np.searchsorted(offs, pos, side="right")) - 1
            pytest.fail(f"{what}: slice {f0}..{f1} differs; first at frame {f}, byte {pos - int(offs[f])} of the frame")


FULL = [
    # name, bits, rate, samples per channel, first frame number, first sample of the synthetic stream
    ("c1_16bit_44k1_60s", 16, 44100, 44100 * 60, 0, 0),
    ("c2_24bit_96k_600s", 24, 96000, 96000 * 600, 0, 0),
    ("c3_32bit_192k_600s", 32, 192000, 192000 * 600, 0, 0),
    # config 5, rank 7 of 8: ceil(843 750 / 8) = 105 469 frames per rank; the last rank holds 105 467, ending at 843 749
    ("c5_24bit_96k_10h_rank7of8", 24, 96000, 96000 * 36000 - 7 * 105469 * BLOCK, 7 * 105469, 7 * 105469 * BLOCK),
]


@pytest.mark.parametrize("name,bits,rate,n,first_frame,first_sample", FULL, ids=[c[0] for c in FULL])
def test_every_frame_of_full_config(zf, oracle, name, bits, rate, n, first_frame, first_sample):
    pcm = zf.synth_pcm(n, rate, bits, first_sample=first_sample)
    frames = (n + BLOCK - 1) // BLOCK
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=2048) as enc:
        got, sizes = enc.encode_pcm(pcm, n, first_frame)
    assert sizes.size == frames
    if name.startswith("c5"):
        assert first_frame + frames - 1 == 843749
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, first_frame, threads=os.cpu_count() or 1)
    _compare_stream(got, sizes, ref, ref_sizes, name)


def test_submit_collect_matches_oracle(zf, oracle):
    """zf_encode_submit / zf_encode_collect (include/zigflac_b200.h): one batch in flight, BUSY while it is; pinned input
    may be overwritten as soon as submit has returned."""
    bits, rate = 24, 96000
    n = 300 * BLOCK + 777
    pcm = zf.synth_pcm(n, rate, bits)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, 5, threads=8)
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=512) as enc:
        for rnd in range(3):
            enc.submit(pcm, n, 5)
            with pytest.raises(zf.FlacGpuError) as ei:
                enc.submit(pcm, n, 5)
            assert ei.value.status == zf.ZF_ERR_BUSY
            got, sizes = enc.collect()
            _compare_stream(got, sizes, ref, ref_sizes, f"submit/collect round {rnd}")
        # page-locked input, refilled right behind submit: the upload must not read the new content
        with zf.HostBuffer(pcm.size) as hb:
            for rnd in range(5):
                hb.array[:] = pcm
                enc.submit(hb.array, n, 5)
                hb.array[:] = 0x55
                got, sizes = enc.collect()
                _compare_stream(got, sizes, ref, ref_sizes, f"pinned submit round {rnd}")
        # the blocking entry on the same handle afterwards
        got, sizes = enc.encode_pcm(pcm, n, 5)
        _compare_stream(got, sizes, ref, ref_sizes, "encode_pcm after submit/collect")


@pytest.mark.parametrize("bits,rate", [(16, 44100), (24, 96000), (32, 192000)])
def test_encode_device_matches_oracle(zf, oracle, bits, rate):
    """zf_encode_device: device pointers in, device pointers out, on a caller stream; consecutive calls on different
    streams are ordered by the library; overflow is reported through the status word and the total."""
    import torch
    dev = torch.device("cuda", 0)
    n = 700 * BLOCK + 1000
    pcm = zf.synth_pcm(n, rate, bits)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, 3, threads=8)
    frames = ref_sizes.size
    d_pcm = torch.from_numpy(pcm).to(dev)
    with zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=frames) as enc:
        cap = enc.max_batch_bytes(frames)
        outs = []
        streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev), None]
        for s in streams:
            d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
            d_sizes = torch.zeros(frames, dtype=torch.int32, device=dev)
            d_total = torch.zeros(1, dtype=torch.int64, device=dev)
            enc.encode_device(d_pcm.data_ptr(), n, 3, d_out.data_ptr(), cap, d_sizes.data_ptr(), d_total.data_ptr(),
                              s.cuda_stream if s is not None else None)
            outs.append((d_out, d_sizes, d_total))
        assert enc.device_status() == 0
        torch.cuda.synchronize(dev)
        for k, (d_out, d_sizes, d_total) in enumerate(outs):
            total = int(d_total.item())
            _compare_stream(d_out[:total].cpu().numpy(), d_sizes.cpu().numpy().astype(np.uint32), ref, ref_sizes,
                            f"encode_device call {k}")
        # capacity one byte short of the stream: the last frame is dropped, the status says so, the total still counts it
        small = int(ref.size) - 1
        d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
        d_sizes = torch.zeros(frames, dtype=torch.int32, device=dev)
        d_total = torch.zeros(1, dtype=torch.int64, device=dev)
        enc.encode_device(d_pcm.data_ptr(), n, 3, d_out.data_ptr(), small, d_sizes.data_ptr(), d_total.data_ptr(), None)
        assert enc.device_status() & 1
        assert int(d_total.item()) == ref.size > small
        head = int(ref.size) - int(ref_sizes[-1])
        assert d_out[:head].cpu().numpy().tobytes() == ref[:head].tobytes()
        assert int(d_out[head:].max().item()) == 0  # nothing of the frame that did not fit
        # submit while nothing is pending works again, and a device call while a batch is submitted is refused
        enc.submit(pcm, n, 3)
        with pytest.raises(zf.FlacGpuError) as ei:
            enc.encode_device(d_pcm.data_ptr(), n, 3, d_out.data_ptr(), cap, d_sizes.data_ptr(), d_total.data_ptr(), None)
        assert ei.value.status == zf.ZF_ERR_BUSY
        got, sizes = enc.collect()
        _compare_stream(got, sizes, ref, ref_sizes, "submit after encode_device")


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0, 1], [1, 0, 1, 0]])
def test_multi_device_whole_file(zf, oracle, devices):
    """zf_encode_wav_memory(devices=[...]): N host threads, N handles, contiguous frame ranges, ordered concatenation,
    sequential min/max replay == the one-thread oracle file."""
    if max(devices) >= _gpu_count():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    bits, rate = 24, 96000
    n = rate * 20 + 1234  # 469 frames: shards of unequal length, the last one ends short
    pcm = zf.synth_pcm(n, rate, bits)
    wav = oracle.make_wav(pcm, 2, bits, rate)
    rc_ref, ref = oracle.wav_to_flac(wav, threads=8)
    rc, got = zf.wav_to_flac(wav, devices=devices)
    assert rc == 0 and rc_ref == 0
    assert got == ref
    d = oracle.decode(got)
    assert d["rc"] == 0 and d["md5_ok"] == 1


def test_cli_binary_exit_codes(zf, oracle, tmp_path):
    """`flac in.wav out.flac` (cli.zig:7-27): exit 0 and the oracle's file; exit 1 on bad usage (cli.zig:17-20);
    exit 2 on a format FLAC cannot carry (wav2flac.zig:24-27)."""
    cli = os.path.join(ROOT, "zig-flac_b200", "flac")
    assert os.path.exists(cli), "run python zig-flac_b200/build.py"
    bits, rate = 16, 44100
    n = rate * 2 + 99
    pcm = zf.synth_pcm(n, rate, bits)
    wav = oracle.make_wav(pcm, 2, bits, rate)
    inp, outp = tmp_path / "in.wav", tmp_path / "out.flac"
    inp.write_bytes(wav)
    r = subprocess.run([cli, str(inp), str(outp)], capture_output=True)
    assert r.returncode == 0, r.stderr
    rc_ref, ref = oracle.wav_to_flac(wav)
    assert outp.read_bytes() == ref
    # usage errors
    for argv in ([cli], [cli, str(inp)]):
        r = subprocess.run(argv, capture_output=True)
        assert r.returncode == 1
        assert b"usage: flac in_file.wav out_file.flac" in r.stderr
    # 9 channels: FLAC carries at most 8 (wav_reader.zig:95-98 -> null -> exit 2)
    import struct
    ch = 9
    body9 = bytes(2 * ch * 16)
    fmt = struct.pack("<HHIIHH", 1, ch, rate, rate * 2 * ch, 2 * ch, 16)
    riff = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(body9)) + body9
    bad = tmp_path / "nine.wav"
    bad.write_bytes(b"RIFF" + struct.pack("<I", len(riff)) + riff)
    r = subprocess.run([cli, str(bad), str(tmp_path / "nine.flac")], capture_output=True)
    assert r.returncode == 2
    assert b"flac does not support this wav format" in r.stderr
    rc_o, _ = oracle.wav_to_flac(bad.read_bytes())
    assert rc_o == 2


@pytest.mark.parametrize("channels", [1, 2, 5])
def test_8bit_wav(zf, oracle, channels):
    """8-bit WAV (wav_reader.zig:71-88 with the subtract-from-the-unshifted-word quirk, frame header depth code 2):
    whole file == the oracle's file; the batched entry on the reader's signed samples == the oracle's frames."""
    rng = np.random.default_rng(80 + channels)
    n = 22050 * 3 + 321
    t = np.arange(n)
    cols = []
    for c in range(channels):
        k = c % 3
        if k == 0:
            v = np.clip(128 + 100 * np.sin(t * 0.01 * (c + 1)) + rng.integers(-2, 3, n), 0, 255)
        elif k == 1:
            v = rng.integers(0, 256, n)
        else:
            v = np.where((t // 900) % 3 == 0, 0, np.where((t // 900) % 3 == 1, 128, 64 + (t % 7)))
        cols.append(v.astype(np.uint8))
    raw = np.stack(cols, axis=1).reshape(-1).copy()
    wav = oracle.make_wav(raw.tobytes(), channels, 8, 22050)
    rc_ref, ref = oracle.wav_to_flac(wav)
    rc, got = zf.wav_to_flac(wav)
    assert rc == 0 and rc_ref == 0
    assert got == ref
    signed = zf.Wav8Reader(channels).convert(raw)
    d = oracle.decode(got)
    assert d["rc"] == 0 and np.array_equal(d["pcm"], signed.astype(np.int32))
    fr, fs = oracle.encode_pcm(raw, n, oracle.config(channels, 8), 22050, 0)
    with zf.Encoder(zf.Config.default(channels, 8), 22050, max_frames_per_batch=7) as enc:
        g, gs = enc.encode_pcm(signed, n, 0)
    assert np.array_equal(gs, fs) and g.tobytes() == fr.tobytes()


@pytest.mark.parametrize("bits,rate,seconds", [(16, 44100, 60), (24, 96000, 12)])
def test_lpc_extension_config4(zf, oracle, bits, rate, seconds):
    """BASELINE config 4 (16-bit / 44.1 kHz, LPC order <= 12 with quantised coefficients).  The reference has no LPC
    (encoder.zig:626-640 is a stub), so the bar is: (1) the GPU stream is bit-identical to the CPU statement of the
    same specification (oracle/zigflac_lpc.h); (2) the independent decoder returns the PCM; (3) the stream is not
    larger than the reference's FIXED-only stream."""
    n = rate * seconds
    pcm = zf.synth_pcm(n, rate, bits)
    with zf.Encoder(zf.Config(2, bits, lpc_order=12), rate, max_frames_per_batch=512) as enc:
        got, sizes = enc.encode_pcm(pcm, n, 0)
    ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits, lpc_order=12), rate, 0, threads=os.cpu_count() or 1)
    _compare_stream(got, sizes, ref, ref_sizes, f"lpc {bits}-bit")
    d = oracle.decode(oracle.wrap_frames(got, 2, bits, rate, total_samples=n), max_frames=1 << 15)
    assert d["rc"] == 0 and d["n_frames"] == sizes.size
    raw = np.frombuffer(pcm, dtype=np.uint8).reshape(-1, bits // 8).astype(np.int32)
    expect = sum(raw[:, k] << (8 * k) for k in range(bits // 8))
    expect = (expect ^ (1 << (bits - 1))) - (1 << (bits - 1))
    assert np.array_equal(d["pcm"], expect)
    assert sum(1 for f in d["frames"] for s in f.sub[:2] if s.type == 3) > sizes.size  # mostly LPC subframes
    fixed, _ = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, 0, threads=os.cpu_count() or 1)
    assert got.size < fixed.size, (got.size, fixed.size)
    # outside the extension's scope: refused, not silently encoded some other way
    for bad in (dict(channels=1, bits=16, lpc_order=12), dict(channels=2, bits=32, lpc_order=12),
                dict(channels=2, bits=16, lpc_order=13), dict(channels=2, bits=16, lpc_order=4, stereo_decorrelation=False)):
        kw = dict(bad)
        ch, b = kw.pop("channels"), kw.pop("bits")
        with pytest.raises(zf.FlacGpuError):
            zf.Encoder(zf.Config(ch, b, **kw), rate)


def test_lpc_input_classes(zf, oracle):
    """Every stereo input class through the LPC kernel: GPU == CPU statement, lossless."""
    import signals
    for bits in (16, 24):
        with zf.Encoder(zf.Config(2, bits, lpc_order=12), 48000, max_frames_per_batch=8) as enc:
            for name, L, R in signals.stereo_classes(bits):
                pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
                ref, rs = oracle.encode_pcm(pcm, L.size, oracle.config(2, bits, lpc_order=12), 48000, 0)
                got, gs = enc.encode_pcm(pcm, L.size, 0)
                assert np.array_equal(rs, gs), (name, bits)
                assert ref.tobytes() == got.tobytes(), (name, bits)
                d = oracle.decode(oracle.wrap_frames(got, 2, bits, 48000))
                assert d["rc"] == 0 and np.array_equal(d["pcm"], signals.interleave([L, R])), (name, bits)


@pytest.mark.parametrize("bits,rate", [(16, 44100), (24, 96000), (32, 192000)])
def test_exact_rice_extension(zf, oracle, bits, rate):
    """zf_config.exact_rice (extension; upstream only has dead code for it, rice.zig:110-245): GPU == the oracle's
    brute-force statement of the rule, lossless, never larger than the estimate's stream."""
    import signals
    n = 40 * BLOCK + 1234
    pcm = zf.synth_pcm(n, rate, bits)
    with zf.Encoder(zf.Config(2, bits, exact_rice=True), rate, max_frames_per_batch=16) as enc:
        got, sizes = enc.encode_pcm(pcm, n, 9)
        ref, ref_sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits, exact_rice=1), rate, 9, threads=os.cpu_count() or 1)
        _compare_stream(got, sizes, ref, ref_sizes, f"exact rice {bits}-bit")
        est, _ = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, 9, threads=os.cpu_count() or 1)
        assert got.size <= est.size
        d = oracle.decode(oracle.wrap_frames(got, 2, bits, rate, total_samples=n), max_frames=1 << 12)
        assert d["rc"] == 0 and d["n_frames"] == sizes.size
        for name, L, R in signals.stereo_classes(bits):
            p = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
            r, rs = oracle.encode_pcm(p, L.size, oracle.config(2, bits, exact_rice=1), rate, 0)
            g, gs = enc.encode_pcm(p, L.size, 0)
            assert np.array_equal(rs, gs) and r.tobytes() == g.tobytes(), (name, bits)
    for bad in (dict(channels=1), dict(lpc_order=8), dict(stereo_decorrelation=False)):
        kw = dict(channels=2, exact_rice=True)
        kw.update(bad)
        ch = kw.pop("channels")
        with pytest.raises(zf.FlacGpuError):
            zf.Encoder(zf.Config(ch, 16 if bits == 32 else bits, **kw), rate)

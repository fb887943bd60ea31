"""GPU tests of the decoder (zf_decoder_* / zf_decode_flac*, zf_kernel_decode.cuh) through the C ABI.

The reference has no decoder (readme.md:33), so the contract is the FLAC format: every stream must decode to the PCM that
went into the encoder, identically to the independent CPU decoder oracle/flac_decode.c, with CRC-16 and MD5 verified --
and any damage must be reported, never decoded silently.
"""
import os
import subprocess

import numpy as np
import pytest

import signals
from test_decode_emu import header_image_stream

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dec(zf):
    with zf.Decoder() as d:
        yield d


def _check(dec, oracle, flac, pcm_bytes, **kw):
    got, info = dec.decode(flac, **kw)
    assert got.tobytes() == bytes(pcm_bytes), info
    return info


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_decode_stereo_classes(zf, oracle, dec, bits):
    """Streams of the CUDA encoder for every input class (CONSTANT / VERBATIM / FIXED, wasted bits, escapes, the four
    stereo assignments, 33-bit side channel): decoder output == encoder input == independent decoder's output."""
    with zf.Encoder(zf.Config.default(2, bits), 44100) as enc:
        for name, L, R in signals.stereo_classes(bits):
            pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
            frames, sizes = enc.encode_pcm(pcm, L.size)
            flac = oracle.wrap_frames(frames, 2, bits, 44100, 4096, L.size)
            info = _check(dec, oracle, flac, pcm)
            assert info["n_frames"] == sizes.size and info["samples_per_channel"] == L.size, name
            ref = oracle.decode(flac)
            assert ref["rc"] == 0, name


def test_decode_channels_blocks_lpc_and_limits(zf, oracle, dec):
    rng = np.random.default_rng(5)
    t = np.arange(3000)
    for channels in (1, 2, 3, 8):
        chans = [(3000 * np.sin(2 * np.pi * (200 + 90 * c) * t / 44100)).astype(np.int64) + rng.integers(-20, 21, t.size) for c in range(channels)]
        pcm = oracle.pcm_bytes_from_int(signals.interleave(chans), 16)
        for block in (4096, 1024, 576, 192, 17):
            frames, _ = oracle.encode_pcm(pcm, t.size, oracle.config(channels, 16, block_size=block), 44100)
            _check(dec, oracle, oracle.wrap_frames(frames, channels, 16, 44100, block, t.size), pcm)
    F = 1 << 23
    t = np.arange(4096 + 1500)
    L = (0.3 * F * np.sin(t * 0.3) + 0.2 * F * np.sin(t * 0.71 + 1) + rng.normal(0, F / 500, t.size)).astype(np.int64)
    R = (0.25 * F * np.sin(t * 0.3 + 0.4) + rng.normal(0, F / 300, t.size)).astype(np.int64)
    pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), 24)
    for kw in (dict(lpc_order=4), dict(lpc_order=12), dict(stereo_decorrelation=0), dict(max_rice_param=3, max_rice_order=2)):
        frames, _ = oracle.encode_pcm(pcm, t.size, oracle.config(2, 24, **kw), 96000)
        _check(dec, oracle, oracle.wrap_frames(frames, 2, 24, 96000, 4096, t.size), pcm)
    with zf.Encoder(zf.Config(2, 24, lpc_order=12), 96000) as enc:  # LPC subframes of the CUDA encoder
        frames, _ = enc.encode_pcm(pcm, t.size)
        flac = oracle.wrap_frames(frames, 2, 24, 96000, 4096, t.size)
        assert 3 in {s.type for f in oracle.decode(flac)["frames"] for s in f.sub[:2]}
        _check(dec, oracle, flac, pcm)
    for n in signals.SHORT_LENGTHS:
        p = oracle.pcm_bytes_from_int(signals.interleave([L[:n], R[:n]]), 24)
        frames, _ = oracle.encode_pcm(p, n, oracle.config(2, 24), 96000)
        _check(dec, oracle, oracle.wrap_frames(frames, 2, 24, 96000, 4096, n), p)


def test_decode_whole_files_md5_and_entries(zf, oracle, dec, tmp_path):
    """wav -> flac (CUDA driver) -> wav through every decode entry: handle, one-shot, device pointers, file + CLI."""
    import torch
    for bits, rate, n in [(16, 44100, 44100 * 3 + 17), (24, 96000, 96000 * 2 + 5), (32, 192000, 70000)]:
        pcm = zf.synth_pcm(n, rate, bits)
        wav = oracle.make_wav(pcm, 2, bits, rate)
        rc, flac = zf.wav_to_flac(wav)
        assert rc == 0
        info = _check(dec, oracle, flac, pcm, check_md5=True, require_md5=True)
        assert info["md5_status"] == 1 and info["sample_rate"] == rate and info["bit_depth"] == bits and info["channels"] == 2
        assert info["streaminfo_samples"] == n and info["launches"] >= 4 and info["kernel_ms"] > 0
        assert zf.flac_stream_info(flac)["pcm_bytes"] == pcm.size
        got, _ = zf.decode_flac(flac, check_md5=True)
        assert got.tobytes() == pcm.tobytes()
        # device-resident: FLAC in HBM -> PCM in HBM
        d_flac = torch.from_numpy(np.frombuffer(flac, dtype=np.uint8).copy()).cuda()
        d_pcm = torch.zeros(pcm.size + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        nbytes, info = dec.decode_device(d_flac.data_ptr(), d_flac.numel(), d_pcm.data_ptr(), d_pcm.numel())
        assert nbytes == pcm.size
        assert d_pcm[:nbytes].cpu().numpy().tobytes() == pcm.tobytes()
        assert int(d_pcm[nbytes:].sum()) == 0  # nothing written behind the end
        # pinned output
        with zf.HostBuffer(pcm.size) as hb:
            got, _ = dec.decode(flac, out=hb.array)
            assert got.tobytes() == pcm.tobytes()
        # file to file, and the CLI
        fin, fout, fcli = tmp_path / "a.flac", tmp_path / "a.wav", tmp_path / "b.wav"
        fin.write_bytes(flac)
        assert zf.decode_file(str(fin), str(fout)) == 0
        assert fout.read_bytes() == wav
        r = subprocess.run([os.path.join(ROOT, "zig-flac_b200", "flac"), "-d", str(fin), str(fcli)], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert fcli.read_bytes() == wav
        # verify after encoding: library entry and `flac -V in.wav out.flac`
        fw = tmp_path / "in.wav"
        fw.write_bytes(wav)
        assert zf.verify_file(str(fw), str(fin)) == 0
        fv = tmp_path / "v.flac"
        r = subprocess.run([os.path.join(ROOT, "zig-flac_b200", "flac"), "-V", str(fw), str(fv)], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert fv.read_bytes() == flac
        other = bytearray(wav)
        other[-1] ^= 1  # a WAV that differs in one bit from what the stream holds
        fo = tmp_path / "other.wav"
        fo.write_bytes(bytes(other))
        assert zf.verify_file(str(fo), str(fin)) == -36
    # 8-bit: the stream holds what the reference reader makes of the bytes (wav_reader.zig:71-88); verify knows
    raw8 = (np.arange(9000) * 7 % 256).astype(np.uint8)
    w8 = oracle.make_wav(raw8, 2, 8, 22050)
    rc, f8 = zf.wav_to_flac(w8)
    assert rc == 0
    (tmp_path / "e.wav").write_bytes(w8)
    (tmp_path / "e.flac").write_bytes(f8)
    assert zf.verify_file(str(tmp_path / "e.wav"), str(tmp_path / "e.flac")) == 0
    # MD5 mismatch is reported
    bad = bytearray(flac)
    bad[30] ^= 0xFF  # inside STREAMINFO's MD5
    got, info = dec.decode(bytes(bad), check_md5=True)
    assert info["md5_status"] == 0 and got.tobytes() == pcm.tobytes()
    with pytest.raises(zf.FlacGpuError) as e:
        dec.decode(bytes(bad), require_md5=True)
    assert e.value.status == -36


def test_decode_header_images_and_damage(zf, oracle, dec):
    """Frame data full of sync codes and complete valid header images (one of them with the expected frame number);
    flipped bits, truncation, foreign data: reported, never decoded silently."""
    for hard in (False, True):
        pcm, flac = header_image_stream(oracle, hard)
        info = _check(dec, oracle, flac, pcm, check_md5=True)
        assert info["n_frames"] == 4 and info["md5_status"] == 1
    rng = np.random.default_rng(3)
    body = len(flac) - 73
    for _ in range(25):
        pos = 73 + int(rng.integers(0, body))
        bad = bytearray(flac)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        with pytest.raises(zf.FlacGpuError) as e:
            dec.decode(bytes(bad))
        assert e.value.status in (-34, -35, -33), (pos, e.value.status)
    for trial in range(45):  # bursts of zeros / ones (endless unary runs), random bytes, truncation: reported, and the device survives
        bad = bytearray(flac)
        at = int(rng.integers(80, len(bad) - 200))
        if trial % 3 == 0:
            bad[at:at + 150] = bytes(150) if trial % 2 else bytes([255]) * 150
        elif trial % 3 == 1:
            bad[at:at + 6] = bytes(int(v) for v in rng.integers(0, 256, 6))
        else:
            bad = bad[:at]
        ref = oracle.decode(bytes(bad))
        try:
            got, _ = dec.decode(bytes(bad))
            assert ref["rc"] == 0 and oracle.pcm_bytes_from_int(ref["pcm"], 16).tobytes() == got.tobytes(), trial
        except zf.FlacGpuError as err:
            assert err.status in (-33, -34, -35), (trial, err.status)
    with pytest.raises(zf.FlacGpuError) as e:
        dec.decode(flac[:-5])
    assert e.value.status == -34 and e.value.info["bad_frame"] == 3
    with pytest.raises(zf.FlacGpuError) as e:
        dec.decode(b"RIFF" + flac[4:])
    assert e.value.status == -32
    with pytest.raises(zf.FlacGpuError) as e:
        dec.decode(flac, out=np.zeros(100, np.uint8))
    assert e.value.status == zf.ZF_ERR_OUT_TOO_SMALL
    got, _ = dec.decode(flac)  # the handle is still good
    assert got.tobytes() == pcm.tobytes()


@pytest.mark.parametrize("name,bits,rate,seconds", [("c1", 16, 44100, 60), ("c2", 24, 96000, 600), ("c3", 32, 192000, 120)])
def test_decode_full_size_roundtrip(zf, oracle, dec, name, bits, rate, seconds):
    """BASELINE-size streams: encode on the GPU, decode on the GPU, PCM identical, MD5 verified (several batches in
    flight on c2 / c3)."""
    n = rate * seconds
    pcm = zf.synth_pcm(n, rate, bits)
    rc, flac = zf.wav_to_flac(oracle.make_wav(pcm, 2, bits, rate))
    assert rc == 0
    got, info = dec.decode(flac, check_md5=True)
    assert info["md5_status"] == 1 and info["n_frames"] == (n + 4095) // 4096
    assert got.size == pcm.size
    assert np.array_equal(got, pcm)


def test_decode_crafted_streams(zf, oracle, dec):
    """Streams built from the format by tests/flac_craft.py with features no encoder here emits: LPC orders 20 and 32,
    partition orders 10 and 12 (single-sample partitions, an empty first partition), 5-bit parameters with escapes of
    width 0, explicit 8- / 16-bit block-size fields, wasted bits under mid/side, CONSTANT and VERBATIM beside LPC,
    metadata blocks full of sync patterns."""
    from test_decode_emu import crafted_streams
    for name, flac, expect, channels, bits in crafted_streams():
        pcm = oracle.pcm_bytes_from_int(expect, bits)
        got, info = dec.decode(flac)
        assert got.tobytes() == pcm.tobytes(), name
        assert info["channels"] == channels and info["bit_depth"] == bits, name
    # the encoder's own empty first partition: block 16, order 4, partition order 2
    rng = np.random.default_rng(16)
    for trial in range(40):
        n = 16 * 5
        x = np.cumsum(np.cumsum(np.cumsum(np.cumsum(rng.integers(-3, 4, n))))) % 20000 - 10000
        pcm = oracle.pcm_bytes_from_int(signals.interleave([x, x // 2]), 16)
        frames, _ = oracle.encode_pcm(pcm, n, oracle.config(2, 16, block_size=16), 44100)
        got, _ = dec.decode(oracle.wrap_frames(frames, 2, 16, 44100, 16, n))
        assert got.tobytes() == pcm.tobytes(), trial

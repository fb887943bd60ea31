"""CPU tests of the oracle (C restatement of the reference) against every pin that exists:
the SURVEY 8-K known-answer vectors, catalogue CRC / RFC 1321 MD5 check values, the regression
digests under tests/golden/, and the independent decoder (lossless round trip, valid CRCs, MD5).
The reference has no tests or vectors of its own (parity unpinned upstream)."""
import hashlib
import json
import os

import numpy as np
import pytest

import signals

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLDEN, "kat_vectors.json")))
DIGESTS = json.load(open(os.path.join(GOLDEN, "oracle_digests.json")))


def test_frame_header_vectors(oracle):
    for v in KAT["frame_headers"]:
        got = oracle.frame_header(v["frame"], v["bits"], v["ch_type"], v["block"], v["rate"]).hex()
        assert got == v["hex"], v


def test_utf8_frame_numbers(oracle):
    for number, hexs in KAT["utf8"]:
        hdr = oracle.frame_header(number, 16, 1, 4096, 44100)
        assert hdr[4:-1].hex() == hexs, number


def test_streaminfo_and_vendor_block(oracle):
    s = KAT["streaminfo"]
    got = oracle.streaminfo_bytes(s["min_block"], s["max_block"], s["min_frame"], s["max_frame"], s["rate"],
                                  s["channels"], s["bits"], s["samples"])
    assert got.hex() == s["hex"]
    assert oracle.vorbis_comment(True).hex() == KAT["vorbis_comment_hex"]


def test_crc_and_md5_check_values(oracle):
    assert oracle.crc8(b"123456789") == int(KAT["crc8_123456789"], 16)
    assert oracle.crc16(b"123456789") == int(KAT["crc16_123456789"], 16)
    for m in (b"", b"a", b"abc", b"message digest", b"abcdefghijklmnopqrstuvwxyz", b"1234567890" * 8, bytes(range(256)) * 50):
        assert oracle.md5(m) == hashlib.md5(m).digest()


def test_clmul_folding_equals_table_crc(oracle):
    """crc16.zig:23-57 (PCLMULQDQ folding) restated with intrinsics == the plain CRC-16 it falls back to."""
    rng = np.random.default_rng(1)
    buf = rng.integers(0, 256, 70000, dtype=np.uint8)
    for n in (0, 1, 15, 16, 63, 64, 65, 79, 80, 81, 127, 128, 1000, 4097, 69999):
        assert oracle.crc16(buf[:n]) == oracle.crc16(buf[:n], clmul=True), n


def test_rice_estimate_precedence_quirk(oracle):
    """param 0 has no -(n >> 1); param > 0 has it."""
    f = oracle.lib().zo_flac_calc_part_size
    assert f(16, 0, 100) == 16 + 200
    assert f(16, 1, 100) == 2 * 16 + 100 - 8
    assert f(16, 3, 100) == 4 * 16 + 25 - 8
    assert f(0, 0, 0) == 0


def test_frame_size_replay_is_order_dependent(oracle):
    """`else if` -- a frame that raises the maximum never lowers the minimum."""
    assert oracle.replay_frame_sizes([100]) == (0xFFFFFF, 100)
    assert oracle.replay_frame_sizes([100, 50]) == (50, 100)
    assert oracle.replay_frame_sizes([50, 100]) == (0xFFFFFF, 100)
    assert oracle.replay_frame_sizes([50, 100, 70, 20]) == (20, 100)


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_input_classes_round_trip_and_digests(oracle, bits):
    for name, L, R in signals.stereo_classes(bits):
        inter = signals.interleave([L, R])
        pcm = oracle.pcm_bytes_from_int(inter, bits)
        out, sizes = oracle.encode_pcm(pcm, L.size, oracle.config(2, bits), 44100)
        assert hashlib.sha256(out.tobytes()).hexdigest() == DIGESTS[f"class_{name}_{bits}"]["flac_sha256"], name
        d = oracle.decode(oracle.wrap_frames(out, 2, bits, 44100))
        assert d["rc"] == 0, (name, d["rc"])
        assert np.array_equal(d["pcm"], inter), name
        assert [f.size for f in d["frames"]] == list(sizes)


def test_expected_branches_are_taken(oracle):
    """The classes really exercise the branches they are named after."""
    def info(L, R, bits):
        pcm = oracle.pcm_bytes_from_int(signals.interleave([L, R]), bits)
        out, _ = oracle.encode_pcm(pcm, len(L), oracle.config(2, bits), 44100)
        return oracle.decode(oracle.wrap_frames(out, 2, bits, 44100))["frames"]
    cls = {n: (L, R) for n, L, R in signals.stereo_classes(16)}
    fr = info(*cls["silence"], 16)
    assert all(s.type == 0 for f in fr for s in f.sub[:2])                      # CONSTANT
    fr = info(*cls["white_noise_full_scale"], 16)
    assert all(s.type == 1 for f in fr for s in f.sub[:2])                      # VERBATIM wins
    fr = info(*cls["wasted_bits"], 16)
    assert any(s.wasted > 0 for f in fr for s in f.sub[:2])
    fr = info(*cls["left_equals_right"], 16)
    assert all(f.ch_assign == 8 for f in fr)                                    # tie -> first minimum = L/S (Q7)
    fr = info(*cls["zero_runs"], 16)
    assert sum(s.n_escape for f in fr for s in f.sub[:2]) > 0                   # escape partitions (Q5)
    cls32 = {n: (L, R) for n, L, R in signals.stereo_classes(32)}
    fr = info(*cls32["full_scale_square"], 32)
    assert any(s.type == 1 for f in fr for s in f.sub[:2])                      # 33-bit side / range check -> VERBATIM


@pytest.mark.parametrize("bits,rate", [(16, 44100), (24, 96000), (32, 192000)])
def test_synthetic_stream_digest_and_threads(oracle, zf, bits, rate):
    n = 4096 * 5 + 1234
    pcm = zf.synth_pcm(n, rate, bits)
    g = DIGESTS[f"synth_{bits}_{rate}_{n}"]
    assert hashlib.sha256(pcm.tobytes()).hexdigest() == g["pcm_sha256"]         # generator is deterministic
    out, sizes = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate)
    assert hashlib.sha256(out.tobytes()).hexdigest() == g["flac_sha256"]
    assert [int(s) for s in sizes] == g["frame_sizes"]
    out8, sizes8 = oracle.encode_pcm(pcm, n, oracle.config(2, bits), rate, threads=4)
    assert out8.tobytes() == out.tobytes() and np.array_equal(sizes, sizes8)    # frame sharding is exact


def test_whole_file_golden_stream(oracle, zf):
    golden = open(os.path.join(GOLDEN, "synth16_6000.flac"), "rb").read()
    pcm = zf.synth_pcm(6000, 44100, 16)
    rc, flac = oracle.wav_to_flac(oracle.make_wav(pcm, 2, 16, 44100))
    assert rc == 0 and flac == golden
    assert flac[:4] == b"fLaC" and flac[42:73].hex() == KAT["vorbis_comment_hex"]
    d = oracle.decode(flac)
    assert d["rc"] == 0 and d["md5_ok"] == 1
    assert d["frames"][0].offset == KAT["first_frame_offset"]
    si = d["streaminfo"]
    assert (si.min_block, si.max_block, si.total_samples, si.channels, si.bits) == (4096, 4096, 6000, 2, 16)
    sizes = [f.size for f in d["frames"]]
    assert (si.min_frame, si.max_frame) == oracle.replay_frame_sizes(sizes)


def test_wav_parser_and_exit_codes(oracle, zf):
    pcm = zf.synth_pcm(100, 48000, 24)
    for ext in (False, True):
        rc, flac = oracle.wav_to_flac(oracle.make_wav(pcm, 2, 24, 48000, extensible=ext))
        assert rc == 0 and oracle.decode(flac)["rc"] == 0
    assert oracle.wav_to_flac(b"RIFX" + bytes(60))[0] == -1          # NotRiffFile
    assert oracle.wav_to_flac(b"RIFF\0\0\0\0WAVX" + bytes(60))[0] == -2  # NotWaveFile
    rc8, flac8 = oracle.wav_to_flac(oracle.make_wav(bytes(200), 2, 8, 8000))   # 8-bit: accepted (wav_reader.zig:71-78)
    assert rc8 == 0 and oracle.decode(flac8)["rc"] == 0 and oracle.decode(flac8)["streaminfo"].bits == 8
    wav12 = oracle.make_wav(bytes(400), 2, 16, 8000)
    wav12 = wav12[:34] + (12).to_bytes(2, "little") + wav12[36:]      # 12 valid bits in 2-byte containers
    assert oracle.wav_to_flac(wav12)[0] == 2                          # frame_writer.zig:221-233 has no code for it


def test_independent_channels_round_trip(oracle):
    rng = np.random.default_rng(3)
    for ch in (1, 3, 8):
        n = 4096 + 55
        planes = [rng.integers(-3000, 3000, n) if c % 2 else np.cumsum(rng.integers(-9, 10, n)) for c in range(ch)]
        inter = signals.interleave(planes)
        out, _ = oracle.encode_pcm(oracle.pcm_bytes_from_int(inter, 16), n, oracle.config(ch, 16), 48000)
        d = oracle.decode(oracle.wrap_frames(out, ch, 16, 48000))
        assert d["rc"] == 0 and np.array_equal(d["pcm"], inter)
        assert all(f.ch_assign == ch - 1 for f in d["frames"])

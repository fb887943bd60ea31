"""CPU tests of the product library's host side and of the C-ABI surface.  No compute entry point is
called: without a GPU they must fail loudly (there is no CPU fallback)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(zf):
    header = open(zf.HEADER_PATH).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(zf_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 25
    lib = C.CDLL(zf.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.zf_abi_version() == 1


def test_static_library_carries_the_same_abi(zf, tmp_path):
    """libzigflac_b200.a (BASELINE north_star: "nvcc-built sm_100a static library, linked from build.zig") defines every
    declared symbol, keeps the oracle out, and links into a program with nothing but the toolkit's static cudart --
    the link line INTEGRATION.md gives for build.zig."""
    import subprocess
    static = os.path.join(os.path.dirname(zf.LIB_PATH), "libzigflac_b200.a")
    assert os.path.exists(static), "run python zig-flac_b200/build.py"
    header = re.sub(r"/\*.*?\*/", "", open(zf.HEADER_PATH).read(), flags=re.S)
    names = set(re.findall(r"\b(zf_[a-z0-9_]+)\s*\(", header))
    syms = subprocess.run(["nm", "--defined-only", static], capture_output=True, text=True).stdout
    defined = set(re.findall(r" T (zf_[a-z0-9_]+)", syms))
    assert not (names - defined), sorted(names - defined)
    assert " zo_" not in syms and "fd_decode" not in syms
    spec_path = os.path.join(os.path.dirname(zf.LIB_PATH), "build.py")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_zf_build_t", spec_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    exe = mod.link_static_check(str(tmp_path / "flac_static"))
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 1 and b"usage: flac in_file.wav out_file.flac" in r.stderr  # cli.zig:17-20


def test_no_cpu_fallback(zf):
    """Constructing an encoder without an sm_100 device must raise, never silently encode on the CPU."""
    if zf.device_available():
        pytest.skip("a GPU is present")
    with pytest.raises(zf.FlacGpuError) as ei:
        zf.Encoder(zf.Config.default(2, 16), 44100)
    assert ei.value.status == zf.ZF_ERR_NO_DEVICE
    rc, out = zf.wav_to_flac(_wav(zf, 16))
    assert rc == zf.ZF_ERR_NO_DEVICE and out is None
    with pytest.raises(zf.FlacGpuError) as e:  # the decoder too
        zf.Decoder()
    assert e.value.status == zf.ZF_ERR_NO_DEVICE
    with pytest.raises(zf.FlacGpuError) as e:
        zf.decode_flac(b"fLaC" + bytes(100))
    assert e.value.status == zf.ZF_ERR_NO_DEVICE


def test_product_does_not_link_the_oracle(zf):
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", zf.LIB_PATH], capture_output=True, text=True).stdout
    assert "zo_" not in syms and "fd_decode" not in syms
    src_dir = os.path.join(os.path.dirname(zf.LIB_PATH), "csrc")
    for f in os.listdir(src_dir):
        if f.endswith((".cu", ".cuh", ".cpp", ".h", ".c")):
            text = open(os.path.join(src_dir, f)).read()
            assert "oracle/" not in text and "zigflac_oracle" not in text, f


def _wav(zf, bits):
    import oracle_lib
    return oracle_lib.make_wav(zf.synth_pcm(500, 44100, bits), 2, bits, 44100)


def test_md5_matches_hashlib(zf):
    rng = np.random.default_rng(0)
    for n in (0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 1000, 100003):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        m = zf.Md5()
        m.update(data[: n // 3])
        m.update(data[n // 3:])
        assert m.final() == hashlib.md5(data.tobytes()).digest(), n


def test_streaminfo_header_and_vendor_block_match_oracle(zf, oracle):
    import io
    si = zf.StreamInfo(44100, 2, 16, 2646000)
    si.c.max_frame_size = 1234
    assert si.bytes() == oracle.streaminfo_bytes(4096, 4096, 0xFFFFFF, 1234, 44100, 2, 16, 2646000)
    for sizes in ([100], [100, 50], [50, 100, 70, 20], [7000, 6500, 8000, 6400, 9000]):
        s2 = zf.StreamInfo(96000, 2, 24, 1000)
        for v in sizes:
            s2.update_frame_size(v)
        assert (s2.min_frame_size, s2.max_frame_size) == oracle.replay_frame_sizes(sizes)
    w = io.BytesIO()
    zf.Encoder.skip_header(w)
    zf.Encoder.write_vorbis_comment(w, True)
    assert w.getvalue() == bytes(42) + oracle.vorbis_comment(True)
    w = io.BytesIO()
    si.set_md5(bytes(range(16)))
    zf.Encoder.write_header(w, si, False)
    assert w.getvalue() == b"fLaC" + bytes([0, 0, 0, 34]) + oracle.streaminfo_bytes(
        4096, 4096, 0xFFFFFF, 1234, 44100, 2, 16, 2646000, bytes(range(16)))


def test_wav_reader_matches_oracle(zf, oracle):
    import struct
    pcm = zf.synth_pcm(300, 96000, 24)
    for ext in (False, True):
        wav = oracle.make_wav(pcm, 2, 24, 96000, extensible=ext)
        r = zf.WavReader(wav)
        fmt = oracle.ZoWavFmt()
        a = np.frombuffer(wav, dtype=np.uint8)
        assert oracle.lib().zo_wav_parse(a.ctypes.data, a.size, C.byref(fmt)) == 0
        assert (r.samples_count, r.sample_rate, r.bit_depth, r.channels, r.bytes_per_sample, r.data_offset, r.data_len) == (
            fmt.samples_count, fmt.sample_rate, fmt.bit_depth, fmt.channels, fmt.bytes_per_sample, fmt.data_offset, fmt.data_len)
        assert r.data().tobytes() == pcm.tobytes()
    # an extra chunk before "data" is skipped; errors map one to one
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 2, 44100, 176400, 4, 16) + b"LIST" + struct.pack("<I", 4) + b"abcd" + \
        b"data" + struct.pack("<I", 8) + bytes(8)
    r = zf.WavReader(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert r.samples_count == 2 and r.data_offset == 12 + 24 + 12 + 8
    for bad, status in ((b"RIFX" + bytes(40), -16), (b"RIFF\0\0\0\0WAVX" + bytes(40), -17)):
        with pytest.raises(zf.FlacGpuError) as ei:
            zf.WavReader(bad)
        assert ei.value.status == status


def test_synth_generator_properties(zf):
    a = zf.synth_pcm(10000, 96000, 24, threads=1)
    b = zf.synth_pcm(10000, 96000, 24, threads=5)
    assert a.tobytes() == b.tobytes()                                   # thread count does not matter
    c = zf.synth_pcm(4000, 96000, 24, first_sample=6000)
    assert c.tobytes() == a[6000 * 6:].tobytes()                          # random access by stream position
    v = a.reshape(-1, 3).astype(np.int32)
    s = v[:, 0] | (v[:, 1] << 8) | (v[:, 2] << 16)
    s = np.where(s >= 1 << 23, s - (1 << 24), s)
    assert np.abs(s).max() < 0.75 * (1 << 23) and np.abs(s).max() > 1000  # no clipping, not silent


def test_reference_arm_maps_no_product_library():
    """bench.py --impl reference times the oracle only: the product .so must not even be mapped into that process, and
    its `config` is the product arm's `config` (the driver compares the two lines)."""
    import json
    import subprocess
    import sys
    code = (
        "import runpy, sys\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '1', '--workload', 'c1_16bit_44k1_60s']\n"
        "try:\n"
        "    runpy.run_path('bench.py', run_name='__main__')\n"
        "except SystemExit:\n"
        "    pass\n"
        "print('MAPS', sorted({l.split()[-1] for l in open('/proc/self/maps') if '.so' in l and ('zig' in l or 'zf_' in l)}))\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, check=True)
    lines = r.stdout.strip().splitlines()
    maps = [l for l in lines if l.startswith("MAPS")][0]
    assert "libzigflac_b200" not in maps and "libzigflac_oracle.so" in maps, maps
    line = json.loads([l for l in lines if l.startswith("{")][0])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.make_config("c1_16bit_44k1_60s", 1)
    assert line["config"]["l2_resident"] is True


def test_wav8_reader_matches_the_oracle_restatement(zf, oracle):
    """One-byte WAV samples (wav_reader.zig:56-90): the reference subtracts 128 from the unshifted plane word, so every
    sample depends on what the plane held one frame earlier.  The product's closed form (zf_wav8_to_samples) against
    the oracle, which restates the reference's word arithmetic on planes it keeps from frame to frame: the oracle's
    stream must decode (independent decoder) to exactly the product reader's samples."""
    rng = np.random.default_rng(8)
    for channels in (1, 2, 3):
        n = 3 * 4096 + 1234
        t = np.arange(n)
        planes = []
        for c in range(channels):
            if c == 0:
                v = np.clip(128 + 90 * np.sin(t * 0.013) + rng.integers(-4, 5, n), 0, 255)
            elif c == 1:
                v = rng.integers(0, 256, n)                       # every byte value, the -128 / borrow wrap included
            else:
                v = np.where((t // 700) % 2 == 0, 0, 128)        # runs of 0x00 (borrow chains) and of 0x80
            planes.append(v.astype(np.uint8))
        raw = np.stack(planes, axis=1).reshape(-1).copy()
        wav = oracle.make_wav(raw.tobytes(), channels, 8, 22050)
        rc, flac = oracle.wav_to_flac(wav)
        assert rc == 0
        d = oracle.decode(flac)
        assert d["rc"] == 0 and d["streaminfo"].bits == 8
        rd = zf.Wav8Reader(channels)
        got = np.concatenate([rd.convert(raw[:5000 * channels]), rd.convert(raw[5000 * channels:])])  # state carries over
        assert np.array_equal(got.astype(np.int32), d["pcm"])

"""ctypes bindings for the CPU oracle and the independent decoder (oracle/_ref/libzigflac_oracle.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_ref", "libzigflac_oracle.so")


class ZoConfig(C.Structure):
    _fields_ = [("block_size", C.c_uint16), ("bit_depth", C.c_uint8), ("channels", C.c_uint8),
                ("stereo_decorrelation", C.c_uint8), ("max_rice_order", C.c_uint8),
                ("max_rice_param", C.c_uint8), ("lpc_order", C.c_uint8), ("exact_rice", C.c_uint8),
                ("reserved", C.c_uint8)]


class ZoFrameInfo(C.Structure):
    _fields_ = [("bit_depth", C.c_uint8), ("channels", C.c_uint8), ("samples_count", C.c_uint16),
                ("sample_rate", C.c_uint32)]


class ZoStreamInfo(C.Structure):
    _fields_ = [("md5", C.c_uint8 * 16), ("interchannel_samples", C.c_uint64),
                ("min_frame_size", C.c_uint32), ("max_frame_size", C.c_uint32),
                ("sample_rate", C.c_uint32), ("min_block_size", C.c_uint16),
                ("max_block_size", C.c_uint16), ("channels", C.c_uint8), ("bit_depth", C.c_uint8)]


class ZoWavFmt(C.Structure):
    _fields_ = [("samples_count", C.c_uint32), ("sample_rate", C.c_uint32), ("bit_depth", C.c_uint16),
                ("channels", C.c_uint16), ("bytes_per_sample", C.c_uint8), ("data_offset", C.c_size_t),
                ("data_len", C.c_uint32)]


class ZoMd5(C.Structure):
    _fields_ = [("s", C.c_uint32 * 4), ("n", C.c_uint64), ("buf", C.c_uint8 * 64)]


class FdStreamInfo(C.Structure):
    _fields_ = [("min_block", C.c_uint32), ("max_block", C.c_uint32), ("min_frame", C.c_uint32),
                ("max_frame", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits", C.c_uint32), ("total_samples", C.c_uint64), ("md5", C.c_uint8 * 16),
                ("first_frame_offset", C.c_size_t)]


class FdSubframeInfo(C.Structure):
    _fields_ = [("type", C.c_uint8), ("order", C.c_uint8), ("wasted", C.c_uint8), ("rice_method", C.c_uint8),
                ("part_order", C.c_uint8), ("n_escape", C.c_uint8), ("reserved", C.c_uint16),
                ("bits", C.c_uint32)]


class FdFrameInfo(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("size", C.c_uint32), ("block_size", C.c_uint32),
                ("sample_rate", C.c_uint32), ("number", C.c_uint64), ("ch_assign", C.c_uint8),
                ("bits", C.c_uint8), ("n_sub", C.c_uint8), ("pad", C.c_uint8), ("sub", FdSubframeInfo * 8)]


_lib = None


def build(force=False):
    """Compile the oracle with its committed Makefile (outputs only into oracle/_ref/)."""
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("zigflac_oracle.c", "flac_decode.c", "zigflac_oracle.h",
                                                  "oracle_cli.c", "Makefile", "zigflac_lpc.h")]
    srcs.append(os.path.join(ROOT, "zig-flac_b200", "csrc", "zf_synth.c"))
    stale = force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return ORACLE_SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(ORACLE_SO)
    u8p = C.POINTER(C.c_uint8)
    L.zo_config_default.argtypes = [C.POINTER(ZoConfig), C.c_uint8, C.c_uint8]
    L.zo_max_frame_bytes.restype = C.c_size_t
    L.zo_max_frame_bytes.argtypes = [C.c_uint16, C.c_uint8, C.c_uint8, C.c_int]
    L.zo_encode_pcm.restype = C.c_size_t
    L.zo_encode_pcm.argtypes = [C.POINTER(ZoConfig), C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                C.c_size_t, C.c_void_p, C.POINTER(C.c_uint32), C.c_int]
    L.zo_streaminfo_init.argtypes = [C.POINTER(ZoStreamInfo)]
    L.zo_streaminfo_update_frame_size.argtypes = [C.POINTER(ZoStreamInfo), C.c_uint32]
    L.zo_streaminfo_bytes.argtypes = [C.POINTER(ZoStreamInfo), u8p]
    L.zo_write_stream_header.restype = C.c_size_t
    L.zo_write_stream_header.argtypes = [C.POINTER(ZoStreamInfo), C.c_int, u8p]
    L.zo_write_vorbis_comment.restype = C.c_size_t
    L.zo_write_vorbis_comment.argtypes = [C.c_int, u8p]
    L.zo_wav_parse.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(ZoWavFmt)]
    L.zo_wav_to_flac.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int]
    L.zo_free.argtypes = [C.c_void_p]
    L.zo_crc8.restype = C.c_uint8
    L.zo_crc8.argtypes = [C.c_void_p, C.c_size_t]
    L.zo_crc16.restype = C.c_uint16
    L.zo_crc16.argtypes = [C.c_uint16, C.c_void_p, C.c_size_t]
    L.zo_crc16_clmul.restype = C.c_uint16
    L.zo_crc16_clmul.argtypes = [C.c_uint16, C.c_void_p, C.c_size_t]
    L.zo_md5_init.argtypes = [C.POINTER(ZoMd5)]
    L.zo_md5_update.argtypes = [C.POINTER(ZoMd5), C.c_void_p, C.c_size_t]
    L.zo_md5_final.argtypes = [C.POINTER(ZoMd5), u8p]
    L.zo_flac_calc_part_size.restype = C.c_uint64
    L.zo_flac_calc_part_size.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.zo_frame_header.restype = C.c_size_t
    L.zo_frame_header.argtypes = [C.c_uint64, C.c_uint8, C.c_uint8, C.c_uint16, C.c_uint32, u8p]
    L.fd_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(FdStreamInfo), C.POINTER(C.c_void_p),
                            C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                            C.POINTER(C.c_int)]
    L.fd_free.argtypes = [C.c_void_p]
    _lib = L
    return L


_synth = None


def synth_pcm(samples, sample_rate, bit_depth, first_sample=0, seed=0x5EED, threads=None):
    """The benchmark's synthetic stereo PCM from oracle/_ref/libzf_synth.so -- the same generator source as
    the product's zf_synth_pcm, built on its own so that the reference arm of bench.py loads no product code."""
    global _synth
    if _synth is None:
        build()
        so = os.path.join(ORACLE_DIR, "_ref", "libzf_synth.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
        _synth = C.CDLL(so)
        _synth.zf_synth_pcm.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
    out = np.empty(samples * 2 * (bit_depth // 8), dtype=np.uint8)
    rc = _synth.zf_synth_pcm(out.ctypes.data, first_sample, samples, sample_rate, bit_depth, seed,
                             threads or (os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("synth_pcm: unsupported format")
    return out


def _buf(b):
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b
    return np.ascontiguousarray(a)


def config(channels=2, bit_depth=16, block_size=4096, stereo_decorrelation=1, max_rice_order=8, max_rice_param=30,
           lpc_order=0, exact_rice=0):
    cfg = ZoConfig()
    lib().zo_config_default(C.byref(cfg), channels, bit_depth)
    cfg.block_size = block_size
    cfg.stereo_decorrelation = stereo_decorrelation
    cfg.max_rice_order = max_rice_order
    cfg.max_rice_param = max_rice_param
    cfg.lpc_order = lpc_order  # 0 = the reference's path; > 0 = LPC extension (oracle/zigflac_lpc.h)
    cfg.exact_rice = exact_rice  # 0 = the reference's estimate; 1 = extension: exact Rice code lengths
    return cfg


def encode_pcm(pcm, samples_per_channel, cfg, sample_rate=44100, first_frame_number=0, threads=1):
    """K x Encoder.writeFrame over raw interleaved PCM bytes -> (frame bytes, frame sizes)."""
    L = lib()
    pcm = _buf(pcm)
    frames = (samples_per_channel + cfg.block_size - 1) // cfg.block_size
    cap = frames * (L.zo_max_frame_bytes(cfg.block_size, cfg.bit_depth, cfg.channels, 1) + 64) + 64
    out = np.empty(cap, dtype=np.uint8)
    sizes = np.zeros(max(frames, 1), dtype=np.uint32)
    n = C.c_uint32(0)
    total = L.zo_encode_pcm(C.byref(cfg), sample_rate, pcm.ctypes.data, samples_per_channel, first_frame_number,
                            out.ctypes.data, cap, sizes.ctypes.data, C.byref(n), threads)
    if total == C.c_size_t(-1).value:
        raise RuntimeError("oracle encode failed")
    return out[:total].copy(), sizes[:n.value].copy()


def wav_to_flac(wav_bytes, threads=1):
    L = lib()
    wav = _buf(wav_bytes)
    p = C.c_void_p()
    n = C.c_size_t()
    rc = L.zo_wav_to_flac(wav.ctypes.data, wav.size, C.byref(p), C.byref(n), threads)
    if rc != 0:
        return rc, None
    out = bytes((C.c_uint8 * n.value).from_address(p.value))
    L.zo_free(p)
    return 0, out


def frame_header(frame_number, bit_depth, ch_type, block_size, sample_rate):
    out = (C.c_uint8 * 16)()
    n = lib().zo_frame_header(frame_number, bit_depth, ch_type, block_size, sample_rate, out)
    return bytes(out[:n])


def crc8(b):
    a = _buf(b)
    return lib().zo_crc8(a.ctypes.data, a.size)


def crc16(b, crc=0, clmul=False):
    a = _buf(b)
    f = lib().zo_crc16_clmul if clmul else lib().zo_crc16
    return f(crc, a.ctypes.data, a.size)


def md5(b):
    a = _buf(b)
    m = ZoMd5()
    L = lib()
    L.zo_md5_init(C.byref(m))
    half = a.size // 3  # exercise the buffering path
    L.zo_md5_update(C.byref(m), a.ctypes.data, half)
    L.zo_md5_update(C.byref(m), a.ctypes.data + half, a.size - half)
    out = (C.c_uint8 * 16)()
    L.zo_md5_final(C.byref(m), out)
    return bytes(out)


def streaminfo_bytes(min_block, max_block, min_frame, max_frame, sample_rate, channels, bit_depth, samples, md5_digest=bytes(16)):
    si = ZoStreamInfo()
    lib().zo_streaminfo_init(C.byref(si))
    si.min_block_size, si.max_block_size = min_block, max_block
    si.min_frame_size, si.max_frame_size = min_frame, max_frame
    si.sample_rate, si.channels, si.bit_depth = sample_rate, channels, bit_depth
    si.interchannel_samples = samples
    for i, v in enumerate(md5_digest):
        si.md5[i] = v
    out = (C.c_uint8 * 34)()
    lib().zo_streaminfo_bytes(C.byref(si), out)
    return bytes(out)


def replay_frame_sizes(sizes):
    """metadata.zig:35-40 replayed sequentially (order dependent) -> (min, max)."""
    si = ZoStreamInfo()
    L = lib()
    L.zo_streaminfo_init(C.byref(si))
    for s in sizes:
        L.zo_streaminfo_update_frame_size(C.byref(si), int(s))
    return si.min_frame_size, si.max_frame_size


def vorbis_comment(last=True):
    out = (C.c_uint8 * 31)()
    n = lib().zo_write_vorbis_comment(1 if last else 0, out)
    return bytes(out[:n])


def decode(flac_bytes, max_frames=1 << 20):
    """Independent decoder -> dict(rc, streaminfo, pcm[int32 interleaved], frames[list], md5_ok)."""
    L = lib()
    a = _buf(flac_bytes)
    si = FdStreamInfo()
    pcm = C.c_void_p()
    ns = C.c_uint64()
    nf = C.c_size_t()
    ok = C.c_int(-2)
    frames = (FdFrameInfo * max_frames)()
    rc = L.fd_decode(a.ctypes.data, a.size, C.byref(si), C.byref(pcm), C.byref(ns), frames, max_frames,
                     C.byref(nf), C.byref(ok))
    res = {"rc": rc, "streaminfo": si, "md5_ok": ok.value, "pcm": None, "frames": []}
    if rc == 0:
        n = ns.value * si.channels
        res["pcm"] = np.ctypeslib.as_array((C.c_int32 * n).from_address(pcm.value)).copy() if n else np.zeros(0, np.int32)
        res["frames"] = [frames[i] for i in range(min(nf.value, max_frames))]
        res["n_frames"] = nf.value
    if pcm.value:
        L.fd_free(pcm)
    return res


def wrap_frames(frames_bytes, channels, bit_depth, sample_rate, block_size=4096, total_samples=0):
    """Prefix a bare frame sequence with 'fLaC' + STREAMINFO (unknown MD5) so decode() accepts it."""
    si = streaminfo_bytes(block_size, block_size, 0, 0, sample_rate, channels, bit_depth, total_samples)
    return b"fLaC" + bytes([0x80, 0, 0, 34]) + si + bytes(frames_bytes)


def pcm_bytes_from_int(samples, bit_depth):
    """int array (interleaved) -> packed little-endian PCM bytes at bit_depth/8 bytes per sample."""
    s = np.asarray(samples, dtype=np.int64)
    nb = bit_depth // 8
    out = np.empty((s.size, nb), dtype=np.uint8)
    u = s.astype(np.uint64)
    for k in range(nb):
        out[:, k] = ((u >> np.uint64(8 * k)) & np.uint64(0xFF)).astype(np.uint8)
    return out.reshape(-1)


def make_wav(pcm, channels, bit_depth, sample_rate, extensible=False):
    pcm = bytes(pcm)
    nb = bit_depth // 8
    block_align = channels * nb
    byte_rate = sample_rate * block_align
    import struct
    if extensible:
        fmt = struct.pack("<HHIIHHHHI16s", 0xFFFE, channels, sample_rate, byte_rate, block_align, bit_depth, 22,
                          bit_depth, 3, bytes([1, 0, 0, 0, 0, 0, 0x10, 0, 0x80, 0, 0, 0xAA, 0, 0x38, 0x9B, 0x71]))
    else:
        fmt = struct.pack("<HHIIHH", 1, channels, sample_rate, byte_rate, block_align, bit_depth)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(pcm)) + pcm
    return b"RIFF" + struct.pack("<I", len(body)) + body

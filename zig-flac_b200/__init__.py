"""zig-flac_b200 -- Python view of the B200-native FLAC encode engine (libzigflac_b200.so).

The product is the C-ABI shared library declared in include/zigflac_b200.h; this module is a thin
ctypes binding that mirrors the reference's Zig module `flac` (src/lib.zig:1-7) for the encode
path, so tests read like the reference's caller (src/cli/wav2flac.zig):

    enc = Encoder(Config.default(channels, bit_depth), sample_rate)   # Encoder.init   encoder.zig:44
    enc.skip_header(w); enc.write_vorbis_comment(w, last=True)        # :177, :211
    size = enc.write_frame(w, frame_idx, planes)                      # writeFrame     :234
    frames, sizes = enc.encode_pcm(raw_bytes, n_samples)              # K x writeFrame (batched)
    enc.write_header(w, streaminfo, last=False)                       # :192

There is no CPU fallback: constructing an Encoder without a usable sm_100 GPU raises FlacGpuError.
Because the directory name carries a hyphen, import it with importlib.import_module("zig-flac_b200")
or through the `zigflac_b200` alias module at the repository root.
"""
import ctypes as C
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzigflac_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "zigflac_b200.h")

ZF_OK = 0
ZF_ERR_NO_DEVICE = -3
ZF_ERR_OUT_TOO_SMALL = -6
ZF_ERR_BUSY = -7


class FlacGpuError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = _lib().zf_strerror(status).decode() if _LIB is not None else str(status)
        cuda = _lib().zf_last_cuda_error().decode() if _LIB is not None else ""
        super().__init__(f"{what}: {msg} ({status}) {cuda}".strip())


class ZfConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("block_size", C.c_uint16), ("bit_depth", C.c_uint8),
                ("channels", C.c_uint8), ("sample_rate", C.c_uint32), ("stereo_decorrelation", C.c_uint8),
                ("max_rice_order", C.c_uint8), ("max_rice_param", C.c_uint8), ("lpc_order", C.c_uint8),
                ("device_id", C.c_int32), ("max_frames_per_batch", C.c_uint32), ("exact_rice", C.c_uint8),
                ("reserved1", C.c_uint8 * 3)]


class ZfStreamInfo(C.Structure):
    _fields_ = [("md5", C.c_uint8 * 16), ("interchannel_samples", C.c_uint64), ("min_frame_size", C.c_uint32),
                ("max_frame_size", C.c_uint32), ("sample_rate", C.c_uint32), ("min_block_size", C.c_uint16),
                ("max_block_size", C.c_uint16), ("channels", C.c_uint8), ("bit_depth", C.c_uint8)]


class ZfMd5(C.Structure):
    _fields_ = [("state", C.c_uint32 * 4), ("length", C.c_uint64), ("buffer", C.c_uint8 * 64)]


class ZfWavFormat(C.Structure):
    _fields_ = [("samples_count", C.c_uint32), ("sample_rate", C.c_uint32), ("bit_depth", C.c_uint16),
                ("channels", C.c_uint16), ("bytes_per_sample", C.c_uint8), ("reserved", C.c_uint8 * 3),
                ("data_offset", C.c_uint64), ("data_len", C.c_uint32)]


_LIB = None


class ZfDecodeInfo(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bit_depth", C.c_uint32),
                ("min_block_size", C.c_uint32), ("max_block_size", C.c_uint32), ("launches", C.c_uint32),
                ("kernel_ms", C.c_float), ("samples_per_channel", C.c_uint64), ("streaminfo_samples", C.c_uint64),
                ("pcm_bytes", C.c_uint64), ("n_frames", C.c_uint64), ("bad_frame", C.c_uint64),
                ("bad_frame_status", C.c_uint32), ("md5_status", C.c_int32), ("md5", C.c_uint8 * 16)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("md5", "struct_size")}
        d["md5"] = bytes(self.md5)
        return d


def build(force=False):
    """Compile the shared library in-tree with nvcc (sm_100a)."""
    spec = importlib.util.spec_from_file_location("_zf_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force)


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python zig-flac_b200/build.py` (needs nvcc); "
                                "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u8p = C.c_void_p, C.POINTER(C.c_uint8)
    L.zf_abi_version.restype = C.c_int
    L.zf_strerror.restype = C.c_char_p
    L.zf_strerror.argtypes = [C.c_int]
    L.zf_last_cuda_error.restype = C.c_char_p
    L.zf_device_check.argtypes = [C.c_int]
    L.zf_config_default.argtypes = [C.POINTER(ZfConfig), C.c_uint8, C.c_uint8, C.c_uint32]
    L.zf_max_frame_bytes.restype = C.c_size_t
    L.zf_max_frame_bytes.argtypes = [C.POINTER(ZfConfig)]
    L.zf_max_batch_bytes.restype = C.c_size_t
    L.zf_max_batch_bytes.argtypes = [C.POINTER(ZfConfig), C.c_uint32]
    L.zf_encoder_create.argtypes = [C.POINTER(ZfConfig), C.POINTER(vp)]
    L.zf_encoder_destroy.argtypes = [vp]
    L.zf_encoder_destroy.restype = None
    L.zf_encode_pcm.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_size_t, C.POINTER(C.c_size_t), vp, C.c_uint32,
                                C.POINTER(C.c_uint32)]
    L.zf_encode_submit.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
    L.zf_encode_collect.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t), vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.zf_encode_device.argtypes = [vp, vp, C.c_uint64, C.c_uint64, vp, C.c_size_t, vp, vp, vp]
    L.zf_encode_device_status.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.zf_write_frame.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint64, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.zf_last_batch_stats.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    L.zf_kernel_times.argtypes = [vp, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint32)]
    L.zf_streaminfo_init.argtypes = [C.POINTER(ZfStreamInfo)]
    L.zf_streaminfo_init.restype = None
    L.zf_streaminfo_update_frame_size.argtypes = [C.POINTER(ZfStreamInfo), C.c_uint32]
    L.zf_streaminfo_update_frame_size.restype = None
    L.zf_streaminfo_bytes.argtypes = [C.POINTER(ZfStreamInfo), u8p]
    L.zf_streaminfo_bytes.restype = None
    L.zf_write_stream_header.restype = C.c_size_t
    L.zf_write_stream_header.argtypes = [C.POINTER(ZfStreamInfo), C.c_int, u8p]
    L.zf_write_vorbis_comment.restype = C.c_size_t
    L.zf_write_vorbis_comment.argtypes = [C.c_int, u8p]
    L.zf_md5_init.argtypes = [C.POINTER(ZfMd5)]
    L.zf_md5_init.restype = None
    L.zf_md5_update.argtypes = [C.POINTER(ZfMd5), vp, C.c_size_t]
    L.zf_md5_update.restype = None
    L.zf_md5_final.argtypes = [C.POINTER(ZfMd5), u8p]
    L.zf_md5_final.restype = None
    L.zf_md5_openssl_available.restype = C.c_int
    L.zf_wav8_state_init.argtypes = [vp, C.c_size_t]
    L.zf_wav8_state_init.restype = None
    L.zf_wav8_to_samples.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp]
    L.zf_wav8_to_samples.restype = None
    L.zf_wav_parse.argtypes = [vp, C.c_size_t, C.POINTER(ZfWavFormat)]
    L.zf_encode_wav_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_int]
    L.zf_encode_wav_memory.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.c_int]
    L.zf_free.argtypes = [vp]
    L.zf_driver_release_cache.restype = None
    L.zf_free.restype = None
    L.zf_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.zf_host_alloc.restype = C.c_int
    L.zf_host_free.argtypes = [vp]
    L.zf_host_free.restype = None
    L.zf_synth_pcm.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
    L.zf_decoder_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.zf_decoder_destroy.argtypes = [vp]
    L.zf_decoder_destroy.restype = None
    L.zf_flac_stream_info.argtypes = [vp, C.c_size_t, C.POINTER(ZfDecodeInfo)]
    L.zf_decode_flac.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_uint32, C.POINTER(ZfDecodeInfo)]
    L.zf_decode_flac_device.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(ZfDecodeInfo)]
    L.zf_decode_flac_memory.argtypes = [vp, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_size_t),
                                        C.POINTER(ZfDecodeInfo)]
    L.zf_decode_flac_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint32]
    L.zf_verify_flac_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    _LIB = L
    return L


def lib():
    return _lib()


def _u8(buf):
    a = buf if isinstance(buf, np.ndarray) else np.frombuffer(buf, dtype=np.uint8)
    if a.dtype != np.uint8:
        a = a.view(np.uint8)
    return np.ascontiguousarray(a)


class Config:
    """Encoder.Config (encoder.zig:609-656); `default` is Config.default(channels, bit_depth) (:642-655)."""

    def __init__(self, channels, bit_depth, block_size=4096, stereo_decorrelation=True, max_rice_order=8,
                 max_rice_param=30, lpc_order=0, exact_rice=False):
        self.channels = channels
        self.bit_depth = bit_depth
        self.block_size = block_size
        self.stereo_decorrelation = stereo_decorrelation
        self.max_rice_order = max_rice_order
        self.max_rice_param = max_rice_param
        self.lpc_order = lpc_order  # 0 = the reference's encoder; 1..12 = LPC extension (no reference counterpart)
        self.exact_rice = exact_rice  # extension: exact Rice code lengths instead of the reference's estimate

    @staticmethod
    def default(channels, bit_depth):
        return Config(channels, bit_depth)

    def to_c(self, sample_rate, device_id=0, max_frames_per_batch=2048):
        c = ZfConfig()
        _lib().zf_config_default(C.byref(c), self.channels, self.bit_depth, sample_rate)
        c.block_size = self.block_size
        c.stereo_decorrelation = 1 if self.stereo_decorrelation else 0
        c.max_rice_order = self.max_rice_order
        c.max_rice_param = self.max_rice_param
        c.lpc_order = self.lpc_order
        c.exact_rice = 1 if self.exact_rice else 0
        c.device_id = device_id
        c.max_frames_per_batch = max_frames_per_batch
        return c


class StreamInfo:
    """metadata.StreamInfo (metadata.zig:22-68)."""

    def __init__(self, sample_rate, channels, bit_depth, interchannel_samples, min_block_size=4096, max_block_size=4096):
        self.c = ZfStreamInfo()
        _lib().zf_streaminfo_init(C.byref(self.c))
        self.c.sample_rate, self.c.channels, self.c.bit_depth = sample_rate, channels, bit_depth
        self.c.interchannel_samples = interchannel_samples
        self.c.min_block_size, self.c.max_block_size = min_block_size, max_block_size

    def update_frame_size(self, frame_size):
        _lib().zf_streaminfo_update_frame_size(C.byref(self.c), int(frame_size))

    def set_md5(self, digest):
        for i, v in enumerate(bytes(digest)):
            self.c.md5[i] = v

    def bytes(self):
        out = (C.c_uint8 * 34)()
        _lib().zf_streaminfo_bytes(C.byref(self.c), out)
        return bytes(out)

    @property
    def min_frame_size(self):
        return self.c.min_frame_size

    @property
    def max_frame_size(self):
        return self.c.max_frame_size


class Md5:
    """Md5 (md5.zig:31)."""

    def __init__(self):
        self.c = ZfMd5()
        _lib().zf_md5_init(C.byref(self.c))

    def update(self, data):
        a = _u8(data)
        _lib().zf_md5_update(C.byref(self.c), a.ctypes.data, a.size)

    def final(self):
        out = (C.c_uint8 * 16)()
        _lib().zf_md5_final(C.byref(self.c), out)
        return bytes(out)


class WavReader:
    """WavReader.init/getFmt + fillSamples source bytes (wav_reader.zig:26-32,116-170)."""

    def __init__(self, file_bytes):
        self.file = _u8(file_bytes)
        fmt = ZfWavFormat()
        rc = _lib().zf_wav_parse(self.file.ctypes.data, self.file.size, C.byref(fmt))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "WavReader")
        self.samples_count, self.sample_rate = fmt.samples_count, fmt.sample_rate
        self.bit_depth, self.channels, self.bytes_per_sample = fmt.bit_depth, fmt.channels, fmt.bytes_per_sample
        self.data_offset, self.data_len = fmt.data_offset, fmt.data_len

    def data(self):
        n = self.samples_count * self.channels * self.bytes_per_sample
        return self.file[self.data_offset:self.data_offset + n]


class Encoder:
    """Encoder (encoder.zig) on the GPU.  `writer` arguments are any object with .write(bytes)."""

    def __init__(self, config, sample_rate, device_id=0, max_frames_per_batch=2048):
        self.config = config
        self.sample_rate = sample_rate
        self.ccfg = config.to_c(sample_rate, device_id, max_frames_per_batch)
        self.handle = C.c_void_p()
        rc = _lib().zf_encoder_create(C.byref(self.ccfg), C.byref(self.handle))
        if rc != ZF_OK:
            self.handle = None
            raise FlacGpuError(rc, "Encoder.init")
        self.md5 = Md5()

    def close(self):
        if getattr(self, "handle", None):
            _lib().zf_encoder_destroy(self.handle)
            self.handle = None

    deinit = close

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # --- stream metadata (host side) ---
    @staticmethod
    def skip_header(writer):
        writer.write(bytes(42))  # encoder.zig:177-185

    @staticmethod
    def write_header(writer, streaminfo, last_metadata=False):
        out = (C.c_uint8 * 42)()
        _lib().zf_write_stream_header(C.byref(streaminfo.c), 1 if last_metadata else 0, out)
        writer.write(bytes(out))

    @staticmethod
    def write_vorbis_comment(writer, last_metadata=True):
        out = (C.c_uint8 * 31)()
        n = _lib().zf_write_vorbis_comment(1 if last_metadata else 0, out)
        writer.write(bytes(out[:n]))

    def finalize_streaminfo_md5(self, streaminfo):
        streaminfo.set_md5(self.md5.final())

    # --- frames ---
    def max_batch_bytes(self, n_frames):
        return _lib().zf_max_batch_bytes(C.byref(self.ccfg), n_frames)

    def write_frame(self, writer, frame_number, planes):
        """Encoder.writeFrame (encoder.zig:234): planes = per-channel int32 arrays; returns the byte count."""
        planes = [np.ascontiguousarray(p, dtype=np.int32) for p in planes]
        n = planes[0].size
        ptrs = (C.c_void_p * len(planes))(*[p.ctypes.data for p in planes])
        cap = self.max_batch_bytes(1)
        out = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t()
        rc = _lib().zf_write_frame(self.handle, ptrs, n, frame_number, out.ctypes.data, cap, C.byref(ln))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "writeFrame")
        if writer is not None:
            writer.write(out[:ln.value].tobytes())
        return ln.value

    def encode_pcm(self, pcm, samples_per_channel, first_frame_number=0, out=None):
        """K x writeFrame over raw interleaved little-endian PCM -> (frame bytes, frame sizes)."""
        pcm = _u8(pcm)
        bs = self.config.block_size
        frames = (samples_per_channel + bs - 1) // bs
        cap = self.max_batch_bytes(frames)
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        sizes = np.zeros(max(frames, 1), dtype=np.uint32)
        ln = C.c_size_t()
        nf = C.c_uint32()
        rc = _lib().zf_encode_pcm(self.handle, pcm.ctypes.data, samples_per_channel, first_frame_number,
                                  out.ctypes.data, out.size, C.byref(ln), sizes.ctypes.data, sizes.size, C.byref(nf))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "encode_pcm")
        return out[:ln.value], sizes[:nf.value]

    def encode_device(self, d_pcm_ptr, samples_per_channel, first_frame_number, d_out_ptr, out_cap, d_sizes_ptr,
                      d_total_ptr, stream_ptr=None):
        rc = _lib().zf_encode_device(self.handle, d_pcm_ptr, samples_per_channel, first_frame_number, d_out_ptr,
                                     out_cap, d_sizes_ptr, d_total_ptr, stream_ptr)
        if rc != ZF_OK:
            raise FlacGpuError(rc, "encode_device")

    def device_status(self):
        """Status flags of the most recent encode_device batch (waits for it)."""
        f = C.c_uint32()
        rc = _lib().zf_encode_device_status(self.handle, C.byref(f))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "encode_device_status")
        return f.value

    def submit(self, pcm, samples_per_channel, first_frame_number=0):
        """zf_encode_submit: upload + kernels of one batch, asynchronous; `pcm` may be reused on return."""
        pcm = _u8(pcm)
        rc = _lib().zf_encode_submit(self.handle, pcm.ctypes.data, samples_per_channel, first_frame_number)
        if rc != ZF_OK:
            raise FlacGpuError(rc, "encode_submit")
        self._submitted = samples_per_channel

    def collect(self, out=None):
        """zf_encode_collect: wait for the submitted batch -> (frame bytes, frame sizes)."""
        bs = self.config.block_size
        frames = (self._submitted + bs - 1) // bs
        if out is None:
            out = np.empty(self.max_batch_bytes(frames), dtype=np.uint8)
        sizes = np.zeros(max(frames, 1), dtype=np.uint32)
        ln = C.c_size_t()
        nf = C.c_uint32()
        rc = _lib().zf_encode_collect(self.handle, out.ctypes.data, out.size, C.byref(ln), sizes.ctypes.data, sizes.size,
                                      C.byref(nf))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "encode_collect")
        return out[:ln.value], sizes[:nf.value]

    def last_batch_stats(self):
        ms = C.c_float()
        n = C.c_uint32()
        _lib().zf_last_batch_stats(self.handle, C.byref(ms), C.byref(n))
        return ms.value, n.value


def _kernel_times(self, cap=512):
    """Durations (ms) of the full-frame kernel launches since the previous call."""
    buf = (C.c_float * cap)()
    n = C.c_uint32()
    rc = _lib().zf_kernel_times(self.handle, buf, cap, C.byref(n))
    if rc != ZF_OK:
        raise FlacGpuError(rc, "kernel_times")
    return [buf[i] for i in range(n.value)]


Encoder.kernel_times = _kernel_times


def wav_to_flac(wav_bytes, devices=None):
    """cli.zig + wav2flac.zig main on in-memory files.  Returns (status, flac bytes or None)."""
    a = _u8(wav_bytes)
    p = C.c_void_p()
    n = C.c_size_t()
    devs = (C.c_int * len(devices))(*devices) if devices else None
    rc = _lib().zf_encode_wav_memory(a.ctypes.data, a.size, C.byref(p), C.byref(n), devs, len(devices) if devices else 0)
    if rc != ZF_OK:
        return rc, None
    out = bytes((C.c_uint8 * n.value).from_address(p.value))
    _lib().zf_free(p)
    return ZF_OK, out


class FlacBuffer:
    """The malloc'd result of zf_encode_wav_memory as a numpy view (no copy); release with close()."""

    def __init__(self, ptr, nbytes):
        self._p = ptr
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(ptr.value))[:nbytes]

    def close(self):
        if self._p is not None:
            self.array = None
            _lib().zf_free(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def wav_to_flac_view(wav_bytes, devices=None):
    """As wav_to_flac, without copying the result into a Python bytes object: (status, FlacBuffer or None)."""
    a = _u8(wav_bytes)
    p = C.c_void_p()
    n = C.c_size_t()
    devs = (C.c_int * len(devices))(*devices) if devices else None
    rc = _lib().zf_encode_wav_memory(a.ctypes.data, a.size, C.byref(p), C.byref(n), devs, len(devices) if devices else 0)
    if rc != ZF_OK:
        return rc, None
    return ZF_OK, FlacBuffer(p, n.value)


class HostBuffer:
    """Page-locked host memory from zf_host_alloc as a numpy uint8 array (`.array`); release with close()."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        rc = _lib().zf_host_alloc(nbytes, C.byref(p))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "zf_host_alloc")
        self._p = p
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(p.value))[:nbytes]

    def close(self):
        if self._p is not None:
            self.array = None
            _lib().zf_host_free(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def encode_file(in_path, out_path, devices=None):
    devs = (C.c_int * len(devices))(*devices) if devices else None
    return _lib().zf_encode_wav_file(os.fsencode(in_path), os.fsencode(out_path), devs, len(devices) if devices else 0)


def synth_pcm(samples, sample_rate, bit_depth, first_sample=0, seed=0x5EED, threads=None, out=None):
    """Deterministic synthetic stereo PCM as interleaved little-endian bytes."""
    n = samples * 2 * (bit_depth // 8)
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    rc = _lib().zf_synth_pcm(out.ctypes.data, first_sample, samples, sample_rate, bit_depth, seed,
                             threads or (os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("synth_pcm: unsupported format")
    return out


class Wav8Reader:
    """WavReader.fillSamples for one-byte containers (wav_reader.zig:56-90): raw WAV bytes -> the signed samples the
    reference's reader produces (zf_wav8_to_samples), with the state it carries from frame to frame."""

    def __init__(self, channels, block_size=4096):
        self.channels, self.block_size, self.pos = channels, block_size, 0
        self.state = np.empty(block_size * channels, dtype=np.uint8)
        _lib().zf_wav8_state_init(self.state.ctypes.data, self.state.size)

    def convert(self, raw):
        raw = _u8(raw)
        n = raw.size // self.channels
        out = np.empty(n * self.channels, dtype=np.int8)
        _lib().zf_wav8_to_samples(raw.ctypes.data, n, self.channels, self.block_size, self.pos, self.state.ctypes.data,
                                  out.ctypes.data)
        self.pos += n
        return out


def device_available(device_id=0):
    return _lib().zf_device_check(device_id) == ZF_OK


# ---- decoder (extension; the reference has none, readme.md:33) ----------------------------------------------------
ZF_DECODE_CHECK_MD5 = 1
ZF_DECODE_REQUIRE_MD5 = 2
ZF_ERR_FLAC_FRAME = -34


def flac_stream_info(flac_bytes):
    """STREAMINFO of a FLAC stream (host only)."""
    a = _u8(flac_bytes)
    info = ZfDecodeInfo(struct_size=C.sizeof(ZfDecodeInfo))
    rc = _lib().zf_flac_stream_info(a.ctypes.data, a.size, C.byref(info))
    if rc != ZF_OK:
        raise FlacGpuError(rc, "flac_stream_info")
    return info.as_dict()


class Decoder:
    """FLAC decoder on the device (zf_decoder_*).  decode() returns (pcm bytes as uint8 array, info dict): interleaved
    little-endian samples of bit_depth / 8 bytes, signed -- the layout Encoder.encode_pcm takes.  Raises FlacGpuError
    without an sm_100 GPU (no CPU fallback) and on any damaged frame."""

    def __init__(self, device_id=0):
        h = C.c_void_p()
        rc = _lib().zf_decoder_create(device_id, C.byref(h))
        if rc != ZF_OK:
            raise FlacGpuError(rc, "zf_decoder_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            _lib().zf_decoder_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def decode(self, flac_bytes, out=None, check_md5=False, require_md5=False):
        a = _u8(flac_bytes)
        info = ZfDecodeInfo(struct_size=C.sizeof(ZfDecodeInfo))
        flags = (ZF_DECODE_CHECK_MD5 if check_md5 else 0) | (ZF_DECODE_REQUIRE_MD5 if require_md5 else 0)
        n = C.c_size_t()
        if out is None:
            rc = _lib().zf_decode_flac(self.handle, a.ctypes.data, a.size, None, 0, C.byref(n), 0, C.byref(info))
            if rc != ZF_OK:
                raise FlacGpuError(rc, "zf_decode_flac (sizes)")
            out = np.empty(n.value, dtype=np.uint8)
        rc = _lib().zf_decode_flac(self.handle, a.ctypes.data, a.size, out.ctypes.data, out.size, C.byref(n), flags, C.byref(info))
        if rc != ZF_OK:
            e = FlacGpuError(rc, "zf_decode_flac")
            e.info = info.as_dict()
            raise e
        return out[:n.value], info.as_dict()

    def decode_device(self, d_flac_ptr, flac_len, d_pcm_ptr, pcm_cap):
        info = ZfDecodeInfo(struct_size=C.sizeof(ZfDecodeInfo))
        n = C.c_size_t()
        rc = _lib().zf_decode_flac_device(self.handle, d_flac_ptr, flac_len, d_pcm_ptr, pcm_cap, C.byref(n), C.byref(info))
        if rc != ZF_OK:
            e = FlacGpuError(rc, "zf_decode_flac_device")
            e.info = info.as_dict()
            raise e
        return n.value, info.as_dict()


def decode_flac(flac_bytes, device_id=0, check_md5=False, require_md5=False):
    """zf_decode_flac_memory: one-shot decode -> (pcm uint8 array, info dict)."""
    a = _u8(flac_bytes)
    p = C.c_void_p()
    n = C.c_size_t()
    info = ZfDecodeInfo(struct_size=C.sizeof(ZfDecodeInfo))
    flags = (ZF_DECODE_CHECK_MD5 if check_md5 else 0) | (ZF_DECODE_REQUIRE_MD5 if require_md5 else 0)
    rc = _lib().zf_decode_flac_memory(a.ctypes.data, a.size, device_id, flags, C.byref(p), C.byref(n), C.byref(info))
    if rc != ZF_OK:
        e = FlacGpuError(rc, "zf_decode_flac_memory")
        e.info = info.as_dict()
        raise e
    out = np.ctypeslib.as_array((C.c_uint8 * max(n.value, 1)).from_address(p.value))[:n.value].copy()
    _lib().zf_free(p)
    return out, info.as_dict()


def decode_file(in_path, out_path, device_id=0, require_md5=True):
    return _lib().zf_decode_flac_file(os.fsencode(in_path), os.fsencode(out_path), device_id,
                                      ZF_DECODE_REQUIRE_MD5 if require_md5 else 0)


def verify_file(wav_path, flac_path, device_id=0):
    """zf_verify_flac_file: decode flac_path on the device and compare with the samples of wav_path."""
    return _lib().zf_verify_flac_file(os.fsencode(wav_path), os.fsencode(flac_path), device_id)

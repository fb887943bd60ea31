/*
 * zigflac_b200.h -- C ABI of the B200-native FLAC encode engine (libzigflac_b200.so).
 *
 * Drop-in boundary for the encode hot path of toastori/zig-flac: everything below sits behind the
 * reference's `Encoder` type (src/lib/encoder.zig) and its one caller, the frame loop of
 * src/cli/wav2flac.zig:66-97.  Each entry point cites the reference interface it replaces.
 *
 * Conventions: plain C, no C++ types or exceptions cross the boundary; every function returns an
 * `int` status (ZF_OK = 0, negative = error, see zf_strerror).  Handles are not thread-safe;
 * distinct handles (one per GPU) are independent.  There is NO CPU fallback: every encode entry
 * fails with ZF_ERR_NO_DEVICE / ZF_ERR_CUDA when no sm_100 device is usable.
 *
 * Output parity: frames are byte-identical to what the reference's Encoder.writeFrame emits for
 * the same samples, frame number and Config (CONSTANT / VERBATIM / FIXED subframes, stereo mode
 * selection, Rice partition search, CRC-8 / CRC-16), quirks included.
 */
#ifndef ZIGFLAC_B200_H
#define ZIGFLAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZF_ABI_VERSION 1

enum {
    ZF_OK = 0,
    ZF_ERR_INVALID_ARG = -1,   /* NULL pointer, zero block size, > 8 channels, ... (asserts encoder.zig:49-51) */
    ZF_ERR_UNSUPPORTED = -2,   /* configuration the reference cannot encode either, or outside this build
                                  (bit depth not 8/16/24/32, block_size > 4096, frame number >= 2^31) */
    ZF_ERR_NO_DEVICE = -3,     /* no CUDA device / not sm_100 -- there is no CPU fallback */
    ZF_ERR_CUDA = -4,          /* CUDA runtime error; zf_last_cuda_error() has the text */
    ZF_ERR_NOMEM = -5,         /* Allocator.Error (encoder.zig:48) */
    ZF_ERR_OUT_TOO_SMALL = -6, /* Writer.Error.WriteFailed (encoder.zig:234): output capacity exceeded */
    ZF_ERR_BUSY = -7,          /* a submitted batch has not been collected yet */
    ZF_ERR_IO = -8,            /* file open/read/write failure (host driver only) */
    ZF_ERR_WAV_NOT_RIFF = -16, /* EncodingError.NotRiffFile     wav_reader.zig:254 */
    ZF_ERR_WAV_NOT_WAVE = -17, /* EncodingError.NotWaveFile     wav_reader.zig:255 */
    ZF_ERR_WAV_EOF = -18,      /* reader hit end of stream while parsing */
    ZF_ERR_WAV_DATA_LEN = -19, /* EncodingError.InvalidDataLen  wav_reader.zig:259 */
    ZF_ERR_WAV_CODEC = -20,    /* EncodingError.UnsupportCodec  wav_reader.zig:261 */
    ZF_ERR_WAV_BIT_DEPTH = -21,/* EncodingError.UnsupportBitDepth wav_reader.zig:263 */
    ZF_ERR_WAV_NO_DATA = -22,  /* EncodingError.DataNotFound    wav_reader.zig:265 */
    ZF_ERR_WAV_BIT_RATE = -23, /* EncodingError.BitRateUnmatch  wav_reader.zig:267 */
    ZF_ERR_WAV_INCOMPLETE = -24,/* StreamError.IncompleteStream  wav_reader.zig:273 */
    ZF_ERR_FLAC_NOT_FLAC = -32, /* decoder: no "fLaC" marker / no STREAMINFO block */
    ZF_ERR_FLAC_TRUNCATED = -33,/* decoder: metadata or the first frame runs past the end of the stream */
    ZF_ERR_FLAC_FRAME = -34,    /* decoder: a frame failed (zf_decode_info.bad_frame / bad_frame_status say which and why) */
    ZF_ERR_FLAC_COUNT = -35,    /* decoder: decoded sample count differs from STREAMINFO's */
    ZF_ERR_FLAC_MD5 = -36       /* decoder: ZF_DECODE_REQUIRE_MD5 and the MD5 of the decoded samples differs */
};

/* Encoder.Config + Config.Feature (encoder.zig:609-656) plus what FrameInfo (encoder.zig:658-663)
 * carries per call in the reference (sample_rate) and the device placement. */
typedef struct zf_config {
    uint32_t struct_size;          /* sizeof(zf_config), for ABI evolution */
    uint16_t block_size;           /* Config.block_size; reference default 4096 (encoder.zig:644) */
    uint8_t bit_depth;             /* 8, 16, 24 or 32; PCM is bit_depth/8 bytes per sample, little-endian, signed (8-bit
                                      too: see zf_wav8_to_samples) */
    uint8_t channels;              /* 1..8; 2 + stereo_decorrelation selects L/R, L/S, S/R, M/S per frame */
    uint32_t sample_rate;          /* FrameInfo.sample_rate.  Rates without a frame-header code of their own (anything but 8, 16,
                                      22.05, 24, 32, 44.1, 48, 88.2, 96, 176.4, 192 kHz) are encoded the way the reference does:
                                      it writes the BLOCK SIZE into the rate trailer (frame_writer.zig:258-262), so
                                      such streams are bug-compatible with upstream and decode only with decoders that take the
                                      rate from STREAMINFO (the one below does) */
    uint8_t stereo_decorrelation;  /* Feature.stereo_decorrelation */
    uint8_t max_rice_order;        /* Feature.max_rice_order, 0..8 */
    uint8_t max_rice_param;        /* Feature.max_rice_param, 1..30 */
    uint8_t lpc_order;             /* 0 = the reference's encoder (fixed predictors only).  1..12: LPC subframes as one more
                                      candidate per channel, up to this order (BASELINE config 4).  The reference has no
                                      LPC (encoder.zig:626-640 is a stub), so this mode has no byte-parity claim: the
                                      stream is standard FLAC, decodes losslessly and is not larger than the FIXED one.
                                      Stereo with decorrelation, 8/16/24-bit only. */
    int32_t device_id;             /* CUDA device ordinal */
    uint32_t max_frames_per_batch; /* capacity of one submit; device buffers are sized from it */
    uint8_t exact_rice;            /* 0 = the reference's Rice size estimate (rice.zig:402-405).  1 = extension: every partition's
                                      parameter by its exact code length (what the reference's dead calcParamExact family,
                                      rice.zig:110-245, set out to do) -- slightly smaller streams, no byte-parity claim.
                                      Stereo with decorrelation, no LPC. */
    uint8_t reserved1[3];
} zf_config;

typedef struct zf_encoder zf_encoder;

/* Encoder.Config.default(channels, bit_depth) -- encoder.zig:642-655. */
int zf_config_default(zf_config *cfg, uint8_t channels, uint8_t bit_depth, uint32_t sample_rate);

/* maxFrameBytes -- encoder.zig:583-595 (upper bound the reference sizes its frame buffer with). */
size_t zf_max_frame_bytes(const zf_config *cfg);
/* Capacity that always suffices for n_frames frames of output from zf_encode_*. */
size_t zf_max_batch_bytes(const zf_config *cfg, uint32_t n_frames);

/* Encoder.init -- encoder.zig:44-118.  Allocates device buffers, pinned staging and streams once. */
int zf_encoder_create(const zf_config *cfg, zf_encoder **out);
/* Encoder.deinit -- encoder.zig:121-164. */
void zf_encoder_destroy(zf_encoder *enc);

/*
 * K x Encoder.writeFrame (encoder.zig:234-284) fused with the sample conversion of
 * WavReader.fillSamples (wav_reader.zig:56-90): `pcm` is the raw little-endian interleaved WAV
 * `data` payload (the very bytes the host MD5s), samples_per_channel inter-channel samples of it.
 * It is cut into ceil(samples/block_size) frames numbered first_frame_number, +1, ...; only the
 * last may be short.  `out` receives the frames back to back, frame_sizes[i] each frame's byte
 * count (what writeFrame returns, for StreamInfo.updateFrameSize -- metadata.zig:35-40).
 * Blocking; host buffers may be pageable (they are staged through pinned memory).
 */
int zf_encode_pcm(zf_encoder *enc, const uint8_t *pcm, uint64_t samples_per_channel, uint64_t first_frame_number,
                  uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *frame_sizes, uint32_t frame_sizes_cap,
                  uint32_t *n_frames);

/* Asynchronous pair for pipelining (one batch in flight per handle): submit copies the input to the
 * device and launches the kernels on the handle's stream; collect waits and copies results back.
 * `pcm` must stay valid until submit returns (pageable memory is staged at once; for page-locked memory
 * submit waits for the upload, not for the kernels); `out`/`frame_sizes` are only touched by collect. */
int zf_encode_submit(zf_encoder *enc, const uint8_t *pcm, uint64_t samples_per_channel, uint64_t first_frame_number);
int zf_encode_collect(zf_encoder *enc, uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *frame_sizes,
                      uint32_t frame_sizes_cap, uint32_t *n_frames);

/*
 * Device-resident variant: d_pcm / d_out / d_frame_sizes / d_total_bytes are device pointers on the
 * encoder's device; nothing is copied to or from the host.  `stream` is a cudaStream_t (NULL = the
 * encoder's own stream).  Asynchronous with respect to the host: results are valid once the stream
 * has been synchronised; *d_total_bytes (uint64 on the device) receives the byte length of d_out.
 */
int zf_encode_device(zf_encoder *enc, const void *d_pcm, uint64_t samples_per_channel, uint64_t first_frame_number,
                     void *d_out, size_t out_cap, uint32_t *d_frame_sizes, uint64_t *d_total_bytes, void *stream);
/*
 * Device-resident batches of one handle share the handle's control block (frame tickets, look-back descriptors,
 * status word): consecutive calls are ordered on the device by the library even when they name different streams,
 * and ZF_ERR_BUSY is returned while a zf_encode_submit batch has not been collected.  Frames that do not fit
 * out_cap are not written, but *d_total_bytes still counts them: total > out_cap means overflow.  This call reports
 * the same from the status word of the most recent device-resident batch (waits for that batch).
 */
#define ZF_STATUS_OUT_OVERFLOW 1u /* out_cap exceeded: Writer.Error.WriteFailed */
#define ZF_STATUS_INTERNAL 2u     /* a frame exceeded maxFrameBytes (cannot happen for valid configurations) */
int zf_encode_device_status(zf_encoder *enc, uint32_t *flags);

/* Encoder.writeFrame itself (encoder.zig:234): ONE frame from planar, sign-extended int32 planes
 * (the layout of Encoder.samples the caller fills through WavReader.fillSamples, wav_reader.zig:44).
 * planes[ch] points at samples_count host samples.  Returns the frame bytes in out, its size in *frame_len. */
int zf_write_frame(zf_encoder *enc, const int32_t *const *planes, uint32_t samples_count, uint64_t frame_number,
                   uint8_t *out, size_t out_cap, size_t *frame_len);

/* Device time (ms, CUDA events on the launch stream) of the kernels of the last completed batch,
 * and how many kernel launches it took. */
int zf_last_batch_stats(zf_encoder *enc, float *kernel_ms, uint32_t *launches);
/* Device durations (ms, CUDA events on the launch stream) of the full-frame encode kernel of the batches
 * enqueued since the previous call (at most 256 per slot are kept); waits for them to finish.  Used by
 * bench.py for the roofline figure. */
int zf_kernel_times(zf_encoder *enc, float *ms, uint32_t cap, uint32_t *n);

const char *zf_strerror(int status);
const char *zf_last_cuda_error(void);
/* 0 if an sm_100 CUDA device is usable, else ZF_ERR_NO_DEVICE / ZF_ERR_CUDA. */
int zf_device_check(int device_id);

/* ---- host-side stream metadata (stays on the host, as in the reference) ------------------------------- */

/* metadata.StreamInfo -- metadata.zig:22-33 */
typedef struct zf_streaminfo {
    uint8_t md5[16];
    uint64_t interchannel_samples;
    uint32_t min_frame_size; /* u24; starts at 0xFFFFFF */
    uint32_t max_frame_size; /* u24; starts at 0 */
    uint32_t sample_rate;
    uint16_t min_block_size;
    uint16_t max_block_size;
    uint8_t channels;
    uint8_t bit_depth;
} zf_streaminfo;

void zf_streaminfo_init(zf_streaminfo *si);
/* StreamInfo.updateFrameSize -- metadata.zig:35-40.  Order dependent: replay in frame order. */
void zf_streaminfo_update_frame_size(zf_streaminfo *si, uint32_t frame_size);
/* StreamInfo.bytes -- metadata.zig:42-68 */
void zf_streaminfo_bytes(const zf_streaminfo *si, uint8_t out[34]);
/* Encoder.writeHeader -- encoder.zig:192-205: "fLaC" + STREAMINFO block, 42 bytes. */
size_t zf_write_stream_header(const zf_streaminfo *si, int last_metadata, uint8_t out[42]);
/* Encoder.writeVorbisComment -- encoder.zig:211-226: vendor-only VORBIS_COMMENT block, 31 bytes. */
size_t zf_write_vorbis_comment(int last_metadata, uint8_t out[31]);

/* Md5 -- md5.zig:31 (std.crypto.hash.Md5); serial, runs on a host thread overlapped with the GPU. */
typedef struct zf_md5 {
    uint32_t state[4];
    uint64_t length;
    uint8_t buffer[64];
} zf_md5;
void zf_md5_init(zf_md5 *m);
void zf_md5_update(zf_md5 *m, const uint8_t *data, size_t len);
void zf_md5_final(zf_md5 *m, uint8_t digest[16]);

/* Optional libcrypto MD5, the counterpart of the reference's `-Dlink_ossl` build option (md5.zig:3-35, build.zig:10-18):
 * the library dlopen()s libcrypto on demand; the whole-file driver uses it when ZF_MD5=openssl is set in the environment.
 * Returns 1 when MD5_Init/Update/Final were found. */
int zf_md5_openssl_available(void);

/* WavReader.fillSamples for one-byte containers (wav_reader.zig:56-90): raw WAV bytes -> the signed samples the reference's
 * reader leaves in Encoder.samples, interleaved, one byte each -- the input of zf_encode_pcm at bit_depth 8.  The
 * reference subtracts 128 from the unshifted plane word (:71-78), so a sample is (int8)(byte - borrow) with borrow = 1
 * iff the plane's previous content at that (channel, index in block) was non-negative (1 in a fresh plane): `state`
 * (block_size * channels bytes, zf_wav8_state_init) carries that from call to call; first_sample is the stream position
 * of raw[0].  Serial by nature, host only. */
void zf_wav8_state_init(uint8_t *state, size_t n);
void zf_wav8_to_samples(const uint8_t *raw, uint64_t samples_per_channel, uint32_t channels, uint32_t block_size,
                        uint64_t first_sample, uint8_t *state, int8_t *out);

/* WavReader.getFmt -- wav_reader.zig:116-170 */
typedef struct zf_wav_format {
    uint32_t samples_count; /* per channel, as the reference computes it (wav_reader.zig:169) */
    uint32_t sample_rate;
    uint16_t bit_depth;
    uint16_t channels;
    uint8_t bytes_per_sample;
    uint8_t reserved[3];
    uint64_t data_offset;
    uint32_t data_len;
} zf_wav_format;
int zf_wav_parse(const uint8_t *file, size_t len, zf_wav_format *fmt);

/*
 * cli.zig:7-27 + wav2flac.zig:10-97: `flac in.wav out.flac` -- parse WAV, encode on the listed GPUs
 * (contiguous frame ranges per device, no inter-GPU communication), MD5 on a host thread, back-patch
 * STREAMINFO.  n_devices = 0 uses device 0.  Returns ZF_OK, or 2 when FLAC cannot carry the format
 * (the reference's exit code, wav2flac.zig:24-27), or a negative error.
 */
int zf_encode_wav_file(const char *in_path, const char *out_path, const int *devices, int n_devices);
/* Same on in-memory buffers; *flac is malloc'd, release with zf_free. */
int zf_encode_wav_memory(const uint8_t *wav, size_t wav_len, uint8_t **flac, size_t *flac_len, const int *devices,
                         int n_devices);
void zf_free(void *p);
/* The whole-file driver keeps its encoder handles and page-locked chunk buffers between calls (creating them costs more than
 * encoding a short file); whole-file calls of one process are serialised.  This releases them. */
void zf_driver_release_cache(void);

/*
 * Page-locked host memory for PCM / FLAC buffers (the counterpart of the sample buffer wav2flac.zig:72 takes from its
 * allocator).  zf_encode_pcm copies straight from / into such buffers; any other host memory is staged through the
 * encoder's own pinned buffers with an extra host copy per batch.  Returns ZF_ERR_NOMEM / ZF_ERR_NO_DEVICE on failure.
 */
int zf_host_alloc(size_t bytes, void **out);
void zf_host_free(void *p);

/* ---- decoder (extension: SURVEY.md 8(f) rank 4) ------------------------------------------------------------
 *
 * The reference has no decoder (readme.md:33 lists it as queued work), so these entries replace nothing upstream; the
 * contract is the FLAC format (RFC 9639).  Fixed-blocksize streams of 8/16/24/32-bit samples, 1..8 channels, every
 * subframe type, i.e. everything the encode entries above (and libFLAC's default settings) produce.  All decoding runs
 * on the device (zf_kernel_decode.cuh: header scan, one thread per frame for the serial bit parse, CRC-16, inter-channel
 * restore and interleave); only the metadata blocks, the chaining of frame headers by frame number and the optional
 * MD5 are host work.  No CPU fallback.
 */
typedef struct zf_decoder zf_decoder;

#define ZF_DECODE_CHECK_MD5 1u    /* MD5 the decoded samples on the host and compare with STREAMINFO (md5_status) */
#define ZF_DECODE_REQUIRE_MD5 2u  /* ... and fail with ZF_ERR_FLAC_MD5 when it differs (a zero STREAMINFO MD5 passes) */

typedef struct zf_decode_info {
    uint32_t struct_size;          /* set by the caller to sizeof(zf_decode_info) */
    uint32_t sample_rate, channels, bit_depth, min_block_size, max_block_size;
    uint32_t launches;             /* kernel launches of the call */
    float kernel_ms;               /* device time of those kernels (CUDA events) */
    uint64_t samples_per_channel;  /* decoded */
    uint64_t streaminfo_samples;   /* STREAMINFO's total (0 = unknown) */
    uint64_t pcm_bytes;            /* samples_per_channel * channels * bit_depth / 8 */
    uint64_t n_frames;
    uint64_t bad_frame;            /* ZF_ERR_FLAC_FRAME: index of the first failing frame ... */
    uint32_t bad_frame_status;     /* ... and why: 1 header, 2 reserved value, 3 range, 4 overrun, 5 length, 6 padding, 7 CRC-16 */
    int32_t md5_status;            /* 1 match, 0 mismatch, -1 not checked or STREAMINFO carries no MD5 */
    uint8_t md5[16];               /* STREAMINFO's */
} zf_decode_info;

int zf_decoder_create(int device_id, zf_decoder **out);
void zf_decoder_destroy(zf_decoder *dec);
/* STREAMINFO of a stream in host memory (host only, no device needed). */
int zf_flac_stream_info(const uint8_t *flac, size_t len, zf_decode_info *info);
/*
 * Decodes a whole FLAC stream from host memory into `pcm`: interleaved little-endian samples of bit_depth / 8 bytes,
 * signed (8-bit too) -- the layout zf_encode_pcm takes.  Page-locked buffers (zf_host_alloc) are copied from / into
 * directly.  ZF_ERR_OUT_TOO_SMALL when pcm_cap < info->pcm_bytes (info is filled in, so the call can be repeated).
 */
int zf_decode_flac(zf_decoder *dec, const uint8_t *flac, size_t len, uint8_t *pcm, size_t pcm_cap, size_t *pcm_len,
                   uint32_t flags, zf_decode_info *info);
/* The same with device pointers on the decoder's device (nothing but the frame table crosses PCIe); MD5 flags are
 * not available here.  Blocking. */
int zf_decode_flac_device(zf_decoder *dec, const void *d_flac, size_t len, void *d_pcm, size_t pcm_cap, size_t *pcm_len,
                          zf_decode_info *info);
/* Convenience: *pcm is malloc'd (zf_free). */
int zf_decode_flac_memory(const uint8_t *flac, size_t len, int device_id, uint32_t flags, uint8_t **pcm, size_t *pcm_len,
                          zf_decode_info *info);
/* `flac -d in.flac out.wav`: canonical PCM WAV (8-bit samples become unsigned, as WAV stores them). */
int zf_decode_flac_file(const char *in_path, const char *out_path, int device_id, uint32_t flags);
/* `flac -V in.wav out.flac` (verify after encoding): decodes out_path on the device and compares the samples with the `data`
 * chunk of in_path byte for byte, CRC-16 of every frame and the STREAMINFO MD5 included.  ZF_OK, ZF_ERR_FLAC_* from the
 * decoder, or ZF_ERR_FLAC_MD5 when the decoded samples differ from the WAV's.  (8-bit WAV goes through the reference
 * reader's conversion, zf_wav8_to_samples, before the comparison.) */
int zf_verify_flac_file(const char *wav_path, const char *flac_path, int device_id);

/* ---- benchmark / test utility --------------------------------------------------------------------------- */

/* Deterministic synthetic stereo PCM: count inter-channel samples starting at stream
 * position first_sample, interleaved little-endian.  Host only; not part of the encode path. */
int zf_synth_pcm(uint8_t *out, uint64_t first_sample, uint64_t count, uint32_t sample_rate, uint32_t bit_depth,
                 uint64_t seed, int n_threads);

int zf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZIGFLAC_B200_H */

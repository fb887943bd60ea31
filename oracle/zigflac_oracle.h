/*
 * zigflac_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the encode hot path of toastori/zig-flac, function by function,
 * each citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this.  The product path
 * (zig-flac_b200/csrc) never links or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- the reference ships no tests, no
 * golden vectors and cannot be compiled here (Zig-only, no Zig toolchain in the image).  The
 * oracle is pinned instead by (1) the known-answer vectors of SURVEY.md 8-K, (2) the catalogue
 * check values of CRC-8/SMBUS, CRC-16/UMTS and RFC 1321 MD5, and (3) an independent
 * specification decoder (oracle/flac_decode.c) that must reproduce the PCM bit-exactly.
 */
#ifndef ZIGFLAC_ORACLE_H
#define ZIGFLAC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZO_MAX_RICE_ORDER 8   /* rice.zig:12 */
#define ZO_MAX_PART 256       /* rice.zig:13 */
#define ZO_FIXED_MAX_ORDER 4  /* fixed.zig:10 */

/* Encoder.Config + Feature, encoder.zig:609-656 */
typedef struct {
    uint16_t block_size;
    uint8_t bit_depth;
    uint8_t channels;
    uint8_t stereo_decorrelation;
    uint8_t max_rice_order;
    uint8_t max_rice_param;
    uint8_t lpc_order; /* 0 = the reference's path (fixed predictors only); 1..32 = LPC extension, zigflac_lpc.h */
    uint8_t exact_rice; /* 0 = the reference's estimate (rice.zig:402-405); 1 = extension: exact code lengths */
    uint8_t reserved;
} zo_config;

/* FrameInfo, encoder.zig:658-663 */
typedef struct {
    uint8_t bit_depth;
    uint8_t channels;
    uint16_t samples_count;
    uint32_t sample_rate;
} zo_frame_info;

/* rice.Config, rice.zig:37-82.  params[i] bit7 = escape, low 7 bits = param / escape width. */
typedef struct {
    uint8_t method; /* 0 = FOUR, 1 = FIVE */
    uint8_t part_order;
    uint8_t params[ZO_MAX_PART];
} zo_rice_config;

enum { ZO_CONSTANT = 0, ZO_VERBATIM = 1, ZO_FIXED = 2, ZO_LPC = 3 /* extension: zigflac_lpc.h */ };

/* SubframeType.Encoding, encoder.zig:678-702, flattened; also the per-subframe decision record
 * tests diff against the GPU path. */
typedef struct {
    uint8_t kind;
    uint8_t waste_bits;
    uint8_t order;
    uint8_t wide;        /* verbatim/constant source is the i64 side plane */
    uint64_t est_bits;   /* the size chooseSubframeEncoding returned */
    int64_t sample;      /* CONSTANT */
    int64_t warmup[4];   /* FIXED */
    const int32_t *samples32;
    const int64_t *samples64;
    const int32_t *residuals;
    uint32_t len;
    zo_rice_config rice;
    uint8_t lpc_shift, lpc_precision; /* ZO_LPC: order in `order`, warm-ups read from the plane */
    int32_t lpc_q[32];
} zo_encoding;

typedef struct {
    uint8_t ch_type;          /* type.zig:1-27: 0..7 independent (channels-1), 8 l_s, 9 s_r, 10 m_s */
    uint8_t n_sub;            /* subframes written */
    uint8_t sub_src[8];       /* which evaluated plane each written subframe came from */
    zo_encoding enc[8];       /* stereo: [left,right,mid,side]; indep: channel order */
} zo_frame_decision;

typedef struct zo_encoder zo_encoder;

/* Encoder.Config.default, encoder.zig:642-655 */
void zo_config_default(zo_config *cfg, uint8_t channels, uint8_t bit_depth);
/* maxFrameBytes, encoder.zig:583-595 (4th argument mis-passed compute_waste_bits=true at :59) */
size_t zo_max_frame_bytes(uint16_t block_size, uint8_t bit_depth, uint8_t channels, int stereo_decorrelation);

/* Encoder.init / deinit, encoder.zig:44-164 */
zo_encoder *zo_encoder_create(const zo_config *cfg);
void zo_encoder_destroy(zo_encoder *e);
/* Encoder.samples[ch] -- the planar i32 plane the caller fills (wav_reader.zig:44) */
int32_t *zo_encoder_samples(zo_encoder *e, int ch);

/* Encoder.writeFrame, encoder.zig:234-284.  Appends one frame to out (capacity out_cap), returns its
 * byte count (u24) or 0 on overflow.  decision (optional) receives the per-subframe choices. */
uint32_t zo_write_frame(zo_encoder *e, uint64_t frame_number, const zo_frame_info *fi, uint8_t *out,
                        size_t out_cap, zo_frame_decision *decision);

/* WavReader.fillSamples body, wav_reader.zig:56-90 minus I/O and MD5: raw little-endian interleaved
 * bytes -> planar sign-extended i32 in the encoder's planes. */
void zo_bytes_to_planes(zo_encoder *e, const uint8_t *bytes, size_t n_samples, uint8_t bit_depth,
                        uint8_t channels);

/* The loop of wav2flac.encode, wav2flac.zig:66-97, over a raw PCM `data` payload held in memory.
 * Frames are independent, so n_threads > 1 shards contiguous frame ranges over host threads (the
 * reference is single-threaded; n_threads = 1 is the like-for-like figure).
 * frame_sizes[] receives each frame's byte count.  Returns total bytes or (size_t)-1 on overflow. */
size_t zo_encode_pcm(const zo_config *cfg, uint32_t sample_rate, const uint8_t *pcm,
                     uint64_t samples_per_channel, uint64_t first_frame_number, uint8_t *out,
                     size_t out_cap, uint32_t *frame_sizes, uint32_t *n_frames, int n_threads);

/* metadata.StreamInfo, metadata.zig:22-68 */
typedef struct {
    uint8_t md5[16];
    uint64_t interchannel_samples;
    uint32_t min_frame_size; /* u24, starts 0xFFFFFF */
    uint32_t max_frame_size; /* u24, starts 0 */
    uint32_t sample_rate;
    uint16_t min_block_size;
    uint16_t max_block_size;
    uint8_t channels;
    uint8_t bit_depth;
} zo_streaminfo;

void zo_streaminfo_init(zo_streaminfo *si);
void zo_streaminfo_update_frame_size(zo_streaminfo *si, uint32_t frame_size); /* metadata.zig:35-40 */
void zo_streaminfo_bytes(const zo_streaminfo *si, uint8_t out[34]);           /* metadata.zig:42-68 */

/* Encoder.writeHeader (encoder.zig:192-205) -> 42 bytes; writeVorbisComment (:211-226) -> 31 bytes */
size_t zo_write_stream_header(const zo_streaminfo *si, int last_metadata, uint8_t out[42]);
size_t zo_write_vorbis_comment(int last_metadata, uint8_t out[31]);

/* WavReader.getFmt, wav_reader.zig:116-170.  Returns 0 or a negative EncodingError. */
typedef struct {
    uint32_t samples_count;
    uint32_t sample_rate;
    uint16_t bit_depth;
    uint16_t channels;
    uint8_t bytes_per_sample;
    size_t data_offset;
    uint32_t data_len;
} zo_wav_fmt;
int zo_wav_parse(const uint8_t *file, size_t len, zo_wav_fmt *fmt);

/* cli.zig + wav2flac.zig main: whole file in memory -> whole .flac in memory.
 * Returns 0, 2 on "flac does not support this wav format", negative on parse errors. */
int zo_wav_to_flac(const uint8_t *wav, size_t wav_len, uint8_t **flac, size_t *flac_len, int n_threads);
void zo_free(void *p);

/* std.hash.crc.Crc8Smbus / Crc16Umts (Zig std, toolchain-bundled), crc16.zig:15-57 */
uint8_t zo_crc8(const uint8_t *p, size_t n);
uint16_t zo_crc16(uint16_t crc, const uint8_t *p, size_t n);
/* restatement of the CLMUL folding path of crc16.zig:23-57 (x86 PCLMULQDQ); must equal zo_crc16 */
uint16_t zo_crc16_clmul(uint16_t crc, const uint8_t *p, size_t n);

/* std.crypto.hash.Md5 (RFC 1321), md5.zig:31 */
typedef struct {
    uint32_t s[4];
    uint64_t n;
    uint8_t buf[64];
} zo_md5;
void zo_md5_init(zo_md5 *m);
void zo_md5_update(zo_md5 *m, const uint8_t *p, size_t n);
void zo_md5_final(zo_md5 *m, uint8_t out[16]);

/* rice.zig:402-405 size estimate with the Zig precedence quirk; exposed for tests */
uint64_t zo_flac_calc_part_size(uint64_t part_size, uint64_t param, uint64_t abs_sum);
/* frame_writer.zig:151-265 header only (known-answer vectors of SURVEY 8-K) */
size_t zo_frame_header(uint64_t frame_number, uint8_t bit_depth, uint8_t ch_type, uint16_t block_size,
                       uint32_t sample_rate, uint8_t out[16]);

#ifdef __cplusplus
}
#endif
#endif

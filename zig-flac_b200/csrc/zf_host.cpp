// zf_host.cpp -- host-side pieces that stay on the CPU, as in the reference: MD5, STREAMINFO,
// the vendor VORBIS_COMMENT block, the RIFF/WAVE parser.  Pure C++, no CUDA.
//
// Reference: src/lib/md5.zig, src/lib/metadata.zig, src/lib/encoder.zig:177-226,
// src/lib/wav_reader.zig:116-170.
#include <string.h>

#include "../../include/zigflac_b200.h"

extern "C" {

int zf_abi_version(void) { return ZF_ABI_VERSION; }

const char *zf_strerror(int status) {
    switch (status) {
        case ZF_OK: return "ok";
        case 2: return "format: flac does not support this wav format";
        case ZF_ERR_INVALID_ARG: return "invalid argument";
        case ZF_ERR_UNSUPPORTED: return "unsupported configuration";
        case ZF_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU fallback)";
        case ZF_ERR_CUDA: return "CUDA runtime error";
        case ZF_ERR_NOMEM: return "out of memory";
        case ZF_ERR_OUT_TOO_SMALL: return "output buffer too small (WriteFailed)";
        case ZF_ERR_BUSY: return "a submitted batch has not been collected";
        case ZF_ERR_IO: return "file I/O error";
        case ZF_ERR_WAV_NOT_RIFF: return "NotRiffFile";
        case ZF_ERR_WAV_NOT_WAVE: return "NotWaveFile";
        case ZF_ERR_WAV_EOF: return "EndOfStream";
        case ZF_ERR_WAV_DATA_LEN: return "InvalidDataLen";
        case ZF_ERR_WAV_CODEC: return "UnsupportCodec";
        case ZF_ERR_WAV_BIT_DEPTH: return "UnsupportBitDepth";
        case ZF_ERR_WAV_NO_DATA: return "DataNotFound";
        case ZF_ERR_WAV_BIT_RATE: return "BitRateUnmatch";
        case ZF_ERR_WAV_INCOMPLETE: return "IncompleteStream";
        case ZF_ERR_FLAC_NOT_FLAC: return "not a FLAC stream";
        case ZF_ERR_FLAC_TRUNCATED: return "FLAC stream truncated";
        case ZF_ERR_FLAC_FRAME: return "FLAC frame failed to decode";
        case ZF_ERR_FLAC_COUNT: return "decoded sample count differs from STREAMINFO";
        case ZF_ERR_FLAC_MD5: return "MD5 of the decoded samples differs from STREAMINFO";
        default: return "unknown status";
    }
}

// ---- MD5 (RFC 1321); md5.zig:31 uses std.crypto.hash.Md5 ----------------------------------------------

static inline uint32_t rotl(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

static void md5_transform(uint32_t st[4], const uint8_t *blk) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8,
        0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340,
        0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87,
        0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
        0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039,
        0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92,
        0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb,
        0xeb86d391};
    uint32_t m[16];
    memcpy(m, blk, 64);  // little-endian host
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#define STEP(f, w, x, y, z, mi, ki, s) w = x + rotl(w + f(x, y, z) + m[mi] + K[ki], s)
#define F1(x, y, z) (z ^ (x & (y ^ z)))
#define F2(x, y, z) (y ^ (z & (x ^ y)))
#define F3(x, y, z) (x ^ y ^ z)
#define F4(x, y, z) (y ^ (x | ~z))
    for (int i = 0; i < 16; i += 4) {
        STEP(F1, a, b, c, d, i, i, 7); STEP(F1, d, a, b, c, i + 1, i + 1, 12);
        STEP(F1, c, d, a, b, i + 2, i + 2, 17); STEP(F1, b, c, d, a, i + 3, i + 3, 22);
    }
    for (int i = 0; i < 16; i += 4) {
        STEP(F2, a, b, c, d, (5 * i + 1) & 15, 16 + i, 5); STEP(F2, d, a, b, c, (5 * (i + 1) + 1) & 15, 17 + i, 9);
        STEP(F2, c, d, a, b, (5 * (i + 2) + 1) & 15, 18 + i, 14); STEP(F2, b, c, d, a, (5 * (i + 3) + 1) & 15, 19 + i, 20);
    }
    for (int i = 0; i < 16; i += 4) {
        STEP(F3, a, b, c, d, (3 * i + 5) & 15, 32 + i, 4); STEP(F3, d, a, b, c, (3 * (i + 1) + 5) & 15, 33 + i, 11);
        STEP(F3, c, d, a, b, (3 * (i + 2) + 5) & 15, 34 + i, 16); STEP(F3, b, c, d, a, (3 * (i + 3) + 5) & 15, 35 + i, 23);
    }
    for (int i = 0; i < 16; i += 4) {
        STEP(F4, a, b, c, d, (7 * i) & 15, 48 + i, 6); STEP(F4, d, a, b, c, (7 * (i + 1)) & 15, 49 + i, 10);
        STEP(F4, c, d, a, b, (7 * (i + 2)) & 15, 50 + i, 15); STEP(F4, b, c, d, a, (7 * (i + 3)) & 15, 51 + i, 21);
    }
#undef STEP
#undef F1
#undef F2
#undef F3
#undef F4
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

void zf_md5_init(zf_md5 *m) {
    m->state[0] = 0x67452301; m->state[1] = 0xefcdab89; m->state[2] = 0x98badcfe; m->state[3] = 0x10325476;
    m->length = 0;
}

void zf_md5_update(zf_md5 *m, const uint8_t *data, size_t len) {
    size_t have = (size_t)(m->length & 63);
    m->length += len;
    if (have) {
        size_t take = 64 - have;
        if (take > len) take = len;
        memcpy(m->buffer + have, data, take);
        data += take; len -= take; have += take;
        if (have < 64) return;
        md5_transform(m->state, m->buffer);
    }
    while (len >= 64) { md5_transform(m->state, data); data += 64; len -= 64; }
    if (len) memcpy(m->buffer, data, len);
}

void zf_md5_final(zf_md5 *m, uint8_t digest[16]) {
    const uint64_t bits = m->length * 8;
    size_t have = (size_t)(m->length & 63);
    uint8_t tail[128];
    memset(tail, 0, sizeof tail);
    memcpy(tail, m->buffer, have);
    tail[have] = 0x80;
    const size_t total = (have < 56) ? 64 : 128;
    for (int i = 0; i < 8; i++) tail[total - 8 + i] = (uint8_t)(bits >> (8 * i));
    md5_transform(m->state, tail);
    if (total == 128) md5_transform(m->state, tail + 64);
    memcpy(digest, m->state, 16);  // little-endian host
}

// ---- STREAMINFO, metadata.zig:22-68 ------------------------------------------------------------------------

void zf_streaminfo_init(zf_streaminfo *si) {
    memset(si, 0, sizeof *si);
    si->min_frame_size = 0xFFFFFF;  // std.math.maxInt(u24), metadata.zig:26
    si->max_frame_size = 0;
}

void zf_streaminfo_update_frame_size(zf_streaminfo *si, uint32_t frame_size) {
    // metadata.zig:35-40: `else if` -- a frame that raises the maximum never lowers the minimum
    if (frame_size > si->max_frame_size) si->max_frame_size = frame_size;
    else if (frame_size < si->min_frame_size) si->min_frame_size = frame_size;
}

void zf_streaminfo_bytes(const zf_streaminfo *si, uint8_t out[34]) {
    out[0] = (uint8_t)(si->min_block_size >> 8); out[1] = (uint8_t)si->min_block_size;
    out[2] = (uint8_t)(si->max_block_size >> 8); out[3] = (uint8_t)si->max_block_size;
    out[4] = (uint8_t)(si->min_frame_size >> 16); out[5] = (uint8_t)(si->min_frame_size >> 8);
    out[6] = (uint8_t)si->min_frame_size;
    out[7] = (uint8_t)(si->max_frame_size >> 16); out[8] = (uint8_t)(si->max_frame_size >> 8);
    out[9] = (uint8_t)si->max_frame_size;
    // 20 bits rate | 3 bits channels-1 | 5 bits depth-1 | 36 bits samples
    const uint64_t packed = ((uint64_t)(si->sample_rate & 0xFFFFF) << 44) | ((uint64_t)((si->channels - 1) & 7) << 41) |
                            ((uint64_t)((si->bit_depth - 1) & 31) << 36) | (si->interchannel_samples & 0xFFFFFFFFFull);
    for (int i = 0; i < 8; i++) out[10 + i] = (uint8_t)(packed >> (56 - 8 * i));
    memcpy(out + 18, si->md5, 16);
}

size_t zf_write_stream_header(const zf_streaminfo *si, int last_metadata, uint8_t out[42]) {
    memcpy(out, "fLaC", 4);                    // encoder.zig:195
    out[4] = (uint8_t)(last_metadata ? 0x80 : 0x00);  // BlockHeader{is_last_block, StreamInfo=0}, :198-201
    out[5] = 0; out[6] = 0; out[7] = 34;       // :202
    zf_streaminfo_bytes(si, out + 8);          // :204
    return 42;
}

size_t zf_write_vorbis_comment(int last_metadata, uint8_t out[31]) {
    static const char vendor[] = "toastori FLAC 0.0.0";  // encoder.zig:212
    const uint32_t vlen = (uint32_t)sizeof(vendor) - 1;
    out[0] = (uint8_t)((last_metadata ? 0x80 : 0x00) | 4);  // VorbisComment = 4
    const uint32_t blen = vlen + 8;                          // :219
    out[1] = (uint8_t)(blen >> 16); out[2] = (uint8_t)(blen >> 8); out[3] = (uint8_t)blen;
    for (int i = 0; i < 4; i++) out[4 + i] = (uint8_t)(vlen >> (8 * i));  // :221 little-endian
    memcpy(out + 8, vendor, vlen);
    memset(out + 8 + vlen, 0, 4);  // no tags, :225
    return 8 + vlen + 4;
}

// ---- WavReader.fillSamples for one-byte containers, wav_reader.zig:56-90 -----------------------------------------
// The reference writes the byte into the top of the plane word, subtracts 128 from the UNSHIFTED word (:71-78) and
// shifts down (:81-88).  The low 24 bits of that word are what the plane held before -- the previous frame's sample
// at this index (possibly shifted by that frame's wasted bits, which keeps its sign) or zero in the fresh allocation --
// so the subtraction borrows one from the byte exactly when that stale value was non-negative:
//     sample = (int8)(byte - borrow),  borrow = [previous sample at this (channel, index) >= 0], 1 before the first frame.
// `state` carries the borrow per (index in block, channel) from call to call.
void zf_wav8_state_init(uint8_t *state, size_t n) { memset(state, 1, n); }

void zf_wav8_to_samples(const uint8_t *raw, uint64_t samples_per_channel, uint32_t channels, uint32_t block_size,
                        uint64_t first_sample, uint8_t *state, int8_t *out) {
    const uint64_t total = samples_per_channel * channels;
    const uint64_t period = (uint64_t)block_size * channels;
    uint64_t s = (first_sample % block_size) * channels;  // position inside the block's state
    for (uint64_t k = 0; k < total; k++) {
        const int8_t v = (int8_t)(uint8_t)(raw[k] - state[s]);
        out[k] = v;
        state[s] = v >= 0 ? 1 : 0;
        if (++s == period) s = 0;
    }
}

// ---- optional libcrypto MD5 (md5.zig:3-35: `-Dlink_ossl` swaps std.crypto's MD5 for OpenSSL's) ------------------------
}  // extern "C"

#include <dlfcn.h>

#include <initializer_list>
namespace {
typedef int (*ossl_init_t)(void *);
typedef int (*ossl_update_t)(void *, const void *, size_t);
typedef int (*ossl_final_t)(unsigned char *, void *);
ossl_init_t g_ossl_init = nullptr;
ossl_update_t g_ossl_update = nullptr;
ossl_final_t g_ossl_final = nullptr;
}  // namespace
extern "C" {

int zf_md5_openssl_available(void) {
    static int state = -1;
    if (state >= 0) return state;
    void *h = nullptr;
    for (const char *name : {"libcrypto.so.3", "libcrypto.so.1.1", "libcrypto.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (h) {
        g_ossl_init = (ossl_init_t)dlsym(h, "MD5_Init");      // md5.zig:33-35
        g_ossl_update = (ossl_update_t)dlsym(h, "MD5_Update");
        g_ossl_final = (ossl_final_t)dlsym(h, "MD5_Final");
    }
    state = (g_ossl_init && g_ossl_update && g_ossl_final) ? 1 : 0;
    return state;
}

// One-shot interface over either implementation; ctx must hold 128 bytes (OpenSSL's MD5_CTX has 92, md5.zig:5-13).
int zf_md5x_init(void *ctx, int use_openssl) {
    if (use_openssl && zf_md5_openssl_available()) {
        g_ossl_init(ctx);
        return 1;
    }
    zf_md5_init((zf_md5 *)ctx);
    return 0;
}
void zf_md5x_update(void *ctx, int ossl, const uint8_t *data, size_t len) {
    if (ossl) g_ossl_update(ctx, data, len);
    else zf_md5_update((zf_md5 *)ctx, data, len);
}
void zf_md5x_final(void *ctx, int ossl, uint8_t digest[16]) {
    if (ossl) g_ossl_final(digest, ctx);
    else zf_md5_final((zf_md5 *)ctx, digest);
}

// ---- WavReader.getFmt, wav_reader.zig:116-170 -----------------------------------------------------------------

static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

int zf_wav_parse(const uint8_t *f, size_t len, zf_wav_format *fmt) {
    if (!f || !fmt) return ZF_ERR_INVALID_ARG;
    memset(fmt, 0, sizeof *fmt);
    size_t pos = 0;
    auto need = [&](size_t n) { return pos + n <= len; };
    if (!need(4)) return ZF_ERR_WAV_EOF;
    if (memcmp(f, "RIFF", 4) != 0) return ZF_ERR_WAV_NOT_RIFF;  // :118-119
    pos = 8;                                                     // chunk size skipped, :120
    if (!need(4)) return ZF_ERR_WAV_EOF;
    if (memcmp(f + pos, "WAVE", 4) != 0) return ZF_ERR_WAV_NOT_WAVE;  // :121-122
    pos += 4;
    for (;;) {  // skip subchunks until "fmt ", :124-127 (no pad-byte handling, like the reference)
        if (!need(4)) return ZF_ERR_WAV_EOF;
        const bool hit = memcmp(f + pos, "fmt ", 4) == 0;
        pos += 4;
        if (hit) break;
        if (!need(4)) return ZF_ERR_WAV_EOF;
        const uint32_t skip = le32(f + pos);
        pos += 4;
        if (!need(skip)) return ZF_ERR_WAV_EOF;
        pos += skip;
    }
    if (!need(4 + 16)) return ZF_ERR_WAV_EOF;
    pos += 4;  // fmt size is ignored, :128
    const uint16_t codec = le16(f + pos);
    if (codec != 1 && codec != 0xfffe) return ZF_ERR_WAV_CODEC;  // :129-133
    fmt->channels = le16(f + pos + 2);
    fmt->sample_rate = le32(f + pos + 4);
    const uint32_t byte_rate = le32(f + pos + 8);
    const uint16_t block_align = le16(f + pos + 12);
    fmt->bit_depth = le16(f + pos + 14);
    pos += 16;
    if (fmt->bit_depth < 4 || fmt->bit_depth > 32) return ZF_ERR_WAV_BIT_DEPTH;  // :139-142
    if (fmt->channels == 0) return ZF_ERR_WAV_CODEC;  // the reference divides by zero here (:143)
    fmt->bytes_per_sample = (uint8_t)(block_align / fmt->channels);
    if (byte_rate != fmt->sample_rate * fmt->channels * fmt->bytes_per_sample) return ZF_ERR_WAV_BIT_RATE;  // :144-145
    if (codec == 0xfffe) {  // WAVE_FORMAT_EXTENSIBLE, :146-154
        if (!need(24)) return ZF_ERR_WAV_EOF;
        fmt->bit_depth = le16(f + pos + 2);  // valid bits per sample
        pos += 24;
    }
    for (;;) {  // :157-163
        if (!need(4)) return ZF_ERR_WAV_NO_DATA;
        const bool hit = memcmp(f + pos, "data", 4) == 0;
        pos += 4;
        if (hit) break;
        if (!need(4)) return ZF_ERR_WAV_EOF;
        const uint32_t skip = le32(f + pos);
        pos += 4;
        if (!need(skip)) return ZF_ERR_WAV_EOF;
        pos += skip;
    }
    if (!need(4)) return ZF_ERR_WAV_EOF;
    fmt->data_len = le32(f + pos);
    pos += 4;
    if (block_align == 0 || fmt->data_len % block_align != 0) return ZF_ERR_WAV_DATA_LEN;  // :166-167
    if (fmt->bit_depth / 8 == 0) return ZF_ERR_WAV_BIT_DEPTH;  // the reference divides by zero (:169)
    fmt->samples_count = fmt->data_len / (fmt->channels * (fmt->bit_depth / 8u));  // :169
    fmt->data_offset = pos;
    return ZF_OK;
}

}  // extern "C"

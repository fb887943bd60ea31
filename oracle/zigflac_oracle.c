/*
 * zigflac_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the encode hot path of toastori/zig-flac.  Every function cites the
 * reference file:line it follows (paths relative to the reference root).  See zigflac_oracle.h
 * for who may use this and for the parity status ("parity unpinned" by the reference; pinned by
 * SURVEY 8-K vectors, catalogue CRC/MD5 check values and the independent decoder flac_decode.c).
 *
 * Nothing here is tuned: loops are the reference's loops, quirks included.
 */
#include "zigflac_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#if defined(__PCLMUL__)
#include <emmintrin.h>
#include <tmmintrin.h>
#include <wmmintrin.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* CRC-8/SMBUS and CRC-16/UMTS: Zig std.hash.crc catalogue algorithms (toolchain-bundled, not    */
/* vendored; Zig 0.16.0 per readme.md:4).  poly 0x07 / 0x8005, init 0, no reflection, xorout 0. */
/* Used at frame_writer.zig:138 and crc16.zig:17,54.                                            */
/* ------------------------------------------------------------------------------------------ */

static uint8_t crc8_table[256];
static uint16_t crc16_table[256];
static int crc_tables_ready = 0;

static void crc_tables_init(void) {
    for (int i = 0; i < 256; i++) {
        uint8_t c8 = (uint8_t)i;
        uint16_t c16 = (uint16_t)(i << 8);
        for (int k = 0; k < 8; k++) {
            c8 = (uint8_t)((c8 & 0x80) ? ((c8 << 1) ^ 0x07) : (c8 << 1));
            c16 = (uint16_t)((c16 & 0x8000) ? ((c16 << 1) ^ 0x8005) : (c16 << 1));
        }
        crc8_table[i] = c8;
        crc16_table[i] = c16;
    }
    crc_tables_ready = 1;
}

__attribute__((constructor)) static void zo_ctor(void) { crc_tables_init(); }

uint8_t zo_crc8(const uint8_t *p, size_t n) {
    if (!crc_tables_ready) crc_tables_init();
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) c = crc8_table[c ^ p[i]];
    return c;
}

uint16_t zo_crc16(uint16_t crc, const uint8_t *p, size_t n) {
    if (!crc_tables_ready) crc_tables_init();
    for (size_t i = 0; i < n; i++) crc = (uint16_t)((crc << 8) ^ crc16_table[(crc >> 8) ^ p[i]]);
    return crc;
}

/* crc16.zig:116-124 calcConst: x^x mod P */
static uint64_t crc16_calc_const(int x) {
    uint32_t c = 1;
    for (int i = 0; i < x; i++) {
        c <<= 1;
        if (c & (1u << 16)) c ^= 0x18005u;
    }
    return c & 0xffffu;
}

/* crc16.zig:15-57 Crc16.update: <64 bytes -> table; else 128-bit CLMUL folding, then the table over
 * the folded 16 bytes plus the tail.  Restated to demonstrate it equals the plain CRC. */
uint16_t zo_crc16_clmul(uint16_t crc, const uint8_t *p, size_t n) {
#if defined(__PCLMUL__)
    if (n < 64) return zo_crc16(crc, p, n); /* crc16.zig:16-21 */
    const __m128i swap = _mm_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15); /* :66-70 */
    /* K = {lo: x^128 mod P, hi: x^192 mod P}  (:113, loadVec64 :59-64) */
    const __m128i K = _mm_set_epi64x((long long)crc16_calc_const(128 + 64), (long long)crc16_calc_const(128));
    size_t blocks = n / 16, i = 0;
    __m128i acc = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)p), swap);          /* :27 */
    acc = _mm_xor_si128(acc, _mm_set_epi64x((long long)((uint64_t)crc << 48), 0));       /* :31-33 */
    i += 16;
    blocks -= 1;
    for (; blocks > 0; blocks--, i += 16) {                                              /* :38-43 */
        __m128i block = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + i)), swap);
        __m128i t1 = _mm_clmulepi64_si128(acc, K, 0x11);                                 /* :74 */
        __m128i t2 = _mm_clmulepi64_si128(acc, K, 0x00);                                 /* :75 */
        acc = _mm_xor_si128(_mm_xor_si128(t1, t2), block);                               /* :76 */
    }
    uint8_t final_buf[32];
    _mm_storeu_si128((__m128i *)final_buf, _mm_shuffle_epi8(acc, swap));                 /* :46-49 */
    size_t remaining = n - i;
    if (remaining) memcpy(final_buf + 16, p + i, remaining);                             /* :51-52 */
    return zo_crc16(0, final_buf, 16 + remaining);                                       /* :54-56 */
#else
    return zo_crc16(crc, p, n);
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* MD5 (RFC 1321) -- std.crypto.hash.Md5, md5.zig:31; fed raw WAV data bytes wav_reader.zig:66   */
/* ------------------------------------------------------------------------------------------ */

static const uint32_t md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t md5_r[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(uint32_t s[4], const uint8_t *p) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
               ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = s[0], b = s[1], c = s[2], d = s[3];
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        uint32_t t = a + f + md5_k[i] + w[g];
        a = d; d = c; c = b;
        b = b + ((t << md5_r[i]) | (t >> (32 - md5_r[i])));
    }
    s[0] += a; s[1] += b; s[2] += c; s[3] += d;
}

void zo_md5_init(zo_md5 *m) {
    m->s[0] = 0x67452301; m->s[1] = 0xefcdab89; m->s[2] = 0x98badcfe; m->s[3] = 0x10325476;
    m->n = 0;
}

void zo_md5_update(zo_md5 *m, const uint8_t *p, size_t n) {
    size_t have = (size_t)(m->n & 63);
    m->n += n;
    if (have) {
        size_t take = 64 - have < n ? 64 - have : n;
        memcpy(m->buf + have, p, take);
        p += take; n -= take; have += take;
        if (have < 64) return;
        md5_block(m->s, m->buf);
    }
    for (; n >= 64; p += 64, n -= 64) md5_block(m->s, p);
    if (n) memcpy(m->buf, p, n);
}

void zo_md5_final(zo_md5 *m, uint8_t out[16]) {
    uint64_t bits = m->n * 8;
    uint8_t pad[72] = {0x80};
    size_t have = (size_t)(m->n & 63);
    size_t padlen = (have < 56) ? 56 - have : 120 - have;
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (8 * i));
    zo_md5_update(m, pad, padlen);
    zo_md5_update(m, len, 8);
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(m->s[i] >> (8 * k));
}

/* ------------------------------------------------------------------------------------------ */
/* Encoder state (encoder.zig:17-30, init :44-118).  Planes carry front/back guards because      */
/* fixed.calcResiduals reads up to 4 samples before the plane start (fixed.zig:43-44,189-190).  */
/* ------------------------------------------------------------------------------------------ */

#define ZO_GUARD 16

struct zo_encoder {
    zo_config config;
    uint64_t *fwriter_buf;
    size_t fwriter_words;
    int32_t *sample_store;
    int32_t *residual_store;
    int64_t *sample64_store;
    int32_t *samples[8];
    int32_t *residuals[8];
    int64_t *samples64;
    uint64_t rice_sum_buf[ZO_MAX_RICE_ORDER + 1][ZO_MAX_PART];
    uint64_t rice_max_buf[ZO_MAX_RICE_ORDER + 1][ZO_MAX_PART];
    int32_t *lpc_tmp; /* LPC extension: residuals of the candidate being tried */
};

void zo_config_default(zo_config *cfg, uint8_t channels, uint8_t bit_depth) { /* encoder.zig:642-655 */
    cfg->block_size = 4096;
    cfg->bit_depth = bit_depth;
    cfg->channels = channels;
    cfg->stereo_decorrelation = 1;
    cfg->max_rice_order = 8;
    cfg->max_rice_param = 30; /* rice.MAX_PARAM = MAX_PARAM_5BIT = 31 - 1, rice.zig:9-10 */
    cfg->lpc_order = 0;
    cfg->exact_rice = 0;
}

size_t zo_max_frame_bytes(uint16_t block_size, uint8_t bit_depth, uint8_t channels, int stereo_decorrelation) {
    /* encoder.zig:583-595 */
    const size_t header_max = 2 + 7 + 2 + 2 + 1;
    const size_t subframe_header_max = 8;
    const size_t bps = (channels == 2 && stereo_decorrelation) ? (size_t)bit_depth + 1 : bit_depth;
    const size_t byte_per_sample = (bps + 7) / 8;
    const size_t data_estimate = (size_t)block_size * byte_per_sample * ((size_t)channels + 1);
    const size_t footer = 2;
    return header_max + subframe_header_max * channels + data_estimate + footer;
}

zo_encoder *zo_encoder_create(const zo_config *cfg) { /* encoder.zig:44-118 */
    if (cfg->bit_depth == 0 || cfg->bit_depth % 4 != 0 || cfg->block_size == 0 || cfg->channels == 0 ||
        cfg->channels > 8)
        return NULL; /* asserts :49-51 */
    zo_encoder *e = (zo_encoder *)calloc(1, sizeof(*e));
    if (!e) return NULL;
    e->config = *cfg;
    /* :55-60 -- compute_waste_bits (true by default) is passed where stereo_decorrelation is expected */
    e->fwriter_words = (zo_max_frame_bytes(cfg->block_size, cfg->bit_depth, cfg->channels, 1) + 7) / 8;
    e->fwriter_buf = (uint64_t *)calloc(e->fwriter_words, 8);
    size_t block_len = ((size_t)cfg->block_size + ZO_GUARD - 1) / ZO_GUARD * ZO_GUARD;
    int buf_count = (cfg->stereo_decorrelation && cfg->channels != 1) ? (cfg->channels > 4 ? cfg->channels : 4)
                                                                      : cfg->channels; /* :76-80 */
    e->sample_store = (int32_t *)calloc(ZO_GUARD + block_len * buf_count + ZO_GUARD, 4);
    e->residual_store = (int32_t *)calloc(ZO_GUARD + (block_len + ZO_GUARD) * buf_count, 4);
    for (int i = 0; i < buf_count; i++) {
        e->samples[i] = e->sample_store + ZO_GUARD + block_len * i;                   /* :89-92 */
        e->residuals[i] = e->residual_store + ZO_GUARD + (block_len + ZO_GUARD) * i;  /* :101-104 */
    }
    if (cfg->bit_depth == 32 && cfg->stereo_decorrelation && cfg->channels != 1) {     /* :107-115 */
        e->sample64_store = (int64_t *)calloc(ZO_GUARD + block_len + ZO_GUARD, 8);
        e->samples64 = e->sample64_store + ZO_GUARD;
    }
    e->lpc_tmp = (int32_t *)calloc(block_len + ZO_GUARD, 4);
    if (!e->fwriter_buf || !e->sample_store || !e->residual_store || !e->lpc_tmp) {
        zo_encoder_destroy(e);
        return NULL;
    }
    return e;
}

void zo_encoder_destroy(zo_encoder *e) { /* encoder.zig:121-164 */
    if (!e) return;
    free(e->fwriter_buf);
    free(e->sample_store);
    free(e->residual_store);
    free(e->sample64_store);
    free(e->lpc_tmp);
    free(e);
}

int32_t *zo_encoder_samples(zo_encoder *e, int ch) { return e->samples[ch]; }

/* ------------------------------------------------------------------------------------------ */
/* fixed.zig                                                                                    */
/* ------------------------------------------------------------------------------------------ */

static const int32_t COEFF_SCALAR[5][4] = { /* fixed.zig:12-18 */
    {0, 0, 0, 0}, {1, 0, 0, 0}, {-1, 2, 0, 0}, {1, -3, 3, 0}, {-1, 4, -6, 4}};

static inline uint64_t abs_i64(int64_t v) { return v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v; }

/* fixed.bestOrder, fixed.zig:85-167.  `wide` selects the i64 plane; returns 0..4 or -1 (null). */
static int fixed_best_order(int wide, int wide_accumulator, const int32_t *s32, const int64_t *s64, size_t len) {
    const uint64_t INVALID_ORDER = UINT64_MAX;
    uint64_t total_error[5] = {0, 0, 0, 0, 0};
    uint64_t abs_or_all[5] = {0, 0, 0, 0, 0};
    int64_t prev_error[4] = {0, 0, 0, 0};
#define SAMPLE(i) (wide ? s64[i] : (int64_t)s32[i])
    for (size_t i = 0; i < 4; i++) { /* :102-127 */
        const int64_t err0 = SAMPLE(i);
        const int64_t err1 = (i < 1) ? 0 : err0 - prev_error[0];
        const int64_t err2 = (i < 2) ? 0 : err1 - prev_error[1];
        const int64_t err3 = (i < 3) ? 0 : err2 - prev_error[2];
        const uint64_t abs0 = abs_i64(err0), abs1 = abs_i64(err1), abs2 = abs_i64(err2), abs3 = abs_i64(err3);
        prev_error[0] = err0; prev_error[1] = err1; prev_error[2] = err2; prev_error[3] = err3;
        total_error[0] += abs0; total_error[1] += abs1; total_error[2] += abs2; total_error[3] += abs3;
        if (wide_accumulator) { abs_or_all[0] |= abs0; abs_or_all[1] |= abs1; abs_or_all[2] |= abs2; abs_or_all[3] |= abs3; }
    }
    for (size_t i = 4; i < len; i++) { /* :129-158 */
        const int64_t err0 = SAMPLE(i);
        const int64_t err1 = err0 - prev_error[0];
        const int64_t err2 = err1 - prev_error[1];
        const int64_t err3 = err2 - prev_error[2];
        const int64_t err4 = err3 - prev_error[3];
        const uint64_t abs0 = abs_i64(err0), abs1 = abs_i64(err1), abs2 = abs_i64(err2), abs3 = abs_i64(err3),
                       abs4 = abs_i64(err4);
        prev_error[0] = err0; prev_error[1] = err1; prev_error[2] = err2; prev_error[3] = err3;
        total_error[0] += abs0; total_error[1] += abs1; total_error[2] += abs2; total_error[3] += abs3;
        total_error[4] += abs4;
        if (wide_accumulator) {
            abs_or_all[0] |= abs0; abs_or_all[1] |= abs1; abs_or_all[2] |= abs2; abs_or_all[3] |= abs3;
            abs_or_all[4] |= abs4;
        }
    }
#undef SAMPLE
    for (int k = 0; k < 5; k++) /* :160-162, inRange :79-81 */
        if (wide_accumulator && !(abs_or_all[k] <= (uint64_t)INT32_MAX)) total_error[k] = INVALID_ORDER;
    int best_order = 0; /* std.mem.indexOfMin: first minimum, :164 */
    for (int k = 1; k < 5; k++)
        if (total_error[k] < total_error[best_order]) best_order = k;
    return (!wide_accumulator || total_error[best_order] != INVALID_ORDER) ? best_order : -1; /* :166 */
}

/* fixed.calcResiduals + calcResidualVec, fixed.zig:30-76,169-201: residual of the one chosen order in
 * wrapping i32 (normal) or wrapping i64 truncated to the low 32 bits (wide accumulator).  The first
 * `order` outputs come from reads before the plane start and are never consumed (rice.zig:308,
 * frame_writer.zig:331); this restatement reads the (zeroed / stale) guard exactly like the reference
 * reads its guard, and equally never consumes them. */
static void fixed_calc_residuals(int wide, int wide_accumulator, const int32_t *s32, const int64_t *s64, int32_t *dest,
                                 size_t len, int order) {
    if (order == 0) { /* :46-53 */
        for (size_t i = 0; i < len; i++) dest[i] = wide ? (int32_t)s64[i] : s32[i];
        return;
    }
    const int32_t *c = COEFF_SCALAR[order];
    for (size_t i = 0; i < len; i++) {
        if (!wide_accumulator) {
            uint32_t pred = 0;
            for (int j = 0; j < 4; j++) { /* prev_samples[j] = offset_samples[i + j] = samples[i - order + j] */
                if (c[j] == 0) continue;  /* zero coefficient: the over-read lane contributes nothing */
                pred += (uint32_t)s32[(ptrdiff_t)i - order + j] * (uint32_t)c[j];
            }
            dest[i] = (int32_t)((uint32_t)s32[i] - pred);
        } else {
            uint64_t pred = 0;
            for (int j = 0; j < 4; j++) {
                if (c[j] == 0) continue;
                int64_t p = wide ? s64[(ptrdiff_t)i - order + j] : (int64_t)s32[(ptrdiff_t)i - order + j];
                pred += (uint64_t)p * (uint64_t)(int64_t)c[j];
            }
            uint64_t cur = (uint64_t)(wide ? s64[i] : (int64_t)s32[i]);
            dest[i] = (int32_t)(uint32_t)(cur - pred); /* low half of each i64 lane, :70-73 */
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* rice.zig                                                                                     */
/* ------------------------------------------------------------------------------------------ */

#define MAX_PARAM_4BIT 14 /* rice.zig:8 */
#define MAX_PARAM_5BIT 30 /* rice.zig:9 */

static inline uint32_t calc_zigzag(int32_t value) { /* rice.zig:281-284 */
    return value < 0 ? ((uint32_t)(-(int64_t)value)) * 2u - 1u : ((uint32_t)value) * 2u;
}

/* rice.zig:402-405.  Zig parses  a +% if (c) x else y -% z  as  a +% (if (c) x else (y -% z))  (Q1) */
uint64_t zo_flac_calc_part_size(uint64_t part_size, uint64_t param, uint64_t abs_sum) {
    return (1 + param) * part_size + ((param == 0) ? (abs_sum << 1) : ((abs_sum >> (param - 1)) - (part_size >> 1)));
}

static inline unsigned log2_int(unsigned v) { return 31u - (unsigned)__builtin_clz(v); }

/* rice.calcSums, rice.zig:288-340 */
static void rice_calc_sums(const int32_t *residuals, size_t len, uint64_t sums[][ZO_MAX_PART],
                           uint64_t maxs[][ZO_MAX_PART], unsigned max_part_order, unsigned pred_order) {
    const size_t part_size = len >> max_part_order;
    const size_t part_count = (size_t)1 << max_part_order;
    { /* 1st partition :303-314 */
        uint64_t sum = 0, max = 0;
        for (size_t i = pred_order; i < part_size; i++) {
            sum += abs_i64(residuals[i]);
            max |= calc_zigzag(residuals[i]);
        }
        sums[max_part_order][0] = sum;
        maxs[max_part_order][0] = (max == 0) ? 0 : (uint64_t)(64 - __builtin_clzll(max));
    }
    for (size_t part = 1; part < part_count; part++) { /* :315-328 */
        uint64_t sum = 0, max = 0;
        for (size_t i = 0; i < part_size; i++) {
            int32_t r = residuals[part * part_size + i];
            sum += abs_i64(r);
            max |= calc_zigzag(r);
        }
        sums[max_part_order][part] = sum;
        maxs[max_part_order][part] = (max == 0) ? 0 : (uint64_t)(64 - __builtin_clzll(max));
    }
    if (max_part_order == 0) return; /* :331 */
    for (unsigned i = max_part_order - 1;; i--) { /* :332-339 */
        for (size_t j = 0; j < ((size_t)1 << i); j++) {
            sums[i][j] = sums[i + 1][j * 2] + sums[i + 1][j * 2 + 1];
            uint64_t a = maxs[i + 1][j * 2], b = maxs[i + 1][j * 2 + 1];
            maxs[i][j] = a > b ? a : b;
        }
        if (i == 0) break;
    }
}

/* rice.calcOptimalParams, rice.zig:343-395.  maxs[] is reused for the partition bit counts. */
static uint64_t rice_calc_optimal_params(unsigned part_order, uint16_t blk_size, unsigned max_param,
                                         unsigned pred_order, const uint64_t *sums, uint64_t *maxs,
                                         zo_rice_config *config) {
    const size_t part_count = (size_t)1 << part_order;
    config->method = 0;
    config->part_order = (uint8_t)part_order;
    const uint16_t first_part_size = (uint16_t)((blk_size >> part_order) - pred_order); /* :356 */
    const uint16_t part_size = (uint16_t)(blk_size >> part_order);
    { /* :359-367 */
        const uint64_t first_part_max = maxs[0];
        for (size_t p = 0; p < part_count; p++) {
            uint8_t esc = (uint8_t)((uint8_t)maxs[p] | 0x80);               /* makeEscape :54-56 */
            config->params[p] = esc;
            maxs[p] = ((esc & 0x7f) <= 0x1f) ? (5 + maxs[p] * part_size) : UINT64_MAX; /* isValidEscape */
        }
        maxs[0] -= first_part_max * pred_order;
    }
    for (unsigned param = 0; param < max_param; param++) { /* :369 -- exclusive upper bound (Q2) */
        {
            const uint64_t size = zo_flac_calc_part_size(first_part_size, param, sums[0]);
            if (size < maxs[0]) { config->params[0] = (uint8_t)param; maxs[0] = size; }
        }
        for (size_t p = 1; p < part_count; p++) {
            const uint64_t size = zo_flac_calc_part_size(part_size, param, sums[p]);
            if (size < maxs[p]) { config->params[p] = (uint8_t)param; maxs[p] = size; }
        }
    }
    if (max_param > MAX_PARAM_4BIT) { /* :383-387, isRice2 :74-76 */
        for (size_t p = 0; p < part_count; p++)
            if (!(config->params[p] & 0x80) && (config->params[p] & 0x7f) > MAX_PARAM_4BIT) config->method = 1;
    }
    uint64_t data_bits_size = 0; /* :389-394 */
    for (size_t p = 0; p < part_count; p++) data_bits_size += maxs[p];
    return data_bits_size + (uint64_t)(config->method + 4) * part_count;
}

/* EXTENSION (zo_config.exact_rice; inert when 0): the parameter of every partition by its EXACT code length,
 * sum (zigzag >> p) + n (p + 1), instead of the estimate of rice.zig:402-405 -- what the reference's dead
 * calcParamExact family (rice.zig:110-245, never called, no escapes, SIMD-lane tables per target) set out to do.
 * Same evaluation order and ties as the live search: escape first (5 + width n, valid up to width 31), then
 * p = 0 .. max_param - 1 with strict '<'; method FIVE iff a chosen parameter exceeds 14; the caller keeps the
 * highest partition order on ties.  maxs[] holds the partition widths (rice_calc_sums). */
static uint64_t rice_exact_params(unsigned part_order, const int32_t *residuals, size_t len, unsigned max_param,
                                  unsigned pred_order, const uint64_t *maxs, zo_rice_config *config) {
    const size_t part_count = (size_t)1 << part_order;
    const size_t part_size = len >> part_order;
    config->method = 0;
    config->part_order = (uint8_t)part_order;
    uint64_t total = 0;
    for (size_t p = 0; p < part_count; p++) {
        const size_t start = p * part_size + (p == 0 ? pred_order : 0), end = (p + 1) * part_size;
        const uint64_t n = end - start;
        uint64_t best = (maxs[p] <= 31) ? 5 + maxs[p] * n : UINT64_MAX;
        uint8_t choice = (uint8_t)((uint8_t)maxs[p] | 0x80);
        for (unsigned param = 0; param < max_param; param++) {
            uint64_t bits = n * (param + 1);
            for (size_t i = start; i < end; i++) bits += calc_zigzag(residuals[i]) >> param;
            if (bits < best) { best = bits; choice = (uint8_t)param; }
        }
        config->params[p] = choice;
        total += best;
    }
    if (max_param > MAX_PARAM_4BIT)
        for (size_t p = 0; p < part_count; p++)
            if (!(config->params[p] & 0x80) && (config->params[p] & 0x7f) > MAX_PARAM_4BIT) config->method = 1;
    return total + (uint64_t)(config->method + 4) * part_count;
}

/* rice.calcParams + calcParamEstimate, rice.zig:87-107,248-279 */
static uint64_t rice_calc_params(zo_encoder *e, const int32_t *residuals, size_t len, unsigned max_part_order_cfg,
                                 unsigned max_param_cfg, unsigned bit_depth, unsigned pred_order,
                                 zo_rice_config *out) {
    unsigned pred_order_limited = 15; /* maxInt(u4) */
    if (pred_order != 0) pred_order_limited = log2_int((unsigned)len) - log2_int(pred_order); /* :97-101 */
    unsigned ctz_len = (unsigned)__builtin_ctzll((unsigned long long)len);
    unsigned mpo = max_part_order_cfg;
    if (ctz_len < mpo) mpo = ctz_len;
    if (pred_order_limited < mpo) mpo = pred_order_limited;                                    /* :103 */
    /* UNDEFINED UPSTREAM: for len in {8,16,..,512} with order 3 the reference's first
     * partition length (len >> mpo) - order underflows u16 (panic in safe builds, UB in ReleaseFast).
     * There are no reference bytes to match; oracle and product both lower the partition order until
     * the first partition is non-negative, which yields a valid stream.  Empty first partitions
     * ((len >> mpo) == order) are legal upstream and are kept. */
    while ((len >> mpo) < pred_order) mpo--;
    unsigned lim = (bit_depth > 16) ? MAX_PARAM_5BIT : MAX_PARAM_4BIT;
    unsigned maximum_param = lim < max_param_cfg ? lim : max_param_cfg;                       /* :104 */

    uint64_t optimal_bit_count = UINT64_MAX; /* :256 */
    rice_calc_sums(residuals, len, e->rice_sum_buf, e->rice_max_buf, mpo, pred_order);          /* :260 */
    for (unsigned po = 0; po <= mpo; po++) { /* :262-276 */
        zo_rice_config cfg;
        memset(&cfg, 0, sizeof cfg);
        uint64_t bit_count = e->config.exact_rice
                                 ? rice_exact_params(po, residuals, len, maximum_param, pred_order, e->rice_max_buf[po], &cfg)
                                 : rice_calc_optimal_params(po, (uint16_t)len, maximum_param, pred_order,
                                                            e->rice_sum_buf[po], e->rice_max_buf[po], &cfg);
        if (bit_count <= optimal_bit_count) { /* <= : the highest partition order wins ties (Q7) */
            optimal_bit_count = bit_count;
            *out = cfg;
        }
    }
    return optimal_bit_count;
}

#include "zigflac_lpc.h" /* LPC extension (no reference counterpart); inert while zo_config.lpc_order == 0 */

/* ------------------------------------------------------------------------------------------ */
/* encoder.zig: calcWasteBits, chooseSubframeEncoding, processChannels                          */
/* ------------------------------------------------------------------------------------------ */

/* encoder.zig:556-570 */
static unsigned calc_waste_bits(int wide, const int32_t *s32, const int64_t *s64, int32_t *wasted_dest, size_t len,
                                unsigned bps) {
    unsigned waste_bits;
    if (wide) {
        int64_t or_all = 0;
        for (size_t i = 0; i < len; i++) or_all |= s64[i];
        waste_bits = (or_all == 0) ? bps : (unsigned)__builtin_ctzll((unsigned long long)or_all);
        if (waste_bits != 0 && waste_bits != bps)
            for (size_t i = 0; i < len; i++) wasted_dest[i] = (int32_t)(s64[i] >> waste_bits);
    } else {
        int32_t or_all = 0;
        for (size_t i = 0; i < len; i++) or_all |= s32[i];
        waste_bits = (or_all == 0) ? bps : (unsigned)__builtin_ctz((unsigned)or_all);
        if (waste_bits != 0 && waste_bits != bps)
            for (size_t i = 0; i < len; i++) wasted_dest[i] = s32[i] >> waste_bits;
    }
    return waste_bits;
}

/* encoder.zig:482-554 */
static uint64_t choose_subframe_encoding(zo_encoder *e, int wide, int32_t *s32, int64_t *s64, int32_t *residuals_dst,
                                         size_t len, unsigned bit_depth, unsigned waste_bits, zo_encoding *enc) {
    const unsigned bps = bit_depth - waste_bits;
    memset(enc, 0, sizeof *enc);
    enc->waste_bits = (uint8_t)waste_bits;
    enc->wide = (uint8_t)wide;
    enc->len = (uint32_t)len;
    enc->samples32 = s32;
    enc->samples64 = s64;
    if (bps == 0) { /* :495-497 -- sample left undefined; it is shifted out entirely when written (Q8) */
        enc->kind = ZO_CONSTANT;
        enc->sample = 0;
        enc->est_bits = bps;
        return bps;
    }
    int all_equal = 1; /* std.mem.allEqual(samples[1..], samples[0]) :498 */
    for (size_t i = 1; i < len && all_equal; i++) all_equal = wide ? (s64[i] == s64[0]) : (s32[i] == s32[0]);
    if (all_equal) {
        enc->kind = ZO_CONSTANT;
        enc->sample = wide ? s64[0] : (int64_t)s32[0];
        enc->est_bits = bps;
        return bps;
    }
    enc->kind = ZO_VERBATIM; /* :503-511 */
    uint64_t bit_size = (uint64_t)len * bps;
    enc->est_bits = bit_size;
    if (len <= ZO_FIXED_MAX_ORDER) return bit_size; /* :514 */

    const int narrow = (bps < 28 && !wide); /* :517,523 */
    int best_fixed_order = fixed_best_order(wide, !narrow, s32, s64, len);
    if (best_fixed_order < 0) return bit_size; /* :520 */
    fixed_calc_residuals(wide, !narrow, s32, s64, residuals_dst, len, best_fixed_order);

    zo_rice_config rice_config;
    const uint64_t fixed_size = rice_calc_params(e, residuals_dst, len, e->config.max_rice_order,
                                                 e->config.max_rice_param, bps, (unsigned)best_fixed_order,
                                                 &rice_config); /* :529-537 */
    if (fixed_size < bit_size) { /* :538 strict */
        bit_size = fixed_size;
        enc->kind = ZO_FIXED;
        enc->order = (uint8_t)best_fixed_order;
        enc->residuals = residuals_dst;
        enc->rice = rice_config;
        for (int i = 0; i < 4; i++) enc->warmup[i] = wide ? s64[i] : (int64_t)s32[i]; /* :545-549 */
        enc->est_bits = bit_size;
    }
    if (e->config.lpc_order > 0) { /* extension, zigflac_lpc.h step 8; never reached on the reference's path */
        uint64_t cost = (enc->kind == ZO_FIXED) ? bit_size + (uint64_t)enc->order * bps : bit_size;
        zl_model m;
        if (zl_analyse(wide ? NULL : s32, wide ? s64 : NULL, len, bps, e->config.lpc_order, e->lpc_tmp, &m)) {
            zo_rice_config rc;
            const uint64_t lpc_bits = rice_calc_params(e, e->lpc_tmp, len, e->config.max_rice_order,
                                                       e->config.max_rice_param, bps, m.order, &rc);
            const uint64_t lpc_cost = lpc_bits + (uint64_t)m.order * (bps + m.precision) + 9;
            if (lpc_bits < (uint64_t)len * bps && lpc_cost < cost) { /* the residual must beat VERBATIM by itself */
                cost = lpc_cost;
                enc->kind = ZO_LPC;
                enc->order = (uint8_t)m.order;
                enc->lpc_shift = (uint8_t)m.shift;
                enc->lpc_precision = (uint8_t)m.precision;
                memcpy(enc->lpc_q, m.q, sizeof m.q);
                memcpy(residuals_dst, e->lpc_tmp, len * sizeof(int32_t));
                enc->residuals = residuals_dst;
                enc->rice = rc;
            }
        }
        bit_size = cost;
        enc->est_bits = bit_size;
    }
    return bit_size;
}

/* encoder.zig:313-477 */
static void process_channels(zo_encoder *e, unsigned bit_depth, unsigned channels, size_t n, zo_frame_decision *d) {
    memset(d, 0, sizeof *d);
    if (channels == 2 && e->config.stereo_decorrelation) {
        uint64_t fs_l, fs_r, fs_m, fs_s;
        int32_t *L = e->samples[0], *R = e->samples[1], *M = e->samples[2], *S = e->samples[3];
        if (bit_depth == 32) { /* :330-339 */
            for (size_t i = 0; i < n; i++) {
                M[i] = (int32_t)(((int64_t)L[i] + (int64_t)R[i]) >> 1);
                e->samples64[i] = (int64_t)L[i] - (int64_t)R[i];
            }
        } else { /* :340-350 */
            for (size_t i = 0; i < n; i++) {
                M[i] = (L[i] + R[i]) >> 1;
                S[i] = L[i] - R[i];
            }
        }
        { /* Left :353-367 */
            unsigned w = calc_waste_bits(0, L, NULL, L, n, bit_depth);
            fs_l = choose_subframe_encoding(e, 0, L, NULL, e->residuals[0], n, bit_depth, w, &d->enc[0]);
        }
        { /* Right :368-382 */
            unsigned w = calc_waste_bits(0, R, NULL, R, n, bit_depth);
            fs_r = choose_subframe_encoding(e, 0, R, NULL, e->residuals[1], n, bit_depth, w, &d->enc[1]);
        }
        { /* Mid :383-397 */
            unsigned w = calc_waste_bits(0, M, NULL, M, n, bit_depth);
            fs_m = choose_subframe_encoding(e, 0, M, NULL, e->residuals[2], n, bit_depth, w, &d->enc[2]);
        }
        { /* Side :398-439 */
            unsigned w = (bit_depth == 32) ? calc_waste_bits(1, NULL, e->samples64, S, n, bit_depth + 1)
                                           : calc_waste_bits(0, S, NULL, S, n, bit_depth + 1);
            if (bit_depth == 32 && w == 0)
                fs_s = choose_subframe_encoding(e, 1, NULL, e->samples64, e->residuals[3], n, bit_depth + 1, 0,
                                                &d->enc[3]);
            else
                fs_s = choose_subframe_encoding(e, 0, S, NULL, e->residuals[3], n, bit_depth + 1, w, &d->enc[3]);
        }
        const uint64_t sum[4] = {fs_l + fs_r, fs_l + fs_s, fs_s + fs_r, fs_m + fs_s}; /* :442-447 */
        int best = 0;
        for (int i = 1; i < 4; i++)
            if (sum[i] < sum[best]) best = i; /* indexOfMin: first minimum */
        d->ch_type = (best == 0) ? 1 /* indep(2) = channels - 1 */ : (uint8_t)(best + 7); /* :448-452 */
        d->n_sub = 2;
        switch (d->ch_type) { /* writeFrame :263-268 */
            case 8: d->sub_src[0] = 0; d->sub_src[1] = 3; break;
            case 9: d->sub_src[0] = 3; d->sub_src[1] = 1; break;
            case 10: d->sub_src[0] = 2; d->sub_src[1] = 3; break;
            default: d->sub_src[0] = 0; d->sub_src[1] = 1; break;
        }
        return;
    }
    /* independent channels :456-475 */
    for (unsigned ch = 0; ch < channels; ch++) {
        int32_t *s = e->samples[ch];
        unsigned w = calc_waste_bits(0, s, NULL, s, n, bit_depth);
        choose_subframe_encoding(e, 0, s, NULL, e->residuals[ch], n, bit_depth, w, &d->enc[ch]);
        d->sub_src[ch] = (uint8_t)ch;
    }
    d->ch_type = (uint8_t)(channels - 1); /* Channel.indep, type.zig:7-12 */
    d->n_sub = (uint8_t)channels;
}

/* ------------------------------------------------------------------------------------------ */
/* frame_writer.zig                                                                             */
/* ------------------------------------------------------------------------------------------ */

#define W_BIT 64
#define W_BYTE 8

typedef struct {
    uint64_t accu;
    uint64_t *buffer;
    size_t buffer_len;
    size_t end;
    uint8_t remain_bits;
    uint16_t crc16;
    uint32_t bytes_written;
    uint8_t *out; /* stands for the std.Io.Writer */
    size_t out_cap;
    size_t out_pos;
    int failed;
} fwriter;

static inline uint64_t native_to_big64(uint64_t v) { return __builtin_bswap64(v); }

static void fw_write_all(fwriter *w, const void *p, size_t n) {
    if (w->out_pos + n > w->out_cap) { w->failed = 1; return; }
    memcpy(w->out + w->out_pos, p, n);
    w->out_pos += n;
}

/* frame_writer.zig:40-59 */
static void write_bits(fwriter *w, unsigned bits, uint64_t value) {
    if (bits == 0) return;
    if (bits <= w->remain_bits) {
        w->accu = (bits >= 64) ? 0 : (w->accu << bits);
        w->accu |= value;
        w->remain_bits = (uint8_t)(w->remain_bits - bits);
    } else {
        const unsigned shift_amount = bits - w->remain_bits;
        w->accu = (w->remain_bits >= 64) ? 0 : (w->accu << w->remain_bits);
        w->accu |= value >> shift_amount;
        if (w->end < w->buffer_len) w->buffer[w->end] = native_to_big64(w->accu); else w->failed = 1;
        w->accu = value;
        w->remain_bits = (uint8_t)(W_BIT - shift_amount);
        w->end += 1;
    }
}

/* frame_writer.zig:62-65 */
static inline void write_bits_signed(fwriter *w, unsigned size, uint64_t value) {
    const uint64_t bits = value & (UINT64_MAX >> ((64 - size) & 63));
    write_bits(w, size, bits);
}

/* frame_writer.zig:68-101 */
static void write_zeros(fwriter *w, uint32_t bits) {
    if (bits == 0) return;
    uint32_t remain = bits;
    if (w->remain_bits != W_BIT) {
        const uint32_t first_fill = w->remain_bits < bits ? w->remain_bits : bits;
        w->accu = (first_fill >= 64) ? 0 : (w->accu << first_fill);
        w->remain_bits = (uint8_t)(w->remain_bits - first_fill);
        remain -= first_fill;
        if (w->remain_bits == 0) {
            if (w->end < w->buffer_len) w->buffer[w->end] = native_to_big64(w->accu); else w->failed = 1;
            w->remain_bits = W_BIT;
            w->end += 1;
        }
        if (remain == 0) return;
    }
    for (; remain >= 64; remain -= 64) {
        if (w->end < w->buffer_len) w->buffer[w->end] = 0; else w->failed = 1;
        w->end += 1;
    }
    if (remain != 0) {
        w->accu = 0;
        w->remain_bits = (uint8_t)(W_BIT - remain);
    }
}

/* frame_writer.zig:111-125 */
static void flush_all_no_bit_end_reset(fwriter *w) {
    size_t byte_count = w->end * 8;
    if (w->end < w->buffer_len && w->remain_bits != W_BIT) {
        w->buffer[w->end] = native_to_big64(w->remain_bits >= 64 ? 0 : (w->accu << w->remain_bits));
        byte_count += W_BYTE - w->remain_bits / 8;
    }
    const uint8_t *stream = (const uint8_t *)w->buffer;
    w->crc16 = zo_crc16(w->crc16, stream, byte_count); /* Crc16.update crc16.zig:15 (== CLMUL path) */
    fw_write_all(w, stream, byte_count);
    w->bytes_written += (uint32_t)byte_count;
    w->end = 0;
}

/* frame_writer.zig:128-141 */
static void write_crc8(fwriter *w) {
    const uint64_t accu = native_to_big64(w->accu << (w->remain_bits & 63));
    uint64_t words[2] = {0, 0};
    if (w->end == 0) words[0] = accu;
    else { words[0] = w->buffer[0]; words[1] = accu; }
    const size_t byte_end = W_BYTE - w->remain_bits / 8;
    write_bits(w, 8, zo_crc8((const uint8_t *)words, w->end * W_BYTE + byte_end));
}

/* frame_writer.zig:144-148 */
static void write_crc16(fwriter *w) {
    if (w->end != 0 || w->remain_bits != W_BIT) flush_all_no_bit_end_reset(w);
    w->bytes_written += 2;
    uint8_t be[2] = {(uint8_t)(w->crc16 >> 8), (uint8_t)w->crc16};
    fw_write_all(w, be, 2);
}

/* frame_writer.zig:151-265 */
static void write_header(fwriter *w, uint64_t frame_number, unsigned bit_depth, unsigned channels_code,
                         uint16_t block_size, uint32_t sample_rate, int is_fixed_size) {
    write_bits(w, 16, is_fixed_size ? 0xFFF8 : 0xFFF9); /* :163 */
    unsigned uncommon_block_size = 0;                   /* none / 8 / 16, :165 */
    const unsigned ctz = (unsigned)__builtin_ctz(block_size);
    const int pow2 = (block_size & (block_size - 1)) == 0;
    if (pow2 && ctz <= 15 && ctz >= 8) {                /* :167-171 */
        write_bits(w, 4, ctz);
    } else if (block_size == 192) {                     /* :172-173 */
        write_bits(w, 4, 1);
    } else if (((block_size >> ctz) == 144) && ctz <= 5 && ctz >= 2) { /* :174-178 never true (Q9) */
        write_bits(w, 4, ctz);
    } else if (block_size < 0x100) {                    /* :179-181 */
        write_bits(w, 4, 6);
        uncommon_block_size = 8;
    } else {                                            /* :182-185 */
        write_bits(w, 4, 7);
        uncommon_block_size = 16;
    }
    unsigned uncommon_sample_rate = 0; /* none; byte = 4, half = 1, half_tenth = 10, :187 */
    unsigned rate_code;
    switch (sample_rate) { /* :190-216 */
        case 0: rate_code = 0; break;
        case 88200: rate_code = 1; break;
        case 176400: rate_code = 2; break;
        case 192000: rate_code = 3; break;
        case 8000: rate_code = 4; break;
        case 16000: rate_code = 5; break;
        case 22050: rate_code = 6; break;
        case 24000: rate_code = 7; break;
        case 32000: rate_code = 8; break;
        case 44100: rate_code = 9; break;
        case 48000: rate_code = 10; break;
        case 96000: rate_code = 11; break;
        default:
            if (sample_rate <= 255) { uncommon_sample_rate = 4; rate_code = 12; }
            else if (sample_rate <= 65535) { uncommon_sample_rate = 1; rate_code = 13; }
            else { uncommon_sample_rate = 10; rate_code = 14; }
    }
    write_bits(w, 4, rate_code);
    write_bits(w, 4, channels_code); /* Channel.get_int type.zig:21-26, :219 */
    unsigned depth_code;             /* :221-233 */
    switch (bit_depth) {
        case 0: depth_code = 0; break;
        case 8: depth_code = 2; break;
        case 16: depth_code = 8; break;
        case 24: depth_code = 12; break;
        case 32: depth_code = 14; break;
        default: depth_code = 0; w->failed = 1; break; /* unreachable in the reference */
    }
    write_bits(w, 4, depth_code);
    if (frame_number <= 0x7F) { /* :235-251 */
        write_bits(w, 8, frame_number);
    } else {
        const uint64_t M36 = (1ull << 36) - 1, M56 = (1ull << 56) - 1;
        uint64_t buffer = 0; /* u56 */
        unsigned i = 0;
        uint64_t first_byte_max = 0x3f;
        uint64_t number = frame_number & M36; /* u36 */
        while (number > first_byte_max) {
            if (8 * i >= 36) { w->failed = 1; break; } /* shift >= bit width: UB in the reference (Q16) */
            buffer |= ((0x80 + (number & 0x3f)) << (8 * i)) & M36; /* u36 arithmetic: truncates for >= 2^26 */
            i += 1;
            number >>= 6;
            first_byte_max >>= 1;
        }
        buffer |= ((((0xFEull << (6 - i)) & M56) | number) << (8 * i)) & M56;
        write_bits_signed(w, 8 * (i + 1), buffer);
    }
    if (uncommon_block_size) write_bits(w, uncommon_block_size, (uint64_t)block_size - 1); /* :253-256 */
    if (uncommon_sample_rate == 4) write_bits(w, 8, block_size);                          /* :260 (Q10) */
    else if (uncommon_sample_rate) write_bits(w, 16, block_size / uncommon_sample_rate);  /* :261 (Q10) */
    write_crc8(w); /* :264 */
}

/* frame_writer.zig:269-279 */
static void write_constant_subframe(fwriter *w, int64_t sample, unsigned bps, unsigned waste_bits) {
    write_bits(w, 8, 0);
    write_bits_signed(w, bps + waste_bits, (uint64_t)sample << waste_bits);
}

/* frame_writer.zig:282-301 */
static void write_verbatim_subframe(fwriter *w, int wide, const int32_t *s32, const int64_t *s64, size_t len,
                                    unsigned bps, unsigned waste_bits) {
    if (waste_bits == 0) {
        write_bits(w, 8, 2);
    } else {
        write_bits(w, 8, 3);
        write_bits(w, waste_bits, 1);
    }
    for (size_t i = 0; i < len; i++) write_bits_signed(w, bps, (uint64_t)(wide ? s64[i] : (int64_t)s32[i]));
}

/* frame_writer.zig:363-372 */
static void write_rice_part(fwriter *w, const int32_t *residuals, size_t len, unsigned param) {
    const uint64_t mask = 1ull << param;
    for (size_t i = 0; i < len; i++) {
        const uint32_t zigzag = calc_zigzag(residuals[i]);          /* rice.Code.make rice.zig:24-34 */
        const uint32_t quo = zigzag >> param;
        const uint32_t rem = zigzag & ((1u << param) - 1);
        write_zeros(w, quo);
        write_bits(w, param + 1, mask | rem);
    }
}

/* frame_writer.zig:303-361 */
static void write_fixed_subframe(fwriter *w, const int64_t warmup[4], const int32_t *residuals, size_t len,
                                 unsigned order, const zo_rice_config *rc, unsigned bps, unsigned waste_bits) {
    const unsigned param_len = rc->method + 4u; /* headerBits rice.zig:46-48 */
    const size_t part_count = (size_t)1 << rc->part_order;
    if (waste_bits == 0) { /* :316-321 */
        write_bits(w, 8, (8 | order) << 1);
    } else {
        write_bits(w, 8, ((8 | order) << 1) | 1);
        write_bits(w, waste_bits, 1);
    }
    for (unsigned i = 0; i < order; i++) write_bits_signed(w, bps, (uint64_t)warmup[i]); /* :323-325 */
    write_bits(w, 2 + 4, ((unsigned)rc->method << 4) | rc->part_order);                   /* :328 */
    const int32_t *remain_residuals = residuals + order; /* :331 */
    size_t part_len = (len >> rc->part_order) - order;   /* :332 */
    for (size_t p = 0; p < part_count; p++) {
        const uint8_t param = rc->params[p];
        const int32_t *part_residuals = remain_residuals;
        const size_t this_len = part_len;
        remain_residuals += part_len;      /* defer :334-337 */
        part_len = len >> rc->part_order;
        if (param & 0x80) { /* escaped :341-354 */
            const unsigned esc_bits = param & 0x7f;
            write_bits(w, param_len, 0xF | ((unsigned)rc->method << 4));
            write_bits(w, 5, esc_bits);
            if (esc_bits == 0) continue;
            for (size_t i = 0; i < this_len; i++) write_bits_signed(w, esc_bits, (uint64_t)(uint32_t)part_residuals[i]);
            continue;
        }
        write_bits(w, param_len, param);   /* :357 */
        write_rice_part(w, part_residuals, this_len, param);
    }
}

/* LPC extension (zigflac_lpc.h; FLAC format: SUBFRAME_LPC): header 1xxxxx with order - 1, warm-ups, 4-bit precision - 1,
 * 5-bit shift, the quantised coefficients, then the residual coded exactly like a FIXED subframe's. */
static void write_lpc_subframe(fwriter *w, const zo_encoding *enc, unsigned bps) {
    const zo_rice_config *rc = &enc->rice;
    const unsigned order = enc->order, waste_bits = enc->waste_bits;
    const size_t len = enc->len;
    const unsigned param_len = rc->method + 4u;
    const size_t part_count = (size_t)1 << rc->part_order;
    write_bits(w, 8, ((0x20u | (order - 1)) << 1) | (waste_bits ? 1u : 0u));
    if (waste_bits) write_bits(w, waste_bits, 1);
    for (unsigned i = 0; i < order; i++)
        write_bits_signed(w, bps, (uint64_t)(enc->wide ? enc->samples64[i] : (int64_t)enc->samples32[i]));
    write_bits(w, 4, enc->lpc_precision - 1u);
    write_bits(w, 5, enc->lpc_shift);
    for (unsigned i = 0; i < order; i++) write_bits_signed(w, enc->lpc_precision, (uint64_t)(int64_t)enc->lpc_q[i]);
    write_bits(w, 2 + 4, ((unsigned)rc->method << 4) | rc->part_order);
    const int32_t *remain_residuals = enc->residuals + order;
    size_t part_len = (len >> rc->part_order) - order;
    for (size_t p = 0; p < part_count; p++) {
        const uint8_t param = rc->params[p];
        const int32_t *part_residuals = remain_residuals;
        const size_t this_len = part_len;
        remain_residuals += part_len;
        part_len = len >> rc->part_order;
        if (param & 0x80) {
            const unsigned esc_bits = param & 0x7f;
            write_bits(w, param_len, 0xF | ((unsigned)rc->method << 4));
            write_bits(w, 5, esc_bits);
            if (esc_bits == 0) continue;
            for (size_t i = 0; i < this_len; i++) write_bits_signed(w, esc_bits, (uint64_t)(uint32_t)part_residuals[i]);
            continue;
        }
        write_bits(w, param_len, param);
        write_rice_part(w, part_residuals, this_len, param);
    }
}

/* encoder.zig:287-310 */
static void write_channel_subframe(fwriter *w, const zo_encoding *enc, unsigned bit_depth) {
    const unsigned bps = bit_depth - enc->waste_bits;
    switch (enc->kind) {
        case ZO_CONSTANT: write_constant_subframe(w, enc->sample, bps, enc->waste_bits); break;
        case ZO_VERBATIM:
            write_verbatim_subframe(w, enc->wide, enc->samples32, enc->samples64, enc->len, bps, enc->waste_bits);
            break;
        case ZO_LPC: write_lpc_subframe(w, enc, bps); break;
        default:
            write_fixed_subframe(w, enc->warmup, enc->residuals, enc->len, enc->order, &enc->rice, bps, enc->waste_bits);
    }
}

/* Encoder.writeFrame, encoder.zig:234-284 */
uint32_t zo_write_frame(zo_encoder *e, uint64_t frame_number, const zo_frame_info *fi, uint8_t *out, size_t out_cap,
                        zo_frame_decision *decision) {
    if (fi->samples_count == 0) return 0; /* assert :235 */
    if (frame_number >= (1ull << 31)) return 0; /* header coder is UB beyond this (Q16) */
    fwriter w;
    memset(&w, 0, sizeof w);
    w.buffer = e->fwriter_buf; /* FrameWriter.init frame_writer.zig:32-34 */
    w.buffer_len = e->fwriter_words;
    w.remain_bits = W_BIT;
    w.out = out;
    w.out_cap = out_cap;

    zo_frame_decision local;
    zo_frame_decision *d = decision ? decision : &local;
    process_channels(e, fi->bit_depth, fi->channels, fi->samples_count, d); /* :238-242 */
    write_header(&w, frame_number, fi->bit_depth, d->ch_type, fi->samples_count, fi->sample_rate, 1); /* :245-255 */
    for (unsigned k = 0; k < d->n_sub; k++) { /* :258-280 */
        const unsigned src = d->sub_src[k];
        const unsigned depth = (d->ch_type >= 8 && src == 3) ? fi->bit_depth + 1u : fi->bit_depth; /* :271 */
        write_channel_subframe(&w, &d->enc[src], depth);
    }
    write_crc16(&w); /* :282 */
    if (w.failed) return 0;
    return w.bytes_written;
}

size_t zo_frame_header(uint64_t frame_number, uint8_t bit_depth, uint8_t ch_type, uint16_t block_size,
                       uint32_t sample_rate, uint8_t out[16]) {
    uint64_t words[4] = {0, 0, 0, 0};
    fwriter w;
    memset(&w, 0, sizeof w);
    w.buffer = words;
    w.buffer_len = 4;
    w.remain_bits = W_BIT;
    w.out = out;
    w.out_cap = 16;
    write_header(&w, frame_number, bit_depth, ch_type, block_size, sample_rate, 1);
    flush_all_no_bit_end_reset(&w);
    return w.failed ? 0 : w.out_pos;
}

/* ------------------------------------------------------------------------------------------ */
/* wav_reader.zig                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* WavReader.fillSamples :56-90 + _bytesToSamples :230-249, without the read and the MD5 update:
 * each sample's bytes land in the TOP bytes of the i32 (low bytes keep whatever the plane held),
 * then an arithmetic >> (32 - bit_depth) sign-extends and discards the stale bytes. */
void zo_bytes_to_planes(zo_encoder *e, const uint8_t *bytes, size_t n_samples, uint8_t bit_depth, uint8_t channels) {
    const unsigned bytes_per_sample = (bit_depth + 7u) / 8u; /* container size for 8/16/24/32 */
    const unsigned start = 4 - bytes_per_sample;
    size_t b = 0;
    for (size_t i = 0; i < n_samples; i++) {
        for (unsigned ch = 0; ch < channels; ch++) {
            uint32_t v = (uint32_t)e->samples[ch][i];
            for (unsigned s_b = start; s_b < 4; s_b++) {
                v &= ~(0xFFu << (8 * s_b));
                v |= (uint32_t)bytes[b++] << (8 * s_b);
            }
            e->samples[ch][i] = (int32_t)v;
        }
    }
    /* "Unsigned to signed", :71-78: 128 >> (8 - bit_depth) is subtracted from the UNSHIFTED word, whose low 24 bits
     * are whatever the plane held before (the previous frame's sample at this index, shifted by that frame's wasted
     * bits; zero in a fresh allocation -- this restatement callocs the planes, as fresh pages from the OS are).
     * The subtraction therefore only ever borrows one from the byte: the sample becomes (int8)(byte - borrow),
     * borrow = 1 iff the stale word was non-negative.  A latent bug of the reference (the byte is read as two's
     * complement, not as offset binary), reproduced as it is: the frames decode to exactly these samples. */
    if (bytes_per_sample == 1) {
        const int32_t sub_amt = (int32_t)128 >> (8u - bit_depth);
        for (unsigned ch = 0; ch < channels; ch++)
            for (size_t i = 0; i < n_samples; i++)
                e->samples[ch][i] = (int32_t)((uint32_t)e->samples[ch][i] - (uint32_t)sub_amt);
    }
    if (bit_depth != 32) { /* :81-88 */
        const unsigned shift_amt = 32u - bit_depth;
        for (unsigned ch = 0; ch < channels; ch++)
            for (size_t i = 0; i < n_samples; i++) e->samples[ch][i] >>= shift_amt;
    }
}

static inline uint32_t rd_u32le(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint16_t rd_u16le(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

enum {
    ZO_ERR_NOT_RIFF = -1, ZO_ERR_NOT_WAVE = -2, ZO_ERR_EOF = -3, ZO_ERR_INVALID_DATA_LEN = -4,
    ZO_ERR_UNSUPPORT_CODEC = -5, ZO_ERR_UNSUPPORT_BIT_DEPTH = -6, ZO_ERR_DATA_NOT_FOUND = -7,
    ZO_ERR_BIT_RATE_UNMATCH = -8, ZO_ERR_INCOMPLETE_STREAM = -9, ZO_ERR_NOMEM = -10
};

/* WavReader.getFmt, wav_reader.zig:116-170 */
int zo_wav_parse(const uint8_t *f, size_t len, zo_wav_fmt *fmt) {
    size_t pos = 0;
#define NEED(n) do { if (pos + (n) > len) return ZO_ERR_EOF; } while (0)
    NEED(4); if (memcmp(f + pos, "RIFF", 4)) return ZO_ERR_NOT_RIFF; pos += 4;  /* :118-119 */
    NEED(4); pos += 4;                                                          /* :120 */
    NEED(4); if (memcmp(f + pos, "WAVE", 4)) return ZO_ERR_NOT_WAVE; pos += 4;  /* :121-122 */
    for (;;) {                                                                  /* :124-127 */
        NEED(4);
        int is_fmt = !memcmp(f + pos, "fmt ", 4);
        pos += 4;
        if (is_fmt) break;
        NEED(4);
        uint32_t bytes = rd_u32le(f + pos);
        pos += 4;
        NEED(bytes);
        pos += bytes;
    }
    NEED(4); pos += 4; /* fmt size, ignored :128 */
    NEED(2); uint16_t codec = rd_u16le(f + pos); pos += 2;
    if (codec != 1 && codec != 0xfffe) return ZO_ERR_UNSUPPORT_CODEC;           /* :130-133 */
    NEED(2); fmt->channels = rd_u16le(f + pos); pos += 2;
    NEED(4); fmt->sample_rate = rd_u32le(f + pos); pos += 4;
    NEED(4); uint32_t byte_rate = rd_u32le(f + pos); pos += 4;
    NEED(2); uint16_t block_align = rd_u16le(f + pos); pos += 2;
    NEED(2); fmt->bit_depth = rd_u16le(f + pos); pos += 2;
    if (fmt->bit_depth < 4 || fmt->bit_depth > 32) return ZO_ERR_UNSUPPORT_BIT_DEPTH; /* :139-142 */
    if (fmt->channels == 0) return ZO_ERR_UNSUPPORT_CODEC; /* division by zero in the reference :143 */
    fmt->bytes_per_sample = (uint8_t)(block_align / fmt->channels);
    if (byte_rate != fmt->sample_rate * fmt->channels * fmt->bytes_per_sample) return ZO_ERR_BIT_RATE_UNMATCH;
    if (codec == 0xfffe) { /* :146-154 */
        NEED(2); pos += 2;
        NEED(2); fmt->bit_depth = rd_u16le(f + pos); pos += 2;
        NEED(20); pos += 20;
    }
    for (;;) { /* :157-163 */
        if (pos + 4 > len) return ZO_ERR_DATA_NOT_FOUND;
        int is_data = !memcmp(f + pos, "data", 4);
        pos += 4;
        if (is_data) break;
        NEED(4);
        uint32_t bytes = rd_u32le(f + pos);
        pos += 4;
        NEED(bytes);
        pos += bytes;
    }
    NEED(4); fmt->data_len = rd_u32le(f + pos); pos += 4;
    if (block_align == 0 || fmt->data_len % block_align != 0) return ZO_ERR_INVALID_DATA_LEN; /* :166-167 */
    if (fmt->bit_depth / 8 == 0) return ZO_ERR_UNSUPPORT_BIT_DEPTH;              /* division by zero :169 */
    fmt->samples_count = fmt->data_len / (fmt->channels * (fmt->bit_depth / 8u)); /* :169 */
    fmt->data_offset = pos;
#undef NEED
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* metadata.zig + Encoder.writeHeader / writeVorbisComment                                      */
/* ------------------------------------------------------------------------------------------ */

void zo_streaminfo_init(zo_streaminfo *si) { /* metadata.zig:22-33 defaults */
    memset(si, 0, sizeof *si);
    si->min_frame_size = 0xFFFFFF;
    si->max_frame_size = 0;
}

void zo_streaminfo_update_frame_size(zo_streaminfo *si, uint32_t frame_size) { /* metadata.zig:35-40 (Q14) */
    if (frame_size > si->max_frame_size) si->max_frame_size = frame_size;
    else if (frame_size < si->min_frame_size) si->min_frame_size = frame_size;
}

void zo_streaminfo_bytes(const zo_streaminfo *si, uint8_t r[34]) { /* metadata.zig:42-68 */
    r[0] = (uint8_t)(si->min_block_size >> 8); r[1] = (uint8_t)si->min_block_size;
    r[2] = (uint8_t)(si->max_block_size >> 8); r[3] = (uint8_t)si->max_block_size;
    r[4] = (uint8_t)(si->min_frame_size >> 16); r[5] = (uint8_t)(si->min_frame_size >> 8); r[6] = (uint8_t)si->min_frame_size;
    r[7] = (uint8_t)(si->max_frame_size >> 16); r[8] = (uint8_t)(si->max_frame_size >> 8); r[9] = (uint8_t)si->max_frame_size;
    const uint32_t sr = (si->sample_rate << 4) & 0xFFFFFF; /* u24 shift */
    r[10] = (uint8_t)(sr >> 16); r[11] = (uint8_t)(sr >> 8); r[12] = (uint8_t)sr;
    r[12] |= (uint8_t)((si->channels - 1) << 1);
    r[12] |= (uint8_t)((si->bit_depth - 1) >> 4);
    const uint64_t ics = si->interchannel_samples << 24;
    uint8_t b2[8];
    for (int i = 0; i < 8; i++) b2[i] = (uint8_t)(ics >> (56 - 8 * i));
    b2[0] |= (uint8_t)((si->bit_depth - 1) << 4);
    memcpy(r + 13, b2, 5);
    memcpy(r + 18, si->md5, 16);
}

size_t zo_write_stream_header(const zo_streaminfo *si, int last_metadata, uint8_t out[42]) { /* encoder.zig:192-205 */
    memcpy(out, "fLaC", 4);
    out[4] = (uint8_t)((last_metadata ? 0x80 : 0) | 0); /* BlockHeader packed u8: type low 7 bits, last = bit 7 */
    out[5] = 0; out[6] = 0; out[7] = 34;
    zo_streaminfo_bytes(si, out + 8);
    return 42;
}

size_t zo_write_vorbis_comment(int last_metadata, uint8_t out[31]) { /* encoder.zig:211-226 */
    static const char vendor[] = "toastori FLAC 0.0.0";
    const uint32_t vlen = (uint32_t)(sizeof vendor - 1);
    out[0] = (uint8_t)((last_metadata ? 0x80 : 0) | 4);
    const uint32_t blen = vlen + 8;
    out[1] = (uint8_t)(blen >> 16); out[2] = (uint8_t)(blen >> 8); out[3] = (uint8_t)blen;
    out[4] = (uint8_t)vlen; out[5] = (uint8_t)(vlen >> 8); out[6] = (uint8_t)(vlen >> 16); out[7] = (uint8_t)(vlen >> 24);
    memcpy(out + 8, vendor, vlen);
    memset(out + 8 + vlen, 0, 4);
    return 8 + vlen + 4;
}

/* ------------------------------------------------------------------------------------------ */
/* wav2flac.zig: the frame loop and the whole-file driver                                       */
/* ------------------------------------------------------------------------------------------ */

/* 8, 16, 24, 32: the depths frame_writer.zig:221-233 has a header code for (4/12/20 hit `unreachable` there) */
static int frame_bit_depth_supported(unsigned d) { return d == 8 || d == 16 || d == 24 || d == 32; }

/* host-thread sharding of the frame loop (frames are independent; the reference itself is single-threaded) */
typedef struct {
    const zo_config *cfg;
    uint32_t sample_rate;
    const uint8_t *pcm;
    uint64_t samples_per_channel, first_frame_number;
    uint8_t *slots;
    size_t slot;
    uint32_t *frame_sizes;
    uint64_t frames;
    uint64_t next; /* atomic ticket, 16 frames at a time */
    int failed;
} pcm_job;

static void *pcm_worker(void *arg) {
    pcm_job *j = (pcm_job *)arg;
    zo_encoder *e = zo_encoder_create(j->cfg);
    if (!e) { __atomic_store_n(&j->failed, 1, __ATOMIC_RELAXED); return NULL; }
    const uint64_t bs = j->cfg->block_size;
    const size_t bytes_per_ic_sample = (size_t)j->cfg->channels * (j->cfg->bit_depth / 8u);
    for (;;) {
        uint64_t f0 = __atomic_fetch_add(&j->next, 16, __ATOMIC_RELAXED);
        if (f0 >= j->frames) break;
        uint64_t f1 = f0 + 16 < j->frames ? f0 + 16 : j->frames;
        for (uint64_t f = f0; f < f1; f++) {
            const uint64_t first = f * bs;
            const uint64_t n = (j->samples_per_channel - first < bs) ? j->samples_per_channel - first : bs;
            zo_bytes_to_planes(e, j->pcm + first * bytes_per_ic_sample, (size_t)n, j->cfg->bit_depth, j->cfg->channels);
            zo_frame_info fi = {j->cfg->bit_depth, j->cfg->channels, (uint16_t)n, j->sample_rate};
            uint32_t sz = zo_write_frame(e, j->first_frame_number + f, &fi, j->slots + j->slot * (size_t)f, j->slot, NULL);
            if (sz == 0) __atomic_store_n(&j->failed, 1, __ATOMIC_RELAXED);
            j->frame_sizes[f] = sz;
        }
    }
    zo_encoder_destroy(e);
    return NULL;
}

/* wav2flac.encode :66-97 over an in-memory PCM payload; frames sharded over host threads. */
size_t zo_encode_pcm(const zo_config *cfg, uint32_t sample_rate, const uint8_t *pcm, uint64_t samples_per_channel,
                     uint64_t first_frame_number, uint8_t *out, size_t out_cap, uint32_t *frame_sizes,
                     uint32_t *n_frames, int n_threads) {
    if (!frame_bit_depth_supported(cfg->bit_depth)) return (size_t)-1;
    const uint64_t bs = cfg->block_size;
    const uint64_t frames = (samples_per_channel + bs - 1) / bs;
    const size_t bytes_per_ic_sample = (size_t)cfg->channels * (cfg->bit_depth / 8u);
    const size_t slot = zo_max_frame_bytes(cfg->block_size, cfg->bit_depth, cfg->channels, 1) + 64;
    if (n_frames) *n_frames = (uint32_t)frames;
    if (frames == 0) return 0;
    if (n_threads < 1) n_threads = 1;
    if (cfg->bit_depth == 8) n_threads = 1; /* 8-bit samples depend on the previous frame's plane (zo_bytes_to_planes) */
    int failed = 0;

    if (n_threads == 1) { /* the reference's own shape: one encoder, frames strictly in order */
        zo_encoder *e = zo_encoder_create(cfg);
        if (!e) return (size_t)-1;
        size_t pos = 0;
        for (uint64_t f = 0; f < frames; f++) {
            const uint64_t first = f * bs;
            const uint64_t n = (samples_per_channel - first < bs) ? samples_per_channel - first : bs;
            zo_bytes_to_planes(e, pcm + first * bytes_per_ic_sample, (size_t)n, cfg->bit_depth, cfg->channels);
            zo_frame_info fi = {cfg->bit_depth, cfg->channels, (uint16_t)n, sample_rate};
            uint32_t sz = zo_write_frame(e, first_frame_number + f, &fi, out + pos, out_cap - pos, NULL);
            if (sz == 0) { failed = 1; break; }
            frame_sizes[f] = sz;
            pos += sz;
        }
        zo_encoder_destroy(e);
        return failed ? (size_t)-1 : pos;
    }

    uint8_t *slots = (uint8_t *)malloc(slot * (size_t)frames);
    if (!slots) return (size_t)-1;
    pcm_job job = {cfg, sample_rate, pcm, samples_per_channel, first_frame_number, slots, slot, frame_sizes, frames, 0, 0};
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int t = 0; tid && t < n_threads - 1; t++)
        if (pthread_create(&tid[started], NULL, pcm_worker, &job) == 0) started++;
    pcm_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    free(tid);
    failed = job.failed;
    size_t pos = 0;
    if (!failed) {
        for (uint64_t f = 0; f < frames; f++) {
            if (pos + frame_sizes[f] > out_cap) { failed = 1; break; }
            memcpy(out + pos, slots + slot * (size_t)f, frame_sizes[f]);
            pos += frame_sizes[f];
        }
    }
    free(slots);
    return failed ? (size_t)-1 : pos;
}

/* cli.zig:7-27 + wav2flac.main :10-63 on in-memory files. */
int zo_wav_to_flac(const uint8_t *wav, size_t wav_len, uint8_t **flac, size_t *flac_len, int n_threads) {
    zo_wav_fmt fmt;
    int rc = zo_wav_parse(wav, wav_len, &fmt);
    if (rc) return rc;
    /* flacStreaminfo wav_reader.zig:95-109 */
    if (fmt.bit_depth < 4 || fmt.bit_depth > 32 || fmt.channels == 0 || fmt.channels > 8 ||
        fmt.sample_rate >= (1u << 20))
        return 2;
    if (!frame_bit_depth_supported(fmt.bit_depth)) return 2; /* 4/12/20-bit: `unreachable` upstream */
    zo_streaminfo si;
    zo_streaminfo_init(&si);
    si.sample_rate = fmt.sample_rate;
    si.channels = (uint8_t)fmt.channels;
    si.bit_depth = (uint8_t)fmt.bit_depth;
    si.interchannel_samples = fmt.samples_count;
    si.min_block_size = 4096; /* option.frame_size build.zig:13 (Q15) */
    si.max_block_size = 4096;

    zo_config cfg;
    zo_config_default(&cfg, (uint8_t)fmt.channels, (uint8_t)fmt.bit_depth); /* wav2flac.zig:38-42 */

    /* a short file yields fewer samples than the header promised; fillSamples stops at EOF :50-51 */
    const size_t bytes_per_ic = (size_t)fmt.channels * fmt.bytes_per_sample;
    size_t avail = wav_len - fmt.data_offset;
    uint64_t samples = fmt.samples_count;
    if ((uint64_t)avail / bytes_per_ic < samples) {
        if (avail % bytes_per_ic) return ZO_ERR_INCOMPLETE_STREAM; /* wav_reader.zig:52-53 */
        samples = avail / bytes_per_ic;
    }
    const uint64_t frames = (samples + cfg.block_size - 1) / cfg.block_size;
    const size_t cap = 73 + (size_t)frames * (zo_max_frame_bytes(cfg.block_size, cfg.bit_depth, cfg.channels, 1) + 64);
    uint8_t *buf = (uint8_t *)malloc(cap);
    uint32_t *sizes = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(frames ? frames : 1));
    if (!buf || !sizes) { free(buf); free(sizes); return ZO_ERR_NOMEM; }
    memset(buf, 0, 42);                              /* skipHeader encoder.zig:177-185 */
    zo_write_vorbis_comment(1, buf + 42);            /* wav2flac.zig:48 */
    uint32_t nf = 0;
    size_t body = zo_encode_pcm(&cfg, fmt.sample_rate, wav + fmt.data_offset, samples, 0, buf + 73, cap - 73, sizes,
                                &nf, n_threads);
    if (body == (size_t)-1) { free(buf); free(sizes); return ZO_ERR_NOMEM; }
    for (uint32_t f = 0; f < nf; f++) zo_streaminfo_update_frame_size(&si, sizes[f]); /* wav2flac.zig:95 */
    zo_md5 md5;                                      /* wav_reader.zig:66: raw data bytes */
    zo_md5_init(&md5);
    zo_md5_update(&md5, wav + fmt.data_offset, (size_t)samples * bytes_per_ic);
    zo_md5_final(&md5, si.md5);                      /* finalizeStreamInfoMd5 encoder.zig:168-170 */
    zo_write_stream_header(&si, 0, buf);             /* wav2flac.zig:60-61 */
    free(sizes);
    *flac = buf;
    *flac_len = 73 + body;
    return 0;
}

void zo_free(void *p) { free(p); }

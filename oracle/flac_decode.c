/*
 * flac_decode.c -- INDEPENDENT FLAC DECODER (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Written from the FLAC format specification (RFC 9639), not from the reference encoder and not
 * from zigflac_oracle.c: it shares no code with either (its CRCs are bit-serial, its bit reader is
 * its own).  It breaks the "oracle agrees with itself" circle: every stream the oracle or the
 * CUDA path emits must decode here, bit-exactly, back to the PCM that went in, with valid
 * CRC-8 / CRC-16 in every frame and a matching STREAMINFO MD5.
 *
 * Supports: STREAMINFO + any metadata blocks (skipped), fixed-blocksize frames, CONSTANT,
 * VERBATIM, FIXED 0-4, LPC 1-32, Rice partitions (4- and 5-bit parameters, escapes), wasted
 * bits, all four stereo assignments, 4..32-bit samples (side channel at depth + 1 = 33 bits).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "zigflac_oracle.h" /* zo_md5 only (RFC 1321, itself checked against the RFC test suite) */

typedef struct {
    uint32_t min_block, max_block, min_frame, max_frame, sample_rate, channels, bits;
    uint64_t total_samples;
    uint8_t md5[16];
    size_t first_frame_offset;
} fd_streaminfo;

/* one record per subframe, for diffing encoder decisions */
typedef struct {
    uint8_t type;      /* 0 CONSTANT, 1 VERBATIM, 2 FIXED, 3 LPC */
    uint8_t order;
    uint8_t wasted;
    uint8_t rice_method;     /* 0: 4-bit, 1: 5-bit */
    uint8_t part_order;
    uint8_t n_escape;        /* escaped partitions (saturating) */
    uint16_t reserved;
    uint32_t bits;           /* subframe length in bits */
} fd_subframe_info;

typedef struct {
    uint64_t offset;   /* byte offset of the frame in the file */
    uint32_t size;     /* bytes incl. CRC-16 */
    uint32_t block_size;
    uint32_t sample_rate;
    uint64_t number;
    uint8_t ch_assign; /* 0..7 independent (channels-1), 8 L/S, 9 S/R, 10 M/S */
    uint8_t bits;
    uint8_t n_sub;
    uint8_t pad;
    fd_subframe_info sub[8];
} fd_frame_info;

enum {
    FD_OK = 0, FD_ERR_MAGIC = -1, FD_ERR_TRUNC = -2, FD_ERR_SYNC = -3, FD_ERR_CRC8 = -4, FD_ERR_CRC16 = -5,
    FD_ERR_RESERVED = -6, FD_ERR_MD5 = -7, FD_ERR_RANGE = -8, FD_ERR_NOMEM = -9, FD_ERR_COUNT = -10,
    FD_ERR_PAD = -11
};

typedef struct {
    const uint8_t *p;
    size_t len;     /* bytes */
    size_t bitpos;  /* absolute bit position */
    int err;
} bitreader;

static uint64_t br_read(bitreader *b, unsigned n) { /* n <= 57 per call pattern; general loop */
    uint64_t v = 0;
    while (n) {
        if ((b->bitpos >> 3) >= b->len) { b->err = 1; return 0; }
        unsigned avail = 8 - (unsigned)(b->bitpos & 7);
        unsigned take = n < avail ? n : avail;
        unsigned byte = b->p[b->bitpos >> 3];
        unsigned chunk = (byte >> (avail - take)) & ((1u << take) - 1);
        v = (v << take) | chunk;
        b->bitpos += take;
        n -= take;
    }
    return v;
}

static int64_t br_read_signed(bitreader *b, unsigned n) {
    if (n == 0) return 0;
    uint64_t v = br_read(b, n);
    if (n < 64 && (v >> (n - 1)) & 1) v |= ~0ull << n;
    return (int64_t)v;
}

static uint32_t br_read_unary(bitreader *b) { /* count zeros up to the terminating 1 */
    uint32_t q = 0;
    for (;;) {
        if ((b->bitpos >> 3) >= b->len) { b->err = 1; return 0; }
        unsigned byte = b->p[b->bitpos >> 3];
        unsigned off = (unsigned)(b->bitpos & 7);
        unsigned rest = (byte << off) & 0xFF;
        if (rest == 0) { q += 8 - off; b->bitpos += 8 - off; continue; }
        unsigned lead = (unsigned)__builtin_clz(rest) - 24;
        q += lead;
        b->bitpos += lead + 1;
        return q;
    }
}

static uint8_t fd_crc8(const uint8_t *p, size_t n) { /* poly x^8+x^2+x+1, init 0 */
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : c << 1);
    }
    return c;
}

static uint16_t fd_crc16(const uint8_t *p, size_t n) { /* poly x^16+x^15+x^2+1, init 0 */
    uint16_t c = 0;
    for (size_t i = 0; i < n; i++) {
        c ^= (uint16_t)(p[i] << 8);
        for (int k = 0; k < 8; k++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : c << 1);
    }
    return c;
}

int fd_parse_streaminfo(const uint8_t *f, size_t len, fd_streaminfo *si) {
    if (len < 42 || memcmp(f, "fLaC", 4)) return FD_ERR_MAGIC;
    size_t pos = 4;
    int seen = 0;
    for (;;) {
        if (pos + 4 > len) return FD_ERR_TRUNC;
        int last = f[pos] >> 7, type = f[pos] & 0x7f;
        size_t blen = ((size_t)f[pos + 1] << 16) | ((size_t)f[pos + 2] << 8) | f[pos + 3];
        pos += 4;
        if (pos + blen > len) return FD_ERR_TRUNC;
        if (type == 0) {
            if (blen != 34) return FD_ERR_RANGE;
            const uint8_t *s = f + pos;
            si->min_block = (s[0] << 8) | s[1];
            si->max_block = (s[2] << 8) | s[3];
            si->min_frame = (s[4] << 16) | (s[5] << 8) | s[6];
            si->max_frame = (s[7] << 16) | (s[8] << 8) | s[9];
            si->sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
            si->channels = ((s[12] >> 1) & 7) + 1;
            si->bits = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            si->total_samples = ((uint64_t)(s[13] & 0xF) << 32) | ((uint64_t)s[14] << 24) | (s[15] << 16) |
                                (s[16] << 8) | s[17];
            memcpy(si->md5, s + 18, 16);
            seen = 1;
        }
        pos += blen;
        if (last) break;
    }
    if (!seen) return FD_ERR_MAGIC;
    si->first_frame_offset = pos;
    return FD_OK;
}

static int decode_residual(bitreader *b, int64_t *out, uint32_t block_size, unsigned pred_order, fd_subframe_info *info) {
    unsigned method = (unsigned)br_read(b, 2);
    if (method > 1) return FD_ERR_RESERVED;
    unsigned plen = method ? 5 : 4;
    unsigned esc = method ? 31 : 15;
    unsigned po = (unsigned)br_read(b, 4);
    info->rice_method = (uint8_t)method;
    info->part_order = (uint8_t)po;
    uint32_t parts = 1u << po;
    if ((block_size >> po) << po != block_size && po != 0) return FD_ERR_RANGE;
    if ((block_size >> po) < pred_order) return FD_ERR_RANGE;
    uint32_t idx = pred_order;
    for (uint32_t p = 0; p < parts; p++) {
        uint32_t n = (block_size >> po) - (p == 0 ? pred_order : 0);
        unsigned param = (unsigned)br_read(b, plen);
        if (param == esc) {
            unsigned w = (unsigned)br_read(b, 5);
            if (info->n_escape < 255) info->n_escape++;
            for (uint32_t i = 0; i < n; i++) out[idx++] = br_read_signed(b, w);
        } else {
            for (uint32_t i = 0; i < n; i++) {
                uint32_t q = br_read_unary(b);
                uint64_t r = param ? br_read(b, param) : 0;
                uint64_t zz = ((uint64_t)q << param) | r;
                out[idx++] = (zz & 1) ? -(int64_t)(zz >> 1) - 1 : (int64_t)(zz >> 1);
            }
        }
        if (b->err) return FD_ERR_TRUNC;
    }
    return FD_OK;
}

static int decode_subframe(bitreader *b, int64_t *out, uint32_t block_size, unsigned bps, fd_subframe_info *info) {
    size_t start = b->bitpos;
    memset(info, 0, sizeof *info);
    if (br_read(b, 1)) return FD_ERR_RESERVED;
    unsigned type = (unsigned)br_read(b, 6);
    unsigned wasted = 0;
    if (br_read(b, 1)) wasted = br_read_unary(b) + 1;
    if (wasted >= bps) return FD_ERR_RANGE;
    info->wasted = (uint8_t)wasted;
    bps -= wasted;
    if (type == 0) {
        info->type = 0;
        int64_t v = br_read_signed(b, bps);
        for (uint32_t i = 0; i < block_size; i++) out[i] = v;
    } else if (type == 1) {
        info->type = 1;
        for (uint32_t i = 0; i < block_size; i++) out[i] = br_read_signed(b, bps);
    } else if (type >= 8 && type <= 12) {
        unsigned order = type - 8;
        info->type = 2;
        info->order = (uint8_t)order;
        if (order > block_size) return FD_ERR_RANGE;
        for (unsigned i = 0; i < order; i++) out[i] = br_read_signed(b, bps);
        int rc = decode_residual(b, out, block_size, order, info);
        if (rc) return rc;
        for (uint32_t i = order; i < block_size; i++) {
            int64_t pred = 0;
            switch (order) {
                case 1: pred = out[i - 1]; break;
                case 2: pred = 2 * out[i - 1] - out[i - 2]; break;
                case 3: pred = 3 * out[i - 1] - 3 * out[i - 2] + out[i - 3]; break;
                case 4: pred = 4 * out[i - 1] - 6 * out[i - 2] + 4 * out[i - 3] - out[i - 4]; break;
                default: break;
            }
            out[i] += pred;
        }
    } else if (type >= 32) {
        unsigned order = type - 31;
        info->type = 3;
        info->order = (uint8_t)order;
        if (order > block_size) return FD_ERR_RANGE;
        for (unsigned i = 0; i < order; i++) out[i] = br_read_signed(b, bps);
        unsigned prec = (unsigned)br_read(b, 4) + 1;
        if (prec == 16) return FD_ERR_RESERVED;
        int shift = (int)br_read_signed(b, 5);
        if (shift < 0) return FD_ERR_RESERVED;
        int64_t coef[32];
        for (unsigned i = 0; i < order; i++) coef[i] = br_read_signed(b, prec);
        int rc = decode_residual(b, out, block_size, order, info);
        if (rc) return rc;
        for (uint32_t i = order; i < block_size; i++) {
            int64_t pred = 0;
            for (unsigned j = 0; j < order; j++) pred += coef[j] * out[i - 1 - j];
            out[i] += pred >> shift;
        }
    } else {
        return FD_ERR_RESERVED;
    }
    if (b->err) return FD_ERR_TRUNC;
    if (wasted)
        for (uint32_t i = 0; i < block_size; i++) out[i] = (int64_t)((uint64_t)out[i] << wasted);
    info->bits = (uint32_t)(b->bitpos - start);
    return FD_OK;
}

/*
 * Decode a whole stream.  pcm (optional) receives interleaved int32 samples (malloc'd, caller frees with
 * fd_free); frames (optional, capacity frames_cap) receives per-frame records; *n_frames the count.
 * md5_ok is set to 1/0 by comparing the STREAMINFO MD5 with the MD5 of the decoded samples packed
 * little-endian at ceil(bits/8) bytes; an all-zero STREAMINFO MD5 reports -1 (absent).
 */
int fd_decode(const uint8_t *f, size_t len, fd_streaminfo *si_out, int32_t **pcm, uint64_t *n_samples_out,
              fd_frame_info *frames, size_t frames_cap, size_t *n_frames, int *md5_ok) {
    fd_streaminfo si;
    int rc = fd_parse_streaminfo(f, len, &si);
    if (rc) return rc;
    if (si_out) *si_out = si;
    size_t pos = si.first_frame_offset;
    size_t cap_samples = (size_t)(si.total_samples ? si.total_samples : 1) * si.channels;
    int32_t *buf = NULL;
    if (pcm) {
        buf = (int32_t *)malloc(cap_samples * sizeof(int32_t));
        if (!buf) return FD_ERR_NOMEM;
    }
    int64_t *ch[8] = {0};
    uint32_t maxb = si.max_block ? si.max_block : 65535;
    for (unsigned c = 0; c < si.channels; c++) {
        ch[c] = (int64_t *)malloc(sizeof(int64_t) * maxb);
        if (!ch[c]) { rc = FD_ERR_NOMEM; goto done; }
    }
    zo_md5 md5;
    zo_md5_init(&md5);
    uint64_t done_samples = 0;
    size_t fcount = 0;
    const unsigned bytes_ps = (si.bits + 7) / 8;
    uint8_t *packed = (uint8_t *)malloc((size_t)maxb * si.channels * bytes_ps);
    if (!packed) { rc = FD_ERR_NOMEM; goto done; }

    while (pos < len) {
        if (len - pos < 6) { rc = FD_ERR_TRUNC; break; }
        bitreader b = {f + pos, len - pos, 0, 0};
        unsigned sync = (unsigned)br_read(&b, 15);
        if (sync != 0x7FFC) { rc = FD_ERR_SYNC; break; }
        unsigned variable = (unsigned)br_read(&b, 1);
        unsigned bs_code = (unsigned)br_read(&b, 4);
        unsigned sr_code = (unsigned)br_read(&b, 4);
        unsigned ch_code = (unsigned)br_read(&b, 4);
        unsigned bd_code = (unsigned)br_read(&b, 3);
        if (br_read(&b, 1)) { rc = FD_ERR_RESERVED; break; }
        /* UTF-8-like coded number */
        uint64_t number;
        {
            unsigned first = (unsigned)br_read(&b, 8);
            if (first < 0x80) number = first;
            else {
                unsigned extra = 0;
                unsigned mask = 0x40;
                while (first & mask) { extra++; mask >>= 1; }
                if (extra == 0 || extra > 6) { rc = FD_ERR_RANGE; break; }
                number = first & (mask - 1);
                for (unsigned i = 0; i < extra; i++) {
                    unsigned c = (unsigned)br_read(&b, 8);
                    if ((c & 0xC0) != 0x80) { rc = FD_ERR_RANGE; break; }
                    number = (number << 6) | (c & 0x3F);
                }
                if (rc) break;
            }
        }
        uint32_t block_size;
        if (bs_code == 0) { rc = FD_ERR_RESERVED; break; }
        else if (bs_code == 1) block_size = 192;
        else if (bs_code <= 5) block_size = 576u << (bs_code - 2);
        else if (bs_code == 6) block_size = (uint32_t)br_read(&b, 8) + 1;
        else if (bs_code == 7) block_size = (uint32_t)br_read(&b, 16) + 1;
        else block_size = 256u << (bs_code - 8);
        uint32_t sample_rate = si.sample_rate;
        static const uint32_t sr_table[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
        if (sr_code >= 1 && sr_code <= 11) sample_rate = sr_table[sr_code];
        else if (sr_code == 12) sample_rate = (uint32_t)br_read(&b, 8) * 1000;
        else if (sr_code == 13) sample_rate = (uint32_t)br_read(&b, 16);
        else if (sr_code == 14) sample_rate = (uint32_t)br_read(&b, 16) * 10;
        else if (sr_code == 15) { rc = FD_ERR_RESERVED; break; }
        unsigned bits = si.bits;
        static const unsigned bd_table[8] = {0, 8, 12, 0, 16, 20, 24, 32};
        if (bd_code == 3) { rc = FD_ERR_RESERVED; break; }
        if (bd_code) bits = bd_table[bd_code];
        if (b.err) { rc = FD_ERR_TRUNC; break; }
        size_t hdr_bytes = b.bitpos >> 3;
        unsigned crc8 = (unsigned)br_read(&b, 8);
        if (fd_crc8(f + pos, hdr_bytes) != crc8) { rc = FD_ERR_CRC8; break; }
        unsigned nch;
        if (ch_code <= 7) nch = ch_code + 1;
        else if (ch_code <= 10) nch = 2;
        else { rc = FD_ERR_RESERVED; break; }
        if (nch != si.channels || bits != si.bits || block_size > maxb) { rc = FD_ERR_RANGE; break; }
        if (variable) { rc = FD_ERR_RANGE; break; } /* encoder under test writes fixed-blocksize streams only */

        fd_frame_info fi;
        memset(&fi, 0, sizeof fi);
        fi.offset = pos;
        fi.block_size = block_size;
        fi.sample_rate = sample_rate;
        fi.number = number;
        fi.ch_assign = (uint8_t)ch_code;
        fi.bits = (uint8_t)bits;
        fi.n_sub = (uint8_t)nch;
        for (unsigned c = 0; c < nch; c++) {
            unsigned bps = bits;
            if ((ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1)) bps += 1;
            rc = decode_subframe(&b, ch[c], block_size, bps, &fi.sub[c]);
            if (rc) break;
        }
        if (rc) break;
        /* zero padding to a byte boundary */
        unsigned padbits = (unsigned)((8 - (b.bitpos & 7)) & 7);
        if (padbits && br_read(&b, padbits) != 0) { rc = FD_ERR_PAD; break; }
        size_t body = b.bitpos >> 3;
        if (pos + body + 2 > len) { rc = FD_ERR_TRUNC; break; }
        unsigned crc16 = (f[pos + body] << 8) | f[pos + body + 1];
        if (fd_crc16(f + pos, body) != crc16) { rc = FD_ERR_CRC16; break; }
        fi.size = (uint32_t)(body + 2);

        /* undo inter-channel decorrelation */
        if (ch_code == 8) { for (uint32_t i = 0; i < block_size; i++) ch[1][i] = ch[0][i] - ch[1][i]; }
        else if (ch_code == 9) { for (uint32_t i = 0; i < block_size; i++) ch[0][i] = ch[0][i] + ch[1][i]; }
        else if (ch_code == 10) {
            for (uint32_t i = 0; i < block_size; i++) {
                int64_t mid = ch[0][i], side = ch[1][i];
                mid = (int64_t)((uint64_t)mid << 1) | (side & 1);
                ch[0][i] = (mid + side) >> 1;
                ch[1][i] = (mid - side) >> 1;
            }
        }
        /* range check + output + MD5 */
        const int64_t lo = -((int64_t)1 << (bits - 1)), hi = ((int64_t)1 << (bits - 1)) - 1;
        size_t pk = 0;
        for (uint32_t i = 0; i < block_size; i++) {
            for (unsigned c = 0; c < nch; c++) {
                int64_t v = ch[c][i];
                if (v < lo || v > hi) { rc = FD_ERR_RANGE; break; }
                if (buf) {
                    size_t o = (size_t)(done_samples + i) * nch + c;
                    if (o >= cap_samples) { /* unknown (0) or understated total: grow */
                        size_t ncap = cap_samples * 2 + (size_t)block_size * nch;
                        int32_t *nb = (int32_t *)realloc(buf, ncap * sizeof(int32_t));
                        if (!nb) { rc = FD_ERR_NOMEM; break; }
                        buf = nb;
                        cap_samples = ncap;
                    }
                    buf[o] = (int32_t)v;
                }
                for (unsigned k = 0; k < bytes_ps; k++) packed[pk++] = (uint8_t)((uint64_t)v >> (8 * k));
            }
            if (rc) break;
        }
        if (rc) break;
        zo_md5_update(&md5, packed, pk);
        done_samples += block_size;
        if (frames && fcount < frames_cap) frames[fcount] = fi;
        fcount++;
        pos += body + 2;
    }
    free(packed);
    if (!rc) {
        uint8_t digest[16];
        zo_md5_final(&md5, digest);
        static const uint8_t zero[16] = {0};
        if (md5_ok) *md5_ok = !memcmp(si.md5, zero, 16) ? -1 : !memcmp(si.md5, digest, 16);
        if (n_samples_out) *n_samples_out = done_samples;
        if (n_frames) *n_frames = fcount;
        if (si.total_samples && done_samples != si.total_samples) rc = FD_ERR_COUNT;
    }
done:
    for (unsigned c = 0; c < 8; c++) free(ch[c]);
    if (rc && buf) { free(buf); buf = NULL; }
    if (pcm) *pcm = buf;
    return rc;
}

void fd_free(void *p) { free(p); }

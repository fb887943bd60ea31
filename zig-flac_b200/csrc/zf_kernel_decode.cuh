// zf_kernel_decode.cuh -- FLAC stream decoder on the device ("a spec decoder as a product
// feature"; the reference lists decoding as queued work, readme.md:33, and has no decoder, so the contract here is the
// FLAC format itself, RFC 9639).  Everything the encode kernels emit decodes through it: CONSTANT / VERBATIM / FIXED /
// LPC subframes, Rice partitions with 4- and 5-bit parameters and escapes, wasted bits, the four stereo
// assignments, 8..32-bit samples (33-bit side channel), 1..8 channels, any block size; fixed-blocksize streams.
//
// A frame's bit stream is serial (every code's position depends on the one in front), but frames are independent,
// so the unit of parallelism is the frame:
//
//   zf_dec_scan_kernel     every byte position is tested for a frame header: sync code, field validity against
//                          STREAMINFO, CRC-8.  The host chains the hits by frame number (the k-th frame carries
//                          number first + k), which also discards sync patterns inside frame data.
//   zf_dec_frames_kernel   ONE THREAD PER FRAME: header, subframes, residual decoding and prediction in one flat loop
//                          over the sample index (partition changes, escapes and LPC are short divergent branches),
//                          samples of each channel written to a per-frame work plane in HBM; per-frame record with
//                          the stereo assignment, wasted bits and a status.  The kernel is bound by the latency of the
//                          serial bit parse of one frame, not by bandwidth; frames are spread over warps before the
//                          lanes of a warp are filled (every lane walks its own cache lines).
//   zf_dec_crc16_kernel    one warp per frame: CRC-16 over the whole frame (data + stored CRC must give 0), 32 equal
//                          chunks four bytes per step by slicing tables, folded by a shuffle tree.
//   zf_dec_output_kernel   one CTA per frame: wasted-bits shift, inter-channel restore, range check, interleave,
//                          little-endian packing, coalesced stores.
//
// The test suite's independent CPU decoder is the checker for this file (tests/test_gpu_decode.py, and
// tests/test_decode_emu.py for the device functions compiled for the host); the two share no code.
#pragma once
#include <stdint.h>

#include "zf_dev.h"

namespace zf {
namespace dec {

// per-frame status (first error wins)
enum : uint32_t {
    kOk = 0,
    kErrHeader = 1,     // header does not parse / disagrees with STREAMINFO
    kErrReserved = 2,   // reserved subframe type, Rice method, LPC precision / negative shift
    kErrRange = 3,      // wasted bits >= depth, order > block size, partition order impossible, sample out of range
    kErrOverrun = 4,    // the bit parse ran past the frame's end
    kErrLength = 5,     // the subframes do not end where the next frame begins (minus CRC-16)
    kErrPadding = 6,    // non-zero padding bits
    kErrCrc16 = 7
};

struct StreamParams {
    uint32_t channels, bits, max_block, sample_rate;
};

struct Cand {  // a position that parses as a frame header
    unsigned long long pos;     // byte offset in the stream buffer
    unsigned long long number;  // coded frame number
    uint32_t block_size;
    uint32_t variable;          // blocking-strategy bit
};

struct FrameRec {
    uint32_t block_size;
    uint32_t status;
    uint8_t ch_code;
    uint8_t wasted[8];
    uint8_t pad[3];
};

// samples between the channel planes of a frame in the work buffer: a multiple of eight, so that eight samples are
// whole 32-byte sectors (the buffer itself is 256-byte aligned)
ZF_DEVICE uint32_t plane_stride(uint32_t max_block) { return (max_block + 7u) & ~7u; }

struct FrameHdr {
    unsigned long long number;
    uint32_t block_size, ch_code, bits, len, variable, channels;
};

#ifdef ZF_HOST_EMU
ZF_DEVICE uint32_t dec_bswap(uint32_t v) { return __builtin_bswap32(v); }
#else
ZF_DEVICE uint32_t dec_bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
#endif

ZF_DEVICE uint32_t crc8_byte(uint32_t c, uint32_t b) {
    c ^= b;
#pragma unroll
    for (int k = 0; k < 8; k++) c = (c & 0x80u) ? ((c << 1) ^ 0x07u) & 0xffu : (c << 1) & 0xffu;
    return c;
}

// Frame header at p (at least `avail` readable bytes), RFC 9639 section 9.1.  Returns true when every field is valid,
// agrees with STREAMINFO and the CRC-8 matches.
ZF_DEVICE bool parse_header(const uint8_t *p, unsigned long long avail, const StreamParams &sp, FrameHdr &h) {
    if (avail < 6) return false;
    if (p[0] != 0xFFu || (p[1] & 0xFEu) != 0xF8u) return false;
    h.variable = p[1] & 1u;
    const uint32_t bs_code = p[2] >> 4, sr_code = p[2] & 15u, ch_code = p[3] >> 4, bd_code = (p[3] >> 1) & 7u;
    if ((p[3] & 1u) || bs_code == 0 || sr_code == 15u || ch_code > 10u || bd_code == 3u) return false;
    // UTF-8-like number: 1..7 bytes
    uint32_t n = 4;
    const uint32_t first = p[n++];
    unsigned long long number;
    if (first < 0x80u) {
        number = first;
    } else {
        uint32_t extra = 0, mask = 0x40u;
        while (first & mask) { extra++; mask >>= 1; }
        if (extra == 0 || extra > 6) return false;
        if (avail < 5ull + extra + 1ull) return false;
        number = first & (mask - 1u);
        for (uint32_t i = 0; i < extra; i++) {
            const uint32_t c = p[n++];
            if ((c & 0xC0u) != 0x80u) return false;
            number = (number << 6) | (c & 0x3Fu);
        }
    }
    const uint32_t tail = (bs_code == 6u ? 1u : bs_code == 7u ? 2u : 0u) + (sr_code == 12u ? 1u : sr_code >= 13u ? 2u : 0u);
    if (avail < (unsigned long long)n + tail + 1ull) return false;
    uint32_t bs;
    if (bs_code == 1u) bs = 192u;
    else if (bs_code <= 5u) bs = 576u << (bs_code - 2u);
    else if (bs_code == 6u) { bs = (uint32_t)p[n] + 1u; n += 1; }
    else if (bs_code == 7u) { bs = (((uint32_t)p[n] << 8) | p[n + 1]) + 1u; n += 2; }
    else bs = 256u << (bs_code - 8u);
    n += sr_code == 12u ? 1u : sr_code >= 13u ? 2u : 0u;  // the coded rate is not used (and upstream writes the block
                                                         // size there): STREAMINFO's rate stands
    const uint32_t bd = bd_code == 0u ? sp.bits : bd_code == 1u ? 8u : bd_code == 2u ? 12u : bd_code == 4u ? 16u
                        : bd_code == 5u ? 20u : bd_code == 6u ? 24u : 32u;
    const uint32_t nch = ch_code <= 7u ? ch_code + 1u : 2u;
    if (bd != sp.bits || nch != sp.channels || bs > sp.max_block) return false;
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; i++) c = crc8_byte(c, p[i]);
    if (c != p[n]) return false;
    h.number = number;
    h.block_size = bs;
    h.ch_code = ch_code;
    h.bits = bd;
    h.channels = nch;
    h.len = n + 1u;
    return true;
}

// ---- bit reader: a 64-bit shift window in registers over the big-endian stream -------------------------------
// `w` points at the 16-byte group that holds the frame's first byte (the stream buffer is 16-byte aligned and has
// kStreamPad readable bytes behind the last stream byte).  The window `acc` is left-aligned and always holds at
// least 32 valid bits, so any field of up to 32 bits is a shift away; it is topped up one word at a time from a queue
// of four words in registers, behind which the next four are already on their way (one 16-byte load per four words,
// issued a whole queue ahead of its use: its latency overlaps the codes in between).  The serial
// chain of a Rice code is  window -> count leading zeros -> add -> shift -> window.
// Every lane of a warp walks its own frame, so some lane enters a new cache line in most iterations: lines are
// requested two ahead of the read position (no destination register, hence no scoreboard to wait for).
constexpr uint32_t kStreamPad = 256;  // zero bytes the caller keeps behind the stream

struct BitR {
    const uint32_t *w;
    unsigned long long acc;
    uint32_t q0, q1, q2, q3;  // the next words to enter the window, as loaded (q0 first); byte-swapped on entry
    uint4 nxt;                // the four words behind them: loaded a whole queue ahead of their use
    uint32_t qn;              // words left in the queue (1..4)
    int32_t nacc;             // valid bits in acc (32..64)
    uint32_t nw;              // index of q0 in the frame's words
    uint32_t wlimit;          // last word index the parse may need
    uint32_t end;             // limiting bit position (relative to w)
    uint32_t err;

    ZF_DEVICE void start(const uint32_t *words, uint32_t bit, uint32_t end_bit) {  // words: 16-byte aligned
        w = words;
        end = end_bit;
        wlimit = (end_bit >> 5) + 2u;
        err = 0;
        const uint32_t i = bit >> 5, sh = bit & 31u;
        acc = (((unsigned long long)dec_bswap(w[i]) << 32) | dec_bswap(w[i + 1])) << sh;
        nacc = 64 - (int32_t)sh;
        nw = i + 2u;
        const uint4 *w4 = reinterpret_cast<const uint4 *>(w);
        const uint4 v = w4[nw >> 2];
        const uint32_t k = nw & 3u;
        q0 = k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
        q1 = k == 0 ? v.y : k == 1 ? v.z : v.w;
        q2 = k == 0 ? v.z : v.w;
        q3 = v.w;
        qn = 4u - k;
        nxt = w4[(nw >> 2) + 1u];
    }
    ZF_DEVICE uint32_t pos() const { return nw * 32u - (uint32_t)nacc; }  // bits consumed, relative to w
    ZF_DEVICE uint32_t peek() const { return (uint32_t)(acc >> 32); }
    ZF_DEVICE void skip(uint32_t n) {  // n <= 32
        acc <<= n;
        nacc -= (int32_t)n;
        if (nacc < 32) {
            acc |= (unsigned long long)dec_bswap(q0) << (32 - nacc);
            nacc += 32;
            nw++;
            q0 = q1; q1 = q2; q2 = q3;
            if (--qn == 0u) {  // nw is a multiple of four here
                q0 = nxt.x; q1 = nxt.y; q2 = nxt.z; q3 = nxt.w;
                qn = 4u;
                if ((nw & 31u) == 0u) {  // once per 128 bytes: the limit (a damaged frame must not lead the reader out of
                                         // the buffer: it goes round in the last words it may touch) and the prefetch
                    if (nw > wlimit) {
                        err = 1;
                        nw -= 32u;
                    }
#ifndef ZF_HOST_EMU
                    else if (nw + 64u <= wlimit) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + nw + 64u));
#endif
                }
                nxt = reinterpret_cast<const uint4 *>(w)[(nw >> 2) + 1u];
            }
        }
    }
    ZF_DEVICE uint32_t read(uint32_t n) {  // n <= 32
        if (n == 0) return 0;
        const uint32_t v = peek() >> (32u - n);
        skip(n);
        return v;
    }
    ZF_DEVICE long long read_signed(uint32_t n) {  // n <= 33
        if (n == 0) return 0;
        unsigned long long v;
        if (n > 32u) {
            const unsigned long long hi = read(n - 32u);
            v = (hi << 32) | read(32);
        } else {
            v = read(n);
        }
        const uint32_t s = 64u - n;
        return (long long)(v << s) >> s;
    }
    ZF_DEVICE uint32_t read_unary() {  // zeros in front of the next one
        uint32_t q = 0;
        for (;;) {
            const uint32_t v = peek();
            if (v) {
                const uint32_t lz = (uint32_t)__clz((int)v);
                skip(lz + 1u);
                return q + lz;
            }
            q += 32u;
            skip(32);
            if (err) return q;
        }
    }
    ZF_DEVICE bool overrun() const { return err != 0 || pos() > end; }
};

// One subframe (RFC 9639 section 9.2) of `bs` samples at `bps` bits into out[0 .. bs).  ST = int32_t for streams of
// up to 24-bit samples (a side channel has 25), long long for 32-bit streams (33).  Samples are stored BEFORE the
// wasted-bits shift (the predictor runs on them); the shift count goes to `wasted`.
//
// `ring` is this lane's column of an 8 x 32 staging array in shared memory (element k at ring[32 k]).  A lane that wrote
// every sample straight to its own plane would re-use the store's source register one iteration later and wait each
// time until the 32-address store has been taken apart; FIXED subframes therefore collect eight samples and write them
// as whole 32-byte sectors.
template <typename ST>
ZF_DEVICE void flush8(ST *dst, const ST *ring) {
    ST v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = ring[32 * k];
#ifdef ZF_HOST_EMU
    for (int k = 0; k < 8; k++) dst[k] = v[k];
#else
    if (sizeof(ST) == 4) {
        uint4 *d = reinterpret_cast<uint4 *>(dst);
        d[0] = make_uint4((uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]);
        d[1] = make_uint4((uint32_t)v[4], (uint32_t)v[5], (uint32_t)v[6], (uint32_t)v[7]);
    } else {
        ulonglong2 *d = reinterpret_cast<ulonglong2 *>(dst);
#pragma unroll
        for (int k = 0; k < 4; k++) d[k] = make_ulonglong2((unsigned long long)v[2 * k], (unsigned long long)v[2 * k + 1]);
    }
#endif
}

template <typename ST>
ZF_DEVICE uint32_t decode_subframe(BitR &br, ST *out, ST *ring, uint32_t bs, uint32_t bps, uint8_t &wasted_out) {
    if (br.read(1)) return kErrReserved;
    const uint32_t type = br.read(6);
    uint32_t wasted = 0;
    if (br.read(1)) wasted = br.read_unary() + 1u;
    wasted_out = (uint8_t)wasted;
    if (br.overrun()) return kErrOverrun;
    if (wasted >= bps) return kErrRange;
    bps -= wasted;
    if (type == 0) {  // CONSTANT
        const ST v = (ST)br.read_signed(bps);
        for (uint32_t i = 0; i < bs; i++) out[i] = v;
        return br.overrun() ? kErrOverrun : kOk;
    }
    if (type == 1) {  // VERBATIM
        for (uint32_t i = 0; i < bs; i++) {
            out[i] = (ST)br.read_signed(bps);
            if (br.err) return kErrOverrun;
        }
        return br.overrun() ? kErrOverrun : kOk;
    }
    uint32_t order, shift = 0;
    bool lpc = false;
    int32_t coef[32];
    if (type >= 8u && type <= 12u) {
        order = type - 8u;
    } else if (type >= 32u) {
        order = type - 31u;
        lpc = true;
    } else {
        return kErrReserved;
    }
    if (order > bs) return kErrRange;
    for (uint32_t i = 0; i < order; i++) out[i] = (ST)br.read_signed(bps);
    if (lpc) {
        const uint32_t prec = br.read(4) + 1u;
        if (prec == 16u) return kErrReserved;
        const long long sh = br.read_signed(5);
        if (sh < 0) return kErrReserved;
        shift = (uint32_t)sh;
        for (uint32_t i = 0; i < order; i++) coef[i] = (int32_t)br.read_signed(prec);
    }
    // residual: method, partition order, then per partition a parameter (or escape + width) and its codes
    const uint32_t method = br.read(2);
    if (method > 1u) return kErrReserved;
    const uint32_t plen = method ? 5u : 4u, escape = method ? 31u : 15u;
    const uint32_t po = br.read(4);
    const uint32_t psize = bs >> po;
    if (po && (psize << po) != bs) return kErrRange;
    if (psize < order) return kErrRange;
    if (br.overrun()) return kErrOverrun;
    // One flat loop over the sample index for all lanes of the warp (their partition sizes differ: a change of
    // partition is a short divergent branch, not a loop boundary).  part_end: end of the partition whose parameter was
    // read last (none yet); the first partition holds psize - order samples and may be EMPTY (psize == order), in which
    // case two parameters stand in front of the first code.
    uint32_t part_end = 0;
    uint32_t k = 0, raw = 0;
    bool esc = false;
#define ZF_DEC_RESIDUAL(R)                                                                       \
    while (i >= part_end) {                                                                      \
        k = br.read(plen);                                                                       \
        esc = k == escape;                                                                       \
        if (esc) raw = br.read(5);                                                               \
        part_end += psize;                                                                       \
    }                                                                                            \
    ST R;                                                                                        \
    if (esc) {                                                                                   \
        R = (ST)br.read_signed(raw);                                                             \
    } else {                                                                                     \
        const uint32_t v = br.peek();                                                            \
        uint32_t zz;                                                                             \
        const uint32_t lz = (uint32_t)__clz((int)v);                                             \
        if (v != 0 && lz + 1u + k <= 32u) { /* the whole code lies in the window */              \
            const uint32_t rem = k ? ((v << lz) << 1) >> (32u - k) : 0u;                         \
            zz = (lz << k) | rem;                                                                \
            br.skip(lz + 1u + k);                                                                \
        } else {                                                                                 \
            const uint32_t q = br.read_unary();                                                  \
            zz = (q << k) | br.read(k);                                                          \
        }                                                                                        \
        R = (ST)(int32_t)((zz >> 1) ^ (0u - (zz & 1u)));                                         \
    }
    if (lpc) {
        for (uint32_t i = order; i < bs; i++) {
            ZF_DEC_RESIDUAL(r)
            long long sum = 0;
            for (uint32_t j = 0; j < order; j++) sum += (long long)coef[j] * (long long)out[i - 1 - j];
            out[i] = r + (ST)(sum >> shift);
        }
    } else {
        // FIXED predictors as running differences: d_j is the j-th difference of the signal at the previous sample; a
        // level above the order passes the residual through (its mask is 0), so one piece of code serves orders 0..4
        ST d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        if (order >= 1) d0 = out[order - 1];
        if (order >= 2) d1 = out[order - 1] - out[order - 2];
        if (order >= 3) d2 = out[order - 1] - 2 * out[order - 2] + out[order - 3];
        if (order >= 4) d3 = out[order - 1] - 3 * out[order - 2] + 3 * out[order - 3] - out[order - 4];
        const ST m0 = order > 0 ? -1 : 0, m1 = order > 1 ? -1 : 0, m2 = order > 2 ? -1 : 0, m3 = order > 3 ? -1 : 0;
        for (uint32_t i = 0; i < order; i++) ring[32u * (i & 7u)] = out[i];  // the warm-ups join the first sector
        for (uint32_t i = order; i < bs; i++) {
            ZF_DEC_RESIDUAL(r)
            d3 = (d3 & m3) + r;
            d2 = (d2 & m2) + d3;
            d1 = (d1 & m1) + d2;
            d0 = (d0 & m0) + d1;
            ring[32u * (i & 7u)] = d0;
            if ((i & 7u) == 7u) flush8<ST>(out + (i - 7u), ring);
        }
        for (uint32_t i = bs & ~7u; i < bs; i++) out[i] = ring[32u * (i & 7u)];
    }
#undef ZF_DEC_RESIDUAL
    return br.overrun() ? kErrOverrun : kOk;
}

// One frame: header, subframes, padding; stream bytes [begin, end) with the CRC-16 in the last two.
template <typename ST>
ZF_DEVICE void decode_frame(const uint8_t *s, unsigned long long begin, unsigned long long end, const StreamParams &sp,
                            ST *work, ST *ring, FrameRec &rec) {
    rec.block_size = 0;
    rec.ch_code = 0;
    for (int c = 0; c < 8; c++) rec.wasted[c] = 0;
    rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
    FrameHdr h;
    if (end < begin + 2 || !parse_header(s + begin, end - begin - 2, sp, h)) {
        rec.status = kErrHeader;
        return;
    }
    rec.block_size = h.block_size;
    rec.ch_code = (uint8_t)h.ch_code;
    BitR br;
    const uint32_t lead = (uint32_t)(begin & 15ull) * 8u;  // bits between the 16-byte boundary and the frame's first byte
    br.start(reinterpret_cast<const uint32_t *>(s + (begin & ~15ull)), lead + h.len * 8u, lead + (uint32_t)(end - 2 - begin) * 8u);
    uint32_t st = kOk;
    for (uint32_t c = 0; c < h.channels && st == kOk; c++) {
        const bool side = (h.ch_code == 8u && c == 1) || (h.ch_code == 9u && c == 0) || (h.ch_code == 10u && c == 1);
        st = decode_subframe<ST>(br, work + (size_t)c * plane_stride(sp.max_block), ring, h.block_size, h.bits + (side ? 1u : 0u),
                                 rec.wasted[c]);
    }
    if (st == kOk) {
        const uint32_t padbits = (8u - (br.pos() & 7u)) & 7u;
        if (padbits && br.read(padbits) != 0) st = kErrPadding;
        else if (br.overrun()) st = kErrOverrun;
        else if (br.pos() != br.end) st = kErrLength;
    }
    rec.status = st;
}

// ---- CRC-16 (poly 0x8005, init 0) ------------------------------------------------------------------------
ZF_DEVICE uint32_t crc16_mulmod(uint32_t a, uint32_t b) {  // a * b mod x^16 + x^15 + x^2 + 1
    uint32_t r = 0;
    for (int i = 15; i >= 0; i--) {
        r = (r & 0x8000u) ? ((r << 1) ^ 0x8005u) & 0xffffu : (r << 1);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
ZF_DEVICE uint32_t crc16_xpow8(unsigned long long n) {  // x^(8 n) mod P
    uint32_t r = 1, base = 0x100u;
    while (n) {
        if (n & 1ull) r = crc16_mulmod(r, base);
        base = crc16_mulmod(base, base);
        n >>= 1;
    }
    return r;
}
ZF_DEVICE uint32_t crc16_table_entry(uint32_t b) {
    uint32_t c = b << 8;
    for (int k = 0; k < 8; k++) c = (c & 0x8000u) ? ((c << 1) ^ 0x8005u) & 0xffffu : (c << 1);
    return c;
}
// tab[k][v] = v * x^(8 k + 16) mod P, k = 0..3: table k + 1 is table k advanced by one zero byte
// CRC-16 of n bytes at p (init 0): bytes up to a 4-byte boundary, then whole words four table look-ups at a time
// (state c = c_hi x^8 + c_lo joins the first two bytes: c x^32 + b0 x^40 + b1 x^32 + b2 x^24 + b3 x^16), then the rest
ZF_DEVICE uint32_t crc16_span(const uint8_t *p, unsigned long long n, const uint16_t (*tab)[256]) {
    uint32_t c = 0;
    unsigned long long i = 0;
    for (; i < n && (((uintptr_t)(p + i)) & 3u) != 0; i++) c = ((c << 8) & 0xffffu) ^ tab[0][((c >> 8) ^ p[i]) & 0xffu];
    for (; i + 4 <= n; i += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);  // little-endian load: byte k of the stream in bits 8k..
        c = tab[3][((c >> 8) ^ w) & 0xffu] ^ tab[2][(c ^ (w >> 8)) & 0xffu] ^ tab[1][(w >> 16) & 0xffu] ^ tab[0][w >> 24];
    }
    for (; i < n; i++) c = ((c << 8) & 0xffffu) ^ tab[0][((c >> 8) ^ p[i]) & 0xffu];
    return c;
}
ZF_DEVICE void crc16_build_tables(uint16_t (*tab)[256], uint32_t b) {  // entry b of the four tables
    uint32_t v = crc16_table_entry(b);
    tab[0][b] = (uint16_t)v;
    for (int k = 1; k < 4; k++) {
        v = ((v << 8) & 0xffffu) ^ crc16_table_entry(v >> 8);
        tab[k][b] = (uint16_t)v;
    }
}

// ---- output: one inter-channel sample ----------------------------------------------------------------------
// Channel values of sample i of a frame after the wasted-bits shift and the inter-channel restore
// (RFC 9639 section 9.2.3 / 4.2); returns false when a value does not fit the stream's depth.
template <typename ST>
ZF_DEVICE bool restore_sample(const ST *work, uint32_t max_block, uint32_t i, const FrameRec &rec, uint32_t channels,
                              uint32_t bits, long long (&v)[8]) {
    for (uint32_t c = 0; c < channels; c++)
        v[c] = (long long)((unsigned long long)(long long)work[(size_t)c * plane_stride(max_block) + i] << rec.wasted[c]);
    if (rec.ch_code == 8u) {
        v[1] = v[0] - v[1];
    } else if (rec.ch_code == 9u) {
        v[0] = v[0] + v[1];
    } else if (rec.ch_code == 10u) {
        const long long side = v[1];
        const long long mid = (long long)((unsigned long long)v[0] << 1) | (side & 1);
        v[0] = (mid + side) >> 1;
        v[1] = (mid - side) >> 1;
    }
    const long long lo = -(1ll << (bits - 1u)), hi = (1ll << (bits - 1u)) - 1;
    bool ok = true;
    for (uint32_t c = 0; c < channels; c++) ok = ok && v[c] >= lo && v[c] <= hi;
    return ok;
}

#ifndef ZF_HOST_EMU
// ===================================================================================================================
// kernels
// ===================================================================================================================

// every byte position of [begin, len): thread t looks at the four positions of word t
__global__ void __launch_bounds__(256) zf_dec_scan_kernel(const uint8_t *s, unsigned long long begin, unsigned long long len,
                                                          StreamParams sp, Cand *cand, uint32_t cap, uint32_t *count) {
    const unsigned long long wi = (unsigned long long)blockIdx.x * 256ull + threadIdx.x + (begin >> 2);
    if (wi * 4ull >= len) return;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
    const uint32_t a = w[wi], b = w[wi + 1];  // little-endian words: byte k of the stream is bits 8k..8k+7
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        const uint32_t two = k == 0 ? a & 0xffffu : k == 1 ? (a >> 8) & 0xffffu : k == 2 ? a >> 16 : (a >> 24) | ((b & 0xffu) << 8);
        if ((two & 0xfeffu) != 0xf8ffu) continue;
        const unsigned long long pos = wi * 4ull + k;
        if (pos < begin || pos + 6ull > len) continue;
        FrameHdr h;
        if (!parse_header(s + pos, len - pos, sp, h)) continue;
        const uint32_t slot = atomicAdd(count, 1u);
        if (slot < cap) {
            Cand c;
            c.pos = pos;
            c.number = h.number;
            c.block_size = h.block_size;
            c.variable = h.variable;
            cand[slot] = c;
        }
    }
}

// One thread per frame, `lpw` of them in a warp (the other lanes leave at once).  Every lane walks its own frame, so
// with 32 busy lanes some lane misses the L1 cache in most iterations and the whole warp waits for L2; the host
// therefore spreads the frames over as many warps as the device holds (a few per scheduler) before it fills the lanes.
template <typename ST>
__global__ void __launch_bounds__(32) zf_dec_frames_kernel(const uint8_t *s, const unsigned long long *fpos, uint32_t n_frames,
                                                           uint32_t lpw, StreamParams sp, ST *work, FrameRec *rec) {
    __shared__ ST ring[8 * 32];
    const uint32_t f = blockIdx.x * lpw + threadIdx.x;
    if (threadIdx.x >= lpw || f >= n_frames) return;
    FrameRec r;
    decode_frame<ST>(s, fpos[f], fpos[f + 1], sp, work + (size_t)f * sp.channels * plane_stride(sp.max_block), ring + threadIdx.x, r);
    rec[f] = r;
}

// x^(8 * 2^i) mod P, i < 32: computed by the host once (zf_decode.cu), passed by value
struct CrcPowers {
    uint16_t pw[32];
};

// One WARP per frame (four frames per CTA share the tables): the frame is cut into 32 chunks of equal length -- it is
// thought of as padded IN FRONT with zero bytes, which do not change a CRC with zero init -- one per lane, four bytes per
// step; the 32 chunk CRCs are then folded by a shuffle tree: CRC(A || B) = CRC(A) x^(8 |B|) + CRC(B), and at level j
// every B is 2^j chunks long, so one multiplier serves the whole level.
__global__ void __launch_bounds__(128) zf_dec_crc16_kernel(const uint8_t *s, const unsigned long long *fpos, uint32_t n_frames,
                                                           CrcPowers pows, FrameRec *rec) {
    __shared__ uint16_t tab[4][256];
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t f = blockIdx.x * 4u + (t >> 5);
    for (uint32_t b = t; b < 256u; b += 128u) crc16_build_tables(tab, b);
    __syncthreads();
    if (f >= n_frames) return;
    const unsigned long long begin = fpos[f], n = fpos[f + 1] - begin;
    const unsigned long long chunk = (n + 31ull) / 32ull, pad = chunk * 32ull - n;
    const unsigned long long lo_v = chunk * lane, hi_v = lo_v + chunk;
    const unsigned long long lo = lo_v > pad ? lo_v - pad : 0ull, hi = hi_v > pad ? hi_v - pad : 0ull;
    uint32_t c = crc16_span(s + begin + lo, hi - lo, tab);
    uint32_t x = 1;  // x^(8 chunk)
    for (uint32_t i = 0; i < 32u; i++)
        if ((chunk >> i) & 1ull) x = crc16_mulmod(x, pows.pw[i]);
#pragma unroll
    for (uint32_t j = 0; j < 5u; j++) {
        const uint32_t other = __shfl_down_sync(0xffffffffu, c, 1u << j);
        if ((lane & ((2u << j) - 1u)) == 0u) c = crc16_mulmod(c, x) ^ other;
        x = crc16_mulmod(x, x);
    }
    if (lane == 0 && c != 0 && rec[f].status == kOk) rec[f].status = kErrCrc16;
}

// one CTA (256 threads) per frame, 256 samples per trip: restore, range check, pack little-endian, store coalesced
template <typename ST>
__global__ void __launch_bounds__(256) zf_dec_output_kernel(const ST *work, FrameRec *rec, const unsigned long long *first_sample,
                                                            StreamParams sp, uint8_t *pcm, unsigned long long pcm_cap) {
    __shared__ __align__(16) uint8_t stage[256 * 8 * 4];
    const uint32_t f = blockIdx.x, t = threadIdx.x;
    const FrameRec r = rec[f];
    if (r.status != kOk) return;
    const uint32_t bytes = sp.bits / 8u, stride = sp.channels * bytes, pstride = plane_stride(sp.max_block);
    const ST *planes = work + (size_t)f * sp.channels * pstride;
    const unsigned long long off0 = first_sample[f] * (unsigned long long)stride;
    if (off0 + (unsigned long long)r.block_size * stride > pcm_cap) return;
    const bool stereo32 = sizeof(ST) == 4 && sp.channels == 2u;
    const int32_t lo = (int32_t)(0xffffffffu << ((sp.bits - 1u) & 31u)), hi = ~lo;  // -2^(bits-1) .. 2^(bits-1) - 1
    bool bad = false;
    for (uint32_t i0 = 0; i0 < r.block_size; i0 += 256u) {
        const uint32_t n = r.block_size - i0 < 256u ? r.block_size - i0 : 256u;
        uint8_t *dst = pcm + off0 + (unsigned long long)i0 * stride;
        const uint32_t total = n * stride;
        if (stereo32) {
            // stereo of up to 24 bits (the common case): 32-bit arithmetic (a side sample has 25 bits, 2 mid + 1 has 26),
            // 8- and 16-bit samples go straight to the stream, 24-bit ones as three half-words through shared memory
            uint32_t lw = 0, rw = 0;
            if (t < n) {
                int32_t a = (int32_t)((uint32_t)(int32_t)planes[i0 + t] << r.wasted[0]);
                int32_t b = (int32_t)((uint32_t)(int32_t)planes[pstride + i0 + t] << r.wasted[1]);
                if (r.ch_code == 8u) {
                    b = a - b;
                } else if (r.ch_code == 9u) {
                    a = a + b;
                } else if (r.ch_code == 10u) {
                    const int32_t mid = (int32_t)((uint32_t)a << 1) | (b & 1);
                    a = (mid + b) >> 1;
                    b = (mid - b) >> 1;
                }
                bad = bad || a < lo || a > hi || b < lo || b > hi;
                lw = (uint32_t)a;
                rw = (uint32_t)b;
            }
            if (bytes == 2u) {
                if (t < n) reinterpret_cast<uint32_t *>(dst)[t] = (lw & 0xffffu) | (rw << 16);  // a multiple of four bytes in
                continue;
            }
            if (bytes == 1u) {
                if (t < n) reinterpret_cast<uint16_t *>(dst)[t] = (uint16_t)((lw & 0xffu) | ((rw & 0xffu) << 8));
                continue;
            }
            if (t < n) {
                uint16_t *h = reinterpret_cast<uint16_t *>(stage) + 3u * t;
                h[0] = (uint16_t)lw;
                h[1] = (uint16_t)(((lw >> 16) & 0xffu) | ((rw & 0xffu) << 8));
                h[2] = (uint16_t)(rw >> 8);
            }
        } else if (t < n) {
            long long v[8];
            if (!restore_sample<ST>(planes, sp.max_block, i0 + t, r, sp.channels, sp.bits, v)) bad = true;
            for (uint32_t c = 0; c < sp.channels; c++)
                for (uint32_t k = 0; k < bytes; k++) stage[t * stride + c * bytes + k] = (uint8_t)((unsigned long long)v[c] >> (8u * k));
        }
        __syncthreads();
        if (((uintptr_t)dst & 15u) == 0) {
            const uint32_t quads = total >> 4;
            const uint4 *sq = reinterpret_cast<const uint4 *>(stage);
            uint4 *dq = reinterpret_cast<uint4 *>(dst);
            for (uint32_t k = t; k < quads; k += 256u) dq[k] = sq[k];
            for (uint32_t k = (quads << 4) + t; k < total; k += 256u) dst[k] = stage[k];
        } else if (((uintptr_t)dst & 3u) == 0) {
            const uint32_t words = total >> 2;
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage);
            uint32_t *dw = reinterpret_cast<uint32_t *>(dst);
            for (uint32_t k = t; k < words; k += 256u) dw[k] = sw[k];
            for (uint32_t k = (words << 2) + t; k < total; k += 256u) dst[k] = stage[k];
        } else {
            for (uint32_t k = t; k < total; k += 256u) dst[k] = stage[k];
        }
        __syncthreads();  // the staging area is written again in the next trip
    }
    if (bad) rec[f].status = kErrRange;  // benign race: every writer stores the same value
}
#endif  // !ZF_HOST_EMU

}  // namespace dec
}  // namespace zf

// zf_kernel.cuh -- shared device code and the GENERAL stereo frame-encode kernel (sm_100a): any block size up to 4096,
// short last frames, any Rice limits, any sample rate.  The dominant kernel for full frames is zf_kernel_v3.cuh.
//
// One thread block (512 threads x 8 samples) encodes one frame; a persistent grid pulls frames from an atomic
// ticket and compacts the variable-length frames into one output stream with a decoupled look-back
// over frame byte sizes (single pass: FLAC bytes are written to HBM exactly once).
//
// Per frame (reference call stack: Encoder.writeFrame, encoder.zig:234-284):
//   load     raw interleaved PCM -> shared memory by a 1-D TMA bulk copy (next frame prefetched),
//            unpacked to int32 registers: 8 consecutive samples + 4 history samples per thread
//   pass 1   L, R, M=(L+R)>>1, S=L-R: OR of samples (wasted bits, encoder.zig:556-570) and the five
//            sum|delta^k x| of fixed.bestOrder (fixed.zig:85-167), block-reduced with REDUX
//   decide   CONSTANT / VERBATIM / fixed order per candidate channel (encoder.zig:482-554)
//   pass 2   residual of the chosen order -> finest-partition abs-sums and zigzag widths
//            (rice.calcSums, rice.zig:288-340), partition tree, closed-form parameter search
//            (rice.calcOptimalParams, rice.zig:343-395) and partition-order pick (rice.zig:262-276)
//   stereo   first minimum of [L+R, L+S, S+R, M+S] estimated bits (encoder.zig:441-452)
//   pack     exact code lengths -> block exclusive scan -> every thread ORs its codewords into a
//            zeroed shared bit buffer (frame_writer.zig:269-372 subframe grammars), header + CRC-8
//            (frame_writer.zig:151-265), CRC-16 by per-chunk table CRC combined with x^(8k) mod P
//   store    look-back for the frame's byte offset, byte-shifted coalesced copy to the stream
//
// All arithmetic is integer and exact; tie-breaking follows SURVEY.md 8-Q.
#pragma once
#include <stdint.h>

#include "zf_dev.h"

namespace zf {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kSpt = 8;                     // samples per thread
constexpr int kMaxBlock = kThreads * kSpt;  // 4096
constexpr int kHalo = 4;
constexpr int kX = kSpt + kHalo;
constexpr int kNodes = 512;  // partition tree: node(level, j) = (1 << level) - 1 + j, 511 used
constexpr int kMaxLevel = 8;
constexpr int kRawPadWords = 8;  // zeroed words in front of the raw PCM (history of thread 0)

constexpr unsigned long long kFlagAggregate = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

enum : uint32_t { kConstant = 0, kVerbatim = 1, kFixed = 2 };
enum : uint32_t { kStatusOutOverflow = 1u, kStatusBitOverflow = 2u };

constexpr unsigned long long kU64Max = ~0ull;

struct FrameJob {
    const uint8_t *pcm;            // PCM of this launch's first frame (interleaved little-endian)
    uint8_t *out;                  // compacted frame stream of the whole batch
    unsigned long long out_cap;
    uint32_t *frame_sizes;         // [batch frames]
    unsigned long long *desc;      // [batch frames] look-back descriptors, zeroed before the batch
    unsigned int *ticket;          // this launch's frame counter, zeroed before the launch
    unsigned int *status;
    unsigned long long *total_bytes;  // written by the batch's last frame
    const uint16_t *pow8;          // x^(8k) mod (x^16+x^15+x^2+1), k < pow8_len
    uint32_t n_frames;             // frames in this launch
    uint32_t frame_base;           // batch index of this launch's first frame
    uint32_t batch_frames;
    unsigned long long first_frame_number;  // FLAC frame number of batch frame 0
    uint32_t block_size;           // samples per channel in every frame of this launch
    uint32_t frame_stride;         // bytes between consecutive frames' PCM
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t max_rice_order;
    uint32_t max_rice_param;
    uint32_t use_tma;
    const uint16_t *lpc_window;    // LPC kernel: integer Welch window of this launch's block size (zf_kernel_lpc.cuh)
    uint32_t lpc_order;            // LPC kernel: maximum order
    uint32_t bit_depth;            // general kernels: sample depth when it is not the container's (8-bit samples travel in
                                   // 16-bit containers); 0 = 8 x container bytes
    uint32_t pdl_trigger;          // general kernels: release the dependent launch at once (the one-CTA last-frame
                                   // launch in front of the persistent full-frame kernel)
};

// ---------------------------------------------------------------------------------------------------
// small integer helpers
// ---------------------------------------------------------------------------------------------------

ZF_DEVICE uint32_t bitlen32(uint32_t v) { return 32u - (uint32_t)__clz((int)v); }
ZF_DEVICE uint32_t bitlen64(unsigned long long v) { return 64u - (uint32_t)__clzll((long long)v); }
ZF_DEVICE uint32_t ctz32(uint32_t v) { return (uint32_t)__ffs((int)v) - 1u; }
ZF_DEVICE uint32_t ctz64(unsigned long long v) { return (uint32_t)__ffsll((long long)v) - 1u; }
ZF_DEVICE uint32_t floor_log2(uint32_t v) { return 31u - (uint32_t)__clz((int)v); }
// rice.calcZigzag, rice.zig:281-284
ZF_DEVICE uint32_t zigzag(int32_t v) { return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); }

template <bool WIDE>
struct Ar;
template <>
struct Ar<false> {
    typedef int32_t T;
    typedef uint32_t U;
};
template <>
struct Ar<true> {
    typedef long long T;
    typedef unsigned long long U;
};

ZF_DEVICE uint32_t uabs(int32_t v) { return v < 0 ? 0u - (uint32_t)v : (uint32_t)v; }
ZF_DEVICE unsigned long long uabs(long long v) { return v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v; }

// exact warp sums through REDUX on 16-bit pieces (32 lanes x 65535 fits 21 bits)
ZF_DEVICE unsigned long long warp_sum(uint32_t v) {
    const uint32_t lo = reduce_add(v & 0xffffu), hi = reduce_add(v >> 16);
    return (unsigned long long)lo + ((unsigned long long)hi << 16);
}
ZF_DEVICE unsigned long long warp_sum(unsigned long long v) {
    return warp_sum((uint32_t)v) + (warp_sum((uint32_t)(v >> 32)) << 32);
}
ZF_DEVICE unsigned long long warp_or(uint32_t v) { return reduce_or(v); }
ZF_DEVICE unsigned long long warp_or(unsigned long long v) {
    return (unsigned long long)reduce_or((uint32_t)v) | ((unsigned long long)reduce_or((uint32_t)(v >> 32)) << 32);
}

// ---------------------------------------------------------------------------------------------------
// shared memory
// ---------------------------------------------------------------------------------------------------

// per candidate channel ("slot": stereo L, R, M, S; or one independent channel)
struct SlotDec {
    unsigned long long est_bits;  // what chooseSubframeEncoding returns (encoder.zig:553)
    uint32_t kind, waste, bps, depth;
    uint32_t order, mpo, po, method;
    uint32_t max_param;  // parameters tried: 0 .. max_param-1 (rice.zig:104,369)
    uint32_t pad;
};

constexpr int kRedVals = 11;  // 5 sums + 5 range ORs + sample OR

template <int BYTES>
struct BitBufWords {
    // frame bytes <= 16 (header) + 2 * (N*(depth+1) + N/2 + 256*10 + 200)/8 + 2
    static constexpr int value = ((16 + 2 * ((kMaxBlock * (8 * BYTES + 1) + kMaxBlock / 2 + 256 * 10 + 200) / 8) + 2 + 64) / 16) * 4;
};

struct SmemCommon {
    unsigned long long mbar;
    unsigned long long out_off;
    unsigned long long red[kWarps][4][kRedVals];
    unsigned long long psum[4][kNodes];
    unsigned long long levelcost[4][kMaxLevel + 1];
    uint32_t pbits[4][kNodes];
    uint32_t levelfive[4];
    unsigned long long mixed[4][32];          // search round A, warp 0: levels 0..4 (heap node = lane)
    unsigned long long wcost[4][2][kWarps];   // per-warp cost sums of the uniform-level warps
    uint32_t wfive[4][2][kWarps];
    uint32_t mixfive[4];
    unsigned long long tot[4][kRedVals];      // block totals of pass 1 (sums, range ORs, sample OR)
    long long warm[2][4];                     // warm-up samples of the two written subframes (thread 0)
    uint32_t hdrw[4][4];                      // frame header words for ch_type 1, 8, 9, 10 (prepared early)
    uint32_t hdr_len;
    SlotDec dec[4];
    uint32_t warp_scan[2][kWarps];
    uint32_t crc_part[kWarps];
    uint32_t cur_frame, next_frame;
    uint32_t sub_slot[2];
    uint32_t ch_type;
    uint32_t total_bits;
    uint32_t sub_bits0;
    uint16_t crc16tab[4][256];
    uint8_t crc8tab[256];
    uint8_t pchoice[4][kNodes];
};

template <int BYTES>
struct SmemStereo {
    SmemCommon c;
    alignas(16) uint32_t raw[kRawPadWords + kMaxBlock * 2 * BYTES / 4];
    alignas(16) uint32_t bits[BitBufWords<BYTES>::value];
};

// ---------------------------------------------------------------------------------------------------
// CRC tables (CRC-8 poly 0x07, CRC-16 poly 0x8005; both init 0, unreflected)
// ---------------------------------------------------------------------------------------------------

ZF_DEVICE void init_tables(SmemCommon &c, int t) {
    if (t < 256) {
        uint32_t v8 = (uint32_t)t, v16 = (uint32_t)t << 8;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v8 = (v8 & 0x80u) ? ((v8 << 1) ^ 0x07u) : (v8 << 1);
            v16 = (v16 & 0x8000u) ? ((v16 << 1) ^ 0x8005u) : (v16 << 1);
        }
        c.crc8tab[t] = (uint8_t)v8;
        c.crc16tab[0][t] = (uint16_t)v16;
    }
    __syncthreads();
    // tab[k][b] = b * x^(8k+16) mod P : advance tab[k-1][b] by one zero byte
    for (int k = 1; k < 4; k++) {
        if (t < 256) {
            const uint32_t prev = c.crc16tab[k - 1][t];
            c.crc16tab[k][t] = (uint16_t)(((prev << 8) & 0xffffu) ^ c.crc16tab[0][prev >> 8]);
        }
        __syncthreads();
    }
}

ZF_DEVICE uint32_t crc16_word(const SmemCommon &c, uint32_t crc, uint32_t be_word) {
    const uint32_t tt = (crc << 16) ^ be_word;
    return (uint32_t)c.crc16tab[3][tt >> 24] ^ c.crc16tab[2][(tt >> 16) & 0xffu] ^ c.crc16tab[1][(tt >> 8) & 0xffu] ^
           c.crc16tab[0][tt & 0xffu];
}
ZF_DEVICE uint32_t crc16_byte(const SmemCommon &c, uint32_t crc, uint32_t byte) {
    return ((crc << 8) & 0xffffu) ^ c.crc16tab[0][((crc >> 8) ^ byte) & 0xffu];
}
// a(x) * b(x) mod P for 16-bit residues
ZF_DEVICE uint32_t crc16_mulmod(const SmemCommon &c, uint32_t a, uint32_t b) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= ((b >> i) & 1u) ? (a << i) : 0u;
    // acc = H * x^16 + Lo with deg H < 15: H * x^16 mod P is the CRC of the two bytes of H
    const uint32_t h = acc >> 16;
    return (acc & 0xffffu) ^ c.crc16tab[1][h >> 8] ^ c.crc16tab[0][h & 0xffu];
}

// ---------------------------------------------------------------------------------------------------
// per-thread bit writer into the zeroed shared bit buffer (big-endian bit order inside 32-bit words).
// A thread owns a contiguous bit range; only its first and last word can be shared with a neighbour,
// so those two use atomicOr and every interior word is a plain store (frame_writer.zig:40-101 semantics:
// MSB-first bit stream; zero runs are simply skipped because the buffer is pre-zeroed).
// ---------------------------------------------------------------------------------------------------

struct BitWriter {
    uint32_t *buf;
    uint32_t w;
    uint32_t acc;
    bool first;

    ZF_DEVICE void init(uint32_t *b, uint32_t bitpos) {
        buf = b;
        w = bitpos >> 5;
        acc = 0;
        first = true;
    }
    ZF_DEVICE void flush_mid() {
        if (first) {
            if (acc) atomicOr(&buf[w], acc);
            first = false;
        } else {
            buf[w] = acc;
        }
    }
    // len in 1..32, val < 2^len, pos >= every earlier position
    ZF_DEVICE void put(uint32_t pos, uint32_t val, uint32_t len) {
        const uint32_t wi = pos >> 5, off = pos & 31u;
        if (wi != w) {
            flush_mid();
            w = wi;
            acc = 0;
        }
        const uint32_t room = 32u - off;
        if (len <= room) {
            acc |= val << (room - len);
        } else {
            const uint32_t spill = len - room;
            acc |= val >> spill;
            flush_mid();
            w = wi + 1;
            acc = val << (32u - spill);
        }
    }
    // len in 1..64
    ZF_DEVICE void put64(uint32_t pos, unsigned long long val, uint32_t len) {
        if (len > 32) {
            put(pos, (uint32_t)(val >> 32), len - 32);
            put(pos + (len - 32), (uint32_t)val, 32);
        } else {
            put(pos, (uint32_t)val, len);
        }
    }
    ZF_DEVICE void finish() {
        if (acc) atomicOr(&buf[w], acc);
    }
};

// ---------------------------------------------------------------------------------------------------
// frame header, frame_writer.zig:151-265 (+ writeCrc8 :128-141); built by one thread
// ---------------------------------------------------------------------------------------------------

ZF_DEVICE uint32_t number_bytes(unsigned long long frame_number) {  // UTF-8-like coder, :235-251
    if (frame_number <= 0x7F) return 1;
    uint32_t i = 0;
    unsigned long long first_byte_max = 0x3f, number = frame_number;
    while (number > first_byte_max) {
        i++;
        number >>= 6;
        first_byte_max >>= 1;
    }
    return i + 1;
}

ZF_DEVICE uint32_t rate_code(uint32_t sample_rate, uint32_t &extra_bytes) {  // :187-217
    extra_bytes = 0;
    switch (sample_rate) {
        case 0: return 0;
        case 88200: return 1;
        case 176400: return 2;
        case 192000: return 3;
        case 8000: return 4;
        case 16000: return 5;
        case 22050: return 6;
        case 24000: return 7;
        case 32000: return 8;
        case 44100: return 9;
        case 48000: return 10;
        case 96000: return 11;
        default: break;
    }
    if (sample_rate <= 255) { extra_bytes = 1; return 12; }
    extra_bytes = 2;
    return sample_rate <= 65535 ? 13 : 14;
}

ZF_DEVICE uint32_t block_size_code(uint32_t n, uint32_t &extra_bytes) {  // :165-185
    extra_bytes = 0;
    const uint32_t tz = ctz32(n);
    if ((n & (n - 1)) == 0 && tz >= 8 && tz <= 15) return tz;
    if (n == 192) return 1;
    // the reference's 144 * 2^v test compares an odd number with 144 and never fires
    if (n < 0x100) { extra_bytes = 1; return 6; }
    extra_bytes = 2;
    return 7;
}

ZF_DEVICE uint32_t header_len(unsigned long long frame_number, uint32_t n, uint32_t sample_rate) {
    uint32_t be, re;
    block_size_code(n, be);
    rate_code(sample_rate, re);
    return 4 + number_bytes(frame_number) + be + re + 1;
}

// header bytes without the CRC-8; returns their count (frame_writer.zig:151-262)
ZF_DEVICE uint32_t build_header(uint8_t (&hb)[16], unsigned long long frame_number, uint32_t depth, uint32_t ch_type,
                                uint32_t n, uint32_t sample_rate) {
    uint32_t len = 0, be, re;
    const uint32_t bsc = block_size_code(n, be), rc = rate_code(sample_rate, re);
    hb[len++] = 0xFF;
    hb[len++] = 0xF8;  // fixed-blocksize stream, :163
    hb[len++] = (uint8_t)((bsc << 4) | rc);
    const uint32_t dc = depth == 8 ? 2u : depth == 16 ? 8u : depth == 24 ? 12u : 14u;  // :221-233
    hb[len++] = (uint8_t)((ch_type << 4) | dc);
    if (frame_number <= 0x7F) {
        hb[len++] = (uint8_t)frame_number;
    } else {
        uint8_t cont[8];
        uint32_t i = 0;
        unsigned long long first_byte_max = 0x3f, number = frame_number;
        while (number > first_byte_max) {
            uint32_t b = 0x80u + (uint32_t)(number & 0x3f);
            if (i == 4) b &= 0x0Fu;  // u36 shift truncation in the reference for numbers >= 2^26
            cont[i] = (uint8_t)b;
            i++;
            number >>= 6;
            first_byte_max >>= 1;
        }
        hb[len++] = (uint8_t)(((0xFEu << (6 - i)) | (uint32_t)number) & 0xFFu);
        for (uint32_t k = i; k > 0; k--) hb[len++] = cont[k - 1];
    }
    if (be == 1) hb[len++] = (uint8_t)(n - 1);
    else if (be == 2) { hb[len++] = (uint8_t)((n - 1) >> 8); hb[len++] = (uint8_t)(n - 1); }
    if (re == 1) {  // :260 writes block_size unmasked in 8 bits: the high byte ORs into the previous byte (Q10)
        hb[len - 1] |= (uint8_t)(n >> 8);
        hb[len++] = (uint8_t)n;
    } else if (re == 2) {  // :261 block_size (code 13) or block_size / 10 (code 14) -- not the sample rate (Q10)
        const uint32_t v = (rc == 13) ? n : n / 10;
        hb[len++] = (uint8_t)(v >> 8);
        hb[len++] = (uint8_t)v;
    }
    return len;
}

// writes header + CRC-8 (:128-141) into the zeroed bit buffer starting at bit 0; returns its length in bytes
ZF_DEVICE uint32_t write_header(SmemCommon &c, uint32_t *bits, unsigned long long frame_number, uint32_t depth,
                                uint32_t ch_type, uint32_t n, uint32_t sample_rate) {
    uint8_t hb[16];
    uint32_t len = build_header(hb, frame_number, depth, ch_type, n, sample_rate);
    uint32_t crc = 0;
    for (uint32_t k = 0; k < len; k++) crc = c.crc8tab[crc ^ hb[k]];
    hb[len++] = (uint8_t)crc;
    for (uint32_t k = 0; k < len; k++) {
        const uint32_t sh = 24u - 8u * (k & 3u);
        atomicOr(&bits[k >> 2], (uint32_t)hb[k] << sh);  // may share its last word with the first subframe
    }
    return len;
}

// Stereo frames: the header is prepared early (one thread, while others decide) for all four channel
// assignments; CRC-8 is linear, so each variant is the base CRC xor the contribution of its ch nibble.
ZF_DEVICE void prepare_stereo_headers(SmemCommon &c, unsigned long long frame_number, uint32_t depth, uint32_t n,
                                      uint32_t sample_rate) {
    uint8_t hb[16];
    const uint32_t len = build_header(hb, frame_number, depth, 0, n, sample_rate);
    uint32_t crc0 = 0;
    for (uint32_t k = 0; k < len; k++) crc0 = c.crc8tab[crc0 ^ hb[k]];
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint32_t k = 0; k < len; k++) w[k >> 2] |= (uint32_t)hb[k] << (24u - 8u * (k & 3u));
    const uint32_t types[4] = {1u, 8u, 9u, 10u};
    for (uint32_t v = 0; v < 4; v++) {
        uint32_t st = c.crc8tab[types[v] << 4];
        for (uint32_t k = 4; k < len; k++) st = c.crc8tab[st];  // the nibble sits in byte 3; len - 4 zero bytes follow
        const uint32_t crc = crc0 ^ st;
        uint32_t ww[4] = {w[0], w[1], w[2], w[3]};
        ww[0] |= types[v] << 4;  // byte 3 is the low byte of word 0
        ww[len >> 2] |= crc << (24u - 8u * (len & 3u));
        for (int k = 0; k < 4; k++) c.hdrw[v][k] = ww[k];
    }
    c.hdr_len = len + 1;
}

// ---------------------------------------------------------------------------------------------------
// PCM unpack: thread t gets samples [16t-4, 16t+16) of both channels (history first)
// ---------------------------------------------------------------------------------------------------

template <int BYTES>
ZF_DEVICE void unpack_stereo(const uint32_t *raw /* points at the pad */, int t, int32_t (&L)[kX], int32_t (&R)[kX]) {
    // stereo sample i occupies bytes [2*BYTES*i, 2*BYTES*(i+1)) after the pad
    if (BYTES == 2) {
        const uint32_t *p = raw + kRawPadWords + (kSpt * t - kHalo);  // one word per inter-channel sample
#pragma unroll
        for (int k = 0; k < kX; k += 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + k);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                L[k + q] = (int32_t)prmt(w[q], 0, 0x9910);
                R[k + q] = (int32_t)prmt(w[q], 0, 0xBB32);
            }
        }
    } else if (BYTES == 3) {
        // 6 bytes per inter-channel sample; kX samples = kX * 6 / 4 words, 8-byte aligned start
        const uint32_t *p = raw + kRawPadWords + (kSpt * t - kHalo) * 6 / 4;
        uint32_t w[kX * 6 / 4];
#pragma unroll
        for (int k = 0; k < kX * 6 / 4; k += 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(p + k);
            w[k] = v.x;
            w[k + 1] = v.y;
        }
        // two inter-channel samples (4 values) per 3 words
#pragma unroll
        for (int k = 0; k < kX / 2; k++) {
            const uint32_t a = w[3 * k], b = w[3 * k + 1], d = w[3 * k + 2];
            L[2 * k] = (int32_t)prmt(a, b, 0xA210);
            R[2 * k] = (int32_t)prmt(a, b, 0xD543);
            L[2 * k + 1] = (int32_t)prmt(b, d, 0xC432);
            R[2 * k + 1] = (int32_t)prmt(d, 0, 0xB321);
        }
    } else {
        const uint32_t *p = raw + kRawPadWords + (kSpt * t - kHalo) * 2;
#pragma unroll
        for (int k = 0; k < kX; k += 2) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + 2 * k);
            L[k] = (int32_t)v.x;
            R[k] = (int32_t)v.y;
            L[k + 1] = (int32_t)v.z;
            R[k + 1] = (int32_t)v.w;
        }
    }
}

// candidate channel `slot` of a stereo frame (encoder.zig:330-350)
template <bool WIDE>
ZF_DEVICE void make_x(uint32_t slot, const int32_t (&L)[kX], const int32_t (&R)[kX], typename Ar<WIDE>::T (&x)[kX]) {
    typedef typename Ar<WIDE>::T T;
    switch (slot) {
        case 0:
#pragma unroll
            for (int i = 0; i < kX; i++) x[i] = L[i];
            break;
        case 1:
#pragma unroll
            for (int i = 0; i < kX; i++) x[i] = R[i];
            break;
        case 2:
#pragma unroll
            for (int i = 0; i < kX; i++) x[i] = ((T)L[i] + (T)R[i]) >> 1;
            break;
        default:
#pragma unroll
            for (int i = 0; i < kX; i++) x[i] = (T)L[i] - (T)R[i];
            break;
    }
}

// ---------------------------------------------------------------------------------------------------
// pass 1: sample OR + five abs-sums (+ range ORs when WIDE), per thread
// ---------------------------------------------------------------------------------------------------

template <bool WIDE>
struct P1 {
    typename Ar<WIDE>::U s[5];
    typename Ar<WIDE>::U rng[5];
    typename Ar<WIDE>::U orv;
};

template <bool WIDE, bool FULL>
ZF_DEVICE void pass1(const typename Ar<WIDE>::T (&x)[kX], uint32_t base, uint32_t n, P1<WIDE> &p) {
    typedef typename Ar<WIDE>::T T;
    typedef typename Ar<WIDE>::U U;
    U s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, orv = 0;
    U r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    const T d32 = x[3] - x[2], d21 = x[2] - x[1], d10 = x[1] - x[0];
    T e1p = d32, e2p = d32 - d21, e3p = (d32 - d21) - (d21 - d10);
#pragma unroll
    for (uint32_t j = 0; j < (uint32_t)kSpt; j++) {
        const T e0 = x[kHalo + j];
        const T e1 = e0 - x[kHalo + j - 1];
        const T e2 = e1 - e1p;
        const T e3 = e2 - e2p;
        const T e4 = e3 - e3p;
        e1p = e1;
        e2p = e2;
        e3p = e3;
        const uint32_t i = base + j;
        const bool v = FULL || (i < n);
        const U a0 = uabs(e0), a1 = uabs(e1), a2 = uabs(e2), a3 = uabs(e3), a4 = uabs(e4);
        // total[k] sums i >= k only (fixed.zig:102-127); j >= k makes the test static for all but thread 0
        if (v) {
            orv |= (U)e0;
            s0 += a0;
            if (WIDE) r0 |= a0;
            if (j >= 1 || i >= 1) { s1 += a1; if (WIDE) r1 |= a1; }
            if (j >= 2 || i >= 2) { s2 += a2; if (WIDE) r2 |= a2; }
            if (j >= 3 || i >= 3) { s3 += a3; if (WIDE) r3 |= a3; }
            if (j >= 4 || i >= 4) { s4 += a4; if (WIDE) r4 |= a4; }
        }
    }
    p.s[0] = s0; p.s[1] = s1; p.s[2] = s2; p.s[3] = s3; p.s[4] = s4;
    p.rng[0] = r0; p.rng[1] = r1; p.rng[2] = r2; p.rng[3] = r3; p.rng[4] = r4;
    p.orv = orv;
}

// residual of a fixed order at x[kHalo + j] (fixed.zig:12-18 coefficients == finite differences)
template <typename T>
ZF_DEVICE T fixed_residual(const T (&x)[kX], uint32_t order, int j) {
    const int i = kHalo + j;
    switch (order) {
        case 0: return x[i];
        case 1: return x[i] - x[i - 1];
        case 2: return x[i] - 2 * x[i - 1] + x[i - 2];
        case 3: return x[i] - 3 * x[i - 1] + 3 * x[i - 2] - x[i - 3];
        default: return x[i] - 4 * x[i - 1] + 6 * x[i - 2] - 4 * x[i - 3] + x[i - 4];
    }
}

// ---------------------------------------------------------------------------------------------------
// decisions after pass 1 (one thread per slot): encoder.zig:482-527, fixed.zig:160-166, rice.zig:97-104
// ---------------------------------------------------------------------------------------------------

// cross-warp step of the pass-1 reduction: thread (slot, k) folds the kWarps partials
ZF_DEVICE void fold_red(SmemCommon &c, int t, uint32_t nslots) {
    if ((uint32_t)t < nslots * kRedVals) {
        const uint32_t s = (uint32_t)t / kRedVals, k = (uint32_t)t % kRedVals;
        unsigned long long v = 0;
        if (k < 5) {
#pragma unroll
            for (int w = 0; w < kWarps; w++) v += c.red[w][s][k];
        } else {
#pragma unroll
            for (int w = 0; w < kWarps; w++) v |= c.red[w][s][k];
        }
        c.tot[s][k] = v;
    }
}

template <bool WIDE>
ZF_DEVICE void decide_slot(SmemCommon &c, uint32_t slot, uint32_t depth_ch, uint32_t n, const FrameJob &job) {
    unsigned long long tot[5], rng[5];
    for (int k = 0; k < 5; k++) { tot[k] = c.tot[slot][k]; rng[k] = c.tot[slot][5 + k]; }
    unsigned long long orv = c.tot[slot][10];
    SlotDec d;
    d.depth = depth_ch;
    // calcWasteBits, encoder.zig:556-570: OR over the samples as integers of the plane's width
    if (!WIDE) orv &= 0xffffffffull;
    d.waste = (orv == 0) ? depth_ch : ctz64(orv);
    d.bps = depth_ch - d.waste;
    d.order = 0; d.mpo = 0; d.po = 0; d.method = 0; d.pad = 0;
    const uint32_t lim = d.bps > 16 ? 30u : 14u;
    d.max_param = lim < job.max_rice_param ? lim : job.max_rice_param;
    if (d.bps == 0) {  // :495-497
        d.kind = kConstant;
        d.est_bits = 0;
    } else if (tot[1] == 0) {  // all samples equal  <=>  sum |x[i]-x[i-1]| == 0, :498-500
        d.kind = kConstant;
        d.est_bits = d.bps;
    } else {
        d.kind = kVerbatim;  // :503-511
        d.est_bits = (unsigned long long)n * d.bps;
        if (n > 4) {  // :514
            // the sums were taken on un-shifted samples; every term is a multiple of 2^waste, so the
            // shift commutes with the sum (and with the OR used for the range test)
            const bool check = WIDE && d.bps >= 28;  // "wide accumulator", :517-520
            for (int k = 0; k < 5; k++) {
                tot[k] >>= d.waste;
                if (check && (rng[k] >> d.waste) > 0x7fffffffull) tot[k] = kU64Max;  // fixed.zig:160-162
            }
            uint32_t best = 0;
            for (uint32_t k = 1; k < 5; k++)
                if (tot[k] < tot[best]) best = k;  // first minimum, fixed.zig:164
            if (!(check && tot[best] == kU64Max)) {  // fixed.zig:166
                d.kind = kFixed;  // tentative: FIXED only if the Rice estimate beats VERBATIM (:538)
                d.order = best;
                // rice.calcParams, rice.zig:97-103
                uint32_t mpo = job.max_rice_order;
                const uint32_t tz = ctz32(n);
                if (tz < mpo) mpo = tz;
                if (best != 0) {
                    const uint32_t lim_o = floor_log2(n) - floor_log2(best);
                    if (lim_o < mpo) mpo = lim_o;
                }
                while ((n >> mpo) < best) mpo--;  // undefined upstream; same rule as the oracle
                d.mpo = mpo;
            }
        }
    }
    c.dec[slot] = d;
}

// ---------------------------------------------------------------------------------------------------
// rice.calcOptimalParams for one partition in closed form (rice.zig:343-395, flacCalcPartSize :402-405).
// Evaluation order escape, p = 0, 1, .. with strict '<' means: escape wins ties, then the lowest p.
// For p >= 1 the estimate f(p) = (1+p) n + (S >> (p-1)) - (n >> 1) is convex in p with
// f(p+1) - f(p) = n - ceil((S >> (p-1)) / 2), so its first minimum is the smallest p with (S >> (p-1)) <= 2n.
// ---------------------------------------------------------------------------------------------------

ZF_DEVICE void best_param(unsigned long long S, uint32_t B, uint32_t n, uint32_t P, uint32_t &choice,
                          unsigned long long &cost) {
    unsigned long long best = (B <= 31u) ? 5ull + (unsigned long long)B * n : kU64Max;
    uint32_t ch = 0x80u | B;
    if (P >= 1) {
        unsigned long long cc = (unsigned long long)n + (S << 1);  // p == 0: no -(n >> 1)
        uint32_t cand = 0;
        if (P >= 2) {
            const unsigned long long n2 = 2ull * n;
            uint32_t q = 0;
            if (S > n2) {
                q = bitlen64(S) - bitlen64(n2);
                if ((S >> q) > n2) q++;
            }
            uint32_t p = q + 1;
            if (p > P - 1) p = P - 1;
            const unsigned long long cp = (unsigned long long)(1 + p) * n + (S >> (p - 1)) - (unsigned long long)(n >> 1);
            if (cp < cc) { cand = p; cc = cp; }
        }
        if (cc < best) { best = cc; ch = cand; }
    }
    choice = ch;
    cost = best;
}

// ---------------------------------------------------------------------------------------------------
// subframe emission: MODE 0 counts the bits of this thread's part of a subframe, MODE 1 writes them.
// Grammar: frame_writer.zig:269-372.
// ---------------------------------------------------------------------------------------------------

template <bool WIDE, bool FULL, int MODE>
ZF_DEVICE uint32_t emit_plain(const typename Ar<WIDE>::T (&x)[kX], int t, uint32_t base, uint32_t n, const SlotDec &d,
                              uint32_t *bits, uint32_t pos) {
    typedef typename Ar<WIDE>::T T;
    BitWriter bw;
    const uint32_t start = pos;
    if (MODE == 1) bw.init(bits, pos);
    if (d.kind == kConstant) {  // :269-279: header 0x00, the un-shifted sample at full depth
        if (t == 0) {
            if (MODE == 1) {
                bw.put(pos, 0, 8);
                const unsigned long long v = (unsigned long long)(long long)x[kHalo] & (kU64Max >> (64 - d.depth));
                bw.put64(pos + 8, v, d.depth);
                bw.finish();
            }
            pos += 8 + d.depth;
        }
        return pos - start;
    }
    const uint32_t unary = d.waste;  // waste-1 zeros then a one; 0 bits when there is no waste
    if (d.kind == kVerbatim) {  // :282-301
        uint32_t p = pos;
        if (t == 0) {
            if (MODE == 1) {
                bw.put(p, d.waste ? 3u : 2u, 8);
                if (d.waste) bw.put(p + 8 + d.waste - 1, 1, 1);
            }
        }
        // every thread's samples sit at a fixed offset: no scan needed
        const uint32_t hdr = 8 + unary;
        if (MODE == 0) {
            uint32_t cnt = 0;
            if (FULL) cnt = kSpt;
            else if (base < n) cnt = (n - base) < (uint32_t)kSpt ? (n - base) : (uint32_t)kSpt;
            return cnt * d.bps + (t == 0 ? hdr : 0);
        }
        p = pos + (t == 0 ? hdr : 0);
        const unsigned long long mask = kU64Max >> (64 - d.bps);
#pragma unroll
        for (int j = 0; j < kSpt; j++) {
            if (FULL || base + j < n) {
                const unsigned long long v = (unsigned long long)((long long)x[kHalo + j] >> d.waste) & mask;
                bw.put64(p, v, d.bps);
                p += d.bps;
            }
        }
        bw.finish();
        return p - start;
    }
    return 0;
}

template <bool WIDE, bool FULL, int MODE>
ZF_DEVICE uint32_t emit_subframe(const typename Ar<WIDE>::T (&x)[kX], int t, uint32_t base, uint32_t n, const SlotDec &d,
                                 const uint8_t *choice_row, uint32_t *bits, uint32_t pos) {
    typedef typename Ar<WIDE>::T T;
    if (d.kind != kFixed) return emit_plain<WIDE, FULL, MODE>(x, t, base, n, d, bits, pos);
    BitWriter bw;
    const uint32_t start = pos;
    if (MODE == 1) bw.init(bits, pos);
    const uint32_t unary = d.waste;
    // FIXED :303-361
    const uint32_t order = d.order, waste = d.waste;
    const uint32_t psz = n >> d.po;
    const uint32_t param_len = 4 + d.method;
    const uint32_t esc_code = d.method ? 31u : 15u;
    if (t == 0) {
        if (MODE == 1) {
            bw.put(pos, ((8u | order) << 1) | (waste ? 1u : 0u), 8);
            if (waste) bw.put(pos + 8 + waste - 1, 1, 1);
        }
        pos += 8 + unary;
        const unsigned long long mask = kU64Max >> (64 - d.bps);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {  // warm-up samples (post-shift), :323-325
            if (k < order) {
                if (MODE == 1) bw.put64(pos, (unsigned long long)((long long)x[kHalo + k] >> waste) & mask, d.bps);
                pos += d.bps;
            }
        }
        if (MODE == 1) bw.put(pos, (d.method << 4) | d.po, 6);  // :328
        pos += 6;
    }
    const bool uniform = FULL || ((psz & (kSpt - 1)) == 0);
    uint32_t choice = 0;
    bool at_start = false;
    uint32_t part = 0, next = 0;  // partition of the sample in hand, first sample of the one behind it
    if (FULL || base < n) {
        part = FULL ? (base >> (12u - d.po)) : base / psz;
        next = (part + 1u) * psz;
        choice = choice_row[part];
        at_start = (base - part * psz) == 0;
    }
#pragma unroll
    for (int j = 0; j < kSpt; j++) {
        const uint32_t i = base + j;
        if (FULL || i < n) {
            bool hdr_here = at_start && j == 0;
            if (!uniform && j > 0) {  // partitions that are not whole numbers of threads: step over the boundary
                hdr_here = i >= next;
                if (hdr_here) {
                    part++;
                    next += psz;
                    choice = choice_row[part];
                }
            }
            const bool esc = (choice & 0x80u) != 0;
            if (hdr_here) {  // partition header: parameter, or escape code + 5-bit width (:341-357)
                if (MODE == 1) {
                    if (esc) {
                        bw.put(pos, esc_code, param_len);
                        bw.put(pos + param_len, choice & 0x7fu, 5);
                    } else {
                        bw.put(pos, choice, param_len);
                    }
                }
                pos += param_len + (esc ? 5u : 0u);
            }
            if (i >= order) {
                const T rt = fixed_residual<T>(x, order, j) >> waste;
                const int32_t r = (int32_t)rt;
                if (esc) {
                    const uint32_t wd = choice & 0x7fu;
                    if (wd) {
                        if (MODE == 1) bw.put(pos, (uint32_t)r & (0xffffffffu >> (32 - wd)), wd);
                        pos += wd;
                    }
                } else {
                    const uint32_t zz = zigzag(r);
                    const uint32_t q = zz >> choice;
                    if (MODE == 1) bw.put(pos + q, (1u << choice) | (zz & ((1u << choice) - 1u)), choice + 1);
                    pos += q + choice + 1;
                }
            }
        }
    }
    if (MODE == 1) bw.finish();
    return pos - start;
}

// block-wide exclusive scan of two values at once; returns totals through tot_a / tot_b
ZF_DEVICE void block_scan2(SmemCommon &c, int t, uint32_t a, uint32_t b, uint32_t &ex_a, uint32_t &ex_b,
                           uint32_t &tot_a, uint32_t &tot_b) {
    const int lane = t & 31, warp = t >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ua; ib += ub; }
    }
    if (lane == 31) { c.warp_scan[0][warp] = ia; c.warp_scan[1][warp] = ib; }
    __syncthreads();
    uint32_t oa = 0, ob = 0, ta = 0, tb = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        const uint32_t wa = c.warp_scan[0][w], wb = c.warp_scan[1][w];
        if (w < warp) { oa += wa; ob += wb; }
        ta += wa;
        tb += wb;
    }
    ex_a = oa + ia - a;
    ex_b = ob + ib - b;
    tot_a = ta;
    tot_b = tb;
    __syncthreads();  // warp_scan may be reused
}

// ---------------------------------------------------------------------------------------------------
// Rice analysis of the tentative FIXED slots: leaves, tree, parameter search, level pick.
// nslots candidate channels are processed together to share the barriers.
// ---------------------------------------------------------------------------------------------------

// One (partition, abs-sum, width) contribution per lane (part == 0xffffffff: none): the warp adds up the contributions of
// every partition its lanes name and issues ONE atomic pair per partition.  (Per-sample 64-bit shared-memory atomics -- a
// CAS loop, thirty-two lanes on one address -- were half of the one-CTA last-frame launch: 128 us for 4080 samples.)
ZF_DEVICE void warp_leaf_add(SmemCommon &c, uint32_t slot, uint32_t leaf0, uint32_t part, unsigned long long sum, uint32_t width,
                             int lane) {
    uint32_t pending = __ballot_sync(0xffffffffu, part != 0xffffffffu);
    while (pending) {  // warp-uniform
        const int src = (int)ctz32(pending);
        const uint32_t p = __shfl_sync(0xffffffffu, part, src);
        const bool in = part == p;
        const unsigned long long s = warp_sum(in ? sum : 0ull);
        const uint32_t w = reduce_max(in ? width : 0u);
        if (lane == src) {
            atomicAdd(&c.psum[slot][leaf0 + p], s);
            atomicMax(&c.pbits[slot][leaf0 + p], w);
        }
        pending &= ~__ballot_sync(0xffffffffu, in);
    }
}

// rice.calcSums leaves (rice.zig:288-340) of this thread's residuals r[j] (sample base + j; entries below the order or
// beyond n are ignored).  Partitions of exactly one thread are stored directly; otherwise a thread's samples are
// grouped into runs of one partition each (at most two when a partition holds eight samples or more) and the runs go
// through warp_leaf_add; partitions smaller than a thread (tiny frames only) fall back to per-sample atomics.
template <bool FULL>
ZF_DEVICE void rice_leaves_from(SmemCommon &c, uint32_t slot, const int32_t (&r)[kSpt], int t, uint32_t base, uint32_t n) {
    const SlotDec &d = c.dec[slot];
    const uint32_t order = d.order, mpo = d.mpo;
    const uint32_t psz = n >> mpo;
    const uint32_t leaf0 = (1u << mpo) - 1u;
    const int lane = t & 31;
    if (FULL || psz == (uint32_t)kSpt) {  // (FULL: 4096 samples, mpo 8 unless max_rice_order is lower -- then psz is a multiple of 8)
        if (psz == (uint32_t)kSpt) {
            unsigned long long sum = 0;
            int32_t mn = 0, mx = 0;
#pragma unroll
            for (int j = 0; j < kSpt; j++) {
                const uint32_t i = base + j;
                if ((FULL || i < n) && i >= order) {
                    sum += uabs(r[j]);
                    mn = r[j] < mn ? r[j] : mn;
                    mx = r[j] > mx ? r[j] : mx;
                }
            }
            const uint32_t zm = zigzag(mn), zx = zigzag(mx);
            if (FULL || base < n) {
                const uint32_t part = FULL ? (base >> (12u - mpo)) : base / psz;
                c.psum[slot][leaf0 + part] = sum;
                c.pbits[slot][leaf0 + part] = bitlen32(zm > zx ? zm : zx);  // bit length of the OR of the zigzags
            }
            return;
        }
    }
    if (psz >= (uint32_t)kSpt) {
        // at most two runs per thread
        uint32_t part[2] = {0xffffffffu, 0xffffffffu};
        unsigned long long sum[2] = {0, 0};
        int32_t mn[2] = {0, 0}, mx[2] = {0, 0};
        const uint32_t p0 = base < n ? base / psz : 0u;
        const uint32_t bound = (p0 + 1u) * psz;  // first sample of the next partition
#pragma unroll
        for (int j = 0; j < kSpt; j++) {
            const uint32_t i = base + j;
            if (i < n) {
                const uint32_t k = i >= bound ? 1u : 0u;
                part[k] = p0 + k;
                if (i >= order) {
                    sum[k] += uabs(r[j]);
                    mn[k] = r[j] < mn[k] ? r[j] : mn[k];
                    mx[k] = r[j] > mx[k] ? r[j] : mx[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t zm = zigzag(mn[k]), zx = zigzag(mx[k]);
            warp_leaf_add(c, slot, leaf0, part[k], sum[k], bitlen32(zm > zx ? zm : zx), lane);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < kSpt; j++) {
        const uint32_t i = base + j;
        if (i < n && i >= order) {
            const uint32_t part = i / psz;
            atomicAdd(&c.psum[slot][leaf0 + part], (unsigned long long)uabs(r[j]));
            atomicMax(&c.pbits[slot][leaf0 + part], bitlen32(zigzag(r[j])));
        }
    }
}

template <bool WIDE, bool FULL>
ZF_DEVICE void rice_leaves(SmemCommon &c, uint32_t slot, const typename Ar<WIDE>::T (&x)[kX], int t, uint32_t base,
                           uint32_t n) {
    typedef typename Ar<WIDE>::T T;
    const SlotDec &d = c.dec[slot];
    int32_t r[kSpt];
#pragma unroll
    for (int j = 0; j < kSpt; j++) r[j] = (int32_t)(fixed_residual<T>(x, d.order, j) >> d.waste);
    rice_leaves_from<FULL>(c, slot, r, t, base, n);
}

// zero the leaves that rice_leaves accumulates into with atomics
ZF_DEVICE void rice_zero_leaves(SmemCommon &c, uint32_t slot, int t, uint32_t n) {
    const SlotDec &d = c.dec[slot];
    if (d.kind != kFixed) return;
    const uint32_t psz = n >> d.mpo;
    if (psz == (uint32_t)kSpt) return;  // direct stores
    const uint32_t leaf0 = (1u << d.mpo) - 1u, cnt = 1u << d.mpo;
    for (uint32_t j = t; j < cnt; j += kThreads) {
        c.psum[slot][leaf0 + j] = 0;
        c.pbits[slot][leaf0 + j] = 0;
    }
}

ZF_DEVICE void rice_tree_and_search(SmemCommon &c, int t, uint32_t n, uint32_t nslots) {
    const int lane = t & 31;
    // partition tree, rice.zig:331-339; all slots advance one level per barrier
    for (uint32_t step = 1; step <= (uint32_t)kMaxLevel; step++) {
        for (uint32_t s = 0; s < nslots; s++) {
            const SlotDec &d = c.dec[s];
            if (d.kind != kFixed || d.mpo < step) continue;
            const uint32_t lvl = d.mpo - step;
            const uint32_t b0 = (1u << lvl) - 1u, b1 = (2u << lvl) - 1u;
            for (uint32_t j = t; j < (1u << lvl); j += kThreads) {
                c.psum[s][b0 + j] = c.psum[s][b1 + 2 * j] + c.psum[s][b1 + 2 * j + 1];
                const uint32_t a = c.pbits[s][b1 + 2 * j], b = c.pbits[s][b1 + 2 * j + 1];
                c.pbits[s][b0 + j] = a > b ? a : b;
            }
        }
        __syncthreads();
    }
    if (t < 4 * (kMaxLevel + 1)) c.levelcost[t / (kMaxLevel + 1)][t % (kMaxLevel + 1)] = 0;
    if (t < 4) c.levelfive[t] = 0;
    __syncthreads();
    // parameter search: node m = 1 .. 2^(mpo+1)-1 in heap numbering (level = floor(log2 m))
    for (uint32_t s = 0; s < nslots; s++) {
        const SlotDec &d = c.dec[s];
        if (d.kind != kFixed) continue;
        const uint32_t last = (2u << d.mpo) - 1u;
        for (uint32_t m0 = 0; m0 <= last; m0 += kThreads) {
            const uint32_t m = m0 + t;
            const bool act = m >= 1 && m <= last;
            uint32_t lvl = 0, choice = 0;
            unsigned long long cost = 0;
            if (act) {
                lvl = floor_log2(m);
                const uint32_t j = m - (1u << lvl);
                const uint32_t psz = n >> lvl;
                const uint32_t cnt = psz - (j == 0 ? d.order : 0u);  // first partition: rice.zig:356,371
                best_param(c.psum[s][m - 1], c.pbits[s][m - 1], cnt, d.max_param, choice, cost);
                c.pchoice[s][m - 1] = (uint8_t)choice;
            }
            const bool five = act && choice < 0x80u && choice > 14u;  // isRice2, rice.zig:74-76
            if (m0 == 0 && t < 32) {
                // warp 0 of the first round holds levels 0..4 mixed: one masked warp sum per level (nobody else touches
                // these levels; 64-bit shared-memory atomics are CAS loops)
#pragma unroll 1
                for (uint32_t q = 0; q < 5; q++) {
                    const unsigned long long ws = warp_sum((act && lvl == q) ? cost : 0ull);
                    const uint32_t wf = reduce_or((five && lvl == q) ? 1u : 0u);
                    if (lane == 0 && q <= d.mpo) {
                        c.levelcost[s][q] += ws;
                        if (wf) atomicOr(&c.levelfive[s], 1u << q);
                    }
                }
            } else {
                // whole warp is in one level (or entirely inactive)
                const unsigned long long wsum = warp_sum(cost);
                const uint32_t wfive = reduce_or(five ? 1u : 0u);
                const uint32_t wl = __shfl_sync(0xffffffffu, lvl, 0);
                const uint32_t wact = __shfl_sync(0xffffffffu, act ? 1u : 0u, 0);
                if (lane == 0 && wact) {
                    atomicAdd(&c.levelcost[s][wl], wsum);
                    if (wfive) atomicOr(&c.levelfive[s], 1u << wl);
                }
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------
// EXTENSION (zf_config.exact_rice): the Rice parameter of every partition by its exact code length
// sum (zigzag >> p) + n (p + 1) instead of the estimate of rice.zig:402-405.  The reference only has dead code for this
// (calcParamExact, rice.zig:110-245: never called, no escapes), so the rule is ours: same evaluation order and ties
// as the live search (escape first, then p = 0 .. P-1 with strict '<').
//
// sum_j (zz_j >> p) = sum_{b >= p} cnt_b 2^(b-p) with cnt_b = how many zigzags have bit b set, so 32 counters per
// partition -- additive over the partition tree -- give the exact length for every p by E_p = cnt_p + 2 E_{p+1}.
// They are kept as sixteen words of two 16-bit counters (bit i low, bit i + 16 high; a frame has at most 4096 samples).
// One candidate at a time through `cnt` (heap numbering as psum / pbits).
// ---------------------------------------------------------------------------------------------------
ZF_DEVICE void rice_exact_slot(SmemCommon &c, uint32_t (*cnt)[16], uint32_t s, const int32_t (&r)[kSpt], int t, uint32_t base,
                               uint32_t n) {
    const int lane = t & 31;
    const SlotDec d = c.dec[s];
    const uint32_t order = d.order, mpo = d.mpo;
    const uint32_t psz = n >> mpo;
    const uint32_t leaf0 = (1u << mpo) - 1u;
    for (uint32_t idx = t; idx < (16u << mpo); idx += kThreads) cnt[leaf0 + (idx >> 4)][idx & 15u] = 0;
    if (t <= kMaxLevel) c.levelcost[s][t] = 0;
    if (t == 0) c.levelfive[s] = 0;
    __syncthreads();
    {   // leaves: a thread's samples fall into runs of one partition each
        uint32_t zz[kSpt], part[kSpt];
        const uint32_t p0 = base < n ? base / psz : 0u;
        uint32_t pp = p0, next = (p0 + 1u) * psz;
#pragma unroll
        for (int j = 0; j < kSpt; j++) {
            const uint32_t i = base + j;
            if (i < n && i >= next) { pp++; next += psz; }
            const bool v = i < n && i >= order;
            zz[j] = v ? zigzag(r[j]) : 0u;
            part[j] = pp;
        }
        // whole warp inside one partition (the usual case for the coarse geometries of short frames): one REDUX per word
        const uint32_t il = (base + kSpt <= n) ? base + kSpt - 1u : n - 1u;  // this thread's last sample
        const bool one = il < (p0 + 1u) * psz;
        const uint32_t lead = __shfl_sync(0xffffffffu, p0, 0);
        const uint32_t live = __ballot_sync(0xffffffffu, base < n);
        const bool uniform = live != 0u && __ballot_sync(0xffffffffu, base < n && one && p0 == lead) == live;
#pragma unroll 1
        for (uint32_t w = 0; w < 16; w++) {
            if (uniform) {
                uint32_t a = 0;
#pragma unroll
                for (int j = 0; j < kSpt; j++) a += (zz[j] >> w) & 0x00010001u;
                a = reduce_add(base < n ? a : 0u);  // at most 32 x 8 per half: no carry between the halves
                if (lane == 0 && a) atomicAdd(&cnt[leaf0 + lead][w], a);
            } else {
                uint32_t a = 0, cur = p0;
#pragma unroll
                for (int j = 0; j < kSpt; j++) {
                    if (part[j] != cur) {
                        if (a) atomicAdd(&cnt[leaf0 + cur][w], a);
                        a = 0;
                        cur = part[j];
                    }
                    a += (zz[j] >> w) & 0x00010001u;
                }
                if (a) atomicAdd(&cnt[leaf0 + cur][w], a);
            }
        }
    }
    __syncthreads();
    for (uint32_t step = 1; step <= mpo; step++) {  // partition tree: counters add
        const uint32_t lvl = mpo - step;
        const uint32_t b0 = (1u << lvl) - 1u, b1 = (2u << lvl) - 1u;
        for (uint32_t idx = t; idx < (16u << lvl); idx += kThreads) {
            const uint32_t j = idx >> 4, w = idx & 15u;
            cnt[b0 + j][w] = cnt[b1 + 2 * j][w] + cnt[b1 + 2 * j + 1][w];
        }
        __syncthreads();
    }
    const uint32_t last = (2u << mpo) - 1u;
    for (uint32_t m0 = 0; m0 <= last; m0 += kThreads) {
        const uint32_t m = m0 + t;
        const bool act = m >= 1 && m <= last;
        uint32_t lvl = 0, choice = 0;
        unsigned long long cost = 0;
        if (act) {
            lvl = floor_log2(m);
            const uint32_t j = m - (1u << lvl);
            const uint32_t ne = (n >> lvl) - (j == 0 ? order : 0u);  // rice.zig:356,371
            uint32_t cb[32];
#pragma unroll
            for (int w = 0; w < 16; w++) {
                const uint32_t v = cnt[m - 1][w];
                cb[w] = v & 0xffffu;
                cb[w + 16] = v >> 16;
            }
            uint32_t B = 0;
#pragma unroll
            for (int b = 0; b < 32; b++)
                if (cb[b]) B = (uint32_t)b + 1u;
            unsigned long long E = 0, bestc = kU64Max;
            uint32_t bestp = 0;
#pragma unroll
            for (int b = 31; b >= 0; b--) {  // downwards with '<=': the lowest parameter wins ties, as the upward '<' scan does
                E = (unsigned long long)cb[b] + 2ull * E;
                if ((uint32_t)b < d.max_param) {
                    const unsigned long long cc = E + (unsigned long long)ne * ((uint32_t)b + 1u);
                    if (cc <= bestc) { bestc = cc; bestp = (uint32_t)b; }
                }
            }
            const unsigned long long esc = (B <= 31u) ? 5ull + (unsigned long long)B * ne : kU64Max;
            if (bestc < esc) { cost = bestc; choice = bestp; }  // the escape wins ties (it is evaluated first upstream)
            else { cost = esc; choice = 0x80u | B; }
            c.pchoice[s][m - 1] = (uint8_t)choice;
        }
        const bool five = act && choice < 0x80u && choice > 14u;
        if (m0 == 0 && t < 32) {
#pragma unroll 1
            for (uint32_t q = 0; q < 5; q++) {
                const unsigned long long ws = warp_sum((act && lvl == q) ? cost : 0ull);
                const uint32_t wf = reduce_or((five && lvl == q) ? 1u : 0u);
                if (lane == 0 && q <= mpo) {
                    c.levelcost[s][q] += ws;
                    if (wf) atomicOr(&c.levelfive[s], 1u << q);
                }
            }
        } else {
            const unsigned long long wsum = warp_sum(cost);
            const uint32_t wfive = reduce_or(five ? 1u : 0u);
            const uint32_t wl = __shfl_sync(0xffffffffu, lvl, 0);
            const uint32_t wact = __shfl_sync(0xffffffffu, act ? 1u : 0u, 0);
            if (lane == 0 && wact) {
                atomicAdd(&c.levelcost[s][wl], wsum);
                if (wfive) atomicOr(&c.levelfive[s], 1u << wl);
            }
        }
    }
    __syncthreads();
}

// pick the partition order (rice.zig:262-276: '<=' keeps the highest on ties) and FIXED vs VERBATIM (:538)
ZF_DEVICE void finish_slot(SmemCommon &c, uint32_t s, uint32_t n) {
    SlotDec &d = c.dec[s];
    if (d.kind != kFixed) return;
    unsigned long long best = kU64Max;
    uint32_t bpo = 0, bmethod = 0;
    for (uint32_t lvl = 0; lvl <= d.mpo; lvl++) {
        const uint32_t method = (c.levelfive[s] >> lvl) & 1u;
        const unsigned long long bc = c.levelcost[s][lvl] + ((unsigned long long)(4 + method) << lvl);  // :394
        if (bc <= best) { best = bc; bpo = lvl; bmethod = method; }
    }
    const unsigned long long verb = (unsigned long long)n * d.bps;
    if (best < verb) {
        d.est_bits = best;
        d.po = bpo;
        d.method = bmethod;
    } else {
        d.kind = kVerbatim;
        d.est_bits = verb;
    }
}

// ---------------------------------------------------------------------------------------------------
// frame finish: CRC-16, look-back, copy-out.  total_bits = bits before byte padding.
// ---------------------------------------------------------------------------------------------------

ZF_DEVICE void finish_frame(SmemCommon &c, uint32_t *bits, int t, const FrameJob &job, uint32_t fidx,
                            uint32_t total_bits, bool fits) {
    const int lane = t & 31, warp = t >> 5;
    const uint32_t fbytes = (total_bits + 7u) >> 3;  // zero padded to a byte (frame_writer.zig:114-117)
    const uint32_t size = fbytes + 2u;
    if (warp == 0) {
        // --- look-back over the frame-size descriptors of the batch (single-pass stream compaction) ---
        unsigned long long excl = 0;
        if (fidx > 0) {
            long long i = (long long)fidx - 1;
            for (;;) {
                const long long idx = i - lane;
                const unsigned long long dsc = (idx >= 0) ? ld_relaxed_gpu(job.desc + idx) : kFlagPrefix;
                const uint32_t flag = (uint32_t)(dsc >> 62);
                const uint32_t pmask = __ballot_sync(0xffffffffu, flag == 2u);
                const uint32_t inval = __ballot_sync(0xffffffffu, flag == 0u);
                const uint32_t first_p = pmask ? ctz32(pmask) : 32u;
                const uint32_t need = first_p >= 31u ? 0xffffffffu : ((2u << first_p) - 1u);
                if (inval & need) continue;  // a predecessor has not published yet: poll again
                const unsigned long long v = ((uint32_t)lane <= first_p) ? (dsc & kValueMask) : 0ull;
                excl += warp_sum(v);
                if (first_p < 32u) break;
                i -= 32;
            }
        }
        if (lane == 0) {
            c.out_off = excl;
            st_relaxed_gpu(job.desc + fidx, kFlagPrefix | (excl + size));
            if (fidx + 1 == job.batch_frames) *job.total_bytes = excl + size;
            if (excl + size > job.out_cap) atomicOr(job.status, kStatusOutOverflow);
        }
    } else {
        // --- CRC-16 over fbytes (warps 1..): 64-byte chunks, each multiplied by x^(8 * bytes-after-it) mod P ---
        uint32_t contrib = 0;
        const uint32_t nchunks = fits ? ((fbytes + 63u) >> 6) : 0u;
        for (uint32_t ck = (uint32_t)t - 32u; ck < nchunks; ck += kThreads - 32) {
            const uint32_t cb = (fbytes - (ck << 6)) < 64u ? (fbytes - (ck << 6)) : 64u;
            const uint32_t *p = bits + (ck << 4);
            uint32_t crc = 0;
            if (cb == 64u) {
#pragma unroll
                for (uint32_t w = 0; w < 16; w += 4) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(p + w);
                    crc = crc16_word(c, crc, v.x);
                    crc = crc16_word(c, crc, v.y);
                    crc = crc16_word(c, crc, v.z);
                    crc = crc16_word(c, crc, v.w);
                }
            } else {
                const uint32_t words = cb >> 2, tail = cb & 3u;
                for (uint32_t w = 0; w < words; w++) crc = crc16_word(c, crc, p[w]);
                for (uint32_t k = 0; k < tail; k++) crc = crc16_byte(c, crc, (p[words] >> (24u - 8u * k)) & 0xffu);
            }
            const uint32_t dist = fbytes - ((ck << 6) + cb);
            contrib ^= crc16_mulmod(c, crc, job.pow8[dist]);
        }
        contrib = reduce_xor(contrib);
        if (lane == 0) c.crc_part[warp] = contrib;
    }
    __syncthreads();
    if (t == 0 && fits) {  // append CRC-16 big-endian at byte fbytes (frame_writer.zig:144-148)
        uint32_t crc = 0;
        for (int w = 1; w < kWarps; w++) crc ^= c.crc_part[w];
        const uint32_t bp = fbytes << 3;
        const uint32_t wi = bp >> 5, off = bp & 31u;
        if (off <= 16u) bits[wi] |= crc << (16u - off);
        else { bits[wi] |= crc >> (off - 16u); bits[wi + 1] |= crc << (48u - off); }
    }
    __syncthreads();
    // --- byte-shifted copy shared -> global: head bytes, aligned 32-bit words, tail bytes ---
    const unsigned long long off = c.out_off;
    if (!fits || off + size > job.out_cap) return;
    uint8_t *dst = job.out + off;
    const uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
    const uint32_t h = head < size ? head : size;
    if ((uint32_t)t < h) dst[t] = (uint8_t)(bits[t >> 2] >> (24u - 8u * (t & 3u)));
    const uint32_t nwords = (size - h) >> 2;
    uint32_t *dw = reinterpret_cast<uint32_t *>(dst + h);
    const uint32_t sh = 8u * h;  // h < 4
    for (uint32_t k = t; k < nwords; k += kThreads) {
        // source bytes h + 4k .. h + 4k + 3 : big-endian words k and k+1
        const uint32_t be = h ? __funnelshift_l(bits[k + 1], bits[k], sh) : bits[k];
        dw[k] = prmt(be, 0, 0x0123);
    }
    const uint32_t done = h + (nwords << 2);
    const uint32_t rem = size - done;  // < 4
    if ((uint32_t)t < rem) {
        const uint32_t b = done + t;
        dst[b] = (uint8_t)(bits[b >> 2] >> (24u - 8u * (b & 3u)));
    }
}

// ---------------------------------------------------------------------------------------------------
// the stereo kernel (2 channels, decorrelation on): candidates L, R, M, S
// ---------------------------------------------------------------------------------------------------

template <int BYTES, bool FULL>
ZF_DEVICE void load_raw_generic(uint32_t *raw, const uint8_t *src, uint32_t nbytes, int t) {
    // partial / odd-sized frames: plain loads (the TMA bulk copy needs 16-byte multiples)
    uint8_t *dst = reinterpret_cast<uint8_t *>(raw + kRawPadWords);
    if ((((uintptr_t)src) & 3u) == 0) {
        const uint32_t words = nbytes >> 2;
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
        for (uint32_t k = t; k < words; k += kThreads) raw[kRawPadWords + k] = s32[k];
        for (uint32_t k = (words << 2) + t; k < nbytes; k += kThreads) dst[k] = src[k];
    } else {
        for (uint32_t k = t; k < nbytes; k += kThreads) dst[k] = src[k];
    }
}

// with the exact Rice search (extension): the partition tree's bit counters behind the ordinary layout
template <int BYTES>
struct SmemStereoExact {
    SmemStereo<BYTES> base;
    alignas(16) uint32_t cnt[kNodes][16];
};

template <int BYTES, bool FULL, bool EXACT = false>
__global__ void __launch_bounds__(kThreads, EXACT ? 1 : 2) zf_encode_stereo_kernel(const FrameJob job) {
    constexpr bool WIDE = (BYTES == 4);
    typedef typename Ar<WIDE>::T T;
    extern __shared__ __align__(16) unsigned char zf_smem[];
    SmemStereo<BYTES> &sm = *reinterpret_cast<SmemStereo<BYTES> *>(zf_smem);
    SmemCommon &c = sm.c;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n = FULL ? (uint32_t)kMaxBlock : job.block_size;
    const uint32_t depth = job.bit_depth ? job.bit_depth : 8u * BYTES;
    const uint32_t frame_bytes = n * 2u * BYTES;
    const uint32_t base = (uint32_t)t * kSpt;
    const bool tma = FULL && job.use_tma;

    if (job.pdl_trigger) pdl_launch_dependents();
    init_tables(c, t);
    if (t < kRawPadWords) sm.raw[t] = 0;
    if (t == 0) {
        if (tma) {
            mbar_init(&c.mbar, 1);
            fence_mbar_init();
        }
        const uint32_t f = atomicAdd(job.ticket, 1u);
        c.cur_frame = f;
        if (tma && f < job.n_frames) {
            mbar_expect_tx(&c.mbar, frame_bytes);
            tma_load_1d(sm.raw + kRawPadWords, job.pcm + (size_t)f * job.frame_stride, frame_bytes, &c.mbar);
        }
    }
    __syncthreads();
    uint32_t phase = 0;

    for (;;) {
        const uint32_t f = c.cur_frame;
        if (f >= job.n_frames) break;
        const uint32_t fidx = job.frame_base + f;
        const unsigned long long frame_number = job.first_frame_number + fidx;

        if (tma) {
            mbar_wait(&c.mbar, phase);
            phase ^= 1u;
        } else {
            load_raw_generic<BYTES, FULL>(sm.raw, job.pcm + (size_t)f * job.frame_stride, frame_bytes, t);
            __syncthreads();
        }
        int32_t L[kX], R[kX];
        unpack_stereo<BYTES>(sm.raw, t, L, R);
        // zero the bit buffer for this frame
        {
            uint4 *bz = reinterpret_cast<uint4 *>(sm.bits);
            const uint4 z = {0, 0, 0, 0};
            for (int k = t; k < BitBufWords<BYTES>::value / 4; k += kThreads) bz[k] = z;
        }
        __syncthreads();  // raw fully consumed: prefetch the next frame into it
        if (t == 0) {
            const uint32_t nf = atomicAdd(job.ticket, 1u);
            c.next_frame = nf;
            if (tma && nf < job.n_frames) {
                fence_proxy_async();
                mbar_expect_tx(&c.mbar, frame_bytes);
                tma_load_1d(sm.raw + kRawPadWords, job.pcm + (size_t)nf * job.frame_stride, frame_bytes, &c.mbar);
            }
        }

        // ---- pass 1 over the four candidates ----
        T x[kX];
#pragma unroll 1
        for (uint32_t s = 0; s < 4; s++) {
            make_x<WIDE>(s, L, R, x);
            P1<WIDE> p;
            pass1<WIDE, FULL>(x, base, n, p);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const unsigned long long ws = warp_sum(p.s[k]);
                if (lane == 0) c.red[warp][s][k] = ws;
                if (WIDE) {
                    const unsigned long long wr = warp_or(p.rng[k]);
                    if (lane == 0) c.red[warp][s][5 + k] = wr;
                } else if (lane == 0) {
                    c.red[warp][s][5 + k] = 0;
                }
            }
            const unsigned long long wo = warp_or(p.orv);
            if (lane == 0) c.red[warp][s][10] = wo;
        }
        __syncthreads();
        fold_red(c, t, 4);
        __syncthreads();
        if (t < 4) decide_slot<WIDE>(c, (uint32_t)t, depth + (t == 3 ? 1u : 0u), n, job);
        __syncthreads();

        // ---- pass 2: Rice analysis of the tentative FIXED candidates ----
        if constexpr (EXACT) {
            uint32_t (*cnt)[16] = reinterpret_cast<SmemStereoExact<BYTES> *>(zf_smem)->cnt;
#pragma unroll 1
            for (uint32_t s = 0; s < 4; s++) {
                if (c.dec[s].kind != kFixed) continue;  // block-uniform
                make_x<WIDE>(s, L, R, x);
                int32_t r[kSpt];
#pragma unroll
                for (int j = 0; j < kSpt; j++) r[j] = (int32_t)(fixed_residual<T>(x, c.dec[s].order, j) >> c.dec[s].waste);
                rice_exact_slot(c, cnt, s, r, t, base, n);
            }
        } else {
        if (!FULL || job.max_rice_order != (uint32_t)kMaxLevel) {
            for (uint32_t s = 0; s < 4; s++) rice_zero_leaves(c, s, t, n);
            __syncthreads();
        }
#pragma unroll 1
        for (uint32_t s = 0; s < 4; s++) {
            if (c.dec[s].kind != kFixed) continue;
            make_x<WIDE>(s, L, R, x);
            rice_leaves<WIDE, FULL>(c, s, x, t, base, n);
        }
        __syncthreads();
        rice_tree_and_search(c, t, n, 4);
        }
        if (t == 0) {
            for (uint32_t s = 0; s < 4; s++) finish_slot(c, s, n);
            // stereo mode: first minimum of [L+R, L+S, S+R, M+S], encoder.zig:441-452
            const unsigned long long el = c.dec[0].est_bits, er = c.dec[1].est_bits, em = c.dec[2].est_bits,
                                     es = c.dec[3].est_bits;
            const unsigned long long sum[4] = {el + er, el + es, es + er, em + es};
            uint32_t best = 0;
            for (uint32_t k = 1; k < 4; k++)
                if (sum[k] < sum[best]) best = k;
            uint32_t a = 0, b = 1, cht = 1;  // Channel.indep(2) = 1, type.zig:7-12
            if (best == 1) { a = 0; b = 3; cht = 8; }
            else if (best == 2) { a = 3; b = 1; cht = 9; }
            else if (best == 3) { a = 2; b = 3; cht = 10; }
            c.sub_slot[0] = a;
            c.sub_slot[1] = b;
            c.ch_type = cht;
        }
        __syncthreads();

        // ---- pack ----
        const uint32_t hdr_bits = 8u * header_len(frame_number, n, job.sample_rate);
        uint32_t len_a = 0, len_b = 0;
#pragma unroll 1
        for (uint32_t k = 0; k < 2; k++) {
            const uint32_t sk = c.sub_slot[k];
            const SlotDec dk = c.dec[sk];
            make_x<WIDE>(sk, L, R, x);
            const uint32_t v = emit_subframe<WIDE, FULL, 0>(x, t, base, n, dk, &c.pchoice[sk][(1u << dk.po) - 1u], sm.bits, 0);
            if (k == 0) len_a = v;
            else len_b = v;
        }
        uint32_t ex_a, ex_b, tot_a, tot_b;
        block_scan2(c, t, len_a, len_b, ex_a, ex_b, tot_a, tot_b);
        const uint32_t total_bits = hdr_bits + tot_a + tot_b;
        const uint32_t fbytes = (total_bits + 7u) >> 3;
        const bool fits = (fbytes + 2u) <= (uint32_t)BitBufWords<BYTES>::value * 4u - 8u;
        if (t == 0) {
            // publish the frame size early so successors can look back through this frame
            const unsigned long long size = fbytes + 2u;
            job.frame_sizes[fidx] = (uint32_t)size;
            if (fidx == 0) st_relaxed_gpu(job.desc, kFlagPrefix | size);
            else st_relaxed_gpu(job.desc + fidx, kFlagAggregate | size);
            if (!fits) atomicOr(job.status, kStatusBitOverflow);
            write_header(c, sm.bits, frame_number, depth, c.ch_type, n, job.sample_rate);
        }
        __syncthreads();
        if (fits) {
#pragma unroll 1
            for (uint32_t k = 0; k < 2; k++) {
                const uint32_t sk = c.sub_slot[k];
                const SlotDec dk = c.dec[sk];
                make_x<WIDE>(sk, L, R, x);
                const uint32_t pos = hdr_bits + (k == 0 ? ex_a : tot_a + ex_b);
                emit_subframe<WIDE, FULL, 1>(x, t, base, n, dk, &c.pchoice[sk][(1u << dk.po) - 1u], sm.bits, pos);
            }
        }
        __syncthreads();
        finish_frame(c, sm.bits, t, job, fidx, total_bits, fits);
        __syncthreads();
        if (t == 0) c.cur_frame = c.next_frame;
        __syncthreads();
    }
    if (t == 0) pdl_wait_primary();
}

// ---------------------------------------------------------------------------------------------------
// The short last frame of a stream is encoded by a one-CTA launch that runs concurrently with the
// full-frame kernel (its own stream, private output); this kernel appends it to the compacted stream.
// ---------------------------------------------------------------------------------------------------
__global__ void zf_append_tail_kernel(const uint8_t *tail, const uint32_t *tail_size, uint8_t *out,
                                      unsigned long long out_cap, unsigned long long *total, uint32_t *frame_sizes,
                                      uint32_t fidx, unsigned int *status) {
    const uint32_t size = *tail_size;
    const unsigned long long off = *total;
    const bool fits = off + size <= out_cap;  // like every frame: one that does not fit is counted, not written
    if (fits) {
        // 16 source bytes per thread and trip (the private buffer is aligned; the destination is wherever the stream ends)
        const uint4 *t16 = reinterpret_cast<const uint4 *>(tail);
        uint8_t *dst = out + off;
        for (uint32_t k = threadIdx.x; 16u * k < size; k += blockDim.x) {
            const uint4 v = t16[k];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const uint32_t left = size - 16u * k;
#pragma unroll
            for (uint32_t b = 0; b < 16; b++)
                if (b < left) dst[16u * k + b] = (uint8_t)(w[b >> 2] >> (8u * (b & 3u)));
        }
    }
    __syncthreads();  // every thread has read *total
    if (threadIdx.x == 0) {
        if (!fits) atomicOr(status, kStatusOutOverflow);
        frame_sizes[fidx] = size;
        *total = off + size;
    }
}

// 8-bit samples (signed, one byte each, as WavReader.fillSamples leaves them -- wav_reader.zig:71-88) -> 16-bit little-endian
// containers, so that the 16-bit kernels encode them with depth 8
__global__ void zf_widen8_kernel(const int8_t *in, int16_t *out, unsigned long long n) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = (int16_t)in[i];
}

}  // namespace zf

// zf_kernel_v3.cuh -- lean full-frame stereo kernel for 16-, 24- and 32-bit PCM (block 4096; any max_rice_order, max_rice_param
// and sample rate).  Same decisions and the same bytes as
// zf_kernel.cuh / the oracle, arranged around the B200's issue limits (the path is integer-issue bound, not HBM
// bound -- DESIGN.md section 4):
//
//   * 256 threads x 16 samples: a thread owns exactly one finest Rice partition (level 8), so leaf statistics
//     never leave registers; 3 CTAs per SM overlap each other's barriers; 80 registers, no spills
//   * samples are NOT kept in registers across phases: every phase re-reads its window from the raw PCM in shared
//     memory; pass 1 streams it four samples at a time through one rolled piece of code (all four candidates side
//     by side) -- the kernel's code size matters as much as its instruction count (32 KB instruction cache)
//   * all arithmetic is 32-bit (|delta^4| < 2^28, 16-sample sums < 2^32); only tree sums above a leaf are 64-bit
//   * decisions that every thread needs (partition order, method, FIXED vs VERBATIM, stereo mode) are computed
//     redundantly by every warp from shared per-level partial sums: no decision barrier, no broadcast
//   * bit writer: right-aligned 64-bit shift register; a word is stored (plain, predicated) by the one thread whose
//     bit range crosses the word's end, partial tail words are ORed in after a barrier, so the buffer is never
//     cleared (only the header words and a partial last word are); the two subframes are written as two interleaved
//     dependency chains, and every kind of field (Rice code, escaped or VERBATIM raw bits) goes through the same
//     branch-free loop -- a raw field is a Rice code with k = 32
//   * the stream is placed in the bit buffer so that it ENDS on a 16-byte boundary (leading zero bytes do not change
//     a CRC with zero init), CRC-16 is computed table-free in GF(2)[x]/(x^15+x+1) x parity
//     (x^16+x^15+x^2+1 = (x+1)(x^15+x+1)): Horner step  a <- a*(x^4+x^2) + word,  x^32 = x^4+x^2 (mod x^15+x+1),
//     two words per dependent step; chunks are combined by a carry-less multiply done with integer multiplies
//   * long dependent chains (parameter search, shuffle tree) run two candidates at a time; FIXED or not, every
//     candidate is evaluated -- no early exits inside a warp
//   * deferred epilogue: a finished frame stays in the bit buffer while the next one is analysed; its output offset
//     comes from a decoupled look-back whose descriptor window is fetched with cp.async (no registers, no waiting),
//     and it is copied out (16-byte stores) just before the bit buffer is needed again
//
// Reference call stack mirrored per frame: Encoder.writeFrame, encoder.zig:234-284 (see zf_kernel.cuh).
#pragma once
#include "zf_kernel.cuh"

#ifdef ZF_HOST_EMU
#include <math.h>
#define ZF_NOINLINE inline
#else
#define ZF_NOINLINE __device__ __noinline__
#endif

namespace zf {
namespace v3 {

#ifdef ZF_HOST_EMU
static unsigned long long g_emu_wide_frames = 0, g_emu_narrow_frames = 0;
#endif

constexpr int kT = 256;          // threads per CTA
constexpr int kS = 16;           // samples per thread == finest partition (4096 >> 8)
constexpr int kH = 4;            // history samples in front of a thread's window
constexpr int kXn = kS + kH;
constexpr int kN = 4096;
constexpr int kW = kT / 32;

// Raw PCM in shared memory.  Thread t reads "its" row (its sixteen inter-channel samples: 64 / 96 / 128 bytes) with
// 16-byte loads, and so do its neighbours at the same moment: rows that simply follow one another would put the eight
// threads of a quarter-warp into 4 / 2 / 1 of the eight 16-byte bank groups (an 8-way conflict on every LDS.128 of
// 32-bit PCM).  So the frame is laid down in groups of 2 / 4 / 1 rows (128 / 384 / 128 bytes) with 16 bytes of padding
// behind every group -- one 1-D bulk copy per group -- which sends any eight consecutive rows to eight different bank
// groups: (4 r + r/2), (6 r + r/4), (9 r) mod 8 are bijections of r mod 8.  A zeroed group in front is "row -1", the
// history of thread 0.
template <int BYTES>
struct Lay {
    static constexpr int kRowW = 8 * BYTES;  // words per row
    static constexpr int kGroupRows = BYTES == 4 ? 1 : BYTES == 3 ? 4 : 2;
    static constexpr int kGroupW = kRowW * kGroupRows;
    static constexpr int kPitchW = kGroupW + 4;
    static constexpr int kGroups = kT / kGroupRows;
    static constexpr int kFrontW = kPitchW;
    static constexpr int kWords = kFrontW + kGroups * kPitchW;
    // word offset of row r >= -1 in raw[]
    static ZF_DEVICE int row(int r) {
        const int q = r + kGroupRows;
        return (q / kGroupRows) * kPitchW + (q % kGroupRows) * kRowW;
    }
};
constexpr int kCrcChunkWords = 16;
// resident CTAs per SM: 16/24-bit: 80 registers, ~66 KB of shared memory each; 32-bit (64-bit chains): 128 registers, ~87 KB
#ifndef ZF_V3_CTAS16
#define ZF_V3_CTAS16 3
#endif
template <int BYTES> struct CtasPerSm { static constexpr int value = BYTES == 4 ? 2 : BYTES == 2 ? ZF_V3_CTAS16 : 3; };

// candidate-channel arithmetic: 32-bit PCM needs 33 bits for the side channel and 37 for its fourth difference
template <int BYTES> struct Arith { typedef int32_t T; typedef uint32_t U; };
template <> struct Arith<4> { typedef long long T; typedef unsigned long long U; };

struct Dec {  // per candidate channel, after pass 1
    uint32_t kind, order, waste, bps, P, est;
    uint32_t pad[2];
};

// scratch of the analysis phases
template <int BYTES>
struct Scratch {
    unsigned long long node[4][256];  // heap nodes 8..255 (levels 3..7): abs-sum | width << 48
    // pass-1 warp partials.  16/24-bit: 5 x (lo, hi) sums + sample OR.  32-bit: 5 x 3 sixteen-bit pieces of the sums,
    // 5 x (lo, hi) ORs of |delta^k x| (range check, fixed.zig:160-162), (lo, hi) sample OR.
    uint32_t red[kW][4][BYTES == 4 ? 28 : 12];
};

template <int BYTES>
struct Smem {
    alignas(16) uint32_t raw[Lay<BYTES>::kWords];
    alignas(16) uint32_t bits[BitBufWords<BYTES>::value + 8];
    alignas(16) uint32_t lvlcost[4][9][20];  // per candidate, per partition order: up to 16 partial cost sums (rows of
                                             // 80 bytes: the pick's eight lanes per candidate read them with LDS.128)
    alignas(16) uint8_t lvlfive[4][9][16];   // ... and whether a parameter > 14 occurs (5-bit method, rice.zig:383-387)
    alignas(16) Scratch<BYTES> sc;
    alignas(16) unsigned long long lbwin[128];  // look-back window: descriptors hi-127 .. hi (fetched by cp.async)
    unsigned long long lb_excl;                // bytes of the predecessors examined so far
    int32_t lb_i;                              // nearest predecessor not examined yet
    uint32_t lb_done;
    unsigned long long mbar;
    unsigned long long out_off;
    Dec dec[4];
    typename Arith<BYTES>::T warm[4][4];  // first four samples of every candidate (warm-ups, CONSTANT value)
    uint32_t scan[2][kW];
    uint32_t crc_part[kW], par_part[kW];
    uint32_t next_frame;
    uint32_t redo;  // 32-bit PCM: the frame overflowed the 32-bit chains, pass 1 is run again in 64 bits
    uint8_t crc8tab[256];
    uint8_t choice[4][512];           // heap node m (1..511) -> Rice parameter, or 0x80 | escape width
};

// ---- small helpers ------------------------------------------------------------------------------------------

// shl.b32 semantics: shift counts >= 32 give 0
ZF_DEVICE uint32_t shl32(uint32_t v, uint32_t s) {
#ifdef ZF_HOST_EMU
    return s >= 32u ? 0u : (v << s);
#else
    uint32_t d;
    asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(s));
    return d;
#endif
}

// shr.b32 semantics: shift counts >= 32 give 0
ZF_DEVICE uint32_t shr32(uint32_t v, uint32_t s) {
#ifdef ZF_HOST_EMU
    return s >= 32u ? 0u : (v >> s);
#else
    uint32_t d;
    asm("shr.b32 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(s));
    return d;
#endif
}

// WHICH: 0 both channels, 1 left only, 2 right only (the other array is left untouched)
template <int BYTES, int WHICH>
ZF_DEVICE void unpack20(const uint32_t *raw, int t, int32_t (&L)[kXn], int32_t (&R)[kXn]) {
    const uint32_t *cur = raw + Lay<BYTES>::row(t), *prev = raw + Lay<BYTES>::row(t - 1);
    if (BYTES == 4) {
        // two words per inter-channel sample; history = the last 32 bytes of the row in front
        const uint4 *pp = reinterpret_cast<const uint4 *>(prev + 24), *pc = reinterpret_cast<const uint4 *>(cur);
#pragma unroll
        for (int k = 0; k < kXn / 2; k++) {
            const uint4 v = k < 2 ? pp[k] : pc[k - 2];
            if (WHICH != 2) { L[2 * k] = (int32_t)v.x; L[2 * k + 1] = (int32_t)v.z; }
            if (WHICH != 1) { R[2 * k] = (int32_t)v.y; R[2 * k + 1] = (int32_t)v.w; }
        }
    } else if (BYTES == 2) {
        // one word per inter-channel sample; history = the last 16 bytes of the row in front
        const uint4 *pp = reinterpret_cast<const uint4 *>(prev + 12), *pc = reinterpret_cast<const uint4 *>(cur);
#pragma unroll
        for (int k = 0; k < kXn / 4; k++) {
            const uint4 v = k < 1 ? pp[0] : pc[k - 1];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (WHICH != 2) L[4 * k + q] = (int32_t)prmt(w[q], 0, 0x9910);
                if (WHICH != 1) R[4 * k + q] = (int32_t)prmt(w[q], 0, 0xBB32);
            }
        }
    } else {
        // 6 bytes per inter-channel sample; the last 32 bytes of the row in front, then the row: the window (4 history +
        // 16 own samples = 120 bytes) starts 8 bytes in
        const uint4 *pp = reinterpret_cast<const uint4 *>(prev + 16), *pc = reinterpret_cast<const uint4 *>(cur);
        uint32_t w[32];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 v = k < 2 ? pp[k] : pc[k - 2];
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < kXn / 2; k++) {  // two inter-channel samples per three words
            const uint32_t a = w[2 + 3 * k], b = w[3 + 3 * k], d = w[4 + 3 * k];
            if (WHICH != 2) {
                L[2 * k] = (int32_t)prmt(a, b, 0xA210);
                L[2 * k + 1] = (int32_t)prmt(b, d, 0xC432);
            }
            if (WHICH != 1) {
                R[2 * k] = (int32_t)prmt(a, b, 0xD543);
                R[2 * k + 1] = (int32_t)prmt(d, 0, 0xB321);
            }
        }
    }
}

// mid = (l + r) >> 1 (encoder.zig:341-349); with 32-bit PCM the sum needs 33 bits, so the floor average is taken without it
template <int BYTES>
ZF_DEVICE int32_t mid32(int32_t l, int32_t r) {
    return BYTES == 4 ? (l & r) + ((l ^ r) >> 1) : (l + r) >> 1;
}

// candidate channel `slot` (uniform over the block) of a stereo frame: 0 L, 1 R, 2 M = (L + R) >> 1, 3 S = L - R
// (encoder.zig:330-350).  Every arm produces x directly, so no register copies are needed to merge them.
// (32-bit PCM comes here only in frames whose side channel and differences fit 32 bits -- see ChainN.)
template <int BYTES>
ZF_DEVICE void load_x(const uint32_t *raw, int t, uint32_t slot, int32_t (&x)[kXn]) {
    if (slot == 0) {
        unpack20<BYTES, 1>(raw, t, x, x);
    } else if (slot == 1) {
        unpack20<BYTES, 2>(raw, t, x, x);
    } else {
        int32_t R[kXn];
        unpack20<BYTES, 0>(raw, t, x, R);
        if (slot == 2) {
#pragma unroll
            for (int i = 0; i < kXn; i++) x[i] = mid32<BYTES>(x[i], R[i]);
        } else {
#pragma unroll
            for (int i = 0; i < kXn; i++) x[i] = x[i] - R[i];
        }
    }
}
// 32-bit PCM: 64-bit candidates (encoder.zig:330-339)
template <int BYTES>
ZF_DEVICE void load_x(const uint32_t *raw, int t, uint32_t slot, long long (&x)[kXn]) {
    int32_t L[kXn], R[kXn];
    unpack20<BYTES, 0>(raw, t, L, R);
#pragma unroll
    for (int i = 0; i < kXn; i++) {
        const long long l = L[i], r = R[i];
        x[i] = slot == 0 ? l : slot == 1 ? r : slot == 2 ? ((l + r) >> 1) : (l - r);
    }
}

// inter-channel samples 4 g .. 4 g + 3 of thread t's row, g = 0..3; g = -1: the last four of the row in front (the
// history; the zeroed front group for thread 0).  16-byte loads only (conflict-free with the padded layout); the
// 24-byte groups of 24-bit PCM are taken out of two aligned chunks.
template <int BYTES>
ZF_DEVICE void load4(const uint32_t *raw, int t, int g, int32_t (&L)[4], int32_t (&R)[4]) {
    const uint32_t *row = raw + (g < 0 ? Lay<BYTES>::row(t - 1) : Lay<BYTES>::row(t));
    const int gg = g < 0 ? 3 : g;
    if (BYTES == 4) {
        const uint4 *p = reinterpret_cast<const uint4 *>(row + 8 * gg);
        const uint4 v0 = p[0], v1 = p[1];
        L[0] = (int32_t)v0.x; R[0] = (int32_t)v0.y; L[1] = (int32_t)v0.z; R[1] = (int32_t)v0.w;
        L[2] = (int32_t)v1.x; R[2] = (int32_t)v1.y; L[3] = (int32_t)v1.z; R[3] = (int32_t)v1.w;
    } else if (BYTES == 2) {
        const uint4 v = *reinterpret_cast<const uint4 *>(row + 4 * gg);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            L[q] = (int32_t)prmt(w[q], 0, 0x9910);
            R[q] = (int32_t)prmt(w[q], 0, 0xBB32);
        }
    } else {
        // bytes 24 g .. 24 g + 23 of the row: chunks (3 g) >> 1 and the next one, from their start (g even) or 8 bytes in
        const uint4 *p = reinterpret_cast<const uint4 *>(row + 4 * ((3 * gg) >> 1));
        const uint4 v0 = p[0], v1 = p[1];
        const bool odd = (gg & 1) != 0;
        const uint32_t w[6] = {odd ? v0.z : v0.x, odd ? v0.w : v0.y, odd ? v1.x : v0.z,
                               odd ? v1.y : v0.w, odd ? v1.z : v1.x, odd ? v1.w : v1.y};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const uint32_t a = w[3 * k], b = w[3 * k + 1], d = w[3 * k + 2];
            L[2 * k] = (int32_t)prmt(a, b, 0xA210);
            R[2 * k] = (int32_t)prmt(a, b, 0xD543);
            L[2 * k + 1] = (int32_t)prmt(b, d, 0xC432);
            R[2 * k + 1] = (int32_t)prmt(d, 0, 0xB321);
        }
    }
}

// fixed.bestOrder's difference chain for one candidate channel, streamed four samples at a time
struct Chain {
    int32_t xp, e1p, e2p, e3p;        // previous sample and differences
    uint32_t s0, s1, s2, s3, s4, orv;  // sum |delta^k x|, OR of the samples

    ZF_DEVICE void init() { xp = e1p = e2p = e3p = 0; s0 = s1 = s2 = s3 = s4 = orv = 0; }
    template <bool ACC>
    ZF_DEVICE void step(int32_t x) {
        const int32_t e1 = x - xp;
        const int32_t e2 = e1 - e1p;
        const int32_t e3 = e2 - e2p;
        const int32_t e4 = e3 - e3p;
        xp = x; e1p = e1; e2p = e2; e3p = e3;
        if (ACC) {
            orv |= (uint32_t)x;
            s0 += uabs(x); s1 += uabs(e1); s2 += uabs(e2); s3 += uabs(e3); s4 += uabs(e4);
        }
    }
    ZF_DEVICE void take_out(uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4) { s1 -= a1; s2 -= a2; s3 -= a3; s4 -= a4; }
    ZF_DEVICE void sums(uint32_t (&o)[5]) const { o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3; o[4] = s4; }
};

// The same in SINGLE PRECISION for 16-bit PCM.  Every quantity is an integer that a float carries exactly: samples and
// the side channel below 2^17, fourth differences below 2^21, a thread's sixteen-term sums at most 2^24 (every partial sum
// of non-negative integers up to 2^24 is representable).  |e| is then a free operand modifier of the accumulating FADD,
// and the FP32 pipe is twice as wide as the integer pipe: 9 instructions per chain and sample instead of 14
// (tools/microbench/chain_bench.cu: 14.0 against 19.0 SM-clocks per warp-sample for the four chains).
struct ChainF {
    float xp, e1p, e2p, e3p, s0, s1, s2, s3, s4;
    uint32_t orv;
    ZF_DEVICE void init() { xp = e1p = e2p = e3p = 0.0f; s0 = s1 = s2 = s3 = s4 = 0.0f; orv = 0; }
    template <bool ACC>
    ZF_DEVICE void step(int32_t xi) {
        const float x = (float)xi;
        const float e1 = x - xp, e2 = e1 - e1p, e3 = e2 - e2p, e4 = e3 - e3p;
        xp = x; e1p = e1; e2p = e2; e3p = e3;
        if (ACC) {
            orv |= (uint32_t)xi;
            s0 += fabsf(x); s1 += fabsf(e1); s2 += fabsf(e2); s3 += fabsf(e3); s4 += fabsf(e4);
        }
    }
    ZF_DEVICE void take_out(uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4) {
        s1 -= (float)a1; s2 -= (float)a2; s3 -= (float)a3; s4 -= (float)a4;
    }
    ZF_DEVICE void sums(uint32_t (&o)[5]) const {
        o[0] = (uint32_t)s0; o[1] = (uint32_t)s1; o[2] = (uint32_t)s2; o[3] = (uint32_t)s3; o[4] = (uint32_t)s4;
    }
};

template <int BYTES> struct Pass1Chain { typedef Chain type; };
#ifndef ZF_V3_INT16_PASS1
template <> struct Pass1Chain<2> { typedef ChainF type; };
#endif

// the same in 64 bits, with the OR of |delta^k x| that fixed.bestOrder's range check needs (fixed.zig:160-162)
struct ChainW {
    typedef long long V;
    static ZF_DEVICE V mid(int32_t l, int32_t r) { return ((V)l + (V)r) >> 1; }
    static ZF_DEVICE V side(int32_t l, int32_t r, bool) { return (V)l - (V)r; }
    ZF_DEVICE uint32_t overflowed() const { return 0; }
    long long xp, e1p, e2p, e3p;
    unsigned long long s0, s1, s2, s3, s4, r0, r1, r2, r3, r4, orv;

    ZF_DEVICE void init() { xp = e1p = e2p = e3p = 0; s0 = s1 = s2 = s3 = s4 = r0 = r1 = r2 = r3 = r4 = orv = 0; }
    template <bool ACC>
    ZF_DEVICE void step(long long x) {
        const long long e1 = x - xp;
        const long long e2 = e1 - e1p;
        const long long e3 = e2 - e2p;
        const long long e4 = e3 - e3p;
        xp = x; e1p = e1; e2p = e2; e3p = e3;
        if (ACC) {
            const unsigned long long a0 = uabs(x), a1 = uabs(e1), a2 = uabs(e2), a3 = uabs(e3), a4 = uabs(e4);
            orv |= (unsigned long long)x;
            s0 += a0; s1 += a1; s2 += a2; s3 += a3; s4 += a4;
            r0 |= a0; r1 |= a1; r2 |= a2; r3 |= a3; r4 |= a4;
        }
    }
    // sample q < 4 of the frame: total[k] and the range OR only count samples i >= k (fixed.zig:102-127)
    template <int Q>
    ZF_DEVICE void step_first(long long x) {
        const long long e1 = x - xp;
        const long long e2 = e1 - e1p;
        const long long e3 = e2 - e2p;
        const long long e4 = e3 - e3p;
        xp = x; e1p = e1; e2p = e2; e3p = e3;
        const unsigned long long a0 = uabs(x), a1 = uabs(e1), a2 = uabs(e2), a3 = uabs(e3), a4 = uabs(e4);
        orv |= (unsigned long long)x;
        s0 += a0; r0 |= a0;
        if (Q >= 1) { s1 += a1; r1 |= a1; }
        if (Q >= 2) { s2 += a2; r2 |= a2; }
        if (Q >= 3) { s3 += a3; r3 |= a3; }
    }
};

// 32-bit PCM in 32-bit registers.  Wrapping 32-bit differences are the true ones as long as every subtraction fits, and
// whether it does is detected exactly (signed overflow of d = a - b: the sign bit of (a ^ b) & (a ^ d)) provided the
// operands were true values -- so ONE flag over all orders says "this frame is what the 64-bit chains would have
// produced"; when it is raised anywhere in the block, pass 1 is redone with ChainW and the frame takes the 64-bit paths.
// Full-scale square waves, noise and 24-bit material left-justified in 32 bits at high level do that; music does not.
// Sums of sixteen terms below 2^31 need 64 bits; the range ORs of fixed.zig:160-162 are 32-bit (bit 31 = out of range).
struct ChainN {
    typedef int32_t V;
    static ZF_DEVICE V mid(int32_t l, int32_t r) { return (l & r) + ((l ^ r) >> 1); }
    // the side channel itself has 33 bits: flag the samples that do not fit (only where they are data: `count`)
    ZF_DEVICE V side(int32_t l, int32_t r, bool count) {
        const int32_t d = l - r;
        if (count) ovf |= (uint32_t)(l ^ r) & (uint32_t)(l ^ d);
        return d;
    }
    ZF_DEVICE uint32_t overflowed() const { return ovf >> 31; }
    int32_t xp, e1p, e2p, e3p;
    unsigned long long s0, s1, s2, s3, s4;
    uint32_t r0, r1, r2, r3, r4, orv, ovf;

    ZF_DEVICE void init() { xp = e1p = e2p = e3p = 0; s0 = s1 = s2 = s3 = s4 = 0; r0 = r1 = r2 = r3 = r4 = orv = ovf = 0; }
    ZF_DEVICE void diffs(int32_t x, int32_t &e1, int32_t &e2, int32_t &e3, int32_t &e4) {
        e1 = x - xp; e2 = e1 - e1p; e3 = e2 - e2p; e4 = e3 - e3p;
    }
    template <bool ACC>
    ZF_DEVICE void step(int32_t x) {
        int32_t e1, e2, e3, e4;
        diffs(x, e1, e2, e3, e4);
        if (ACC) {
            ovf |= ((uint32_t)(x ^ xp) & (uint32_t)(x ^ e1)) | ((uint32_t)(e1 ^ e1p) & (uint32_t)(e1 ^ e2));
            ovf |= ((uint32_t)(e2 ^ e2p) & (uint32_t)(e2 ^ e3)) | ((uint32_t)(e3 ^ e3p) & (uint32_t)(e3 ^ e4));
            const uint32_t a0 = uabs(x), a1 = uabs(e1), a2 = uabs(e2), a3 = uabs(e3), a4 = uabs(e4);
            orv |= (uint32_t)x;
            s0 += a0; s1 += a1; s2 += a2; s3 += a3; s4 += a4;
            r0 |= a0; r1 |= a1; r2 |= a2; r3 |= a3; r4 |= a4;
        }
        xp = x; e1p = e1; e2p = e2; e3p = e3;
    }
    // sample q < 4 of the frame: total[k] and the range OR only count samples i >= k (fixed.zig:102-127); the differences
    // against the zero pad are not data, so they raise no flag
    template <int Q>
    ZF_DEVICE void step_first(int32_t x) {
        int32_t e1, e2, e3, e4;
        diffs(x, e1, e2, e3, e4);
        const uint32_t a0 = uabs(x), a1 = uabs(e1), a2 = uabs(e2), a3 = uabs(e3);
        orv |= (uint32_t)x;
        s0 += a0; r0 |= a0;
        if (Q >= 1) { s1 += a1; r1 |= a1; ovf |= (uint32_t)(x ^ xp) & (uint32_t)(x ^ e1); }
        if (Q >= 2) { s2 += a2; r2 |= a2; ovf |= (uint32_t)(e1 ^ e1p) & (uint32_t)(e1 ^ e2); }
        if (Q >= 3) { s3 += a3; r3 |= a3; ovf |= (uint32_t)(e2 ^ e2p) & (uint32_t)(e2 ^ e3); }
        xp = x; e1p = e1; e2p = e2; e3p = e3;
    }
};

// `order` passes of in-place differencing: afterwards x[i] is the order-th difference for i >= order
// (one rolled pass per order: the low positions a pass would not need are differenced too -- nobody reads them -- so
// that every pass is the same piece of code)
template <typename TT>
ZF_DEVICE void diff_in_place(TT (&x)[kXn], uint32_t order) {
#pragma unroll 1
    for (uint32_t k = 0; k < order; k++) {
#pragma unroll
        for (int i = kXn - 1; i > 0; i--) x[i] -= x[i - 1];
    }
}

// rice.calcOptimalParams for one partition (rice.zig:343-395, flacCalcPartSize :402-405) in closed form, as
// zf::best_param (proof there), specialised: width B <= 31 always holds below 32-bit PCM, P >= 2, and the winning
// cost is bounded by the escape cost 5 + B n < 2^18, so costs are 32-bit (a losing candidate may saturate).
ZF_DEVICE void best_param_nw(unsigned long long S, uint32_t B, uint32_t n, uint32_t P, uint32_t &choice, uint32_t &cost) {
    uint32_t best = 5u + B * n;
    uint32_t ch = 0x80u | B;
    const uint32_t n2 = 2u * n;
    uint32_t q = 0;
    if (S > n2) {
        q = bitlen64(S) - bitlen32(n2);
        if ((S >> q) > n2) q++;
    }
    uint32_t p = q + 1u;
    if (p > P - 1u) p = P - 1u;
    const unsigned long long sh = S >> ((p - 1u) & 63u);
    const uint32_t sh32 = sh > 0x3fffffffull ? 0x3fffffffu : (uint32_t)sh;
    uint32_t cc = (1u + p) * n + sh32 - (n >> 1);
    if (p == 0u) cc = 0xffffffffu;  // max_rice_param 1: only parameter 0 is tried (rice.zig:369)
    uint32_t cand = p;
    const uint32_t c0 = S > 0x0fffffffull ? 0x7fffffffu : n + 2u * (uint32_t)S;  // p == 0: no -(n >> 1)
    if (c0 <= cc) { cc = c0; cand = 0; }                                         // lowest p wins ties
    if (cc < best) { best = cc; ch = cand; }                                     // escape wins ties
    choice = ch;
    cost = best;
}
// 32-bit PCM: a width of 32 makes the escape invalid (rice.zig:361-366); everything else as above
ZF_DEVICE void best_param_w(unsigned long long S, uint32_t B, uint32_t n, uint32_t P, uint32_t &choice, uint32_t &cost) {
    uint32_t best = B <= 31u ? 5u + B * n : 0xffffffffu;
    uint32_t ch = 0x80u | B;
    const uint32_t n2 = 2u * n;
    uint32_t q = 0;
    if (S > n2) {
        q = bitlen64(S) - bitlen32(n2);
        if ((S >> q) > n2) q++;
    }
    uint32_t p = q + 1u;
    if (p > P - 1u) p = P - 1u;
    const unsigned long long sh = S >> ((p - 1u) & 63u);
    const uint32_t sh32 = sh > 0x3fffffffull ? 0x3fffffffu : (uint32_t)sh;
    uint32_t cc = (1u + p) * n + sh32 - (n >> 1);
    if (p == 0u) cc = 0xffffffffu;  // max_rice_param 1: only parameter 0 is tried (rice.zig:369)
    uint32_t cand = p;
    const uint32_t c0 = S > 0x0fffffffull ? 0x7fffffffu : n + 2u * (uint32_t)S;
    if (c0 <= cc) { cc = c0; cand = 0; }
    if (cc < best) { best = cc; ch = cand; }
    choice = ch;
    cost = best;
}
// the same for abs-sums below 2^32 (every leaf; every node of ordinary audio)
ZF_DEVICE void best_param_32(uint32_t S, uint32_t B, uint32_t n, uint32_t P, uint32_t &choice, uint32_t &cost) {
    uint32_t best = 5u + B * n;
    uint32_t ch = 0x80u | B;
    const uint32_t n2 = 2u * n;
    uint32_t q = 0;
    if (S > n2) {
        q = bitlen32(S) - bitlen32(n2);
        if ((S >> q) > n2) q++;
    }
    uint32_t p = q + 1u;
    if (p > P - 1u) p = P - 1u;
    const uint32_t sh = S >> ((p - 1u) & 31u);
    const uint32_t sh32 = sh > 0x3fffffffu ? 0x3fffffffu : sh;
    uint32_t cc = (1u + p) * n + sh32 - (n >> 1);
    if (p == 0u) cc = 0xffffffffu;  // max_rice_param 1: only parameter 0 is tried (rice.zig:369)
    uint32_t cand = p;
    const uint32_t c0 = S > 0x0fffffffu ? 0x7fffffffu : n + 2u * S;
    if (c0 <= cc) { cc = c0; cand = 0; }
    if (cc < best) { best = cc; ch = cand; }
    choice = ch;
    cost = best;
}

// ---- bit writer: right-aligned 64-bit shift register ------------------------------------------------------------
// Branch-free.  A completed 32-bit word is written with a plain predicated store: exactly one thread completes a
// given word (the one whose bit range crosses the word's end), and it holds every bit of the word that lies in its
// own range while the bits in front of its range are still zero in its register.  What precedes (the tail of the
// previous thread, the frame header) is ORed into the word after a barrier (or_tail()).
struct BitW {
#ifdef ZF_HOST_EMU
    uint32_t *wp;
#else
    uint32_t wp;  // shared-window address
#endif
    uint32_t hi, lo, nb;  // nb < 32 pending bits in the low end of hi:lo

    ZF_DEVICE void init(uint32_t *bits, uint32_t bitpos) {
#ifdef ZF_HOST_EMU
        wp = bits + (bitpos >> 5);
#else
        wp = smem_addr(bits) + ((bitpos >> 5) << 2);
#endif
        nb = bitpos & 31u;
        hi = 0;
        lo = 0;
    }
    // append the fl-bit field val (fl 0..32, val < 2^fl)
    ZF_DEVICE void put(uint32_t val, uint32_t fl) {
        hi = __funnelshift_lc(lo, hi, fl);
        lo = shl32(lo, fl) | val;
        nb += fl;
#ifdef ZF_HOST_EMU
        if (nb >= 32u) {
            *wp = __funnelshift_r(lo, hi, nb - 32u);
            wp++;
        }
#else
        const uint32_t word = __funnelshift_r(lo, hi, nb);  // the shift uses nb mod 32
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ge.u32 p, %1, 32;\n"
            "@p st.shared.b32 [%0], %2;\n"
            "@p add.u32 %0, %0, 4;\n"
            "}\n"
            : "+r"(wp)
            : "r"(nb), "r"(word)
            : "memory");
#endif
        nb &= 31u;
    }
    // q zero bits, then the len-bit field val (len 1..31)
    ZF_DEVICE void put_code(uint32_t q, uint32_t val, uint32_t len) {
        if (q + len > 32u) {  // long unary run (rare)
#pragma unroll 1
            while (q >= 32u) { put(0, 32); q -= 32u; }
            put(0, q);
            q = 0;
        }
        put(val, q + len);
    }
    // the pending bits, left-aligned in their word; ORed into the buffer after the barrier that ends the stores
    ZF_DEVICE void or_tail() const {
        if (nb) {
            const uint32_t w = lo << (32u - nb);
#ifdef ZF_HOST_EMU
            atomicOr(wp, w);
#else
            asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wp), "r"(w) : "memory");
#endif
        }
    }
};

// ---- CRC-16 (poly 0x8005) in the factor ring ----------------------------------------------------------------------
// one folding round modulo Q = x^15 + x + 1: degree d -> max(14, d - 14)
ZF_DEVICE uint32_t q_fold(uint32_t v) {
    const uint32_t h = v >> 15;
    return (v & 0x7fffu) ^ h ^ (h << 1);
}
// Carry-less 15 x 15 -> 29-bit product by integer multiplies "with holes": both operands are split by bit position mod 4,
// so at most four partial products meet in any 4-bit slot of a product and no carry leaves the slot; the slot's low bit is
// the parity, i.e. the GF(2) sum.  Sixteen IMADs (the FMA pipe is idle in this kernel) instead of a 15-trip shift/select loop.
ZF_DEVICE uint32_t q_mulmod(uint32_t a, uint32_t b) {  // a, b < 2^15
    const uint32_t a0 = a & 0x1111u, a1 = a & 0x2222u, a2 = a & 0x4444u, a3 = a & 0x8888u;
    const uint32_t b0 = b & 0x1111u, b1 = b & 0x2222u, b2 = b & 0x4444u, b3 = b & 0x8888u;
    const uint32_t z0 = (a0 * b0) ^ (a1 * b3) ^ (a2 * b2) ^ (a3 * b1);
    const uint32_t z1 = (a0 * b1) ^ (a1 * b0) ^ (a2 * b3) ^ (a3 * b2);
    const uint32_t z2 = (a0 * b2) ^ (a1 * b1) ^ (a2 * b0) ^ (a3 * b3);
    const uint32_t z3 = (a0 * b3) ^ (a1 * b2) ^ (a2 * b1) ^ (a3 * b0);
    const uint32_t acc = (z0 & 0x11111111u) | (z1 & 0x22222222u) | (z2 & 0x44444444u) | (z3 & 0x88888888u);
    return q_fold(acc);  // 29-bit product: one round suffices
}

// what one thread knows about the subframe it helps to write
struct Sub {
    uint32_t kind, order, waste, bps, po, method;
    uint32_t depth_ch;
    uint32_t choice;
    uint32_t maxq;  // largest unary quotient among this thread's codes
    uint32_t vhi;   // VERBATIM samples of 33 bits (32-bit side channel): bit j = bit 32 of sample j
    bool at_start;
};

// One subframe, counting side.  Fills v[] with what the writing side needs (zigzagged residuals for FIXED, plain residuals
// where the partition is escaped, shifted samples for VERBATIM) and returns this thread's bit count.  frame_writer.zig:269-372.
template <int BYTES, bool W64>
ZF_DEVICE uint32_t sub_count(const Smem<BYTES> &sm, int t, uint32_t slot, Sub &u, uint32_t (&v)[kS]) {
    typedef typename Arith<W64 ? 4 : 2>::T XT;
    u.choice = 0;
    u.maxq = 0;
    u.vhi = 0;
    u.at_start = false;
#pragma unroll
    for (int j = 0; j < kS; j++) v[j] = 0;
    if (__builtin_expect(u.kind == kConstant, 0)) return t == 0 ? 8u + u.depth_ch : 0u;
    XT x[kXn];
    load_x<BYTES>(sm.raw, t, slot, x);
    if (__builtin_expect(u.kind == kVerbatim, 0)) {
#pragma unroll
        for (int j = 0; j < kS; j++) {
            const XT sv = x[kH + j] >> u.waste;
            v[j] = (uint32_t)sv;
            if (BYTES == 4) u.vhi |= (sv < 0 ? 1u : 0u) << j;  // bit 32 of a 33-bit field = the sign
        }
        return (uint32_t)kS * u.bps + (t == 0 ? 8u + u.waste : 0u);
    }
    diff_in_place(x, u.order);
    const uint32_t sh = 8u - u.po;  // threads per partition = 1 << sh
    u.choice = sm.choice[slot][(1u << u.po) + ((uint32_t)t >> sh)];
    u.at_start = ((uint32_t)t & ((1u << sh) - 1u)) == 0;
    const bool esc = (u.choice & 0x80u) != 0;
    {   // zigzagged residuals (low 32 bits, fixed.zig:70-73); an escaped partition keeps them as they are (:341-354)
        const uint32_t zs = esc ? 0u : 1u, zm = esc ? 0u : 0xffffffffu;
#pragma unroll
        for (int j = 0; j < kS; j++) {
            const int32_t r = (int32_t)(x[kH + j] >> u.waste);
            v[j] = ((uint32_t)r << zs) ^ ((uint32_t)(r >> 31) & zm);
        }
    }
    const uint32_t jstart = (t == 0) ? u.order : 0u;
    const uint32_t cnt = (uint32_t)kS - jstart;
    uint32_t bits = (t == 0) ? 8u + u.waste + u.order * u.bps + 6u : 0u;
    if (u.at_start) bits += 4u + u.method + (esc ? 5u : 0u);
    if (__builtin_expect(esc, 0)) return bits + (u.choice & 0x7fu) * cnt;
    uint32_t qs = 0, mq = 0;
#pragma unroll
    for (int j = 0; j < kS; j++) {
        uint32_t q = v[j] >> u.choice;
        if (j < 4) q = ((uint32_t)j >= jstart) ? q : 0u;
        qs += q;
        mq = q > mq ? q : mq;
    }
    u.maxq = mq;
    return bits + qs + cnt * (u.choice + 1u);
}

// everything of a subframe in front of this thread's residual codes (frame_writer.zig:269-329, :341-357)
// a field of up to 33 bits (32-bit side channel)
ZF_DEVICE void put_wide(BitW &bw, long long value, uint32_t width) {
    if (width > 32u) {
        bw.put((uint32_t)((unsigned long long)value >> 32) & (0xffffffffu >> (64u - width)), width - 32u);
        bw.put((uint32_t)value, 32);
    } else {
        bw.put((uint32_t)value & (0xffffffffu >> (32u - width)), width);
    }
}

// a sample field: up to 33 bits with 64-bit candidates, at most 25 bits otherwise
ZF_DEVICE void put_sample(BitW &bw, long long value, uint32_t width) { put_wide(bw, value, width); }
ZF_DEVICE void put_sample(BitW &bw, int32_t value, uint32_t width) {
    bw.put((uint32_t)value & (0xffffffffu >> (32u - width)), width);
}

template <typename TT>
ZF_DEVICE void sub_head(BitW &bw, const TT (&warm)[4], int t, const Sub &u) {
    if (u.kind == kConstant) {  // :269-279: 0x00, then the un-shifted sample at full depth
        if (t == 0) {
            bw.put(0, 8);
            put_sample(bw, warm[0], u.depth_ch);
        }
        return;
    }
    if (u.kind == kVerbatim) {  // :282-301
        if (t == 0) {
            bw.put(u.waste ? 3u : 2u, 8);
            if (u.waste) bw.put(1u, u.waste);
        }
        return;
    }
    const uint32_t param_len = 4u + u.method;
    if (t == 0) {  // :303-329
        bw.put(((8u | u.order) << 1) | (u.waste ? 1u : 0u), 8);
        if (u.waste) bw.put(1u, u.waste);
#pragma unroll 1
        for (uint32_t k = 0; k < u.order; k++) put_sample(bw, (TT)(warm[k] >> u.waste), u.bps);
        bw.put((u.method << 4) | u.po, 6);
    }
    if (u.at_start) {  // :341-357
        if (u.choice & 0x80u) {
            bw.put(u.method ? 31u : 15u, param_len);
            bw.put(u.choice & 0x7fu, 5);
        } else {
            bw.put(u.choice, param_len);
        }
    }
}

// One subframe's fields of this thread where a codeword may not fit one 32-bit field (long unary run) or the samples have
// 33 bits (VERBATIM 32-bit side channel); the caller's branch-free loop handles everything else.  Field parameters as there.
template <bool WIDE>
ZF_DEVICE void sub_body(BitW &bw, const Sub &u, const uint32_t (&v)[kS], uint32_t k, uint32_t one, uint32_t m, uint32_t len0,
                        uint32_t jstart) {
    // rare path: compact code (a 4-trip loop; v[] must stay in registers, so select instead of indexing)
#pragma unroll 1
    for (int j0 = 0; j0 < kS; j0 += 4) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            uint32_t z = v[jj];
            z = j0 == 4 ? v[4 + jj] : z;
            z = j0 == 8 ? v[8 + jj] : z;
            z = j0 == 12 ? v[12 + jj] : z;
            if ((uint32_t)(j0 + jj) >= jstart) {
                if (WIDE && len0 > 32u) {  // :282-301, 33-bit samples
                    bw.put((u.vhi >> (j0 + jj)) & 1u, 1);
                    bw.put(z, 32);
                } else {
                    bw.put_code(shr32(z, k), one | (z & m), len0);
                }
            }
        }
    }
}

ZF_DEVICE void block_scan2(uint32_t (&scan)[2][kW], int t, uint32_t a, uint32_t b, uint32_t &ex_a, uint32_t &ex_b,
                           uint32_t &tot_a, uint32_t &tot_b) {
    const int lane = t & 31, warp = t >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += ua; ib += ub; }
    }
    if (lane == 31) { scan[0][warp] = ia; scan[1][warp] = ib; }
    __syncthreads();
    uint32_t oa = 0, ob = 0, ta = 0, tb = 0;
#pragma unroll
    for (int w = 0; w < kW; w++) {
        const uint32_t wa = scan[0][w], wb = scan[1][w];
        if (w < warp) { oa += wa; ob += wb; }
        ta += wa;
        tb += wb;
    }
    ex_a = oa + ia - a;
    ex_b = ob + ib - b;
    tot_a = ta;
    tot_b = tb;
}

// ---- deferred epilogue -------------------------------------------------------------------------------------------
// A finished frame stays in the bit buffer while the CTA analyses its next frame; its output offset is found in the
// meantime by a decoupled look-back over the frame-size descriptors (128 per round, fetched by cp.async so that no
// registers are held while a round trip to L2 is in flight), and the frame is copied out just before the bit buffer
// is needed again.  Nobody ever waits for a predecessor that is merely a little behind.
struct Pend {
    uint32_t valid, fidx, fbytes, lead, fits;
};

// the few FrameJob fields the epilogue needs (a kernel parameter cannot be passed on by reference cheaply)
struct LbArgs {
    unsigned long long *desc;
    unsigned long long *total_bytes;
    unsigned int *status;
    unsigned long long out_cap;
    uint32_t batch_frames;
};

template <int BYTES>
ZF_DEVICE void lb_issue(Smem<BYTES> &sm, const LbArgs &job, int lane) {
    const int hi = sm.lb_i | 1, base = hi - 127;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int idx = base + 64 * h + 2 * lane;
        if (idx >= 0) cp_async16_cg(&sm.lbwin[64 * h + 2 * lane], job.desc + idx);
    }
    cp_async_commit();
}

template <int BYTES>
ZF_DEVICE void lb_done(Smem<BYTES> &sm, const LbArgs &job, int lane, const Pend &P, unsigned long long excl) {
    if (lane == 0) {
        const unsigned long long size = P.fbytes + 2u;
        sm.out_off = excl;
        sm.lb_done = 1;
        st_relaxed_gpu(job.desc + P.fidx, kFlagPrefix | (excl + size));
        if (P.fidx + 1 == job.batch_frames) *job.total_bytes = excl + size;
        if (excl + size > job.out_cap) atomicOr(job.status, kStatusOutOverflow);
    }
    __syncwarp();
}

// warp 0: begin the look-back of the pending frame
template <int BYTES>
ZF_NOINLINE void lb_start(Smem<BYTES> &sm, const LbArgs job, int lane, const Pend P) {
    if (P.fidx == 0) {
        lb_done(sm, job, lane, P, 0);
        return;
    }
    if (lane == 0) {
        sm.lb_i = (int32_t)P.fidx - 1;
        sm.lb_excl = 0;
        sm.lb_done = 0;
    }
    __syncwarp();
    lb_issue(sm, job, lane);
}

// one warp: take the round that is in flight (waits for it if it has not landed), start the next one if needed.
// Lane l examines descriptors hi - 4 l - j, j = 0 (nearest) .. 3, without branches: two 16-byte loads, then the first
// prefix / unpublished flag of the lane decides which of its values count.
template <int BYTES>
ZF_NOINLINE void lb_step(Smem<BYTES> &sm, const LbArgs job, int lane, const Pend P) {
    cp_async_wait_all();
    __syncwarp();
    const int i = sm.lb_i;
    const unsigned long long excl0 = sm.lb_excl;
    const int hi = i | 1, base = hi - 127;
    const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(&sm.lbwin[124 - 4 * lane]);
    const ulonglong2 far2 = wp[0], near2 = wp[1];
    const unsigned long long dd[4] = {near2.y, near2.x, far2.y, far2.x};  // nearest first
    unsigned long long sum = 0;
    bool open = true, found = false, invalid = false;  // open: neither a prefix nor a gap met so far
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int idx = hi - 4 * lane - j;
        unsigned long long d = dd[j];
        d = idx < 0 ? kFlagPrefix : d;      // in front of the first frame: prefix 0
        d = idx > i ? kFlagAggregate : d;   // the pending frame itself (hi = i + 1): nothing
        const uint32_t flag = (uint32_t)(d >> 62);
        sum += open ? (d & kValueMask) : 0ull;
        found = found || (open && flag == 2u);
        invalid = invalid || (open && flag == 0u);
        open = open && flag == 1u;
    }
    const uint32_t pmask = __ballot_sync(0xffffffffu, found);
    const uint32_t imask = __ballot_sync(0xffffffffu, invalid);
    const uint32_t first_p = pmask ? ctz32(pmask) : 32u;
    const uint32_t need = first_p >= 31u ? 0xffffffffu : ((2u << first_p) - 1u);
    // lanes in front of the first prefix lane hold aggregates only (4 frames < 2^32 bytes); the prefix lane's sum is 64-bit
    const uint32_t agg = reduce_add(((uint32_t)lane < first_p) ? (uint32_t)sum : 0u);
    const unsigned long long pre = __shfl_sync(0xffffffffu, sum, (int)(first_p & 31u));
    const unsigned long long excl = excl0 + agg + (first_p < 32u ? pre : 0ull);
    __syncwarp();  // every lane has read the window and the state
    if (imask & need) {  // a predecessor has not published its size yet: fetch the same window again
        lb_issue(sm, job, lane);
        return;
    }
    if (first_p < 32u) {
        lb_done(sm, job, lane, P, excl);
        return;
    }
    if (lane == 0) {
        sm.lb_excl = excl;
        sm.lb_i = base - 1;
    }
    __syncwarp();
    lb_issue(sm, job, lane);
}

// all threads: copy the pending frame to its place in the output stream (byte-shifted, 16-byte coalesced stores);
// the CRC-16 is assembled from the per-warp partials and goes straight to global memory
template <int BYTES>
ZF_NOINLINE void copy_out(const Smem<BYTES> &sm, uint8_t *out, unsigned long long out_cap, int t, const Pend P) {
    const unsigned long long off = sm.out_off;
    const uint32_t fbytes = P.fbytes, lead = P.lead, size = fbytes + 2u;
    if (!P.fits || off + size > out_cap) return;
    uint8_t *dst = out + off;
    const uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    const uint32_t h = head < fbytes ? head : fbytes;
    if ((uint32_t)t < h) {
        const uint32_t b = lead + (uint32_t)t;
        dst[t] = (uint8_t)(sm.bits[b >> 2] >> (24u - 8u * (b & 3u)));
    }
    const uint32_t nq = (fbytes - h) >> 4;
    uint4 *dq = reinterpret_cast<uint4 *>(dst + h);
    const uint32_t o = lead + h, w0 = o >> 2, sh = 8u * (o & 3u);
#pragma unroll 1
    for (uint32_t k = t; k < nq; k += kT) {
        const uint32_t *p = sm.bits + w0 + 4u * k;
        const uint32_t b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3], b4 = p[4];
        uint4 v;
        v.x = prmt(__funnelshift_l(b1, b0, sh), 0, 0x0123);
        v.y = prmt(__funnelshift_l(b2, b1, sh), 0, 0x0123);
        v.z = prmt(__funnelshift_l(b3, b2, sh), 0, 0x0123);
        v.w = prmt(__funnelshift_l(b4, b3, sh), 0, 0x0123);
        dq[k] = v;
    }
    const uint32_t done = h + (nq << 4);
    const uint32_t rem = fbytes - done;  // < 16
    if ((uint32_t)t < rem) {
        const uint32_t b = lead + done + (uint32_t)t;
        dst[done + t] = (uint8_t)(sm.bits[b >> 2] >> (24u - 8u * (b & 3u)));
    }
    if (t == kT - 1) {  // CRC-16 big-endian after the padded frame (frame_writer.zig:144-148)
        uint32_t a = 0, par = 0;
#pragma unroll
        for (int w = 0; w < kW; w++) { a ^= sm.crc_part[w]; par ^= sm.par_part[w]; }
        a = q_fold((a << 2) ^ (a << 1));  // * x^16 = x^2 + x  (mod x^15 + x + 1)
        const uint32_t flip = ((uint32_t)__popc(a) ^ (uint32_t)__popc(par)) & 1u;
        const uint32_t crc = a ^ (flip ? 0x8003u : 0u);  // CRT with the parity (mod x + 1)
        dst[fbytes] = (uint8_t)(crc >> 8);
        dst[fbytes + 1] = (uint8_t)crc;
    }
}

// Frame header (frame_writer.zig:151-265) + CRC-8 (:128-141), one byte per lane of one warp.  Covers what this kernel
// is launched for: block size 4096 (code 12, no block-size trailer), frame numbers below 2^31, any sample rate: one
// from the table has a code of its own (1..11), any other takes code 12 / 13 / 14 and a trailer of `rate_extra`
// bytes -- into which the reference writes the BLOCK SIZE (frame_writer.zig:258-262), reproduced here.
// Returns this lane's byte; `len` is the header length including the CRC-8.
ZF_DEVICE uint32_t header_byte(const uint8_t *crc8tab, int lane, unsigned long long frame_number, uint32_t depth,
                               uint32_t ch_type, uint32_t rate_code_v, uint32_t rate_extra, uint32_t &len) {
    const uint32_t fn = (uint32_t)frame_number;
    // UTF-8-like number coder (:235-251): i continuation bytes of 6 bits, lowest group last
    const uint32_t i = fn < 0x80u ? 0u : fn < 0x800u ? 1u : fn < 0x10000u ? 2u : fn < 0x200000u ? 3u : fn < 0x4000000u ? 4u : 5u;
    len = 4u + 1u + i + rate_extra + 1u;
    const uint32_t dc = depth == 16 ? 8u : depth == 24 ? 12u : 14u;  // :221-233
    uint32_t b = 0;
    const uint32_t k = (uint32_t)lane;
    if (k == 0) b = 0xFFu;
    else if (k == 1) b = 0xF8u;  // fixed-blocksize stream, :163
    else if (k == 2) b = (12u << 4) | rate_code_v;
    else if (k == 3) b = (ch_type << 4) | dc;
    else if (k == 4) b = i == 0 ? fn : (((0xFEu << (6u - i)) | (fn >> (6u * i))) & 0xFFu);
    else if (k <= 4u + i) {
        const uint32_t gi = 4u + i - k;  // group index, 0 = least significant
        b = 0x80u + ((fn >> (6u * gi)) & 0x3fu);
        if (gi == 4u) b &= 0x0Fu;  // u36 shift truncation in the reference for numbers >= 2^26
    } else if (k <= 4u + i + rate_extra) {
        // :260-261: the block size where the rate should be -- 8 bits (code 12), 16 bits (13), block size / 10 (14)
        const uint32_t v = rate_code_v == 14u ? (uint32_t)kN / 10u : (uint32_t)kN;
        b = (rate_extra == 2u && k == 5u + i) ? (v >> 8) : (v & 0xFFu);
    }
    // :260 writes the 13-bit block size through an 8-bit field: its high bits OR into the byte in front (Q10)
    if (rate_extra == 1u && k == 4u + i) b |= (uint32_t)kN >> 8;
    // CRC-8 (poly 0x07, init 0) is linear: byte k contributes T^(len-1-k)[byte], T = one table step
    uint32_t v = (k < len - 1u) ? b : 0u;
    const uint32_t steps = len - 1u - (k < len - 1u ? k : len - 1u);
#pragma unroll 1
    for (uint32_t r = 0; r < steps; r++) v = crc8tab[v];
    const uint32_t crc = reduce_xor(v);
    if (k == len - 1u) b = crc;
    return b;
}

// Pass 1 of 32-bit PCM: fixed.bestOrder's chains (fixed.zig:85-167) two candidates at a time -- trip 0: L, R; trip 1: M, S --
// with CH = ChainN (32-bit registers, exact overflow flag) or ChainW (64-bit).  Leaves per warp and candidate in sc.red:
// 5 x 3 sixteen-bit pieces of the sums, 5 x (lo, hi) range ORs, (lo, hi) sample OR, overflow flag.
template <int BYTES, typename CH>
ZF_DEVICE void pass1_pairs(Smem<BYTES> &sm, int t, int lane, int warp) {
    typedef typename CH::V V;
#pragma unroll 1
    for (uint32_t trip = 0; trip < 2; trip++) {
        CH ca, cb;
        ca.init(); cb.init();
        {
            int32_t L[4], R[4];
            load4<BYTES>(sm.raw, t, -1, L, R);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                ca.template step<false>(trip ? CH::mid(L[q], R[q]) : (V)L[q]);
                cb.template step<false>(trip ? cb.side(L[q], R[q], false) : (V)R[q]);
            }
        }
        const uint32_t slot_a = trip ? 2u : 0u, slot_b = trip ? 3u : 1u;
#pragma unroll 1
        for (int g = 0; g < kS / 4; g++) {
            int32_t L[4], R[4];
            load4<BYTES>(sm.raw, t, g, L, R);
            V xa[4], xb[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                xa[q] = trip ? CH::mid(L[q], R[q]) : (V)L[q];
                xb[q] = trip ? cb.side(L[q], R[q], true) : (V)R[q];
            }
            if (t == 0 && g == 0) {  // the frame's first samples; they are the warm-ups too
                ca.template step_first<0>(xa[0]); ca.template step_first<1>(xa[1]);
                ca.template step_first<2>(xa[2]); ca.template step_first<3>(xa[3]);
                cb.template step_first<0>(xb[0]); cb.template step_first<1>(xb[1]);
                cb.template step_first<2>(xb[2]); cb.template step_first<3>(xb[3]);
#pragma unroll
                for (int q = 0; q < 4; q++) { sm.warm[slot_a][q] = xa[q]; sm.warm[slot_b][q] = xb[q]; }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    ca.template step<true>(xa[q]);
                    cb.template step<true>(xb[q]);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const CH &C = h ? cb : ca;
            uint32_t *rp = &sm.sc.red[warp][h ? slot_b : slot_a][0];
            const unsigned long long sv[5] = {C.s0, C.s1, C.s2, C.s3, C.s4};
            const unsigned long long rv[5] = {C.r0, C.r1, C.r2, C.r3, C.r4};
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const uint32_t p0 = reduce_add((uint32_t)sv[k] & 0xffffu);
                const uint32_t p1 = reduce_add((uint32_t)(sv[k] >> 16) & 0xffffu);
                const uint32_t p2 = reduce_add((uint32_t)(sv[k] >> 32));  // per-thread sums < 2^42
                const uint32_t o0 = reduce_or((uint32_t)rv[k]), o1 = reduce_or((uint32_t)(rv[k] >> 32));
                if (lane == 0) {
                    rp[3 * k] = p0; rp[3 * k + 1] = p1; rp[3 * k + 2] = p2;
                    rp[15 + 2 * k] = o0; rp[16 + 2 * k] = o1;
                }
            }
            // the sample OR only serves ctz / == 0: a 33-bit side sample has a one among its low 32 bits (|S| < 2^32)
            const unsigned long long ov = (unsigned long long)(long long)C.orv;
            const uint32_t w0 = reduce_or((uint32_t)ov), w1 = reduce_or((uint32_t)(ov >> 32));
            const uint32_t fl = reduce_or(C.overflowed());
            if (lane == 0) { rp[25] = w0; rp[26] = w1; rp[27] = fl; }
        }
    }
}

// All warps (converged): warp w brings in rows 32 w .. 32 w + 31 of a frame, one bulk copy per group of the padded layout,
// issued by an elected lane from uniform registers.  Thread 0 arms the barrier with the frame's byte count (a copy that
// completes first only drives the transaction count negative for a moment; the phase cannot end before that arrival).
template <int BYTES>
ZF_DEVICE void fetch_frame(Smem<BYTES> &sm, const uint8_t *frame_pcm, int warp) {
    typedef Lay<BYTES> LY;
    constexpr int kPer = 32 / LY::kGroupRows;
    const int wu = __shfl_sync(0xffffffffu, warp, 0);
    fence_proxy_async();  // the reads of the frame that is being replaced were generic-proxy accesses
    uint32_t *dst = sm.raw + LY::kFrontW + wu * (kPer * LY::kPitchW);
    const uint8_t *src = frame_pcm + (size_t)wu * (size_t)(kPer * LY::kGroupW * 4);
#pragma unroll 4
    for (int k = 0; k < kPer; k++)
        tma_load_1d_elect(dst + k * LY::kPitchW, src + (size_t)k * (size_t)(LY::kGroupW * 4), (uint32_t)LY::kGroupW * 4u, &sm.mbar);
}

template <int BYTES>
__global__ void __launch_bounds__(kT, CtasPerSm<BYTES>::value) zf_encode_stereo_v3_kernel(const FrameJob job) {
    extern __shared__ __align__(16) unsigned char zf_smem[];
    Smem<BYTES> &sm = *reinterpret_cast<Smem<BYTES> *>(zf_smem);
    Scratch<BYTES> &sc = sm.sc;
    constexpr bool WIDE = BYTES == 4;
    typedef typename Arith<BYTES>::T T;
    typedef typename Arith<BYTES>::U U;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr uint32_t depth = 8u * BYTES;
    constexpr uint32_t frame_bytes = (uint32_t)kN * 2u * BYTES;
    const bool tma = job.use_tma != 0;

    {   // once per CTA: CRC-8 table, zero pad, bit buffer and level partials, barrier, first frame
        uint32_t v8 = (uint32_t)t;
#pragma unroll
        for (int k = 0; k < 8; k++) v8 = (v8 & 0x80u) ? ((v8 << 1) ^ 0x07u) : (v8 << 1);
        sm.crc8tab[t] = (uint8_t)v8;
        if (t < Lay<BYTES>::kFrontW) sm.raw[t] = 0;  // "row -1": the history of thread 0
        uint4 *bz = reinterpret_cast<uint4 *>(sm.bits);
        const uint4 z = {0, 0, 0, 0};
        for (int k = t; k < (BitBufWords<BYTES>::value + 8) / 4; k += kT) bz[k] = z;
        for (int k = t; k < 4 * 9 * 20; k += kT) (&sm.lvlcost[0][0][0])[k] = 0;
        for (int k = t; k < 4 * 9 * 16; k += kT) (&sm.lvlfive[0][0][0])[k] = 0;
        if (t == 0) {
            if (tma) {
                mbar_init(&sm.mbar, 1);
                fence_mbar_init();
            }
            sm.next_frame = atomicAdd(job.ticket, 1u);
        }
    }
    __syncthreads();
    uint32_t phase = 0;
    uint32_t f = __shfl_sync(0xffffffffu, sm.next_frame, 0);
    if (tma && f < job.n_frames) {
        if (t == 0) mbar_expect_tx(&sm.mbar, frame_bytes);
        fetch_frame<BYTES>(sm, job.pcm + (size_t)f * job.frame_stride, warp);
    }
    uint32_t rate_extra;
    const uint32_t rate_code_v = rate_code(job.sample_rate, rate_extra);
    LbArgs lba;
    lba.desc = job.desc; lba.total_bytes = job.total_bytes; lba.status = job.status; lba.out_cap = job.out_cap;
    lba.batch_frames = job.batch_frames;
    Pend P;
    P.valid = 0; P.fidx = 0; P.fbytes = 0; P.lead = 0; P.fits = 0;

    while (f < job.n_frames) {
        const uint32_t fidx = job.frame_base + f;
        const unsigned long long frame_number = job.first_frame_number + fidx;
        if (__builtin_expect(tma, 1)) {
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        } else {
            // source not 16-byte aligned (a caller's device pointer): plain loads into the same padded layout
            const uint8_t *src = job.pcm + (size_t)f * job.frame_stride;
            constexpr uint32_t kRowW = (uint32_t)Lay<BYTES>::kRowW;
            if ((((uintptr_t)src) & 3u) == 0) {
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
                for (uint32_t k = t; k < frame_bytes / 4u; k += kT) sm.raw[Lay<BYTES>::row((int)(k / kRowW)) + k % kRowW] = s32[k];
            } else {
                uint8_t *dst = reinterpret_cast<uint8_t *>(sm.raw);
                for (uint32_t k = t; k < frame_bytes; k += kT)
                    dst[4u * (uint32_t)Lay<BYTES>::row((int)(k / (4u * kRowW))) + k % (4u * kRowW)] = src[k];
            }
            __syncthreads();
        }

        // ================= pass 1: the four candidates =================
        // fixed.bestOrder (fixed.zig:85-167) + calcWasteBits' OR (encoder.zig:556-570), streamed: five trips over one
        // piece of code, four inter-channel samples per trip (the first trip runs the chains over the history only),
        // all four candidate channels L, R, M = (L + R) >> 1, S = L - R side by side.
        bool wide_frame = false;  // 32-bit PCM only: this frame needs the 64-bit paths
        // 16/24-bit PCM: this thread's sum |delta^k x| of every candidate -- they ARE the finest partitions' abs-sums of
        // rice.calcSums for whichever order is chosen (a thread's sixteen samples are one finest partition), so pass 2 need
        // not add them up again
        uint32_t ks[WIDE ? 1 : 4][WIDE ? 1 : 5];
#pragma unroll 1
        for (;;) {
        if constexpr (WIDE) {
            // 32-bit PCM: two candidates at a time (registers), two trips over the samples -- in 32-bit registers with
            // exact overflow detection (ChainN) first; a frame that raises the flag comes back here for the 64-bit chains
            if (wide_frame) pass1_pairs<BYTES, ChainW>(sm, t, lane, warp);
            else pass1_pairs<BYTES, ChainN>(sm, t, lane, warp);
        } else
        {
            typename Pass1Chain<BYTES>::type c0, c1, c2, c3;
            c0.init(); c1.init(); c2.init(); c3.init();
            {
                int32_t L[4], R[4];
                load4<BYTES>(sm.raw, t, -1, L, R);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    c0.template step<false>(L[q]);
                    c1.template step<false>(R[q]);
                    c2.template step<false>((L[q] + R[q]) >> 1);
                    c3.template step<false>(L[q] - R[q]);
                }
            }
#pragma unroll 1
            for (int g = 0; g < kS / 4; g++) {
                int32_t L[4], R[4];
                load4<BYTES>(sm.raw, t, g, L, R);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    c0.template step<true>(L[q]);
                    c1.template step<true>(R[q]);
                    c2.template step<true>((L[q] + R[q]) >> 1);
                    c3.template step<true>(L[q] - R[q]);
                }
            }
            if (t == 0) {
                // total[k] only counts i >= k (fixed.zig:102-127): take the first terms out again (the history of
                // thread 0 is the zero pad, so the chain produced exactly these terms); keep the warm-up samples
                int32_t L[4], R[4];
                load4<BYTES>(sm.raw, 0, 0, L, R);
#define ZF3_FIX(C, SLOT, EXPR)                                                          \
    {                                                                                   \
        int32_t f1p = 0, f2p = 0, f3p = 0, fxp = 0;                                     \
        _Pragma("unroll") for (int q = 0; q < 4; q++) {                                 \
            const int32_t x = (EXPR);                                                   \
            const int32_t e1 = x - fxp, e2 = e1 - f1p, e3 = e2 - f2p, e4 = e3 - f3p;    \
            fxp = x; f1p = e1; f2p = e2; f3p = e3;                                      \
            C.take_out(q < 1 ? uabs(e1) : 0u, q < 2 ? uabs(e2) : 0u, q < 3 ? uabs(e3) : 0u, uabs(e4)); \
            sm.warm[SLOT][q] = x;                                                       \
        }                                                                               \
    }
                ZF3_FIX(c0, 0, L[q]) ZF3_FIX(c1, 1, R[q]) ZF3_FIX(c2, 2, (L[q] + R[q]) >> 1) ZF3_FIX(c3, 3, L[q] - R[q])
#undef ZF3_FIX
            }
#define ZF3_RED(C, SLOT)                                                                         \
    {                                                                                            \
        uint32_t *rp = &sc.red[warp][SLOT][0];                                                   \
        uint32_t sv[5];                                                                          \
        C.sums(sv);                                                                              \
        _Pragma("unroll") for (int k = 0; k < 5; k++) {                                          \
            ks[SLOT][k] = sv[k];                                                                 \
            if (BYTES == 2) { /* 16 x 2^20 x 32 lanes < 2^32 */                                  \
                const uint32_t ws = reduce_add(sv[k]);                                           \
                if (lane == 0) { rp[2 * k] = ws; rp[2 * k + 1] = 0; }                            \
            } else {                                                                             \
                const uint32_t lo = reduce_add(sv[k] & 0xffffu), hi = reduce_add(sv[k] >> 16);   \
                if (lane == 0) { rp[2 * k] = lo; rp[2 * k + 1] = hi; }                           \
            }                                                                                    \
        }                                                                                        \
        const uint32_t wo = reduce_or(C.orv);                                                    \
        if (lane == 0) rp[10] = wo;                                                              \
    }
            ZF3_RED(c0, 0) ZF3_RED(c1, 1) ZF3_RED(c2, 2) ZF3_RED(c3, 3)
#undef ZF3_RED
        }
        __syncthreads();
        if (!wide_frame) {
            // the next frame's ticket: every thread has read the current one by now, and warp 1 only waits for warp 0 here
            if (t == 32) sm.next_frame = atomicAdd(job.ticket, 1u);
            // ... and so does the last warp: it fetches the look-back window of the previous frame and examines it (round 1)
            if (P.valid && warp == kW - 1) {
                lb_start(sm, lba, lane, P);
                if (!sm.lb_done) lb_step(sm, lba, lane, P);
            }
        }
        // ---- decide (one lane per candidate): encoder.zig:482-527, fixed.zig:160-166, rice.zig:97-104 ----
        if (warp == 0) {
            // lane 6 s + k folds value k of candidate s over the warps (k < 5: sum |delta^k x|, k = 5: sample OR); the
            // six values of a candidate are then exchanged inside its lane group
            const uint32_t s = (uint32_t)lane / 6u, kk = (uint32_t)lane % 6u;
            unsigned long long mine = 0, mine_rng = 0;
            uint32_t ovf_any = 0;
            if (lane < 24) {
                if constexpr (WIDE) {
                    if (kk < 5) {
                        unsigned long long p0 = 0, p1 = 0, p2 = 0;
                        uint32_t o0 = 0, o1 = 0;
#pragma unroll
                        for (int w = 0; w < kW; w++) {
                            p0 += sc.red[w][s][3 * kk]; p1 += sc.red[w][s][3 * kk + 1]; p2 += sc.red[w][s][3 * kk + 2];
                            o0 |= sc.red[w][s][15 + 2 * kk]; o1 |= sc.red[w][s][16 + 2 * kk];
                        }
                        mine = p0 + (p1 << 16) + (p2 << 32);
                        mine_rng = (unsigned long long)o0 | ((unsigned long long)o1 << 32);
                    } else {
                        uint32_t o0 = 0, o1 = 0;
#pragma unroll
                        for (int w = 0; w < kW; w++) { o0 |= sc.red[w][s][25]; o1 |= sc.red[w][s][26]; ovf_any |= sc.red[w][s][27]; }
                        mine = (unsigned long long)o0 | ((unsigned long long)o1 << 32);
                    }
                } else {
                    if (kk < 5) {
                        unsigned long long lo = 0, hi = 0;
#pragma unroll
                        for (int w = 0; w < kW; w++) { lo += sc.red[w][s][2 * kk]; hi += sc.red[w][s][2 * kk + 1]; }
                        mine = lo + (hi << 16);
                    } else {
                        uint32_t o = 0;
#pragma unroll
                        for (int w = 0; w < kW; w++) o |= sc.red[w][s][10];
                        mine = o;
                    }
                }
            }
            unsigned long long tot[5], rng[5];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                tot[k] = __shfl_sync(0xffffffffu, mine, (int)(6u * s) + k);
                rng[k] = WIDE ? __shfl_sync(0xffffffffu, mine_rng, (int)(6u * s) + k) : 0ull;
            }
            const unsigned long long orv = __shfl_sync(0xffffffffu, mine, (int)(6u * s) + 5);
            if (lane < 24 && kk == 0) {
            const uint32_t depth_ch = depth + (s == 3 ? 1u : 0u);
            Dec d;
            d.pad[0] = d.pad[1] = 0;
            d.waste = (orv == 0) ? depth_ch : (WIDE ? ctz64(orv) : ctz32((uint32_t)orv));
            d.bps = depth_ch - d.waste;
            d.order = 0;
            const uint32_t lim = d.bps > 16 ? 30u : 14u;
            d.P = lim < job.max_rice_param ? lim : job.max_rice_param;
            if (d.bps == 0) {  // :495-497
                d.kind = kConstant;
                d.est = 0;
            } else if (tot[1] == 0) {  // all samples equal, :498-500
                d.kind = kConstant;
                d.est = d.bps;
            } else {
                // every |delta^k x| is a multiple of 2^waste, so the shift commutes with the sums (and the range ORs)
                const bool check = WIDE && d.bps >= 28;  // "wide accumulator", encoder.zig:517-520
                uint32_t best = 0;
                unsigned long long bv = 0;
#pragma unroll
                for (uint32_t k = 0; k < 5; k++) {
                    unsigned long long v = tot[k] >> d.waste;
                    if (check && (rng[k] >> d.waste) > 0x7fffffffull) v = kU64Max;  // fixed.zig:160-162
                    if (k == 0 || v < bv) { bv = v; best = k; }                     // first minimum, fixed.zig:164
                }
                d.est = (uint32_t)kN * d.bps;
                if (check && bv == kU64Max) {  // no usable order: VERBATIM (fixed.zig:166, encoder.zig:521-527)
                    d.kind = kVerbatim;
                } else {
                    d.kind = kFixed;  // tentative: FIXED only if the Rice estimate beats VERBATIM (:538)
                    d.order = best;
                }
            }
            sm.dec[s] = d;
            }
            if constexpr (WIDE) {
                const uint32_t any = __ballot_sync(0xffffffffu, ovf_any != 0);
                if (lane == 0) sm.redo = (!wide_frame && any) ? 1u : 0u;
            }
        }
        __syncthreads();
        if constexpr (WIDE) {
            if (!wide_frame && sm.redo) {  // block-uniform
                wide_frame = true;
                continue;
            }
        }
        break;
        }  // pass 1 + decide
        bool use_wide = false;
        if constexpr (WIDE) use_wide = wide_frame;
        uint32_t lsum[4] = {0, 0, 0, 0};  // the leaf abs-sum of every candidate at its chosen order
        if constexpr (!WIDE) {
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                const uint32_t o = sm.dec[s4].order;
                lsum[s4] = o == 0 ? ks[s4][0] : o == 1 ? ks[s4][1] : o == 2 ? ks[s4][2] : o == 3 ? ks[s4][3] : ks[s4][4];
            }
        }
#ifdef ZF_HOST_EMU
        if (t == 0) (use_wide ? g_emu_wide_frames : g_emu_narrow_frames)++;  // test harness: which path a frame took
#endif

        // ================= pass 2: leaf statistics of the chosen order, level-8 search, tree levels 7..3 ==========
        {
// leaf statistics of one candidate: residual of the chosen order in place, then abs-sum, minimum and maximum of the
// thread's sixteen residuals (partition 0 skips the warm-ups, rice.zig:308)
#define ZF3_LEAF(SLOT, X, SUM, MN, MX)                                                              \
    {                                                                                               \
        const uint32_t order = sm.dec[SLOT].order;                                                  \
        diff_in_place(X, order);                                                                    \
        const uint32_t jstart = (t == 0) ? order : 0u;                                              \
        _Pragma("unroll") for (int j = 0; j < kS; j++) {                                            \
            int32_t r = X[kH + j];                                                                  \
            if (j < 4) r = ((uint32_t)j >= jstart) ? r : 0;                                         \
            SUM += uabs(r);                                                                         \
            MN = r < MN ? r : MN;                                                                   \
            MX = r > MX ? r : MX;                                                                   \
        }                                                                                           \
    }
// the same without the abs-sum (16/24-bit PCM: pass 1 has it)
#define ZF3_LEAF_MM(SLOT, X, MN, MX)                                                                \
    {                                                                                               \
        const uint32_t order = sm.dec[SLOT].order;                                                  \
        diff_in_place(X, order);                                                                    \
        const uint32_t jstart = (t == 0) ? order : 0u;                                              \
        _Pragma("unroll") for (int j = 0; j < kS; j++) {                                            \
            int32_t r = X[kH + j];                                                                  \
            if (j < 4) r = ((uint32_t)j >= jstart) ? r : 0;                                         \
            MN = r < MN ? r : MN;                                                                   \
            MX = r > MX ? r : MX;                                                                   \
        }                                                                                           \
    }
            if (use_wide) {
                // 64-bit residual chains, one candidate per trip; the residual itself is the low 32 bits (fixed.zig:70-73)
#pragma unroll 1
                for (uint32_t s = 0; s < 4; s++) {
                    if (sm.dec[s].kind != kFixed) continue;
                    const uint32_t order = sm.dec[s].order, waste = sm.dec[s].waste, P = sm.dec[s].P;
                    long long x[kXn];
                    load_x<BYTES>(sm.raw, t, s, x);
                    diff_in_place(x, order);
                    const uint32_t jstart = (t == 0) ? order : 0u;
                    int32_t mn = 0, mx = 0;
                    unsigned long long S = 0;
#pragma unroll
                    for (int j = 0; j < kS; j++) {
                        int32_t r = (int32_t)(x[kH + j] >> waste);
                        if (j < 4) r = ((uint32_t)j >= jstart) ? r : 0;  // partition 0 skips the warm-ups, rice.zig:308
                        S += uabs(r);
                        mn = r < mn ? r : mn;
                        mx = r > mx ? r : mx;
                    }
                    const uint32_t zm = zigzag(mn), zx = zigzag(mx);
                    uint32_t B = bitlen32(zm > zx ? zm : zx);
                    uint32_t choice, cost;
                    best_param_w(S, B, (uint32_t)kS - jstart, P, choice, cost);
                    sm.choice[s][256 + t] = (uint8_t)choice;
                    const uint32_t wc = reduce_add(cost);
                    const uint32_t wf = __ballot_sync(0xffffffffu, choice < 0x80u && choice > 14u);
                    if (lane == 0) { sm.lvlcost[s][8][warp] = wc; sm.lvlfive[s][8][warp] = wf ? 1 : 0; }
#pragma unroll
                    for (int lv = 7; lv >= 3; lv--) {
                        const int stride = 1 << (7 - lv);
                        S += __shfl_xor_sync(0xffffffffu, S, stride);
                        const uint32_t ob = __shfl_xor_sync(0xffffffffu, B, stride);
                        B = ob > B ? ob : B;
                        if ((lane & (2 * stride - 1)) == 0)
                            sc.node[s][(1u << lv) + ((uint32_t)t >> (8 - lv))] = S | ((unsigned long long)B << 48);
                    }
                }
            } else
#pragma unroll 1
            for (uint32_t it = 0; it < 2; it++) {
                int32_t A[kXn], B[kXn];
                unpack20<BYTES, 0>(sm.raw, t, A, B);
                if (it == 0) {
#pragma unroll
                    for (int i = 0; i < kXn; i++) {
                        const int32_t sd = A[i] - B[i];
                        B[i] = mid32<BYTES>(A[i], B[i]);
                        A[i] = sd;
                    }
                }
                const uint32_t slot_a = it ? 0u : 3u, slot_b = it ? 1u : 2u;
                const bool fix_a = sm.dec[slot_a].kind == kFixed, fix_b = sm.dec[slot_b].kind == kFixed;
                typename Arith<BYTES>::U sum_a = 0, sum_b = 0;  // sixteen residuals below 2^31 each with 32-bit PCM
                int32_t mn_a = 0, mx_a = 0, mn_b = 0, mx_b = 0;
                if constexpr (WIDE) {
                    if (fix_a) ZF3_LEAF(slot_a, A, sum_a, mn_a, mx_a)
                    if (fix_b) ZF3_LEAF(slot_b, B, sum_b, mn_b, mx_b)
                } else {
                    if (fix_a) ZF3_LEAF_MM(slot_a, A, mn_a, mx_a)
                    if (fix_b) ZF3_LEAF_MM(slot_b, B, mn_b, mx_b)
                    sum_a = it ? lsum[0] : lsum[3];
                    sum_b = it ? lsum[1] : lsum[2];
                }
                // width, leaf parameter, tree levels 7..3 of both candidates (FIXED or not: what the others yield is never
                // looked at), interleaved: the chains are long and dependent
#pragma unroll
                for (uint32_t h = 0; h < 2; h++) {
                    const uint32_t slot = h ? slot_b : slot_a;
                    const uint32_t order = sm.dec[slot].order, waste = sm.dec[slot].waste, P = sm.dec[slot].P;
                    const uint32_t jstart = (t == 0) ? order : 0u;
                    const int32_t mn = (h ? mn_b : mn_a) >> waste, mx = (h ? mx_b : mx_a) >> waste;
                    const uint32_t zm = zigzag(mn), zx = zigzag(mx);
                    uint32_t B = bitlen32(zm > zx ? zm : zx);        // bit length of the OR of the zigzags
                    const typename Arith<BYTES>::U S32 = (h ? sum_b : sum_a) >> waste;  // rice.calcSums, rice.zig:288-340
                    uint32_t choice, cost;
                    if constexpr (WIDE) best_param_w(S32, B, (uint32_t)kS - jstart, P, choice, cost);
                    else best_param_32(S32, B, (uint32_t)kS - jstart, P, choice, cost);
                    sm.choice[slot][256 + t] = (uint8_t)choice;
                    const uint32_t wc = reduce_add(cost);
                    const uint32_t wf = __ballot_sync(0xffffffffu, choice < 0x80u && choice > 14u);
                    if (lane == 0) { sm.lvlcost[slot][8][warp] = wc; sm.lvlfive[slot][8][warp] = wf ? 1 : 0; }
                    unsigned long long S = S32;
#pragma unroll
                    for (int lv = 7; lv >= 3; lv--) {
                        const int stride = 1 << (7 - lv);
                        S += __shfl_xor_sync(0xffffffffu, S, stride);
                        const uint32_t ob = __shfl_xor_sync(0xffffffffu, B, stride);
                        B = ob > B ? ob : B;
                        if ((lane & (2 * stride - 1)) == 0)
                            sc.node[slot][(1u << lv) + ((uint32_t)t >> (8 - lv))] = S | ((unsigned long long)B << 48);
                    }
                }
            }
#undef ZF3_LEAF
#undef ZF3_LEAF_MM
        }
        // round 2 of the look-back, if round 1 did not reach a prefix: its window has arrived during pass 2
        if (P.valid && warp == kW - 1 && !sm.lb_done) lb_step(sm, lba, lane, P);
        __syncthreads();
        // ================= round B: heap nodes 1..255 (levels 0..7), one per thread =================
        {
            const uint32_t m = (uint32_t)t;
            const uint32_t lvl = m ? floor_log2(m) : 0u;
            const uint32_t j = m - (1u << lvl);
            // every candidate, FIXED or not (what a CONSTANT / VERBATIM candidate yields is never looked at): without the
            // early exit two candidates' dependent chains interleave
#pragma unroll 2
            for (uint32_t s = 0; s < 4; s++) {
                const Dec &d = sm.dec[s];
                uint32_t choice = 0, cost = 0;
                unsigned long long S = 0;
                uint32_t B = 0;
                if (m >= 1) {
                    if (lvl < 3) {  // levels 2..0 straight from the eight level-3 nodes (heap 8..15): all eight are
                                    // loaded (independent, broadcast) and the ones of this node's span are kept
                        const ulonglong2 *np = reinterpret_cast<const ulonglong2 *>(&sc.node[s][8]);
                        unsigned long long nv[8];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const ulonglong2 v2 = np[k];
                            nv[2 * k] = v2.x;
                            nv[2 * k + 1] = v2.y;
                        }
#pragma unroll
                        for (uint32_t k = 0; k < 8; k++) {
                            const bool in = (k >> (3u - lvl)) == j;
                            S += in ? (nv[k] & 0xffffffffffffull) : 0ull;
                            const uint32_t b = in ? (uint32_t)(nv[k] >> 48) : 0u;
                            B = b > B ? b : B;
                        }
                    } else {
                        const unsigned long long v = sc.node[s][m];
                        S = v & 0xffffffffffffull;
                        B = (uint32_t)(v >> 48);
                    }
                }
                const uint32_t cnt = ((uint32_t)kN >> lvl) - (j == 0 ? d.order : 0u);  // rice.zig:356,371
                if constexpr (WIDE) {
                    best_param_w(S, B, cnt, d.P, choice, cost);
                } else {
                    if (__builtin_expect(__any_sync(0xffffffffu, (S >> 32) != 0), 0)) best_param_nw(S, B, cnt, d.P, choice, cost);
                    else best_param_32((uint32_t)S, B, cnt, d.P, choice, cost);
                }
                if (m >= 1) sm.choice[s][m] = (uint8_t)choice;
                else cost = 0;
                const bool five = m >= 1 && choice < 0x80u && choice > 14u;  // isRice2, rice.zig:74-76
                const uint32_t fm = __ballot_sync(0xffffffffu, five);
                if (warp == 0) {  // heap nodes 1..31: levels 0..4 -- each node's cost is one partial of its level
                    if (m >= 1) {
                        sm.lvlcost[s][lvl][j] = cost;
                        sm.lvlfive[s][lvl][j] = five ? 1 : 0;
                    }
                } else {  // warp 1: level 5; warps 2-3: level 6; warps 4-7: level 7
                    const uint32_t wsum = reduce_add(cost);
                    const uint32_t wl = floor_log2((uint32_t)warp);
                    if (lane == 0) {
                        sm.lvlcost[s][5u + wl][(uint32_t)warp - (1u << wl)] = wsum;
                        sm.lvlfive[s][5u + wl][(uint32_t)warp - (1u << wl)] = fm ? 1 : 0;
                    }
                }
            }
        }
        if (P.valid && warp == kW - 1) {
            while (!sm.lb_done) lb_step(sm, lba, lane, P);
        }
        __syncthreads();
        // the previous frame leaves the bit buffer (the scan's barrier orders these reads before the next stores; the
        // buffer is not cleared: see below)
        if (P.valid) copy_out(sm, job.out, job.out_cap, t, P);
        // ---- every warp: partition order per candidate (rice.zig:262-276: '<=' keeps the highest order on ties),
        //      FIXED only with a strictly smaller estimate than VERBATIM (encoder.zig:538), then the stereo mode:
        //      first minimum of [L+R, L+S, S+R, M+S] (encoder.zig:441-452).  Results are uniform over the block. ----
        Sub ua, ub;
        uint32_t ch_type = 1, sa = 0, sb = 1;
        {
            // lane 8 s + i looks at partition order i + 1 of candidate s (its 16 partial sums), lane 8 s also at order 0;
            // key = (bits << 5) | (15 - order) << 1 | method: the minimum is the smallest size, then the highest order
            uint32_t key = 0xffffffffu;
            {
                const uint32_t s = (uint32_t)lane >> 3, l = ((uint32_t)lane & 7u) + 1u;
                if (sm.dec[s].kind == kFixed) {
                    const uint4 *cp = reinterpret_cast<const uint4 *>(&sm.lvlcost[s][l][0]);
                    const uint4 c0 = cp[0], c1 = cp[1], c2 = cp[2], c3 = cp[3];
                    const uint4 fv = *reinterpret_cast<const uint4 *>(&sm.lvlfive[s][l][0]);
                    const uint32_t cost = (c0.x + c0.y + c0.z + c0.w) + (c1.x + c1.y + c1.z + c1.w) +
                                          (c2.x + c2.y + c2.z + c2.w) + (c3.x + c3.y + c3.z + c3.w);
                    const uint32_t method = (fv.x | fv.y | fv.z | fv.w) ? 1u : 0u;
                    const uint32_t bc = cost + ((4u + method) << l);  // :394
                    key = (bc << 5) | ((15u - l) << 1) | method;
                    if (l > job.max_rice_order) key = 0xffffffffu;  // rice.calcParams: orders above Config.max_rice_order are not tried
                    if (l == 1u) {
                        const uint32_t m0 = sm.lvlfive[s][0][0] ? 1u : 0u;
                        const uint32_t k0 = ((sm.lvlcost[s][0][0] + 4u + m0) << 5) | (15u << 1) | m0;
                        key = k0 < key ? k0 : key;
                    }
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other < key ? other : key;
                }
            }
            uint32_t kind[4], est[4], pom[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                kind[s] = sm.dec[s].kind;
                est[s] = sm.dec[s].est;
                pom[s] = 0;
                const uint32_t bk = __shfl_sync(0xffffffffu, key, 8 * s);
                if (kind[s] == kFixed) {
                    const uint32_t best = bk >> 5;
                    if (best < est[s]) { est[s] = best; pom[s] = (15u - ((bk >> 1) & 15u)) | ((bk & 1u) << 4); }
                    else kind[s] = kVerbatim;
                }
            }
            uint32_t bestv = est[0] + est[1], mode = 0;
            if (est[0] + est[3] < bestv) { bestv = est[0] + est[3]; mode = 1; }
            if (est[3] + est[1] < bestv) { bestv = est[3] + est[1]; mode = 2; }
            if (est[2] + est[3] < bestv) { bestv = est[2] + est[3]; mode = 3; }
            // Channel codes: indep(2) = 1, L/S 8, S/R 9, M/S 10 (type.zig:1-27)
            uint32_t ka = kind[0], kb = kind[1], pa = pom[0], pb = pom[1];
            if (mode == 1) { sb = 3; ch_type = 8; kb = kind[3]; pb = pom[3]; }
            else if (mode == 2) { sa = 3; ch_type = 9; ka = kind[3]; pa = pom[3]; }
            else if (mode == 3) { sa = 2; sb = 3; ch_type = 10; ka = kind[2]; pa = pom[2]; kb = kind[3]; pb = pom[3]; }
            ua.kind = ka; ua.po = pa & 15u; ua.method = pa >> 4;
            ub.kind = kb; ub.po = pb & 15u; ub.method = pb >> 4;
            ua.order = sm.dec[sa].order; ua.waste = sm.dec[sa].waste; ua.bps = sm.dec[sa].bps;
            ub.order = sm.dec[sb].order; ub.waste = sm.dec[sb].waste; ub.bps = sm.dec[sb].bps;
            ua.depth_ch = depth + (sa == 3 ? 1u : 0u);
            ub.depth_ch = depth + (sb == 3 ? 1u : 0u);
        }

        // ================= pack: count, scan, publish, write =================
        uint32_t va[kS], vb[kS];
        uint32_t len_a, len_b;
        if (use_wide) {
            len_a = sub_count<BYTES, true>(sm, t, sa, ua, va);
            len_b = sub_count<BYTES, true>(sm, t, sb, ub, vb);
        } else {
            len_a = sub_count<BYTES, false>(sm, t, sa, ua, va);
            len_b = sub_count<BYTES, false>(sm, t, sb, ub, vb);
        }
        uint32_t ex_a, ex_b, tot_a, tot_b;
        block_scan2(sm.scan, t, len_a, len_b, ex_a, ex_b, tot_a, tot_b);
        // 4 fixed bytes + the UTF-8-like frame number + the trailer of a sample rate without a code + CRC-8
        const uint32_t fn32 = (uint32_t)frame_number;
        const uint32_t hdr_bits = 8u * (6u + rate_extra + (fn32 >= 0x80u) + (fn32 >= 0x800u) + (fn32 >= 0x10000u) +
                                        (fn32 >= 0x200000u) + (fn32 >= 0x4000000u));
        const uint32_t total_bits = hdr_bits + tot_a + tot_b;
        const uint32_t fbytes = (total_bits + 7u) >> 3;
        const uint32_t size = fbytes + 2u;
        const uint32_t lead = (16u - (fbytes & 15u)) & 15u;  // leading zero bytes: the frame ends on a 16-byte boundary
        const bool fits = (lead + fbytes + 2u) <= (uint32_t)BitBufWords<BYTES>::value * 4u;
        if (t == 0) {
            job.frame_sizes[fidx] = size;
            if (fidx == 0) st_relaxed_gpu(job.desc, kFlagPrefix | (unsigned long long)size);
            else st_relaxed_gpu(job.desc + fidx, kFlagAggregate | (unsigned long long)size);
            if (!fits) atomicOr(job.status, kStatusBitOverflow);
        }
        {   // every thread has taken what it needs from the raw PCM (the scan's barrier): fetch the next frame
            const uint32_t nf = __shfl_sync(0xffffffffu, sm.next_frame, 0);
            if (tma && nf < job.n_frames) {
                if (t == 0) mbar_expect_tx(&sm.mbar, frame_bytes);
                fetch_frame<BYTES>(sm, job.pcm + (size_t)nf * job.frame_stride, warp);
            }
        }
        BitW wa, wb;
        wa.init(sm.bits, 8u * lead + hdr_bits + ex_a);
        wb.init(sm.bits, 8u * lead + hdr_bits + tot_a + ex_b);
        if (fits) {
            // Every word that some thread completes is stored whole; only words nobody completes depend on their old
            // content, because they only receive ORs (after the next barrier): the words in front of the first subframe
            // bit (leading zero bytes, frame header) and the frame's last word if it is partial.  Clear exactly those.
            {
                const uint32_t first_bit = 8u * lead + hdr_bits, end_bit = 8u * lead + total_bits;
                if ((uint32_t)t < (first_bit >> 5)) sm.bits[t] = 0;               // at most 8 words
                if (t == kT - 1 && (end_bit & 31u)) sm.bits[end_bit >> 5] = 0;
            }
#pragma unroll 1
            for (int ch = 0; ch < 2; ch++) {  // one copy of the header writer
                BitW w = ch ? wb : wa;
                sub_head(w, sm.warm[ch ? sb : sa], t, ch ? ub : ua);
                if (ch) wb = w;
                else wa = w;
            }
            // One field per sample: a Rice codeword (:363-372: q zeros, a one, k remainder bits), or w raw bits where the
            // partition is escaped (:341-354) or the subframe VERBATIM (:282-301), or nothing (CONSTANT).  As
            // (v >> k) zeros, then `one | (v & m)` in len0 bits: Rice k, 1 << k, (1 << k) - 1, k + 1; raw 32, 0, mask(w), w.
            uint32_t ka, onea, ma, lena, ja, kb, oneb, mb, lenb, jb;
#define ZF3_FIELD(U, K, ONE, M, LEN, JS)                                                                   \
    {                                                                                                      \
        const bool rice = U.kind == kFixed && !(U.choice & 0x80u);                                         \
        const uint32_t w = U.kind == kFixed ? (U.choice & 0x7fu) : U.kind == kVerbatim ? U.bps : 0u;       \
        K = rice ? w : 32u;                                                                                \
        ONE = rice ? 1u << (w & 31u) : 0u;                                                                 \
        M = rice ? ONE - 1u : ~shl32(0xffffffffu, w);                                                      \
        LEN = rice ? w + 1u : w;                                                                           \
        JS = (t == 0 && U.kind == kFixed) ? U.order : 0u;                                                  \
    }
            ZF3_FIELD(ua, ka, onea, ma, lena, ja)
            ZF3_FIELD(ub, kb, oneb, mb, lenb, jb)
#undef ZF3_FIELD
            const bool fast = ua.maxq + lena <= 32u && ub.maxq + lenb <= 32u;  // false for 33-bit samples as well
            if (__builtin_expect(fast, 1)) {  // every field has at most 32 bits: two independent chains, no branches
                // Four trips over one piece of code (the arrays rotate by four registers per trip): code size counts as
                // much as instruction count here
#pragma unroll 1
                for (int j0 = 0; j0 < kS; j0 += 4) {
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        // the warm-up samples of thread 0 are not coded: an empty field is a no-op
                        const bool oa = (uint32_t)(j0 + jj) >= ja, ob = (uint32_t)(j0 + jj) >= jb;
                        wa.put(oa ? (onea | (va[jj] & ma)) : 0u, oa ? (shr32(va[jj], ka) + lena) : 0u);
                        wb.put(ob ? (oneb | (vb[jj] & mb)) : 0u, ob ? (shr32(vb[jj], kb) + lenb) : 0u);
                    }
#pragma unroll
                    for (int jj = 0; jj < kS - 4; jj++) { va[jj] = va[jj + 4]; vb[jj] = vb[jj + 4]; }
                }
            } else {
#pragma unroll 1
                for (int ch = 0; ch < 2; ch++) {  // one copy of the general writer
                    uint32_t v[kS];
#pragma unroll
                    for (int j = 0; j < kS; j++) v[j] = ch ? vb[j] : va[j];
                    BitW w = ch ? wb : wa;
                    sub_body<WIDE>(w, ch ? ub : ua, v, ch ? kb : ka, ch ? oneb : onea, ch ? mb : ma, ch ? lenb : lena,
                                   ch ? jb : ja);
                    if (ch) wb = w;
                    else wa = w;
                }
            }
        }
        // frame header, one byte per lane of the last warp: worked out in front of the barrier (between the two barriers
        // the other seven warps would only wait for it), ORed in behind it
        uint32_t hbyte = 0, hlen = 0;
        if (fits && warp == kW - 1) hbyte = header_byte(sm.crc8tab, lane, frame_number, depth, ch_type, rate_code_v, rate_extra, hlen);
        __syncthreads();  // all complete words stored
        if (fits) {
            wa.or_tail();
            wb.or_tail();
            if ((uint32_t)lane < hlen) {  // hlen = 0 in the other warps
                const uint32_t b = lead + (uint32_t)lane;
                atomicOr(&sm.bits[b >> 2], hbyte << (24u - 8u * (b & 3u)));
            }
        }
        __syncthreads();

        // ================= CRC-16 of the finished frame (all warps); its copy-out is deferred =================
        {
            const uint32_t nwords_crc = (lead + fbytes) >> 2;  // a multiple of 4: the frame ends on a 16-byte boundary
            uint32_t acc_q = 0, par = 0;
            if (fits) {
                for (uint32_t j = (uint32_t)t; j * (uint32_t)kCrcChunkWords < nwords_crc; j += kT) {
                    const uint32_t end = nwords_crc - j * (uint32_t)kCrcChunkWords;  // exclusive
                    uint32_t a = 0;
                    // x^(8 * bytes after the chunk) mod P: an L2 round trip, started before the words are folded
                    const uint32_t pw = job.pow8[j * (uint32_t)kCrcChunkWords * 4u];
                    if (end >= (uint32_t)kCrcChunkWords) {
                        const uint4 *p = reinterpret_cast<const uint4 *>(sm.bits + (end - (uint32_t)kCrcChunkWords));
                        uint4 v = p[0];
#pragma unroll 1
                        for (int k = 1; k <= kCrcChunkWords / 4; k++) {
                            const uint4 nx = p[k];  // next four words meanwhile (the last trip reads into the padding)
                            // two words per dependent step: a x^64 + w0 x^32 + w1 with x^64 = x^8 + x^4 and
                            // x^32 = x^4 + x^2 (mod Q); the w0 term does not depend on a
                            const uint32_t f0 = q_fold(v.x), f2 = q_fold(v.z);
                            const uint32_t u0 = (f0 << 4) ^ (f0 << 2) ^ v.y, u2 = (f2 << 4) ^ (f2 << 2) ^ v.w;
                            a = q_fold((a << 8) ^ (a << 4) ^ u0);
                            a = q_fold((a << 8) ^ (a << 4) ^ u2);
                            par ^= v.x ^ v.y ^ v.z ^ v.w;
                            v = nx;
                        }
                    } else {
                        for (uint32_t k = 0; k < end; k++) {
                            const uint32_t v = sm.bits[k];
                            a = q_fold((a << 4) ^ (a << 2) ^ v);
                            par ^= v;
                        }
                    }
                    a = q_fold(a);  // < 2^15
                    acc_q ^= q_mulmod(a, q_fold(pw));
                }
            }
            acc_q = reduce_xor(acc_q);
            par = reduce_xor(par);
            if (lane == 0) { sm.crc_part[warp] = acc_q; sm.par_part[warp] = par; }
        }
        P.valid = 1; P.fidx = fidx; P.fbytes = fbytes; P.lead = lead; P.fits = fits ? 1u : 0u;
        f = sm.next_frame;  // written behind this frame's first barrier
    }
    if (P.valid) {  // the CTA's last frame
        if (warp == 0) {
            lb_start(sm, lba, lane, P);
            while (!sm.lb_done) lb_step(sm, lba, lane, P);
        }
        __syncthreads();
        copy_out(sm, job.out, job.out_cap, t, P);
    }
    if (t == 0) pdl_wait_primary();
}

}  // namespace v3
}  // namespace zf

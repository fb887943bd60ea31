// zf_kernel_v3.cuh -- lean full-frame stereo kernel for 16- and 24-bit PCM (block 4096, max_rice_order 8,
// max_rice_param 30).  Same decisions and the same bytes as zf_kernel.cuh / the oracle, arranged around the
// B200's issue limits (the path is integer-issue bound, not HBM bound -- DESIGN.md section 4):
//
//   * 256 threads x 16 samples: a thread owns exactly one finest Rice partition (level 8), so leaf statistics
//     never leave registers; 3 CTAs per SM overlap each other's barriers
//   * samples are NOT kept in registers across phases: every phase re-unpacks its 20-sample window from the raw
//     PCM in shared memory (8 LDS.128 + PRMT), which keeps the kernel spill-free at 80 registers
//   * all arithmetic is 32-bit (|delta^4| < 2^28, 16-sample sums < 2^32); only tree sums above a leaf are 64-bit
//   * pass 1 keeps the five per-leaf abs-sums in registers, so pass 2 only needs min/max of the chosen order
//   * one barrier per decision: decisions are taken by one warp and broadcast through shared memory
//   * bit writer: right-aligned 64-bit shift register, one funnel shift + one shared atomicOr per 32 output bits
//   * the stream is placed in the bit buffer so that it ENDS on a 16-byte boundary (leading zero bytes do not change
//     a CRC with zero init), CRC-16 is computed table-free in GF(2)[x]/(x^15+x+1) x parity
//     (x^16+x^15+x^2+1 = (x+1)(x^15+x+1)): Horner step  a <- a*(x^4+x^2) + word,  x^32 = x^4+x^2 (mod x^15+x+1)
//
// Reference call stack mirrored per frame: Encoder.writeFrame, encoder.zig:234-284 (see zf_kernel.cuh).
#pragma once
#include "zf_kernel.cuh"

namespace zf {
namespace v3 {

constexpr int kT = 256;          // threads per CTA
constexpr int kS = 16;           // samples per thread == finest partition (4096 >> 8)
constexpr int kH = 4;            // history samples in front of a thread's window
constexpr int kXn = kS + kH;
constexpr int kN = 4096;
constexpr int kW = kT / 32;
constexpr int kPadWords = 8;     // 32 zero bytes in front of the PCM: history of thread 0
constexpr int kCrcChunkWords = 32;
#ifndef ZF_V3_MIN_CTAS
#define ZF_V3_MIN_CTAS 3
#endif

struct Dec {  // per candidate channel
    uint32_t kind, order, waste, bps, P, po, method, est;
};

template <int BYTES>
struct Smem {
    alignas(16) uint32_t raw[kPadWords + kN * 2 * BYTES / 4];
    alignas(16) uint32_t bits[BitBufWords<BYTES>::value + 8];
    unsigned long long mbar;
    unsigned long long out_off;
    unsigned long long node[4][256];  // heap nodes 8..255 (levels 3..7): abs-sum | width << 48
    uint32_t red[kW][4][12];          // pass-1 warp partials: 5 x (lo, hi) + sample OR
    uint32_t mixed[4][32];            // round-B costs of heap nodes 1..31 (levels 0..4)
    uint32_t wcostA[4][kW];           // level-8 cost per warp
    uint32_t wcostB[4][kW];           // round-B cost per warp (warps 1..7 hold levels 5, 6, 6, 7, 7, 7, 7)
    uint32_t wfiveA[4][kW], wfiveB[4][kW];
    uint32_t mixfive[4];
    Dec dec[4];
    int32_t warm[4][4];               // first four samples of every candidate (warm-ups, CONSTANT value)
    uint32_t scan[2][kW];
    uint32_t crc_part[kW], par_part[kW];
    uint32_t cur_frame, next_frame;
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

template <int BYTES>
ZF_DEVICE void unpack20(const uint32_t *raw, int t, int32_t (&L)[kXn], int32_t (&R)[kXn]) {
    if (BYTES == 2) {
        // one word per inter-channel sample; the window starts 4 samples before 16 t
        const uint4 *p = reinterpret_cast<const uint4 *>(raw + kPadWords + kS * t - kH);
#pragma unroll
        for (int k = 0; k < kXn / 4; k++) {
            const uint4 v = p[k];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                L[4 * k + q] = (int32_t)prmt(w[q], 0, 0x9910);
                R[4 * k + q] = (int32_t)prmt(w[q], 0, 0xBB32);
            }
        }
    } else {
        // 6 bytes per inter-channel sample; 128 bytes from byte 96 t - 32 of the PCM (16-byte aligned), the window
        // (4 history + 16 own samples = 120 bytes) starts 8 bytes in
        const uint4 *p = reinterpret_cast<const uint4 *>(raw + kPadWords + 24 * t - 8);
        uint32_t w[32];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 v = p[k];
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < kXn / 2; k++) {  // two inter-channel samples per three words
            const uint32_t a = w[2 + 3 * k], b = w[3 + 3 * k], d = w[4 + 3 * k];
            L[2 * k] = (int32_t)prmt(a, b, 0xA210);
            R[2 * k] = (int32_t)prmt(a, b, 0xD543);
            L[2 * k + 1] = (int32_t)prmt(b, d, 0xC432);
            R[2 * k + 1] = (int32_t)prmt(d, 0, 0xB321);
        }
    }
}

// candidate channel SLOT of a stereo frame: 0 L, 1 R, 2 M = (L + R) >> 1, 3 S = L - R (encoder.zig:330-350)
template <int SLOT>
ZF_DEVICE void make_x(const int32_t (&L)[kXn], const int32_t (&R)[kXn], int32_t (&x)[kXn]) {
#pragma unroll
    for (int i = 0; i < kXn; i++) {
        if (SLOT == 0) x[i] = L[i];
        else if (SLOT == 1) x[i] = R[i];
        else if (SLOT == 2) x[i] = (L[i] + R[i]) >> 1;
        else x[i] = L[i] - R[i];
    }
}

struct P1 {
    uint32_t s[5];
    uint32_t orv;
};

// fixed.bestOrder (fixed.zig:85-167) on a 16-sample window: sum |delta^k x| for k = 0..4, and the OR of the samples
// (calcWasteBits, encoder.zig:556-570).  total[k] only counts i >= k (fixed.zig:102-127): thread 0 takes its
// first terms out again (its history is zero, so the terms are what the chain below produced).
ZF_DEVICE void pass1(const int32_t (&x)[kXn], int t, P1 &p) {
    const int32_t d32 = x[3] - x[2], d21 = x[2] - x[1], d10 = x[1] - x[0];
    int32_t e1p = d32, e2p = d32 - d21, e3p = (d32 - d21) - (d21 - d10);
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, orv = 0;
#pragma unroll
    for (int j = 0; j < kS; j++) {
        const int32_t e0 = x[kH + j];
        const int32_t e1 = e0 - x[kH + j - 1];
        const int32_t e2 = e1 - e1p;
        const int32_t e3 = e2 - e2p;
        const int32_t e4 = e3 - e3p;
        e1p = e1; e2p = e2; e3p = e3;
        orv |= (uint32_t)e0;
        s0 += uabs(e0); s1 += uabs(e1); s2 += uabs(e2); s3 += uabs(e3); s4 += uabs(e4);
    }
    if (t == 0) {
        int32_t f1p = 0, f2p = 0, f3p = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int32_t e0 = x[kH + j];
            const int32_t e1 = e0 - x[kH + j - 1];
            const int32_t e2 = e1 - f1p;
            const int32_t e3 = e2 - f2p;
            const int32_t e4 = e3 - f3p;
            f1p = e1; f2p = e2; f3p = e3;
            if (j < 1) s1 -= uabs(e1);
            if (j < 2) s2 -= uabs(e2);
            if (j < 3) s3 -= uabs(e3);
            s4 -= uabs(e4);
        }
    }
    p.s[0] = s0; p.s[1] = s1; p.s[2] = s2; p.s[3] = s3; p.s[4] = s4;
    p.orv = orv;
}

// `order` passes of in-place differencing: afterwards x[i] is the order-th difference for i >= order
ZF_DEVICE void diff_in_place(int32_t (&x)[kXn], uint32_t order) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if ((uint32_t)k < order) {
#pragma unroll
            for (int i = kXn - 1; i > k; i--) x[i] -= x[i - 1];
        }
    }
}

ZF_DEVICE uint32_t sel5(const uint32_t (&v)[5], uint32_t k) {
    uint32_t r = v[0];
    r = k == 1 ? v[1] : r;
    r = k == 2 ? v[2] : r;
    r = k == 3 ? v[3] : r;
    r = k == 4 ? v[4] : r;
    return r;
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
    const unsigned long long sh = S >> (p - 1u);
    const uint32_t sh32 = sh > 0x3fffffffull ? 0x3fffffffu : (uint32_t)sh;
    uint32_t cc = (1u + p) * n + sh32 - (n >> 1);
    uint32_t cand = p;
    const uint32_t c0 = S > 0x0fffffffull ? 0x7fffffffu : n + 2u * (uint32_t)S;  // p == 0: no -(n >> 1)
    if (c0 <= cc) { cc = c0; cand = 0; }                                         // lowest p wins ties
    if (cc < best) { best = cc; ch = cand; }                                     // escape wins ties
    choice = ch;
    cost = best;
}

// ---- bit writer: right-aligned 64-bit shift register ------------------------------------------------------------
// Branch-free: the flush is a predicated shared-memory reduction (the bit buffer is zeroed, so OR == store, and the
// first/last words a thread touches may be shared with its neighbours).
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
            atomicOr(wp, __funnelshift_r(lo, hi, nb - 32u));
            wp++;
        }
#else
        const uint32_t word = __funnelshift_r(lo, hi, nb);  // the shift uses nb mod 32
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ge.u32 p, %1, 32;\n"
            "@p red.shared.or.b32 [%0], %2;\n"
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
            while (q >= 32u) { put(0, 32); q -= 32u; }
            put(0, q);
            q = 0;
        }
        put(val, q + len);
    }
    ZF_DEVICE void finish() {
#ifdef ZF_HOST_EMU
        if (nb) atomicOr(wp, lo << (32u - nb));
#else
        if (nb) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wp), "r"(lo << (32u - nb)) : "memory");
#endif
    }
};

// ---- CRC-16 (poly 0x8005) in the factor ring ----------------------------------------------------------------------
// one folding round modulo Q = x^15 + x + 1: degree d -> max(14, d - 14)
ZF_DEVICE uint32_t q_fold(uint32_t v) {
    const uint32_t h = v >> 15;
    return (v & 0x7fffu) ^ h ^ (h << 1);
}
ZF_DEVICE uint32_t q_mulmod(uint32_t a, uint32_t b) {  // a, b < 2^15
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) acc ^= ((b >> i) & 1u) ? (a << i) : 0u;
    return q_fold(acc);  // 29-bit product: one round suffices
}

// candidate channel `slot` (runtime, uniform over the block) of this thread's window
template <int BYTES>
ZF_DEVICE void load_x(const uint32_t *raw, int t, uint32_t slot, int32_t (&x)[kXn]) {
    int32_t L[kXn], R[kXn];
    unpack20<BYTES>(raw, t, L, R);
    if (slot == 0) make_x<0>(L, R, x);
    else if (slot == 1) make_x<1>(L, R, x);
    else if (slot == 2) make_x<2>(L, R, x);
    else make_x<3>(L, R, x);
}

// what one thread knows about the subframe it helps to write
struct Sub {
    uint32_t kind, order, waste, bps, po, method;
    uint32_t depth_ch;
    uint32_t choice;
    uint32_t maxq;  // largest unary quotient among this thread's codes
    bool at_start;
};

// One subframe, counting side.  Fills v[] with what the writing side needs (zigzagged residuals for FIXED, shifted
// samples for VERBATIM) and returns this thread's bit count.  frame_writer.zig:269-372.
template <int BYTES>
ZF_DEVICE uint32_t sub_count(const Smem<BYTES> &sm, int t, uint32_t slot, Sub &u, uint32_t (&v)[kS]) {
    const Dec &d = sm.dec[slot];
    u.kind = d.kind; u.order = d.order; u.waste = d.waste; u.bps = d.bps; u.po = d.po; u.method = d.method;
    u.depth_ch = 8u * BYTES + (slot == 3 ? 1u : 0u);
    u.choice = 0;
    u.maxq = 0;
    u.at_start = false;
#pragma unroll
    for (int j = 0; j < kS; j++) v[j] = 0;
    if (u.kind == kConstant) return t == 0 ? 8u + u.depth_ch : 0u;
    int32_t x[kXn];
    load_x<BYTES>(sm.raw, t, slot, x);
    if (u.kind == kVerbatim) {
#pragma unroll
        for (int j = 0; j < kS; j++) v[j] = (uint32_t)(x[kH + j] >> u.waste);
        return (uint32_t)kS * u.bps + (t == 0 ? 8u + u.waste : 0u);
    }
    diff_in_place(x, u.order);
#pragma unroll
    for (int j = 0; j < kS; j++) v[j] = zigzag(x[kH + j] >> u.waste);
    const uint32_t sh = 8u - u.po;  // threads per partition = 1 << sh
    u.choice = sm.choice[slot][(1u << u.po) + ((uint32_t)t >> sh)];
    u.at_start = ((uint32_t)t & ((1u << sh) - 1u)) == 0;
    const uint32_t jstart = (t == 0) ? u.order : 0u;
    const uint32_t cnt = (uint32_t)kS - jstart;
    uint32_t bits = (t == 0) ? 8u + u.waste + u.order * u.bps + 6u : 0u;
    const bool esc = (u.choice & 0x80u) != 0;
    if (u.at_start) bits += 4u + u.method + (esc ? 5u : 0u);
    if (esc) return bits + (u.choice & 0x7fu) * cnt;
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

ZF_DEVICE void sub_write(uint32_t *bitbuf, const int32_t (&warm)[4], int t, const Sub &u, const uint32_t (&v)[kS], uint32_t pos) {
    BitW bw;
    bw.init(bitbuf, pos);
    if (u.kind == kConstant) {  // :269-279: 0x00, then the un-shifted sample at full depth
        if (t == 0) {
            bw.put(0, 8);
            bw.put((uint32_t)warm[0] & (0xffffffffu >> (32u - u.depth_ch)), u.depth_ch);
            bw.finish();
        }
        return;
    }
    const uint32_t smask = 0xffffffffu >> (32u - u.bps);
    if (u.kind == kVerbatim) {  // :282-301
        if (t == 0) {
            bw.put(u.waste ? 3u : 2u, 8);
            if (u.waste) bw.put(1u, u.waste);
        }
#pragma unroll
        for (int j = 0; j < kS; j++) bw.put(v[j] & smask, u.bps);
        bw.finish();
        return;
    }
    const uint32_t param_len = 4u + u.method;
    if (t == 0) {  // :303-329
        bw.put(((8u | u.order) << 1) | (u.waste ? 1u : 0u), 8);
        if (u.waste) bw.put(1u, u.waste);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++)
            if (k < u.order) bw.put((uint32_t)(warm[k] >> u.waste) & smask, u.bps);
        bw.put((u.method << 4) | u.po, 6);
    }
    const bool esc = (u.choice & 0x80u) != 0;
    if (u.at_start) {  // :341-357
        if (esc) {
            bw.put(u.method ? 31u : 15u, param_len);
            bw.put(u.choice & 0x7fu, 5);
        } else {
            bw.put(u.choice, param_len);
        }
    }
    const uint32_t jstart = (t == 0) ? u.order : 0u;
    if (esc) {
        const uint32_t wd = u.choice & 0x7fu;
        if (wd) {
            const uint32_t m = 0xffffffffu >> (32u - wd);
#pragma unroll
            for (int j = 0; j < kS; j++) {
                if (j >= 4 || (uint32_t)j >= jstart) {
                    const uint32_t r = (v[j] >> 1) ^ (0u - (v[j] & 1u));  // undo the zigzag
                    bw.put(r & m, wd);
                }
            }
        }
    } else {
        const uint32_t k = u.choice, one = 1u << k, m = one - 1u, len = k + 1u;
        if (u.maxq + len <= 32u) {  // every codeword fits one 32-bit field (all but pathological partitions)
#pragma unroll
            for (int j = 0; j < kS; j++) {
                const uint32_t q = v[j] >> k;
                if (j < 4) {
                    if ((uint32_t)j >= jstart) bw.put(one | (v[j] & m), q + len);
                } else {
                    bw.put(one | (v[j] & m), q + len);  // :363-372: q zeros, a one, k remainder bits
                }
            }
        } else {
#pragma unroll 1
            for (int j0 = 0; j0 < kS; j0 += 4) {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    // v[] must stay in registers: select instead of indexing
                    uint32_t z = v[jj];
                    z = j0 == 4 ? v[4 + jj] : z;
                    z = j0 == 8 ? v[8 + jj] : z;
                    z = j0 == 12 ? v[12 + jj] : z;
                    if ((uint32_t)(j0 + jj) >= jstart) bw.put_code(z >> k, one | (z & m), len);
                }
            }
        }
    }
    bw.finish();
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

// count + scan + publish + write of one frame for the chosen channel pair (sa, sb)
template <int BYTES>
ZF_DEVICE void pack_frame(Smem<BYTES> &sm, int t, const FrameJob &job, uint32_t fidx, unsigned long long frame_number,
                          uint32_t sa, uint32_t sb, uint32_t ch_type, uint32_t &total_bits_out, uint32_t &lead_out,
                          bool &fits_out) {
    uint32_t va[kS], vb[kS];
    Sub ua, ub;
    const uint32_t len_a = sub_count<BYTES>(sm, t, sa, ua, va);
    const uint32_t len_b = sub_count<BYTES>(sm, t, sb, ub, vb);
    uint32_t ex_a, ex_b, tot_a, tot_b;
    block_scan2(sm.scan, t, len_a, len_b, ex_a, ex_b, tot_a, tot_b);
    const uint32_t hdr_bits = 8u * header_len(frame_number, (uint32_t)kN, job.sample_rate);
    const uint32_t total_bits = hdr_bits + tot_a + tot_b;
    const uint32_t fbytes = (total_bits + 7u) >> 3;
    const uint32_t lead = (16u - (fbytes & 15u)) & 15u;  // leading zero bytes: the frame ends on a 16-byte boundary
    const bool fits = (lead + fbytes + 2u) <= (uint32_t)BitBufWords<BYTES>::value * 4u;
    if (t == 0) {
        const unsigned long long size = fbytes + 2u;
        job.frame_sizes[fidx] = (uint32_t)size;
        if (fidx == 0) st_relaxed_gpu(job.desc, kFlagPrefix | size);
        else st_relaxed_gpu(job.desc + fidx, kFlagAggregate | size);
        if (!fits) atomicOr(job.status, kStatusBitOverflow);
    }
    if (fits) {
        const uint32_t p0 = 8u * lead;
        if (t == kT - 1) {  // frame header + CRC-8 (frame_writer.zig:151-265, :128-141)
            uint8_t hb[16];
            uint32_t len = build_header(hb, frame_number, 8u * BYTES, ch_type, (uint32_t)kN, job.sample_rate);
            uint32_t crc = 0;
            for (uint32_t k = 0; k < len; k++) crc = sm.crc8tab[crc ^ hb[k]];
            hb[len++] = (uint8_t)crc;
            for (uint32_t k = 0; k < len; k++) {
                const uint32_t b = lead + k;
                atomicOr(&sm.bits[b >> 2], (uint32_t)hb[k] << (24u - 8u * (b & 3u)));
            }
        }
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++) {  // one copy of the writer code
            uint32_t v[kS];
#pragma unroll
            for (int j = 0; j < kS; j++) v[j] = ch ? vb[j] : va[j];
            const Sub u = ch ? ub : ua;
            sub_write(sm.bits, sm.warm[ch ? sb : sa], t, u, v, p0 + hdr_bits + (ch ? tot_a + ex_b : ex_a));
        }
    }
    total_bits_out = total_bits;
    lead_out = lead;
    fits_out = fits;
}

template <int BYTES>
__global__ void __launch_bounds__(kT, ZF_V3_MIN_CTAS) zf_encode_stereo_v3_kernel(const FrameJob job) {
    extern __shared__ __align__(16) unsigned char zf_smem[];
    Smem<BYTES> &sm = *reinterpret_cast<Smem<BYTES> *>(zf_smem);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr uint32_t depth = 8u * BYTES;
    constexpr uint32_t frame_bytes = (uint32_t)kN * 2u * BYTES;
    const bool tma = job.use_tma != 0;

    {   // once per CTA: CRC-8 table, zero pad and bit buffer, barrier, first frame
        uint32_t v8 = (uint32_t)t;
#pragma unroll
        for (int k = 0; k < 8; k++) v8 = (v8 & 0x80u) ? ((v8 << 1) ^ 0x07u) : (v8 << 1);
        sm.crc8tab[t] = (uint8_t)v8;
        if (t < kPadWords) sm.raw[t] = 0;
        uint4 *bz = reinterpret_cast<uint4 *>(sm.bits);
        const uint4 z = {0, 0, 0, 0};
        for (int k = t; k < (BitBufWords<BYTES>::value + 8) / 4; k += kT) bz[k] = z;
        if (t == 0) {
            if (tma) {
                mbar_init(&sm.mbar, 1);
                fence_mbar_init();
            }
            const uint32_t f = atomicAdd(job.ticket, 1u);
            sm.cur_frame = f;
            if (tma && f < job.n_frames) {
                mbar_expect_tx(&sm.mbar, frame_bytes);
                tma_load_1d(sm.raw + kPadWords, job.pcm + (size_t)f * job.frame_stride, frame_bytes, &sm.mbar);
            }
        }
    }
    __syncthreads();
    uint32_t phase = 0;

    for (;;) {
        const uint32_t f = sm.cur_frame;
        if (f >= job.n_frames) break;
        const uint32_t fidx = job.frame_base + f;
        const unsigned long long frame_number = job.first_frame_number + fidx;
        if (tma) {
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        } else {
            const uint8_t *src = job.pcm + (size_t)f * job.frame_stride;
            uint8_t *dst = reinterpret_cast<uint8_t *>(sm.raw + kPadWords);
            if ((((uintptr_t)src) & 3u) == 0) {
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
                for (uint32_t k = t; k < frame_bytes / 4u; k += kT) sm.raw[kPadWords + k] = s32[k];
            } else {
                for (uint32_t k = t; k < frame_bytes; k += kT) dst[k] = src[k];
            }
            __syncthreads();
        }

        // ================= pass 1: four candidates, sums kept in registers =================
        uint32_t ks[4][5];
        {
            int32_t L[kXn], R[kXn];
            unpack20<BYTES>(sm.raw, t, L, R);
#define ZF3_PASS1(SLOT)                                                                     \
    {                                                                                       \
        int32_t x[kXn];                                                                     \
        make_x<SLOT>(L, R, x);                                                              \
        P1 p;                                                                               \
        pass1(x, t, p);                                                                     \
        if (t == 0) {                                                                       \
            _Pragma("unroll") for (int k = 0; k < 4; k++) sm.warm[SLOT][k] = x[kH + k];     \
        }                                                                                   \
        _Pragma("unroll") for (int k = 0; k < 5; k++) {                                     \
            ks[SLOT][k] = p.s[k];                                                           \
            if (BYTES == 2) {                                                               \
                const uint32_t ws = reduce_add(p.s[k]);                                     \
                if (lane == 0) { sm.red[warp][SLOT][2 * k] = ws; sm.red[warp][SLOT][2 * k + 1] = 0; } \
            } else {                                                                        \
                const uint32_t lo = reduce_add(p.s[k] & 0xffffu), hi = reduce_add(p.s[k] >> 16); \
                if (lane == 0) { sm.red[warp][SLOT][2 * k] = lo; sm.red[warp][SLOT][2 * k + 1] = hi; } \
            }                                                                               \
        }                                                                                   \
        const uint32_t wo = reduce_or(p.orv);                                               \
        if (lane == 0) sm.red[warp][SLOT][10] = wo;                                         \
    }
            ZF3_PASS1(2) ZF3_PASS1(3) ZF3_PASS1(0) ZF3_PASS1(1)
#undef ZF3_PASS1
        }
        __syncthreads();
        // ---- decide (one lane per candidate): encoder.zig:482-527, fixed.zig:160-166, rice.zig:97-104 ----
        if (warp == 0 && lane < 4) {
            const uint32_t s = (uint32_t)lane;
            unsigned long long tot[5];
            uint32_t orv = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                unsigned long long lo = 0, hi = 0;
#pragma unroll
                for (int w = 0; w < kW; w++) { lo += sm.red[w][s][2 * k]; hi += sm.red[w][s][2 * k + 1]; }
                tot[k] = lo + (hi << 16);
            }
#pragma unroll
            for (int w = 0; w < kW; w++) orv |= sm.red[w][s][10];
            const uint32_t depth_ch = depth + (s == 3 ? 1u : 0u);
            Dec d;
            d.waste = (orv == 0) ? depth_ch : ctz32(orv);
            d.bps = depth_ch - d.waste;
            d.order = 0; d.po = 0; d.method = 0;
            const uint32_t lim = d.bps > 16 ? 30u : 14u;
            d.P = lim < job.max_rice_param ? lim : job.max_rice_param;
            if (d.bps == 0) {  // :495-497
                d.kind = kConstant;
                d.est = 0;
            } else if (tot[1] == 0) {  // all samples equal, :498-500
                d.kind = kConstant;
                d.est = d.bps;
            } else {
                // every |delta^k x| is a multiple of 2^waste, so the shift commutes with the sums
                uint32_t best = 0;
                unsigned long long bv = tot[0] >> d.waste;
#pragma unroll
                for (uint32_t k = 1; k < 5; k++) {
                    const unsigned long long v = tot[k] >> d.waste;
                    if (v < bv) { bv = v; best = k; }  // first minimum, fixed.zig:164
                }
                d.kind = kFixed;  // tentative: FIXED only if the Rice estimate beats VERBATIM (:538)
                d.order = best;
                d.est = (uint32_t)kN * d.bps;
            }
            sm.dec[s] = d;
        }
        __syncthreads();

        // ================= pass 2: leaf statistics of the chosen order, level-8 search, tree levels 7..3 ==========
        {
            uint32_t order[4], lsum[4];
            bool fixedk[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                fixedk[s] = sm.dec[s].kind == kFixed;
                order[s] = sm.dec[s].order;
                lsum[s] = sel5(ks[s], order[s]);
            }
            int32_t L[kXn], R[kXn];
            unpack20<BYTES>(sm.raw, t, L, R);
#define ZF3_LEAF(SLOT)                                                                              \
    if (fixedk[SLOT]) {                                                                             \
        const uint32_t waste = sm.dec[SLOT].waste, P = sm.dec[SLOT].P;                              \
        int32_t x[kXn];                                                                             \
        make_x<SLOT>(L, R, x);                                                                      \
        diff_in_place(x, order[SLOT]);                                                              \
        const uint32_t jstart = (t == 0) ? order[SLOT] : 0u;                                        \
        int32_t mn = 0, mx = 0;                                                                     \
        _Pragma("unroll") for (int j = 0; j < kS; j++) {                                            \
            int32_t r = x[kH + j];                                                                  \
            if (j < 4) r = ((uint32_t)j >= jstart) ? r : 0;                                         \
            mn = r < mn ? r : mn;                                                                   \
            mx = r > mx ? r : mx;                                                                   \
        }                                                                                           \
        mn >>= waste;                                                                               \
        mx >>= waste;                                                                               \
        const uint32_t zm = zigzag(mn), zx = zigzag(mx);                                            \
        uint32_t B = bitlen32(zm > zx ? zm : zx);      /* bit length of the OR of the zigzags */    \
        unsigned long long S = lsum[SLOT] >> waste;    /* rice.calcSums, rice.zig:288-340 */        \
        uint32_t choice, cost;                                                                      \
        best_param_nw(S, B, (uint32_t)kS - jstart, P, choice, cost);                                \
        sm.choice[SLOT][256 + t] = (uint8_t)choice;                                                 \
        const uint32_t wc = reduce_add(cost);                                                       \
        const uint32_t wf = __ballot_sync(0xffffffffu, choice < 0x80u && choice > 14u);             \
        if (lane == 0) { sm.wcostA[SLOT][warp] = wc; sm.wfiveA[SLOT][warp] = wf; }                  \
        _Pragma("unroll") for (int lv = 7; lv >= 3; lv--) {                                         \
            const int stride = 1 << (7 - lv);                                                       \
            S += __shfl_xor_sync(0xffffffffu, S, stride);                                           \
            const uint32_t ob = __shfl_xor_sync(0xffffffffu, B, stride);                            \
            B = ob > B ? ob : B;                                                                    \
            if ((lane & (2 * stride - 1)) == 0)                                                     \
                sm.node[SLOT][(1u << lv) + ((uint32_t)t >> (8 - lv))] = S | ((unsigned long long)B << 48); \
        }                                                                                           \
    }
            ZF3_LEAF(2) ZF3_LEAF(3) ZF3_LEAF(0) ZF3_LEAF(1)
#undef ZF3_LEAF
        }
        __syncthreads();
        // ================= round B: heap nodes 1..255 (levels 0..7), one per thread =================
#pragma unroll 1
        for (uint32_t s = 0; s < 4; s++) {
            const Dec &d = sm.dec[s];
            if (d.kind != kFixed) continue;
            const uint32_t m = (uint32_t)t;
            uint32_t choice = 0, cost = 0;
            if (m >= 1) {
                const uint32_t lvl = floor_log2(m);
                const uint32_t j = m - (1u << lvl);
                unsigned long long S;
                uint32_t B;
                if (lvl < 3) {  // levels 2..0 straight from the eight level-3 nodes (heap 8..15)
                    const uint32_t span = 8u >> lvl;
                    S = 0;
                    B = 0;
                    for (uint32_t k = 0; k < span; k++) {
                        const unsigned long long v = sm.node[s][8u + j * span + k];
                        S += v & 0xffffffffffffull;
                        const uint32_t b = (uint32_t)(v >> 48);
                        B = b > B ? b : B;
                    }
                } else {
                    const unsigned long long v = sm.node[s][m];
                    S = v & 0xffffffffffffull;
                    B = (uint32_t)(v >> 48);
                }
                const uint32_t cnt = ((uint32_t)kN >> lvl) - (j == 0 ? d.order : 0u);  // rice.zig:356,371
                best_param_nw(S, B, cnt, d.P, choice, cost);
                sm.choice[s][m] = (uint8_t)choice;
            }
            const bool five = m >= 1 && choice < 0x80u && choice > 14u;  // isRice2, rice.zig:74-76
            const uint32_t fm = __ballot_sync(0xffffffffu, five);
            if (warp == 0) {
                sm.mixed[s][lane] = cost;
                if (lane == 0) sm.mixfive[s] = fm;
            } else {
                const uint32_t wsum = reduce_add(cost);
                if (lane == 0) { sm.wcostB[s][warp] = wsum; sm.wfiveB[s][warp] = fm; }
            }
        }
        __syncthreads();
        // ---- partition order per candidate (rice.zig:262-276: '<=' keeps the highest order on ties); FIXED needs a
        //      strictly smaller estimate than VERBATIM (encoder.zig:538) ----
        if (warp < 4 && sm.dec[warp].kind == kFixed) {
            const uint32_t s = (uint32_t)warp;
            uint32_t key = 0xffffffffu, method = 0;
            if (lane <= 8) {
                uint32_t cost = 0, fv = 0;
                if (lane <= 4) {
                    for (uint32_t m = 1u << lane; m < (2u << lane); m++) cost += sm.mixed[s][m];
                    fv = (sm.mixfive[s] >> (1u << lane)) & ((1u << (1u << lane)) - 1u);
                } else if (lane <= 7) {
                    const uint32_t w0 = 1u << (lane - 5), w1 = 2u << (lane - 5);
                    for (uint32_t w = w0; w < w1; w++) { cost += sm.wcostB[s][w]; fv |= sm.wfiveB[s][w]; }
                } else {
                    for (uint32_t w = 0; w < (uint32_t)kW; w++) { cost += sm.wcostA[s][w]; fv |= sm.wfiveA[s][w]; }
                }
                method = fv ? 1u : 0u;
                const uint32_t bc = cost + ((4u + method) << lane);  // :394
                key = (bc << 4) | (15u - (uint32_t)lane);            // minimum cost, then the highest level
            }
            const uint32_t bk = reduce_min(key);
            const uint32_t bpo = 15u - (bk & 15u);
            const uint32_t bmethod = __shfl_sync(0xffffffffu, method, (int)bpo);
            if (lane == 0) {
                Dec &d = sm.dec[s];
                const uint32_t best = bk >> 4;
                if (best < d.est) { d.est = best; d.po = bpo; d.method = bmethod; }
                else d.kind = kVerbatim;
            }
        }
        __syncthreads();

        // ================= stereo mode: first minimum of [L+R, L+S, S+R, M+S], encoder.zig:441-452 =================
        uint32_t total_bits, lead;
        bool fits;
        {
            const uint32_t el = sm.dec[0].est, er = sm.dec[1].est, em = sm.dec[2].est, es = sm.dec[3].est;
            uint32_t bestv = el + er, mode = 0;
            if (el + es < bestv) { bestv = el + es; mode = 1; }
            if (es + er < bestv) { bestv = es + er; mode = 2; }
            if (em + es < bestv) { bestv = em + es; mode = 3; }
            // Channel codes: indep(2) = 1, L/S 8, S/R 9, M/S 10 (type.zig:1-27)
            uint32_t sa = 0, sb = 1, ch_type = 1;
            if (mode == 1) { sb = 3; ch_type = 8; }
            else if (mode == 2) { sa = 3; ch_type = 9; }
            else if (mode == 3) { sa = 2; sb = 3; ch_type = 10; }
            pack_frame<BYTES>(sm, t, job, fidx, frame_number, sa, sb, ch_type, total_bits, lead, fits);
        }
        __syncthreads();  // all codewords placed; the raw PCM is no longer needed
        if (t == 0) {
            const uint32_t nf = atomicAdd(job.ticket, 1u);
            sm.next_frame = nf;
            if (tma && nf < job.n_frames) {
                fence_proxy_async();
                mbar_expect_tx(&sm.mbar, frame_bytes);
                tma_load_1d(sm.raw + kPadWords, job.pcm + (size_t)nf * job.frame_stride, frame_bytes, &sm.mbar);
            }
        }

        // ================= finish: look-back (warp 0) | CRC-16 (warps 1..7) =================
        const uint32_t fbytes = (total_bits + 7u) >> 3;
        const uint32_t size = fbytes + 2u;
        const uint32_t nwords_crc = (lead + fbytes) >> 2;  // a multiple of 4: the frame ends on a 16-byte boundary
        if (warp == 0) {
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
                sm.out_off = excl;
                st_relaxed_gpu(job.desc + fidx, kFlagPrefix | (excl + size));
                if (fidx + 1 == job.batch_frames) *job.total_bytes = excl + size;
                if (excl + size > job.out_cap) atomicOr(job.status, kStatusOutOverflow);
            }
        } else {
            uint32_t acc_q = 0, par = 0;
            if (fits) {
                for (uint32_t j = (uint32_t)t - 32u; j * (uint32_t)kCrcChunkWords < nwords_crc; j += kT - 32) {
                    const uint32_t end = nwords_crc - j * (uint32_t)kCrcChunkWords;  // exclusive
                    uint32_t a = 0;
                    if (end >= (uint32_t)kCrcChunkWords) {
                        const uint4 *p = reinterpret_cast<const uint4 *>(sm.bits + (end - (uint32_t)kCrcChunkWords));
#pragma unroll
                        for (int k = 0; k < kCrcChunkWords / 4; k++) {
                            const uint4 v = p[k];
                            a = q_fold((a << 4) ^ (a << 2) ^ v.x);
                            a = q_fold((a << 4) ^ (a << 2) ^ v.y);
                            a = q_fold((a << 4) ^ (a << 2) ^ v.z);
                            a = q_fold((a << 4) ^ (a << 2) ^ v.w);
                            par ^= v.x ^ v.y ^ v.z ^ v.w;
                        }
                    } else {
                        for (uint32_t k = 0; k < end; k++) {
                            const uint32_t v = sm.bits[k];
                            a = q_fold((a << 4) ^ (a << 2) ^ v);
                            par ^= v;
                        }
                    }
                    a = q_fold(a);                                                // < 2^15
                    const uint32_t pw = job.pow8[j * (uint32_t)kCrcChunkWords * 4u];  // x^(8 * bytes after the chunk) mod P
                    acc_q ^= q_mulmod(a, q_fold(pw));
                }
            }
            acc_q = reduce_xor(acc_q);
            par = reduce_xor(par);
            if (lane == 0) { sm.crc_part[warp] = acc_q; sm.par_part[warp] = par; }
        }
        __syncthreads();
        // ---- copy-out: byte-shifted, word-coalesced; the CRC-16 goes straight to global memory ----
        {
            const unsigned long long off = sm.out_off;
            if (fits && off + size <= job.out_cap) {
                uint8_t *dst = job.out + off;
                const uint32_t head = (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u);
                const uint32_t h = head < fbytes ? head : fbytes;
                if ((uint32_t)t < h) {
                    const uint32_t b = lead + (uint32_t)t;
                    dst[t] = (uint8_t)(sm.bits[b >> 2] >> (24u - 8u * (b & 3u)));
                }
                const uint32_t nw = (fbytes - h) >> 2;
                uint32_t *dw = reinterpret_cast<uint32_t *>(dst + h);
                const uint32_t o = lead + h, w0 = o >> 2, sh = 8u * (o & 3u);
                for (uint32_t k = t; k < nw; k += kT) {
                    const uint32_t be = __funnelshift_l(sm.bits[w0 + k + 1], sm.bits[w0 + k], sh);
                    dw[k] = prmt(be, 0, 0x0123);
                }
                const uint32_t done = h + (nw << 2);
                const uint32_t rem = fbytes - done;  // < 4
                if ((uint32_t)t < rem) {
                    const uint32_t b = lead + done + (uint32_t)t;
                    dst[done + t] = (uint8_t)(sm.bits[b >> 2] >> (24u - 8u * (b & 3u)));
                }
                if (t == kT - 1) {  // CRC-16 big-endian after the padded frame (frame_writer.zig:144-148)
                    uint32_t a = 0, par = 0;
#pragma unroll
                    for (int w = 1; w < kW; w++) { a ^= sm.crc_part[w]; par ^= sm.par_part[w]; }
                    a = q_fold((a << 2) ^ (a << 1));  // * x^16 = x^2 + x  (mod x^15 + x + 1)
                    const uint32_t flip = ((uint32_t)__popc(a) ^ (uint32_t)__popc(par)) & 1u;
                    const uint32_t crc = a ^ (flip ? 0x8003u : 0u);  // CRT with the parity (mod x + 1)
                    dst[fbytes] = (uint8_t)(crc >> 8);
                    dst[fbytes + 1] = (uint8_t)crc;
                }
            }
        }
        __syncthreads();
        {   // zero what this frame used of the bit buffer
            uint4 *bz = reinterpret_cast<uint4 *>(sm.bits);
            const uint4 z = {0, 0, 0, 0};
            const uint32_t n4 = fits ? ((nwords_crc + 2u + 3u) >> 2) : (uint32_t)(BitBufWords<BYTES>::value + 8) / 4u;
            for (uint32_t k = t; k < n4; k += kT) bz[k] = z;
        }
        if (t == 0) sm.cur_frame = sm.next_frame;
        __syncthreads();
    }
}

}  // namespace v3
}  // namespace zf

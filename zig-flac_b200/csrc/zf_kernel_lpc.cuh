// zf_kernel_lpc.cuh -- stereo frame-encode kernel with LPC subframes (BASELINE.json config 4: "LPC order 12 with quantised
// coefficients"), 8/16/24-bit samples, any block size up to 4096, maximum order <= 12.
//
// THERE IS NO REFERENCE LPC: encoder.zig:626-640 is an unused enum, the union arm at :694-699 is commented out and
// readme.md:24-27 lists linear prediction as unfinished.  This kernel therefore follows a specification of its own,
// "zf-LPC v1" -- every arithmetic step is fixed (integer window and autocorrelation, Levinson-Durbin in IEEE double with
// one rounding per operation, deterministic order choice and quantisation) so that the result is bit-identical to the
// CPU statement of the same specification the tests check it against, and the stream is standard FLAC (SUBFRAME_LPC).
// Everything the reference does have is reused unchanged: the FIXED / VERBATIM / CONSTANT analysis of zf_kernel.cuh runs
// first, LPC is one more candidate per channel, and the Rice partition / parameter search is the reference's
// (rice.zig) with pred_order = the LPC order.  With zf_config.lpc_order == 0 this kernel is never launched.
//
// zf-LPC v1 per candidate channel (x = samples after the wasted-bits shift, bps bits, N >= 64 samples, max order M):
//   window   W[i] = 16384 - floor((2i - (N-1))^2 * 16384 / (N-1)^2) (host table), xw = ((x >> sh) * W) >> 14, sh = max(0, bps-24)
//   R[l]     = sum_{i >= l} xw[i] xw[i-l], l = 0..M, exact in int64 (any summation order)
//   a[]      Levinson-Durbin in double (lpc_levinson below: the order of operations is the specification)
//   order    smallest o with err[o] <= err[omax] (1 + delta)^(omax - o), delta = (bps + P) 2 ln 2 / N
//   q[]      P = 14 (bps <= 17) or 15 bits, shift = P - exponent(max |a|) <= 15, error feedback, floor(v + 0.5)
//   r[i]     = x[i] - ((sum_j q[j] x[i-1-j]) >> shift) in int64; all must fit 32 bits
//   choice   LPC iff rice_bits + o (bps + P) + 9 < bits of the FIXED/VERBATIM alternative + its order * bps
#pragma once
#include "zf_kernel.cuh"

namespace zf {
namespace lpc {

constexpr int kMaxOrder = 12;
constexpr int kPlanePad = 16;  // zero words in front of the plane: history of the first samples
constexpr uint32_t kLpc = 3;   // SlotDec.kind beside kConstant / kVerbatim / kFixed
constexpr uint32_t kMinBlock = 64;

struct Model {
    int32_t q[kMaxOrder];
    uint32_t valid, order, shift, precision;
};

template <int BYTES>
struct SmemLpc {
    SmemCommon c;
    alignas(16) uint32_t raw[kRawPadWords + kMaxBlock * 2 * BYTES / 4];
    alignas(16) uint32_t bits[BitBufWords<BYTES>::value];
    alignas(16) int32_t plane[kPlanePad + kMaxBlock];
    uint16_t win[kMaxBlock];
    long long R[4][kMaxOrder + 1];
    unsigned long long lred[kWarps][kMaxOrder + 1];
    Model model[4];
    SlotDec fx_dec[4];
    uint8_t fx_choice[4][kNodes];
    long long warm[2][kMaxOrder];
    uint32_t bad[4];
    uint32_t keep_fixed[4];
};

ZF_DEVICE int exponent_of(double v) {  // v = m 2^e, m in [0.5, 1), v positive and normal
#ifdef ZF_HOST_EMU
    unsigned long long b;
    memcpy(&b, &v, 8);
#else
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
#endif
    return (int)((b >> 52) & 0x7ffu) - 1022;
}

ZF_DEVICE double floor_d(double v) {
#ifdef ZF_HOST_EMU
    return __builtin_floor(v);
#else
    return floor(v);
#endif
}

// Levinson-Durbin, order choice and quantisation (zf-LPC v1 steps 3-5).  One thread; the library is compiled with
// -fmad=false, so every * + - / below is one IEEE double operation, exactly as written.
ZF_DEVICE void lpc_levinson(const long long *R, uint32_t M, uint32_t n, uint32_t bps, Model &m) {
    double lpcv[kMaxOrder], coef[kMaxOrder][kMaxOrder], error[kMaxOrder];
    m.valid = 0;
    m.order = 0; m.shift = 0; m.precision = 0;
    for (int j = 0; j < kMaxOrder; j++) m.q[j] = 0;
    if (R[0] == 0) return;
    double err = (double)R[0];
    uint32_t omax = 0;
    for (uint32_t i = 0; i < M; i++) {
        double r = -(double)R[i + 1];
        for (uint32_t j = 0; j < i; j++) r = r - lpcv[j] * (double)R[i - j];
        r = r / err;
        lpcv[i] = r;
        for (uint32_t j = 0; j < (i >> 1); j++) {
            const double tmp = lpcv[j];
            lpcv[j] = lpcv[j] + r * lpcv[i - 1 - j];
            lpcv[i - 1 - j] = lpcv[i - 1 - j] + r * tmp;
        }
        if (i & 1u) lpcv[i >> 1] = lpcv[i >> 1] + lpcv[i >> 1] * r;
        err = err * (1.0 - r * r);
        if (!(err > 0.0)) break;
        for (uint32_t j = 0; j <= i; j++) coef[i][j] = -lpcv[j];
        error[i] = err;
        omax = i + 1;
    }
    if (omax == 0) return;
    const uint32_t P = bps <= 17u ? 14u : 15u;
    const double delta = ((double)(bps + P) * 1.3862943611198906) / (double)n;
    uint32_t o = omax;
    double thr = error[omax - 1];
    for (uint32_t k = omax - 1; k >= 1; k--) {
        thr = thr * (1.0 + delta);
        if (error[k - 1] <= thr) o = k;
    }
    double cmax = 0.0;
    for (uint32_t j = 0; j < o; j++) {
        const double cj = coef[o - 1][j];
        const double a = cj < 0.0 ? -cj : cj;
        if (a > cmax) cmax = a;
    }
    if (!(cmax > 0.0) || !(cmax < 1e300)) return;
    int shift = (int)P - exponent_of(cmax);
    if (shift > 15) shift = 15;
    if (shift < 0) return;
    const double qmax = (double)((1 << (P - 1)) - 1), qmin = -(double)(1 << (P - 1));
    const double scale = (double)(1 << shift);
    double e = 0.0;
    for (uint32_t j = 0; j < o; j++) {
        e = e + coef[o - 1][j] * scale;
        double v = floor_d(e + 0.5);
        if (v > qmax) v = qmax;
        if (v < qmin) v = qmin;
        m.q[j] = (int32_t)v;
        e = e - v;
    }
    m.order = o;
    m.shift = (uint32_t)shift;
    m.precision = P;
    m.valid = 1;
}

// this thread's eight samples of candidate `slot` after the wasted-bits shift
ZF_DEVICE void shifted8(uint32_t slot, const int32_t (&L)[kX], const int32_t (&R)[kX], uint32_t waste, int32_t (&xs)[kSpt]) {
    int32_t x[kX];
    make_x<false>(slot, L, R, x);
#pragma unroll
    for (int j = 0; j < kSpt; j++) xs[j] = x[kHalo + j] >> waste;
}

// LPC residuals of this thread's samples from the plane (which holds the shifted samples, zeros in front); returns
// whether all of them fit 32 bits.  Entries below the order are unused.
ZF_DEVICE bool lpc_residual8(const int32_t *plane, const Model &m, uint32_t base, uint32_t n, int32_t (&r)[kSpt]) {
    bool ok = true;
    int32_t q[kMaxOrder];
#pragma unroll
    for (int k = 0; k < kMaxOrder; k++) q[k] = m.q[k];  // zero beyond the order
    const int32_t *p = plane + kPlanePad + base;
    int32_t h[kMaxOrder + kSpt];  // plane[base - 12 .. base + 7]
#pragma unroll
    for (int k = 0; k < kMaxOrder + kSpt; k++) h[k] = p[k - kMaxOrder];
#pragma unroll
    for (int j = 0; j < kSpt; j++) {
        long long sum = 0;
#pragma unroll
        for (int k = 0; k < kMaxOrder; k++) sum += (long long)q[k] * (long long)h[kMaxOrder + j - 1 - k];
        const long long v = (long long)h[kMaxOrder + j] - (sum >> m.shift);
        r[j] = (int32_t)v;
        const uint32_t i = base + j;
        if (i < n && i >= m.order && (v > 0x7fffffffll || v < -0x7fffffffll)) ok = false;
    }
    return ok;
}

// SUBFRAME_LPC: header 1xxxxx with order - 1 (+ wasted-bits flag and unary count), warm-ups, 4-bit precision - 1, 5-bit
// shift, coefficients, then the residual exactly as in a FIXED subframe (frame_writer.zig:328-361).  MODE 0 counts.
template <int MODE>
ZF_DEVICE uint32_t emit_lpc(const int32_t (&r)[kSpt], const long long *warm, const Model &m, int t, uint32_t base, uint32_t n,
                            const SlotDec &d, const uint8_t *choice_row, uint32_t *bits, uint32_t pos) {
    BitWriter bw;
    const uint32_t start = pos;
    if (MODE == 1) bw.init(bits, pos);
    const uint32_t order = d.order, waste = d.waste;
    const uint32_t psz = n >> d.po;
    const uint32_t param_len = 4 + d.method;
    const uint32_t esc_code = d.method ? 31u : 15u;
    if (t == 0) {
        if (MODE == 1) {
            bw.put(pos, ((0x20u | (order - 1u)) << 1) | (waste ? 1u : 0u), 8);
            if (waste) bw.put(pos + 8 + waste - 1, 1, 1);
        }
        pos += 8 + waste;
        const unsigned long long mask = kU64Max >> (64 - d.bps);
        for (uint32_t k = 0; k < order; k++) {
            if (MODE == 1) bw.put64(pos, (unsigned long long)warm[k] & mask, d.bps);
            pos += d.bps;
        }
        if (MODE == 1) {
            bw.put(pos, m.precision - 1u, 4);
            bw.put(pos + 4, m.shift, 5);
        }
        pos += 9;
        for (uint32_t k = 0; k < order; k++) {
            if (MODE == 1) bw.put(pos, (uint32_t)m.q[k] & (0xffffffffu >> (32 - m.precision)), m.precision);
            pos += m.precision;
        }
        if (MODE == 1) bw.put(pos, (d.method << 4) | d.po, 6);
        pos += 6;
    }
    uint32_t part = base < n ? base / psz : 0u, next = (part + 1u) * psz;
#pragma unroll
    for (int j = 0; j < kSpt; j++) {
        const uint32_t i = base + j;
        if (i < n) {
            if (i >= next) { part++; next += psz; }
            const uint32_t choice = choice_row[part];
            const bool esc = (choice & 0x80u) != 0;
            if (i == next - psz) {
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
                if (esc) {
                    const uint32_t wd = choice & 0x7fu;
                    if (wd) {
                        if (MODE == 1) bw.put(pos, (uint32_t)r[j] & (0xffffffffu >> (32 - wd)), wd);
                        pos += wd;
                    }
                } else {
                    const uint32_t zz = zigzag(r[j]);
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

template <int BYTES>
__global__ void __launch_bounds__(kThreads, 2) zf_encode_stereo_lpc_kernel(const FrameJob job) {
    static_assert(BYTES == 2 || BYTES == 3, "LPC kernel: 16- and 24-bit containers");
    extern __shared__ __align__(16) unsigned char zf_smem[];
    SmemLpc<BYTES> &sm = *reinterpret_cast<SmemLpc<BYTES> *>(zf_smem);
    SmemCommon &c = sm.c;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t n = job.block_size;
    const uint32_t depth = job.bit_depth ? job.bit_depth : 8u * BYTES;
    const uint32_t frame_bytes = n * 2u * BYTES;
    const uint32_t base = (uint32_t)t * kSpt;
    const uint32_t M = job.lpc_order < (uint32_t)kMaxOrder ? job.lpc_order : (uint32_t)kMaxOrder;
    const bool lpc_on = n >= kMinBlock && M > 0;

    if (job.pdl_trigger) pdl_launch_dependents();
    init_tables(c, t);
    if (t < kRawPadWords) sm.raw[t] = 0;
    if (t < kPlanePad) sm.plane[t] = 0;
    for (uint32_t i = t; i < (uint32_t)kMaxBlock; i += kThreads) sm.win[i] = (lpc_on && i < n) ? job.lpc_window[i] : (uint16_t)0;
    if (t == 0) c.cur_frame = atomicAdd(job.ticket, 1u);
    __syncthreads();

    for (;;) {
        const uint32_t f = c.cur_frame;
        if (f >= job.n_frames) break;
        const uint32_t fidx = job.frame_base + f;
        const unsigned long long frame_number = job.first_frame_number + fidx;

        load_raw_generic<BYTES, false>(sm.raw, job.pcm + (size_t)f * job.frame_stride, frame_bytes, t);
        __syncthreads();
        int32_t L[kX], R[kX];
        unpack_stereo<BYTES>(sm.raw, t, L, R);
        {
            uint4 *bz = reinterpret_cast<uint4 *>(sm.bits);
            const uint4 z = {0, 0, 0, 0};
            for (int k = t; k < BitBufWords<BYTES>::value / 4; k += kThreads) bz[k] = z;
        }
        if (t == 0) c.next_frame = atomicAdd(job.ticket, 1u);

        // ---- the reference's analysis: CONSTANT / VERBATIM / FIXED per candidate (as zf_encode_stereo_kernel) ----
        int32_t x[kX];
#pragma unroll 1
        for (uint32_t s = 0; s < 4; s++) {
            make_x<false>(s, L, R, x);
            P1<false> p;
            pass1<false, false>(x, base, n, p);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const unsigned long long ws = warp_sum(p.s[k]);
                if (lane == 0) { c.red[warp][s][k] = ws; c.red[warp][s][5 + k] = 0; }
            }
            const unsigned long long wo = warp_or(p.orv);
            if (lane == 0) c.red[warp][s][10] = wo;
        }
        __syncthreads();
        fold_red(c, t, 4);
        __syncthreads();
        if (t < 4) decide_slot<false>(c, (uint32_t)t, depth + (t == 3 ? 1u : 0u), n, job);
        __syncthreads();
        for (uint32_t s = 0; s < 4; s++) rice_zero_leaves(c, s, t, n);
        __syncthreads();
#pragma unroll 1
        for (uint32_t s = 0; s < 4; s++) {
            if (c.dec[s].kind != kFixed) continue;
            make_x<false>(s, L, R, x);
            rice_leaves<false, false>(c, s, x, t, base, n);
        }
        __syncthreads();
        rice_tree_and_search(c, t, n, 4);
        if (t == 0)
            for (uint32_t s = 0; s < 4; s++) finish_slot(c, s, n);
        __syncthreads();
        // keep what the reference's path decided
        if (t < 4) { sm.fx_dec[t] = c.dec[t]; sm.bad[t] = 0; sm.keep_fixed[t] = 1; }
        for (uint32_t s = 0; s < 4; s++) sm.fx_choice[s][t] = c.pchoice[s][t];
        __syncthreads();

        if (lpc_on) {
            // ---- windowed autocorrelation of the four candidates (exact integers) ----
#pragma unroll 1
            for (uint32_t s = 0; s < 4; s++) {
                const SlotDec fd = sm.fx_dec[s];
                if (fd.kind == kConstant) continue;  // block-uniform
                int32_t xs[kSpt];
                shifted8(s, L, R, fd.waste, xs);
                const uint32_t sh = fd.bps > 24u ? fd.bps - 24u : 0u;
#pragma unroll
                for (int j = 0; j < kSpt; j++) {
                    const uint32_t i = base + j;
                    const long long w = i < n ? (long long)sm.win[i] : 0ll;
                    sm.plane[kPlanePad + i] = (int32_t)(((long long)(xs[j] >> sh) * w) >> 14);
                }
                __syncthreads();
                {
                    const int32_t *p = sm.plane + kPlanePad + base;
                    int32_t h[kMaxOrder + kSpt];
#pragma unroll
                    for (int k = 0; k < kMaxOrder + kSpt; k++) h[k] = p[k - kMaxOrder];
                    long long acc[kMaxOrder + 1];
#pragma unroll
                    for (int l = 0; l <= kMaxOrder; l++) {
                        long long a = 0;
#pragma unroll
                        for (int j = 0; j < kSpt; j++) a += (long long)h[kMaxOrder + j] * (long long)h[kMaxOrder + j - l];
                        acc[l] = a;
                    }
#pragma unroll
                    for (int l = 0; l <= kMaxOrder; l++) {
                        const unsigned long long ws = warp_sum((unsigned long long)acc[l]);  // exact modulo 2^64
                        if (lane == 0) sm.lred[warp][l] = ws;
                    }
                }
                __syncthreads();
                if (t <= kMaxOrder) {
                    unsigned long long v = 0;
#pragma unroll
                    for (int w = 0; w < kWarps; w++) v += sm.lred[w][t];
                    sm.R[s][t] = (long long)v;
                }
                __syncthreads();
            }
            // ---- Levinson-Durbin, order, quantisation: one thread per candidate ----
            if (t < 4) {
                const SlotDec fd = sm.fx_dec[t];
                Model m;
                m.valid = 0;
                if (fd.kind != kConstant) lpc_levinson(sm.R[t], M, n, fd.bps, m);
                else { m.order = 0; m.shift = 0; m.precision = 0; for (int j = 0; j < kMaxOrder; j++) m.q[j] = 0; }
                sm.model[t] = m;
                SlotDec d = fd;
                if (m.valid) {  // the Rice search of the LPC residual: rice.calcParams (rice.zig:97-103) with this order
                    d.kind = kFixed;
                    d.order = m.order;
                    uint32_t mpo = job.max_rice_order;
                    const uint32_t tz = ctz32(n);
                    if (tz < mpo) mpo = tz;
                    const uint32_t lim_o = floor_log2(n) - floor_log2(m.order);
                    if (lim_o < mpo) mpo = lim_o;
                    while ((n >> mpo) < m.order) mpo--;
                    d.mpo = mpo;
                    d.po = 0; d.method = 0;
                } else {
                    d.kind = kVerbatim;  // not searched
                }
                c.dec[t] = d;
            }
            __syncthreads();
            for (uint32_t s = 0; s < 4; s++) rice_zero_leaves(c, s, t, n);  // leaves that accumulate
            __syncthreads();
#pragma unroll 1
            for (uint32_t s = 0; s < 4; s++) {
                if (!sm.model[s].valid) continue;  // block-uniform
                int32_t xs[kSpt];
                shifted8(s, L, R, sm.fx_dec[s].waste, xs);
#pragma unroll
                for (int j = 0; j < kSpt; j++) sm.plane[kPlanePad + base + j] = (base + j < n) ? xs[j] : 0;
                __syncthreads();
                int32_t r[kSpt];
                if (!lpc_residual8(sm.plane, sm.model[s], base, n, r)) atomicOr(&sm.bad[s], 1u);
                rice_leaves_from<false>(c, s, r, t, base, n);
                __syncthreads();
            }
            rice_tree_and_search(c, t, n, 4);
            if (t == 0)
                for (uint32_t s = 0; s < 4; s++) finish_slot(c, s, n);
            __syncthreads();
        }
        // ---- LPC or the reference's choice per candidate, then the stereo mode (encoder.zig:441-452 on these costs) ----
        if (t == 0) {
            unsigned long long est[4];
            for (uint32_t s = 0; s < 4; s++) {
                const SlotDec fd = sm.fx_dec[s];
                // in LPC mode a predictor's warm-up samples count (the reference's estimate leaves them out)
                unsigned long long cost = fd.est_bits + (fd.kind == kFixed ? (unsigned long long)fd.order * fd.bps : 0ull);
                bool use_lpc = false;
                if (lpc_on && sm.model[s].valid && !sm.bad[s] && c.dec[s].kind == kFixed) {
                    const Model &m = sm.model[s];
                    const unsigned long long lc = c.dec[s].est_bits + (unsigned long long)m.order * (fd.bps + m.precision) + 9ull;
                    if (lc < cost) { cost = lc; use_lpc = true; }
                }
                if (use_lpc) {
                    c.dec[s].kind = kLpc;
                    sm.keep_fixed[s] = 0;
                } else {
                    c.dec[s] = fd;
                    sm.keep_fixed[s] = 1;
                }
                c.dec[s].est_bits = cost;
                est[s] = cost;
            }
            const unsigned long long sum[4] = {est[0] + est[1], est[0] + est[3], est[3] + est[1], est[2] + est[3]};
            uint32_t best = 0;
            for (uint32_t k = 1; k < 4; k++)
                if (sum[k] < sum[best]) best = k;
            uint32_t a = 0, b = 1, cht = 1;
            if (best == 1) { a = 0; b = 3; cht = 8; }
            else if (best == 2) { a = 3; b = 1; cht = 9; }
            else if (best == 3) { a = 2; b = 3; cht = 10; }
            c.sub_slot[0] = a;
            c.sub_slot[1] = b;
            c.ch_type = cht;
        }
        __syncthreads();
        for (uint32_t s = 0; s < 4; s++)
            if (sm.keep_fixed[s]) c.pchoice[s][t] = sm.fx_choice[s][t];
        __syncthreads();

        // ---- pack ----
        const uint32_t hdr_bits = 8u * header_len(frame_number, n, job.sample_rate);
        int32_t ra[kSpt], rb[kSpt];
#pragma unroll
        for (int j = 0; j < kSpt; j++) { ra[j] = 0; rb[j] = 0; }
#pragma unroll 1
        for (uint32_t k = 0; k < 2; k++) {  // LPC residuals of the subframes that will be written, warm-ups to shared memory
            const uint32_t sk = c.sub_slot[k];
            if (c.dec[sk].kind != kLpc) continue;  // block-uniform
            int32_t xs[kSpt];
            shifted8(sk, L, R, c.dec[sk].waste, xs);
#pragma unroll
            for (int j = 0; j < kSpt; j++) {
                sm.plane[kPlanePad + base + j] = (base + j < n) ? xs[j] : 0;
                if (base + j < (uint32_t)kMaxOrder) sm.warm[k][base + j] = xs[j];
            }
            __syncthreads();
            if (k == 0) lpc_residual8(sm.plane, sm.model[sk], base, n, ra);
            else lpc_residual8(sm.plane, sm.model[sk], base, n, rb);
            __syncthreads();
        }
        uint32_t len_a = 0, len_b = 0;
#pragma unroll 1
        for (uint32_t k = 0; k < 2; k++) {
            const uint32_t sk = c.sub_slot[k];
            const SlotDec dk = c.dec[sk];
            uint32_t v;
            if (dk.kind == kLpc) {
                v = emit_lpc<0>(k ? rb : ra, sm.warm[k], sm.model[sk], t, base, n, dk, &c.pchoice[sk][(1u << dk.po) - 1u], sm.bits, 0);
            } else {
                make_x<false>(sk, L, R, x);
                v = emit_subframe<false, false, 0>(x, t, base, n, dk, &c.pchoice[sk][(1u << dk.po) - 1u], sm.bits, 0);
            }
            if (k == 0) len_a = v;
            else len_b = v;
        }
        uint32_t ex_a, ex_b, tot_a, tot_b;
        block_scan2(c, t, len_a, len_b, ex_a, ex_b, tot_a, tot_b);
        const uint32_t total_bits = hdr_bits + tot_a + tot_b;
        const uint32_t fbytes = (total_bits + 7u) >> 3;
        const bool fits = (fbytes + 2u) <= (uint32_t)BitBufWords<BYTES>::value * 4u - 8u;
        if (t == 0) {
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
                const uint32_t pos = hdr_bits + (k == 0 ? ex_a : tot_a + ex_b);
                if (dk.kind == kLpc) {
                    emit_lpc<1>(k ? rb : ra, sm.warm[k], sm.model[sk], t, base, n, dk, &c.pchoice[sk][(1u << dk.po) - 1u], sm.bits, pos);
                } else {
                    make_x<false>(sk, L, R, x);
                    emit_subframe<false, false, 1>(x, t, base, n, dk, &c.pchoice[sk][(1u << dk.po) - 1u], sm.bits, pos);
                }
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

}  // namespace lpc
}  // namespace zf

// zf_kernel_full.cuh -- fast path for the common case: stereo, decorrelation on, block size 4096
// (= 512 threads x 8 samples, every Rice leaf partition of 16 samples is one pair of threads), all frames
// full, max_rice_order = 8.  Same algorithm and same decisions as zf_kernel.cuh, arranged for fewer
// instructions and barriers:
//   * residuals by differencing the thread's 20-sample window in place `order` times (no per-sample switch)
//   * Rice leaves stay in registers; levels 7..3 of the partition tree are built with warp shuffles
//   * the two chosen channels' residuals are kept in registers across the block scan, so the bit-writing
//     pass only zigzags and places codewords
//   * a 64-bit accumulator bit writer: one shared-memory atomicOr per 32 output bits
#pragma once
#include "zf_kernel.cuh"

namespace zf {

template <bool WIDE, int SLOT>
ZF_DEVICE void make_x_c(const int32_t (&L)[kX], const int32_t (&R)[kX], typename Ar<WIDE>::T (&x)[kX]) {
    typedef typename Ar<WIDE>::T T;
#pragma unroll
    for (int i = 0; i < kX; i++) {
        if (SLOT == 0) x[i] = L[i];
        else if (SLOT == 1) x[i] = R[i];
        else if (SLOT == 2) x[i] = ((T)L[i] + (T)R[i]) >> 1;
        else x[i] = (T)L[i] - (T)R[i];
    }
}

template <bool WIDE>
ZF_DEVICE void make_x_rt(uint32_t slot, const int32_t (&L)[kX], const int32_t (&R)[kX], typename Ar<WIDE>::T (&x)[kX]) {
    if (slot == 0) make_x_c<WIDE, 0>(L, R, x);
    else if (slot == 1) make_x_c<WIDE, 1>(L, R, x);
    else if (slot == 2) make_x_c<WIDE, 2>(L, R, x);
    else make_x_c<WIDE, 3>(L, R, x);
}

// After `order` passes x[i] holds the order-th finite difference for every i >= order (fixed.zig:12-18).
template <typename T>
ZF_DEVICE void diff_in_place(T (&x)[kX], uint32_t order) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if ((uint32_t)k < order) {
#pragma unroll
            for (int i = kX - 1; i > k; i--) x[i] -= x[i - 1];
        }
    }
}

// MSB-first bit writer over a 64-bit accumulator: bits of word `w` sit in the high half.
struct BitOut {
    uint32_t *buf;
    uint32_t w, nb;
    unsigned long long acc;

    ZF_DEVICE void init(uint32_t *b, uint32_t bitpos) {
        buf = b;
        w = bitpos >> 5;
        nb = bitpos & 31u;
        acc = 0;
    }
    // q zero bits (the buffer is pre-zeroed: just advance), then `len` (1..32) bits of val (< 2^len).
    // Common case q + len <= 32: one (q + len)-bit field, a single flush test, no divergent control flow.
    ZF_DEVICE void put(uint32_t q, uint32_t val, uint32_t len) {
        uint32_t fl = q + len;
        if (fl > 32u) {  // long zero run (rare): leave the current word, skip whole words
            nb += q;
            const uint32_t hi = (uint32_t)(acc >> 32);
            if (nb >= 32u) {
                if (hi) atomicOr(&buf[w], hi);
                acc = 0;
                w += nb >> 5;
                nb &= 31u;
            }
            fl = len;
        }
        acc |= (unsigned long long)val << (64u - nb - fl);
        nb += fl;
        const bool full = nb >= 32u;
        if (full) atomicOr(&buf[w], (uint32_t)(acc >> 32));
        acc = full ? (acc << 32) : acc;
        w += full ? 1u : 0u;
        nb &= 31u;
    }
    ZF_DEVICE void put64(unsigned long long val, uint32_t len) {  // len 1..64
        if (len > 32u) {
            put(0, (uint32_t)(val >> 32), len - 32u);
            put(0, (uint32_t)val, 32u);
        } else {
            put(0, (uint32_t)val, len);
        }
    }
    ZF_DEVICE void finish() {
        const uint32_t hi = (uint32_t)(acc >> 32);
        if (hi) atomicOr(&buf[w], hi);
    }
};

// header bits of a FIXED subframe written by thread 0: type byte, wasted-bits unary, warm-ups, Rice header
ZF_DEVICE uint32_t fixed_head_bits(const SlotDec &d) { return 8u + d.waste + d.order * d.bps + 6u; }

// Bits this thread contributes to a FIXED subframe whose residuals r[0..15] it holds (frame_writer.zig:303-372).
ZF_DEVICE uint32_t count_fixed_full(const int32_t (&r)[kSpt], int t, const SlotDec &d, uint32_t choice, bool at_start) {
    const uint32_t jstart = (t == 0) ? d.order : 0u;
    const uint32_t cnt = (uint32_t)kSpt - jstart;
    uint32_t bits = (t == 0) ? fixed_head_bits(d) : 0u;
    const bool esc = (choice & 0x80u) != 0;
    if (at_start) bits += 4u + d.method + (esc ? 5u : 0u);
    if (esc) return bits + (choice & 0x7fu) * cnt;
    uint32_t qs = 0;
#pragma unroll
    for (int j = 0; j < kSpt; j++) {
        const uint32_t q = zigzag(r[j]) >> choice;
        qs += ((uint32_t)j >= jstart) ? q : 0u;
    }
    return bits + qs + cnt * (choice + 1u);
}

ZF_DEVICE void write_fixed_full(const int32_t (&r)[kSpt], const long long (&warm)[4], int t, const SlotDec &d,
                                uint32_t choice, bool at_start, uint32_t *bits, uint32_t pos) {
    BitOut bo;
    bo.init(bits, pos);
    const uint32_t param_len = 4u + d.method;
    if (t == 0) {
        bo.put(0, ((8u | d.order) << 1) | (d.waste ? 1u : 0u), 8);           // :316-321
        if (d.waste) bo.put(d.waste - 1u, 1u, 1);
        const unsigned long long mask = kU64Max >> (64u - d.bps);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++)
            if (k < d.order) bo.put64((unsigned long long)warm[k] & mask, d.bps);  // :323-325
        bo.put(0, (d.method << 4) | d.po, 6);                                  // :328
    }
    const bool esc = (choice & 0x80u) != 0;
    if (at_start) {                                                             // :341-357
        if (esc) {
            bo.put(0, d.method ? 31u : 15u, param_len);
            bo.put(0, choice & 0x7fu, 5);
        } else {
            bo.put(0, choice, param_len);
        }
    }
    const uint32_t jstart = (t == 0) ? d.order : 0u;
    if (esc) {
        const uint32_t wd = choice & 0x7fu;
        if (wd) {
            const uint32_t m = 0xffffffffu >> (32u - wd);
#pragma unroll
            for (int j = 0; j < kSpt; j++)
                if ((uint32_t)j >= jstart) bo.put(0, (uint32_t)r[j] & m, wd);
        }
    } else {
        const uint32_t one = 1u << choice, m = one - 1u, len = choice + 1u;
#pragma unroll
        for (int j = 0; j < kSpt; j++) {
            if ((uint32_t)j >= jstart) {
                const uint32_t zz = zigzag(r[j]);
                bo.put(zz >> choice, one | (zz & m), len);                       // :363-372
            }
        }
    }
    bo.finish();
}

template <int BYTES>
__global__ void __launch_bounds__(kThreads, BYTES == 4 ? 1 : 2) zf_encode_stereo_full_kernel(const FrameJob job) {
    constexpr bool WIDE = (BYTES == 4);
    typedef typename Ar<WIDE>::T T;
    extern __shared__ __align__(16) unsigned char zf_smem[];
    SmemStereo<BYTES> &sm = *reinterpret_cast<SmemStereo<BYTES> *>(zf_smem);
    SmemCommon &c = sm.c;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr uint32_t n = (uint32_t)kMaxBlock;
    constexpr uint32_t depth = 8u * BYTES;
    constexpr uint32_t frame_bytes = n * 2u * BYTES;
    const uint32_t base = (uint32_t)t * kSpt;
    const bool tma = job.use_tma != 0;

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
            load_raw_generic<BYTES, true>(sm.raw, job.pcm + (size_t)f * job.frame_stride, frame_bytes, t);
            __syncthreads();
        }
        int32_t L[kX], R[kX];
        unpack_stereo<BYTES>(sm.raw, t, L, R);
        {
            uint4 *bz = reinterpret_cast<uint4 *>(sm.bits);
            const uint4 z = {0, 0, 0, 0};
            for (int k = t; k < BitBufWords<BYTES>::value / 4; k += kThreads) bz[k] = z;
        }
        __syncthreads();  // raw consumed: prefetch the next frame
        if (t == 0) {
            const uint32_t nf = atomicAdd(job.ticket, 1u);
            c.next_frame = nf;
            if (tma && nf < job.n_frames) {
                fence_proxy_async();
                mbar_expect_tx(&c.mbar, frame_bytes);
                tma_load_1d(sm.raw + kRawPadWords, job.pcm + (size_t)nf * job.frame_stride, frame_bytes, &c.mbar);
            }
        }

        // ---- pass 1 ----
#define ZF_PASS1(SLOT)                                                          \
    {                                                                           \
        T x[kX];                                                                \
        make_x_c<WIDE, SLOT>(L, R, x);                                          \
        P1<WIDE> p;                                                             \
        pass1<WIDE, true>(x, base, n, p);                                       \
        _Pragma("unroll") for (int k = 0; k < 5; k++) {                         \
            const unsigned long long ws = warp_sum(p.s[k]);                     \
            const unsigned long long wr = WIDE ? warp_or(p.rng[k]) : 0ull;      \
            if (lane == 0) { c.red[warp][SLOT][k] = ws; c.red[warp][SLOT][5 + k] = wr; } \
        }                                                                       \
        const unsigned long long wo = warp_or(p.orv);                           \
        if (lane == 0) c.red[warp][SLOT][10] = wo;                              \
    }
        ZF_PASS1(0) ZF_PASS1(1) ZF_PASS1(2) ZF_PASS1(3)
#undef ZF_PASS1
        __syncthreads();
        fold_red(c, t, 4);
        __syncthreads();
        if (t < 4) decide_slot<WIDE>(c, (uint32_t)t, depth + (t == 3 ? 1u : 0u), n, job);
        __syncthreads();

        // ---- pass 2: leaves in registers, tree levels 7..3 by shuffles ----
#define ZF_LEAVES(SLOT)                                                                          \
    if (c.dec[SLOT].kind == kFixed) {                                                            \
        const uint32_t order = c.dec[SLOT].order, waste = c.dec[SLOT].waste;                     \
        T x[kX];                                                                                 \
        make_x_c<WIDE, SLOT>(L, R, x);                                                           \
        diff_in_place<T>(x, order);                                                              \
        const uint32_t jstart = (t == 0) ? order : 0u;                                           \
        unsigned long long S = 0;                                                                \
        int32_t mn = 0, mx = 0;                                                                  \
        _Pragma("unroll") for (int j = 0; j < kSpt; j++) {                                       \
            int32_t r = (int32_t)(x[kHalo + j] >> waste);                                        \
            r = ((uint32_t)j >= jstart) ? r : 0;                                                 \
            S += uabs(r);                                                                        \
            mn = r < mn ? r : mn;                                                                \
            mx = r > mx ? r : mx;                                                                \
        }                                                                                        \
        const uint32_t zm = zigzag(mn), zx = zigzag(mx);                                         \
        uint32_t B = bitlen32(zm > zx ? zm : zx);                                                \
        /* leaf (level 8) = 16 samples = this thread and its xor-1 neighbour; levels 7..4 inside the warp */ \
        _Pragma("unroll") for (int lv = 8; lv >= 4; lv--) {                                      \
            const int stride = 1 << (8 - lv);                                                    \
            S += __shfl_xor_sync(0xffffffffu, S, stride);                                        \
            const uint32_t ob = __shfl_xor_sync(0xffffffffu, B, stride);                         \
            B = ob > B ? ob : B;                                                                 \
            if ((lane & (2 * stride - 1)) == 0) {                                                \
                const uint32_t node = (1u << lv) - 1u + ((uint32_t)t >> (9 - lv));               \
                c.psum[SLOT][node] = S;                                                          \
                c.pbits[SLOT][node] = B;                                                         \
            }                                                                                    \
        }                                                                                        \
    }
        ZF_LEAVES(0) ZF_LEAVES(1) ZF_LEAVES(2) ZF_LEAVES(3)
#undef ZF_LEAVES
        __syncthreads();
        // ---- parameter search: heap node m = t (1..511); warp 0 holds levels 0..4, warps 1, 2-3, 4-7, 8-15
        //      hold levels 5, 6, 7, 8 ----
#pragma unroll 1
        for (uint32_t s = 0; s < 4; s++) {
            const SlotDec &d = c.dec[s];
            if (d.kind != kFixed) continue;
            const uint32_t m = (uint32_t)t;
            const bool act = m >= 1;
            uint32_t choice = 0;
            unsigned long long cost = 0;
            if (act) {
                const uint32_t lvl = floor_log2(m);
                const uint32_t j = m - (1u << lvl);
                unsigned long long S;
                uint32_t B;
                if (lvl < 4) {  // levels 3..0 are summed straight from the sixteen level-4 nodes (heap 16..31)
                    const uint32_t span = 1u << (4 - lvl);
                    S = 0;
                    B = 0;
                    for (uint32_t k = 0; k < span; k++) {
                        S += c.psum[s][15 + j * span + k];
                        const uint32_t b = c.pbits[s][15 + j * span + k];
                        B = b > B ? b : B;
                    }
                } else {
                    S = c.psum[s][m - 1];
                    B = c.pbits[s][m - 1];
                }
                const uint32_t cnt = (n >> lvl) - (j == 0 ? d.order : 0u);  // rice.zig:356,371
                best_param(S, B, cnt, d.max_param, choice, cost);
                c.pchoice[s][m - 1] = (uint8_t)choice;
            }
            const bool five = act && choice < 0x80u && choice > 14u;  // isRice2, rice.zig:74-76
            if (warp == 0) {
                c.mixed[s][lane] = cost;
                const uint32_t fm = __ballot_sync(0xffffffffu, five);
                if (lane == 0) c.mixfive[s] = fm;
            } else {
                const unsigned long long wsum = warp_sum(cost);
                const uint32_t wf = reduce_or(five ? 1u : 0u);
                if (lane == 0) { c.wcost[s][0][warp] = wsum; c.wfive[s][0][warp] = wf; }
            }
        }
        __syncthreads();
        // ---- partition order per slot (rice.zig:262-276, '<=': highest order wins ties), FIXED vs VERBATIM (:538) ----
        if (warp < 4 && c.dec[warp].kind == kFixed) {
            const uint32_t s = (uint32_t)warp;
            unsigned long long bc = kU64Max;
            uint32_t method = 0;
            if (lane <= kMaxLevel) {
                unsigned long long cost = 0;
                uint32_t fv = 0;
                if (lane <= 4) {
                    for (uint32_t m = 1u << lane; m < (2u << lane); m++) cost += c.mixed[s][m];
                    fv = (c.mixfive[s] >> (1u << lane)) & ((1u << (1u << lane)) - 1u);
                } else {
                    const uint32_t w0 = 1u << (lane - 5), w1 = 2u << (lane - 5);
                    for (uint32_t w = w0; w < w1; w++) { cost += c.wcost[s][0][w]; fv |= c.wfive[s][0][w]; }
                }
                method = fv ? 1u : 0u;
                bc = cost + ((unsigned long long)(4u + method) << lane);  // :394
            }
            unsigned long long best = kU64Max;
            uint32_t bpo = 0, bmethod = 0;
            for (uint32_t lvl = 0; lvl <= (uint32_t)kMaxLevel; lvl++) {
                const unsigned long long v = __shfl_sync(0xffffffffu, bc, (int)lvl);
                const uint32_t mv = __shfl_sync(0xffffffffu, method, (int)lvl);
                if (v <= best) { best = v; bpo = lvl; bmethod = mv; }
            }
            if (lane == 0) {
                SlotDec &d = c.dec[s];
                const unsigned long long verb = (unsigned long long)n * d.bps;
                if (best < verb) { d.est_bits = best; d.po = bpo; d.method = bmethod; }
                else { d.kind = kVerbatim; d.est_bits = verb; }
            }
        }
        __syncthreads();
        uint32_t sa = 0, sb = 1, ch_type = 1;  // Channel.indep(2) = 1, type.zig:7-12
        {   // stereo mode: first minimum of [L+R, L+S, S+R, M+S], encoder.zig:441-452 (every thread, same result)
            const unsigned long long el = c.dec[0].est_bits, er = c.dec[1].est_bits, em = c.dec[2].est_bits,
                                     es = c.dec[3].est_bits;
            unsigned long long bestv = el + er;
            if (el + es < bestv) { bestv = el + es; sa = 0; sb = 3; ch_type = 8; }
            if (es + er < bestv) { bestv = es + er; sa = 3; sb = 1; ch_type = 9; }
            if (em + es < bestv) { bestv = em + es; sa = 2; sb = 3; ch_type = 10; }
        }

        // ---- pack: count, scan, write ----
        const uint32_t hdr_bits = 8u * header_len(frame_number, n, job.sample_rate);
        int32_t ra[kSpt], rb[kSpt];
        long long warm_a[4] = {0, 0, 0, 0}, warm_b[4] = {0, 0, 0, 0};
        uint32_t len_a, len_b, choice_a = 0, choice_b = 0;
        bool start_a = false, start_b = false;
        const SlotDec da = c.dec[sa], db = c.dec[sb];
#define ZF_COUNT(SK, DK, RK, WARM, CHOICE, START, LEN)                                                  \
    if (DK.kind == kFixed) {                                                                          \
        T x[kX];                                                                                      \
        make_x_rt<WIDE>(SK, L, R, x);                                                                 \
        if (t == 0) {                                                                                 \
            _Pragma("unroll") for (int k = 0; k < 4; k++) WARM[k] = (long long)x[kHalo + k] >> DK.waste; \
        }                                                                                             \
        diff_in_place<T>(x, DK.order);                                                                \
        _Pragma("unroll") for (int j = 0; j < kSpt; j++) RK[j] = (int32_t)(x[kHalo + j] >> DK.waste); \
        const uint32_t part = base >> (12u - DK.po);                                                  \
        CHOICE = c.pchoice[SK][(1u << DK.po) - 1u + part];                                            \
        START = (base & ((n >> DK.po) - 1u)) == 0;                                                    \
        LEN = count_fixed_full(RK, t, DK, CHOICE, START);                                             \
    } else {                                                                                          \
        T x[kX];                                                                                      \
        make_x_rt<WIDE>(SK, L, R, x);                                                                 \
        LEN = emit_plain<WIDE, true, 0>(x, t, base, n, DK, sm.bits, 0);         \
    }
        ZF_COUNT(sa, da, ra, warm_a, choice_a, start_a, len_a)
        ZF_COUNT(sb, db, rb, warm_b, choice_b, start_b, len_b)
#undef ZF_COUNT
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
        }
        // the frame header is written by the last thread while the others already place codewords
        if (t == kThreads - 1) write_header(c, sm.bits, frame_number, depth, ch_type, n, job.sample_rate);
        if (fits) {
#define ZF_WRITE(SK, DK, RK, WARM, CHOICE, START, POS)                                                \
    if (DK.kind == kFixed) {                                                                          \
        write_fixed_full(RK, WARM, t, DK, CHOICE, START, sm.bits, POS);                               \
    } else {                                                                                          \
        T x[kX];                                                                                      \
        make_x_rt<WIDE>(SK, L, R, x);                                                                 \
        emit_plain<WIDE, true, 1>(x, t, base, n, DK, sm.bits, POS);             \
    }
            ZF_WRITE(sa, da, ra, warm_a, choice_a, start_a, hdr_bits + ex_a)
            ZF_WRITE(sb, db, rb, warm_b, choice_b, start_b, hdr_bits + tot_a + ex_b)
#undef ZF_WRITE
        }
        __syncthreads();
        finish_frame(c, sm.bits, t, job, fidx, total_bits, fits);
        __syncthreads();
        if (t == 0) c.cur_frame = c.next_frame;
        __syncthreads();
    }
    if (t == 0) pdl_wait_primary();
}

}  // namespace zf

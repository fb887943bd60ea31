// zf_kernel_indep.cuh -- independent-channel frames: mono, 3..8 channels, or stereo with decorrelation
// switched off (reference: processChannels `else` arm, encoder.zig:456-475; subframes written in
// channel order, encoder.zig:258-261).  Same per-channel machinery as the stereo kernel, applied to
// one channel after another; samples are gathered straight from global memory (stride = channels).
#pragma once
#include "zf_kernel.cuh"

namespace zf {

struct SmemIndep {
    SmemCommon c;
    alignas(16) uint32_t bits[1];  // dynamic
};

// words of bit buffer for C channels (same bound as BitBufWords, per channel)
__host__ __device__ inline uint32_t indep_bit_words(int bytes, int channels) {
    const uint32_t per_ch = (kMaxBlock * (8u * bytes) + kMaxBlock / 2 + 256u * 10u + 200u) / 8u;
    return ((16u + (uint32_t)channels * per_ch + 2u + 64u) / 16u) * 4u;
}
__host__ inline size_t indep_smem_bytes(int bytes, int channels) {
    return sizeof(SmemCommon) + 16 + (size_t)indep_bit_words(bytes, channels) * 4u;
}

template <int BYTES>
ZF_DEVICE int32_t load_sample(const uint8_t *p) {
    if (BYTES == 2) return (int32_t)(int16_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
    if (BYTES == 3) return ((int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24))) >> 8;
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}

template <int BYTES>
__global__ void __launch_bounds__(kThreads, 2) zf_encode_indep_kernel(const FrameJob job) {
    constexpr bool WIDE = (BYTES == 4);
    typedef typename Ar<WIDE>::T T;
    extern __shared__ __align__(16) unsigned char zf_smem[];
    SmemIndep &sm = *reinterpret_cast<SmemIndep *>(zf_smem);
    SmemCommon &c = sm.c;
    uint32_t *bits = sm.bits;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (job.pdl_trigger) pdl_launch_dependents();
    const uint32_t n = job.block_size;
    const uint32_t nch = job.channels;
    const uint32_t depth = job.bit_depth ? job.bit_depth : 8u * BYTES;
    const uint32_t base = (uint32_t)t * kSpt;
    const uint32_t bit_words = indep_bit_words(BYTES, (int)nch);

    init_tables(c, t);
    if (t == 0) c.cur_frame = atomicAdd(job.ticket, 1u);
    __syncthreads();

    for (;;) {
        const uint32_t f = c.cur_frame;
        if (f >= job.n_frames) break;
        const uint32_t fidx = job.frame_base + f;
        const unsigned long long frame_number = job.first_frame_number + fidx;
        const uint8_t *src = job.pcm + (size_t)f * job.frame_stride;
        {
            uint4 *bz = reinterpret_cast<uint4 *>(bits);
            const uint4 z = {0, 0, 0, 0};
            for (uint32_t k = t; k < bit_words / 4; k += kThreads) bz[k] = z;
        }
        __syncthreads();
        if (t == 0) {
            c.next_frame = atomicAdd(job.ticket, 1u);
            write_header(c, bits, frame_number, depth, nch - 1u /* Channel.indep, type.zig:7-12 */, n, job.sample_rate);
        }
        uint32_t bitpos = 8u * header_len(frame_number, n, job.sample_rate);
        bool fits = true;

#pragma unroll 1
        for (uint32_t ch = 0; ch < nch; ch++) {
            T x[kX];
#pragma unroll
            for (int k = 0; k < kX; k++) {
                const int i = (int)base - kHalo + k;
                x[k] = (i >= 0 && (uint32_t)i < n) ? (T)load_sample<BYTES>(src + ((size_t)i * nch + ch) * BYTES) : (T)0;
            }
            P1<WIDE> p;
            pass1<WIDE, false>(x, base, n, p);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const unsigned long long ws = warp_sum(p.s[k]);
                const unsigned long long wr = WIDE ? warp_or(p.rng[k]) : 0ull;
                if (lane == 0) { c.red[warp][0][k] = ws; c.red[warp][0][5 + k] = wr; }
            }
            const unsigned long long wo = warp_or(p.orv);
            if (lane == 0) c.red[warp][0][10] = wo;
            __syncthreads();
            fold_red(c, t, 1);
            __syncthreads();
            if (t == 0) decide_slot<WIDE>(c, 0, depth, n, job);
            __syncthreads();
            rice_zero_leaves(c, 0, t, n);
            __syncthreads();
            if (c.dec[0].kind == kFixed) rice_leaves<WIDE, false>(c, 0, x, t, base, n);
            __syncthreads();
            rice_tree_and_search(c, t, n, 1);
            if (t == 0) finish_slot(c, 0, n);
            __syncthreads();
            const SlotDec d = c.dec[0];
            const uint8_t *row = &c.pchoice[0][(1u << d.po) - 1u];
            const uint32_t len = emit_subframe<WIDE, false, 0>(x, t, base, n, d, row, bits, 0);
            uint32_t ex, ex2, tot, tot2;
            block_scan2(c, t, len, 0u, ex, ex2, tot, tot2);
            fits = fits && ((bitpos + tot + 7u) / 8u + 2u) <= bit_words * 4u - 8u;
            if (fits) emit_subframe<WIDE, false, 1>(x, t, base, n, d, row, bits, bitpos + ex);
            bitpos += tot;
            __syncthreads();
        }
        if (t == 0) {
            const unsigned long long size = ((bitpos + 7u) >> 3) + 2u;
            job.frame_sizes[fidx] = (uint32_t)size;
            if (fidx == 0) st_relaxed_gpu(job.desc, kFlagPrefix | size);
            else st_relaxed_gpu(job.desc + fidx, kFlagAggregate | size);
            if (!fits) atomicOr(job.status, kStatusBitOverflow);
        }
        __syncthreads();
        finish_frame(c, bits, t, job, fidx, bitpos, fits);
        __syncthreads();
        if (t == 0) c.cur_frame = c.next_frame;
        __syncthreads();
    }
    if (t == 0) pdl_wait_primary();
}

}  // namespace zf

// zf_capi.cu -- C ABI of the encode engine: encoder handle, batch submit/collect, device-resident
// entry, the per-frame writeFrame mirror.  Declared in include/zigflac_b200.h.
//
// The reference's Encoder.writeFrame (encoder.zig:234) is synchronous and per frame; a GPU needs many
// frames in flight, so the boundary is a batch: K frames of raw PCM in, K frames + K sizes out.
// writeFrame is the K = 1 case (zf_write_frame).  No CPU fallback exists anywhere in this file.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <vector>

#include "../../include/zigflac_b200.h"
#include "zf_kernel.cuh"
#include "zf_kernel_indep.cuh"
#include "zf_kernel_v3.cuh"
#include "zf_kernel_lpc.cuh"

namespace {

thread_local char g_cuda_err[256] = "";

#define ZF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                                    \
            return ZF_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

constexpr int kPow8Len = 1 << 18;
constexpr int kSlots = 4;  // upload of batch b+1, kernels of batch b and downloads of batches b-1, b-2 overlap: a batch's
                           // buffers are held until its download has arrived, and the host must not wait for that before
                           // it queues the next upload
constexpr int kRing = 256;  // covers the largest frame of the multi-channel path (8 ch x 4096 x 33 bits)

struct Slot {  // one in-flight batch: device buffers + pinned staging
    uint8_t *d_pcm = nullptr;
    uint8_t *d_wide = nullptr;  // 8-bit input widened to 16-bit containers
    uint8_t *d_out = nullptr;
    uint8_t *h_pcm = nullptr;   // pinned
    uint8_t *h_out = nullptr;   // pinned
    uint32_t *d_sizes = nullptr;
    uint32_t *h_sizes = nullptr;  // pinned
    unsigned char *d_ctl_block = nullptr;   // one allocation, one memset per batch: [ctl 16 B][tail meta 32 B][pad][descriptors]
    unsigned long long *d_desc = nullptr;
    unsigned int *d_ctl = nullptr;          // [0],[1] tickets, [2] status
    unsigned long long *d_total = nullptr;
    unsigned long long *h_total = nullptr;  // pinned: [0] total, [1] status
    cudaStream_t stream = nullptr;
    uint16_t *d_win_tail = nullptr;          // LPC: window of the short last frame
    uint32_t win_tail_len = 0;
    uint8_t *d_tail = nullptr;               // private output of the last-frame launch
    unsigned long long *d_tail_meta = nullptr;  // [0] desc, [1] total, [2] lo32: frame size
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_done = nullptr;
    cudaEvent_t ev_up = nullptr, ev_down = nullptr;  // upload / download finished: keep the batches' copies in order
    cudaEvent_t kev[2 * kRing] = {};  // start/stop pairs around the full-frame kernel of recent batches
    uint32_t kev_count = 0;           // pairs recorded since the last zf_kernel_times()
    size_t pcm_cap = 0, out_cap = 0;
    uint32_t frames = 0;  // frames of the batch in flight
    bool busy = false;      // kernels submitted, results not fetched yet
    bool draining = false;  // output copy in flight
    uint8_t *copy_dst = nullptr;  // pageable destination of the staged output copy
    uint32_t *sizes_dst = nullptr;  // caller's frame_sizes (filled from the pinned copy when the download has arrived)
    size_t copy_len = 0;
    bool have_io = false;
};

}  // namespace

struct zf_encoder {
    zf_config cfg;
    int sm_count = 0;
    int launches_last = 0;
    float kernel_ms_last = 0.f;
    uint16_t *d_pow8 = nullptr;
    Slot slot[kSlots];  // submit/collect use slot 0 only; zf_encode_pcm rotates through all of them
    cudaStream_t s_up = nullptr, s_down = nullptr;  // all uploads / all downloads, each in batch order on its own stream
    cudaEvent_t ev_dev = nullptr;  // end of the last zf_encode_device batch: it shares slot 0's control block and descriptors
    bool dev_pending = false;
    size_t frame_pcm_bytes = 0;   // one frame of the caller's PCM
    size_t frame_dev_bytes = 0;   // ... and as the kernels read it (differs for 8-bit samples)
    size_t max_frame_bytes = 0;
    bool stereo = false;
    int occ_gen = 0, occ_v3 = 0, occ_lpc = 0, occ_exact = 0;
    size_t smem_lpc = 0, smem_exact = 0;
    uint16_t *d_win = nullptr;  // LPC: window of a full block
    size_t smem_stereo = 0;
    size_t smem_v3 = 0;
    size_t smem_indep = 0;
    // ZF_TRACE=1 (development aid): per-batch device timeline of zf_encode_pcm, printed to stderr
    bool trace = false;
    bool no_taper = false;
    std::vector<cudaEvent_t> tr_ev;  // [batch][kTracePoints]
    std::vector<double> tr_host;     // [batch][3]: submit returned, fetch returned, finish returned (ms)
    std::chrono::steady_clock::time_point tr_h0;
    uint64_t tr_batch = 0;
};

namespace {

enum { kTrUp0 = 0, kTrUp1, kTrKern1, kTrSmall1, kTrDown0, kTrDown1, kTracePoints };
void tr_mark(zf_encoder *e, cudaStream_t s, int point) {
    if (!e->trace) return;
    const size_t i = e->tr_batch * kTracePoints + point;
    if (i < e->tr_ev.size()) cudaEventRecord(e->tr_ev[i], s);
}
void tr_host(zf_encoder *e, uint64_t batch, int point) {
    if (!e->trace || batch * 3 + point >= e->tr_host.size()) return;
    e->tr_host[batch * 3 + point] =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - e->tr_h0).count();
}

bool depth_ok(unsigned d) { return d == 8 || d == 16 || d == 24 || d == 32; }
// bytes per sample inside the kernels: 8-bit samples are widened to 16-bit containers on the device first
int container_bytes(const zf_config &cfg) { return cfg.bit_depth == 8 ? 2 : cfg.bit_depth / 8; }

size_t max_frame_bytes_of(const zf_config *cfg) {
    // encoder.zig:583-595 with the reference's own call-site quirk (:59 passes compute_waste_bits = true)
    const size_t header_max = 2 + 7 + 2 + 2 + 1, subframe_header_max = 8, footer = 2;
    const size_t bps = (cfg->channels == 2) ? (size_t)cfg->bit_depth + 1 : cfg->bit_depth;
    const size_t byte_per_sample = (bps + 7) / 8;
    return header_max + subframe_header_max * cfg->channels + (size_t)cfg->block_size * byte_per_sample * (cfg->channels + 1u) +
           footer;
}

template <int BYTES>
int setup_stereo_kernel(zf_encoder *e, int *occ) {
    void (*k)(const zf::FrameJob) = zf::zf_encode_stereo_kernel<BYTES, false>;
    const size_t smem = sizeof(zf::SmemStereo<BYTES>);
    ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, zf::kThreads, smem));
    e->smem_stereo = smem;
    return ZF_OK;
}

template <int BYTES>
int setup_exact_kernel(zf_encoder *e, int *occ) {
    void (*k)(const zf::FrameJob) = zf::zf_encode_stereo_kernel<BYTES, false, true>;
    const size_t smem = sizeof(zf::SmemStereoExact<BYTES>);
    ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, zf::kThreads, smem));
    e->smem_exact = smem;
    return ZF_OK;
}

template <int BYTES>
int setup_v3_kernel(zf_encoder *e, int *occ) {
    auto k = zf::v3::zf_encode_stereo_v3_kernel<BYTES>;
    const size_t smem = sizeof(zf::v3::Smem<BYTES>);
    ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, zf::v3::kT, smem));
    e->smem_v3 = smem;
    return ZF_OK;
}

template <int BYTES>
int setup_lpc_kernel(zf_encoder *e, int *occ) {
    auto k = zf::lpc::zf_encode_stereo_lpc_kernel<BYTES>;
    const size_t smem = sizeof(zf::lpc::SmemLpc<BYTES>);
    ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, zf::kThreads, smem));
    e->smem_lpc = smem;
    return ZF_OK;
}

// zf-LPC v1 window (zf_kernel_lpc.cuh): W[i] = 16384 - floor((2i - (n-1))^2 * 16384 / (n-1)^2)
void lpc_window(uint32_t n, std::vector<uint16_t> &w) {
    w.assign(n, 16384);
    if (n <= 1) return;
    const unsigned long long den = (unsigned long long)(n - 1) * (n - 1);
    for (uint32_t i = 0; i < n; i++) {
        const long long d = 2ll * i - (long long)(n - 1);
        w[i] = (uint16_t)(16384u - (uint32_t)((((unsigned long long)(d * d)) << 14) / den));
    }
}

template <int BYTES>
int setup_indep_kernel(zf_encoder *e, int *occ) {
    auto k = zf::zf_encode_indep_kernel<BYTES>;
    const size_t smem = zf::indep_smem_bytes(BYTES, e->cfg.channels);
    ZF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, zf::kThreads, smem));
    e->smem_indep = smem;
    return ZF_OK;
}

int setup_kernels(zf_encoder *e) {
    const int bytes = container_bytes(e->cfg);
    int rc = ZF_OK;
    if (e->stereo) {
        if (bytes == 2) { rc = setup_stereo_kernel<2>(e, &e->occ_gen); if (!rc) rc = setup_v3_kernel<2>(e, &e->occ_v3); }
        else if (bytes == 3) { rc = setup_stereo_kernel<3>(e, &e->occ_gen); if (!rc) rc = setup_v3_kernel<3>(e, &e->occ_v3); }
        else { rc = setup_stereo_kernel<4>(e, &e->occ_gen); if (!rc) rc = setup_v3_kernel<4>(e, &e->occ_v3); }
    } else {
        if (bytes == 2) rc = setup_indep_kernel<2>(e, &e->occ_gen);
        else if (bytes == 3) rc = setup_indep_kernel<3>(e, &e->occ_gen);
        else rc = setup_indep_kernel<4>(e, &e->occ_gen);
    }
    if (!rc && e->cfg.exact_rice) {
        if (bytes == 2) rc = setup_exact_kernel<2>(e, &e->occ_exact);
        else if (bytes == 3) rc = setup_exact_kernel<3>(e, &e->occ_exact);
        else rc = setup_exact_kernel<4>(e, &e->occ_exact);
        if (!rc && e->occ_exact < 1) {
            snprintf(g_cuda_err, sizeof g_cuda_err, "exact-search kernel does not fit on an SM");
            rc = ZF_ERR_CUDA;
        }
    }
    if (!rc && e->cfg.lpc_order) {
        if (bytes == 2) rc = setup_lpc_kernel<2>(e, &e->occ_lpc);
        else rc = setup_lpc_kernel<3>(e, &e->occ_lpc);
        if (!rc && e->occ_lpc < 1) {
            snprintf(g_cuda_err, sizeof g_cuda_err, "LPC kernel does not fit on an SM");
            rc = ZF_ERR_CUDA;
        }
    }
    if (rc) return rc;
    if (e->occ_gen < 1) {
        snprintf(g_cuda_err, sizeof g_cuda_err, "kernel does not fit on an SM (occupancy 0)");
        return ZF_ERR_CUDA;
    }
    return ZF_OK;
}

// `overlap`: programmatic dependent launch -- the kernel may start as soon as the kernel in front of it in the stream
// has released its dependents (the one-CTA last-frame launch does so at once), instead of after its end
template <typename K>
void launch_k(K kernel, int grid, int block, size_t smem, cudaStream_t s, bool overlap, const zf::FrameJob &job) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = overlap ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, job);
}

template <int BYTES>
void launch_stereo(int grid, size_t smem, cudaStream_t s, bool overlap, const zf::FrameJob &job) {
    launch_k(zf::zf_encode_stereo_kernel<BYTES, false>, grid, zf::kThreads, smem, s, overlap, job);
}

// the general kernels: LPC, stereo of any geometry, independent channels
void launch_one(zf_encoder *e, int grid, cudaStream_t s, const zf::FrameJob &job, bool overlap = false) {
    const int bytes = container_bytes(e->cfg);
    if (e->cfg.lpc_order) {
        if (bytes == 2) launch_k(zf::lpc::zf_encode_stereo_lpc_kernel<2>, grid, zf::kThreads, e->smem_lpc, s, overlap, job);
        else launch_k(zf::lpc::zf_encode_stereo_lpc_kernel<3>, grid, zf::kThreads, e->smem_lpc, s, overlap, job);
        return;
    }
    if (e->cfg.exact_rice) {
        if (bytes == 2) launch_k(zf::zf_encode_stereo_kernel<2, false, true>, grid, zf::kThreads, e->smem_exact, s, overlap, job);
        else if (bytes == 3) launch_k(zf::zf_encode_stereo_kernel<3, false, true>, grid, zf::kThreads, e->smem_exact, s, overlap, job);
        else launch_k(zf::zf_encode_stereo_kernel<4, false, true>, grid, zf::kThreads, e->smem_exact, s, overlap, job);
        return;
    }
    if (e->stereo) {
        if (bytes == 2) launch_stereo<2>(grid, e->smem_stereo, s, overlap, job);
        else if (bytes == 3) launch_stereo<3>(grid, e->smem_stereo, s, overlap, job);
        else launch_stereo<4>(grid, e->smem_stereo, s, overlap, job);
    } else {
        if (bytes == 2) launch_k(zf::zf_encode_indep_kernel<2>, grid, zf::kThreads, e->smem_indep, s, overlap, job);
        else if (bytes == 3) launch_k(zf::zf_encode_indep_kernel<3>, grid, zf::kThreads, e->smem_indep, s, overlap, job);
        else launch_k(zf::zf_encode_indep_kernel<4>, grid, zf::kThreads, e->smem_indep, s, overlap, job);
    }
}

// Enqueue the kernels of one batch on `s`.  All pointers are device pointers.
int launch_batch(zf_encoder *e, Slot &sl, const uint8_t *d_pcm, uint64_t samples, uint64_t first_frame_number, uint8_t *d_out,
                 size_t out_cap, uint32_t *d_sizes, unsigned long long *d_total, cudaStream_t s, int *launches) {
    const uint32_t bs = e->cfg.block_size;
    const uint64_t frames = (samples + bs - 1) / bs;
    const uint64_t full = samples / bs;
    const uint32_t tail = (uint32_t)(samples - full * bs);
    *launches = 0;
    if (frames == 0) {
        ZF_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s));
        return ZF_OK;
    }
    if (frames > e->cfg.max_frames_per_batch) return ZF_ERR_INVALID_ARG;
    if (first_frame_number + frames > (1ull << 31)) return ZF_ERR_UNSUPPORTED;  // header coder is UB upstream (Q16)
    ZF_CUDA(cudaMemsetAsync(sl.d_ctl_block, 0, 64 + sizeof(unsigned long long) * frames, s));  // tickets, status, descriptors
    ZF_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s));
    zf::FrameJob job;
    memset(&job, 0, sizeof job);
    job.out = d_out;
    job.out_cap = out_cap;
    job.frame_sizes = d_sizes;
    job.desc = sl.d_desc;
    job.status = sl.d_ctl + 2;
    job.total_bytes = d_total;
    job.pow8 = e->d_pow8;
    job.batch_frames = (uint32_t)frames;
    job.first_frame_number = first_frame_number;
    job.frame_stride = (uint32_t)e->frame_dev_bytes;
    job.bit_depth = e->cfg.bit_depth;
    job.lpc_order = e->cfg.lpc_order;
    job.lpc_window = e->d_win;
    job.sample_rate = e->cfg.sample_rate;
    job.channels = e->cfg.channels;
    job.max_rice_order = e->cfg.max_rice_order;
    job.max_rice_param = e->cfg.max_rice_param;
    if (e->cfg.bit_depth == 8) {  // one signed byte per sample in, 16-bit containers for the kernels
        if (!sl.d_wide) ZF_CUDA(cudaMalloc(&sl.d_wide, (size_t)e->cfg.max_frames_per_batch * e->frame_dev_bytes));
        const unsigned long long count = samples * e->cfg.channels;
        const int blocks = (int)std::min<unsigned long long>((count + 255) / 256, 4096);
        zf::zf_widen8_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const int8_t *>(d_pcm), reinterpret_cast<int16_t *>(sl.d_wide), count);
        d_pcm = sl.d_wide;
        (*launches)++;
    }
    // the fast kernel covers the reference's default block: 4096-sample frames of 16/24/32-bit samples, any Rice limits
    const bool fast = e->stereo && bs == (uint32_t)zf::kMaxBlock && e->cfg.bit_depth != 8 &&
                      !e->cfg.lpc_order && !e->cfg.exact_rice;
    if (e->cfg.lpc_order && tail && sl.win_tail_len != tail) {  // the short last frame has a window of its own length
        std::vector<uint16_t> w;
        lpc_window(tail, w);
        if (!sl.d_win_tail) ZF_CUDA(cudaMalloc(&sl.d_win_tail, sizeof(uint16_t) * zf::kMaxBlock));
        ZF_CUDA(cudaMemcpyAsync(sl.d_win_tail, w.data(), sizeof(uint16_t) * tail, cudaMemcpyHostToDevice, s));  // pageable: staged at once
        sl.win_tail_len = tail;
    }
    // the 1-D TMA bulk copy needs 16-byte aligned sources; frame strides are multiples of 16 already
    job.use_tma = ((uintptr_t)d_pcm & 15u) == 0 ? 1u : 0u;
    // The short last frame is a one-CTA launch of the general kernel IN FRONT of the persistent full-frame kernel, which is
    // launched as its programmatic dependent: the short-frame CTA releases its dependents at once, so both run together
    // (launched after, or on another stream, it would only get an SM when the persistent kernel ends).
    const bool split_tail = tail && full;
    const uint32_t ring = sl.kev_count % kRing;
    if (full && !sl.kev[2 * ring]) {  // timing events of the full-frame kernel: created on first use
        ZF_CUDA(cudaEventCreate(&sl.kev[2 * ring]));
        ZF_CUDA(cudaEventCreate(&sl.kev[2 * ring + 1]));
    }
    if (full) ZF_CUDA(cudaEventRecord(sl.kev[2 * ring], s));
    if (split_tail) {
        zf::FrameJob tj = job;
        tj.pcm = d_pcm + full * e->frame_dev_bytes;
        tj.out = sl.d_tail;
        tj.out_cap = e->max_frame_bytes + 64;
        tj.frame_sizes = reinterpret_cast<uint32_t *>(sl.d_tail_meta + 2);
        tj.desc = sl.d_tail_meta;
        tj.total_bytes = sl.d_tail_meta + 1;
        tj.n_frames = 1;
        tj.frame_base = 0;
        tj.batch_frames = 1;
        tj.first_frame_number = first_frame_number + full;
        tj.block_size = tail;
        tj.lpc_window = sl.d_win_tail;
        tj.ticket = sl.d_ctl + 1;
        tj.pdl_trigger = 1;
        launch_one(e, 1, s, tj);
        (*launches)++;
    }
    if (full) {
        job.pcm = d_pcm;
        job.n_frames = (uint32_t)full;
        job.frame_base = 0;
        job.block_size = bs;
        job.ticket = sl.d_ctl + 0;
        if (split_tail) job.batch_frames = (uint32_t)full;  // the full-frame kernel closes its own total
        // full 4096-sample stereo frames with the reference's partition depth: the lean 256-thread kernel (zf_kernel_v3.cuh)
        // (32-bit PCM with a low max_rice_param stays with the general kernel: where no escape is possible -- residuals of
        // 32 bits -- and no high parameter either, partition costs exceed the lean kernel's 32-bit cost arithmetic)
        const bool v3 = fast && e->occ_v3 > 0 && (e->cfg.bit_depth != 32 || e->cfg.max_rice_param >= 20);
        const int occ = e->cfg.lpc_order ? e->occ_lpc : e->cfg.exact_rice ? e->occ_exact : v3 ? e->occ_v3 : e->occ_gen;
        const int grid = (int)std::min<uint64_t>(full, (uint64_t)e->sm_count * occ);
        if (v3) {
            if (e->cfg.bit_depth == 16) launch_k(zf::v3::zf_encode_stereo_v3_kernel<2>, grid, zf::v3::kT, e->smem_v3, s, split_tail, job);
            else if (e->cfg.bit_depth == 24) launch_k(zf::v3::zf_encode_stereo_v3_kernel<3>, grid, zf::v3::kT, e->smem_v3, s, split_tail, job);
            else launch_k(zf::v3::zf_encode_stereo_v3_kernel<4>, grid, zf::v3::kT, e->smem_v3, s, split_tail, job);
        } else {
            launch_one(e, grid, s, job, split_tail);
        }
        ZF_CUDA(cudaEventRecord(sl.kev[2 * ring + 1], s));
        sl.kev_count++;
        (*launches)++;
    }
    if (split_tail) {
        zf::zf_append_tail_kernel<<<1, 512, 0, s>>>(sl.d_tail, reinterpret_cast<const uint32_t *>(sl.d_tail_meta + 2), d_out,
                                                    out_cap, d_total, d_sizes, (uint32_t)full, sl.d_ctl + 2);
        (*launches)++;
    } else if (tail) {  // a stream shorter than one block: the short frame is the whole batch
        job.pcm = d_pcm;
        job.n_frames = 1;
        job.frame_base = 0;
        job.block_size = tail;
        job.lpc_window = sl.d_win_tail;
        job.ticket = sl.d_ctl + 1;
        launch_one(e, 1, s, job);
        (*launches)++;
    }
    ZF_CUDA(cudaGetLastError());
    return ZF_OK;
}

int ensure_io(zf_encoder *e, Slot &sl) {
    if (sl.have_io) return ZF_OK;
    const size_t frames = e->cfg.max_frames_per_batch;
    sl.pcm_cap = frames * e->frame_pcm_bytes;
    sl.out_cap = frames * e->max_frame_bytes + 64;
    ZF_CUDA(cudaMalloc(&sl.d_pcm, sl.pcm_cap));
    ZF_CUDA(cudaMalloc(&sl.d_out, sl.out_cap));
    sl.have_io = true;
    return ZF_OK;
}

// Page-locked staging for callers whose own buffers are pageable; allocated on first need only (page-locking memory is
// slow -- of the order of a GB/s -- and callers that bring zf_host_alloc buffers never need it).
int ensure_staging_in(Slot &sl) {
    if (!sl.h_pcm) ZF_CUDA(cudaMallocHost(&sl.h_pcm, sl.pcm_cap));
    return ZF_OK;
}
int ensure_staging_out(Slot &sl) {
    if (!sl.h_out) ZF_CUDA(cudaMallocHost(&sl.h_out, sl.out_cap));
    return ZF_OK;
}

int slot_init(zf_encoder *e, Slot &sl) {
    const size_t frames = e->cfg.max_frames_per_batch;
    ZF_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    ZF_CUDA(cudaMalloc(&sl.d_tail, e->max_frame_bytes + 64));
    ZF_CUDA(cudaEventCreate(&sl.ev_start));
    ZF_CUDA(cudaEventCreate(&sl.ev_stop));
    ZF_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    ZF_CUDA(cudaEventCreateWithFlags(&sl.ev_up, cudaEventDisableTiming));
    ZF_CUDA(cudaEventCreateWithFlags(&sl.ev_down, cudaEventDisableTiming));
    ZF_CUDA(cudaMalloc(&sl.d_sizes, sizeof(uint32_t) * frames));
    ZF_CUDA(cudaMalloc(&sl.d_ctl_block, 64 + sizeof(unsigned long long) * frames));
    sl.d_ctl = reinterpret_cast<unsigned int *>(sl.d_ctl_block);
    sl.d_tail_meta = reinterpret_cast<unsigned long long *>(sl.d_ctl_block + 16);
    sl.d_desc = reinterpret_cast<unsigned long long *>(sl.d_ctl_block + 64);
    ZF_CUDA(cudaMalloc(&sl.d_total, sizeof(unsigned long long)));
    ZF_CUDA(cudaMallocHost(&sl.h_sizes, sizeof(uint32_t) * frames));
    ZF_CUDA(cudaMallocHost(&sl.h_total, sizeof(unsigned long long) * 2));
    return ZF_OK;
}

void slot_free(Slot &sl) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    cudaFree(sl.d_tail);
    cudaFree(sl.d_win_tail);
    cudaFree(sl.d_wide);
    cudaFree(sl.d_pcm); cudaFree(sl.d_out); cudaFree(sl.d_sizes); cudaFree(sl.d_ctl_block); cudaFree(sl.d_total);
    cudaFreeHost(sl.h_pcm); cudaFreeHost(sl.h_out); cudaFreeHost(sl.h_sizes); cudaFreeHost(sl.h_total);
    if (sl.ev_start) cudaEventDestroy(sl.ev_start);
    if (sl.ev_stop) cudaEventDestroy(sl.ev_stop);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    if (sl.ev_up) cudaEventDestroy(sl.ev_up);
    if (sl.ev_down) cudaEventDestroy(sl.ev_down);
    for (int i = 0; i < 2 * kRing; i++) if (sl.kev[i]) cudaEventDestroy(sl.kev[i]);
    if (sl.stream) cudaStreamDestroy(sl.stream);
    sl = Slot();
}

bool is_pinned_or_device_visible(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// H2D + kernels + D2H of the small results for one batch held in a slot (asynchronous).
// The two words the host needs before it can start the download (bytes produced, status flags) go straight to pinned host
// memory: a copy-engine transfer of 12 bytes took 20-115 us while the link was busy with the next upload.
__global__ void zf_publish_kernel(const unsigned long long *d_total, const unsigned int *d_status, unsigned long long *h_pub) {
    h_pub[0] = *d_total;
    h_pub[1] = *d_status;
}

// Upload (upload stream) + kernels (the slot's stream) for one batch held in a slot (asynchronous).  Uploads and downloads
// of all batches each run in order on a stream of their own: issued on per-slot streams, copies of both directions blocked
// one another on the copy engines (an upload waited for the download of the batch two before it).
int slot_submit(zf_encoder *e, Slot &sl, const uint8_t *pcm, uint64_t samples, uint64_t first_frame_number) {
    int rc = ensure_io(e, sl);
    if (rc) return rc;
    const uint32_t bs = e->cfg.block_size;
    const uint64_t frames = (samples + bs - 1) / bs;
    if (frames > e->cfg.max_frames_per_batch) return ZF_ERR_INVALID_ARG;
    const size_t bytes = (size_t)samples * e->cfg.channels * (e->cfg.bit_depth / 8);
    const uint8_t *src = pcm;
    if (bytes && !is_pinned_or_device_visible(pcm)) {  // pageable caller memory: stage through pinned
        rc = ensure_staging_in(sl);
        if (rc) return rc;
        memcpy(sl.h_pcm, pcm, bytes);
        src = sl.h_pcm;
    }
    tr_mark(e, e->s_up, kTrUp0);
    if (bytes) ZF_CUDA(cudaMemcpyAsync(sl.d_pcm, src, bytes, cudaMemcpyHostToDevice, e->s_up));
    ZF_CUDA(cudaEventRecord(sl.ev_up, e->s_up));
    tr_mark(e, e->s_up, kTrUp1);
    ZF_CUDA(cudaStreamWaitEvent(sl.stream, sl.ev_up, 0));
    if (&sl == &e->slot[0] && e->dev_pending) {  // a zf_encode_device batch may still be using this slot's control block
        ZF_CUDA(cudaStreamWaitEvent(sl.stream, e->ev_dev, 0));
        e->dev_pending = false;
    }
    ZF_CUDA(cudaEventRecord(sl.ev_start, sl.stream));
    int launches = 0;
    rc = launch_batch(e, sl, sl.d_pcm, samples, first_frame_number, sl.d_out, sl.out_cap, sl.d_sizes, sl.d_total, sl.stream,
                      &launches);
    if (rc) return rc;
    ZF_CUDA(cudaEventRecord(sl.ev_stop, sl.stream));
    tr_mark(e, sl.stream, kTrKern1);
    e->launches_last = launches + 1;
    zf_publish_kernel<<<1, 1, 0, sl.stream>>>(sl.d_total, sl.d_ctl + 2, sl.h_total);
    ZF_CUDA(cudaGetLastError());
    ZF_CUDA(cudaEventRecord(sl.ev_done, sl.stream));
    tr_mark(e, sl.stream, kTrSmall1);
    sl.frames = (uint32_t)frames;
    sl.busy = true;
    return ZF_OK;
}

// Wait for the batch's kernels, check the status, and start the copy of exactly the produced bytes (asynchronous).
int slot_fetch(zf_encoder *e, Slot &sl, uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *frame_sizes,
               uint32_t frame_sizes_cap, uint32_t *n_frames) {
    if (!sl.busy) return ZF_ERR_INVALID_ARG;
    sl.busy = false;
    ZF_CUDA(cudaEventSynchronize(sl.ev_done));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sl.ev_start, sl.ev_stop) == cudaSuccess) e->kernel_ms_last = ms;
    const unsigned int status = (unsigned int)sl.h_total[1];
    if (status & zf::kStatusOutOverflow) return ZF_ERR_OUT_TOO_SMALL;
    if (status) {
        snprintf(g_cuda_err, sizeof g_cuda_err, "kernel status flags 0x%x", status);
        return ZF_ERR_CUDA;
    }
    const size_t total = (size_t)sl.h_total[0];
    if (n_frames) *n_frames = sl.frames;
    if (out_len) *out_len = total;
    if (sl.frames > frame_sizes_cap) return ZF_ERR_OUT_TOO_SMALL;
    if (total > out_cap) return ZF_ERR_OUT_TOO_SMALL;
    sl.copy_dst = nullptr;
    sl.copy_len = 0;
    sl.sizes_dst = frame_sizes;
    tr_mark(e, e->s_down, kTrDown0);
    if (total) {
        if (is_pinned_or_device_visible(out)) {
            ZF_CUDA(cudaMemcpyAsync(out, sl.d_out, total, cudaMemcpyDeviceToHost, e->s_down));
        } else {  // pageable caller memory: through the pinned staging buffer, copied on in slot_finish
            const int rs = ensure_staging_out(sl);
            if (rs) return rs;
            ZF_CUDA(cudaMemcpyAsync(sl.h_out, sl.d_out, total, cudaMemcpyDeviceToHost, e->s_down));
            sl.copy_dst = out;
            sl.copy_len = total;
        }
    }
    if (frame_sizes && sl.frames)
        ZF_CUDA(cudaMemcpyAsync(sl.h_sizes, sl.d_sizes, sizeof(uint32_t) * sl.frames, cudaMemcpyDeviceToHost, e->s_down));
    ZF_CUDA(cudaEventRecord(sl.ev_down, e->s_down));
    tr_mark(e, e->s_down, kTrDown1);
    sl.draining = true;
    return ZF_OK;
}

// Wait until the batch's output has arrived; the slot is free afterwards.
int slot_finish(zf_encoder *, Slot &sl) {
    if (!sl.draining) return ZF_OK;
    sl.draining = false;
    ZF_CUDA(cudaEventSynchronize(sl.ev_down));
    if (sl.sizes_dst && sl.frames) memcpy(sl.sizes_dst, sl.h_sizes, sizeof(uint32_t) * sl.frames);
    if (sl.copy_dst && sl.copy_len) memcpy(sl.copy_dst, sl.h_out, sl.copy_len);
    return ZF_OK;
}

int slot_collect(zf_encoder *e, Slot &sl, uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *frame_sizes,
                 uint32_t frame_sizes_cap, uint32_t *n_frames) {
    int rc = slot_fetch(e, sl, out, out_cap, out_len, frame_sizes, frame_sizes_cap, n_frames);
    if (rc) return rc;
    return slot_finish(e, sl);
}

}  // namespace

// zf_decode.cu reports through the same thread-local text (library-internal, not part of the ABI)
__attribute__((visibility("hidden"))) void zf_internal_set_error(const char *msg) {
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s", msg);
}

extern "C" {

const char *zf_last_cuda_error(void) { return g_cuda_err; }

int zf_device_check(int device_id) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        snprintf(g_cuda_err, sizeof g_cuda_err, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return ZF_ERR_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= n) return ZF_ERR_NO_DEVICE;
    cudaDeviceProp p;
    ZF_CUDA(cudaGetDeviceProperties(&p, device_id));
    if (p.major != 10) {
        snprintf(g_cuda_err, sizeof g_cuda_err, "device %d is sm_%d%d; this library carries sm_100a code only", device_id,
                 p.major, p.minor);
        return ZF_ERR_NO_DEVICE;
    }
    return ZF_OK;
}

int zf_config_default(zf_config *cfg, uint8_t channels, uint8_t bit_depth, uint32_t sample_rate) {
    if (!cfg) return ZF_ERR_INVALID_ARG;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = sizeof *cfg;
    cfg->block_size = 4096;  // encoder.zig:644
    cfg->bit_depth = bit_depth;
    cfg->channels = channels;
    cfg->sample_rate = sample_rate;
    cfg->stereo_decorrelation = 1;  // :649
    cfg->max_rice_order = 8;        // :651
    cfg->max_rice_param = 30;       // rice.MAX_PARAM, rice.zig:9-10
    cfg->device_id = 0;
    cfg->max_frames_per_batch = 2048;
    return ZF_OK;
}

size_t zf_max_frame_bytes(const zf_config *cfg) { return cfg ? max_frame_bytes_of(cfg) : 0; }

size_t zf_max_batch_bytes(const zf_config *cfg, uint32_t n_frames) {
    return cfg ? (size_t)n_frames * max_frame_bytes_of(cfg) + 64 : 0;
}

int zf_encoder_create(const zf_config *cfg, zf_encoder **out) {
    if (!cfg || !out) return ZF_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->struct_size != sizeof(zf_config)) return ZF_ERR_INVALID_ARG;
    // asserts of Encoder.init, encoder.zig:49-51
    if (cfg->block_size == 0 || cfg->channels == 0 || cfg->channels > 8 || cfg->bit_depth == 0 || cfg->bit_depth % 4 != 0)
        return ZF_ERR_INVALID_ARG;
    if (!depth_ok(cfg->bit_depth)) return ZF_ERR_UNSUPPORTED;          // 4/12/20-bit: `unreachable` upstream (frame_writer.zig:221-233)
    if (cfg->block_size > zf::kMaxBlock) return ZF_ERR_UNSUPPORTED;    // one CTA holds at most 4096 samples/channel
    if (cfg->max_rice_order > 8) return ZF_ERR_UNSUPPORTED;            // rice.MAX_ORDER = 8 (rice.zig:12)
    if (cfg->max_rice_param == 0 || cfg->max_rice_param > 30) return ZF_ERR_UNSUPPORTED;  // 0: overflow upstream
    if (cfg->max_frames_per_batch == 0) return ZF_ERR_INVALID_ARG;
    if (cfg->exact_rice) {  // the exact-search extension: the general stereo kernel only
        if (cfg->exact_rice > 1 || cfg->lpc_order || cfg->channels != 2 || !cfg->stereo_decorrelation) return ZF_ERR_UNSUPPORTED;
    }
    if (cfg->lpc_order) {  // the LPC extension (no reference counterpart): stereo with decorrelation, 8/16/24-bit, order <= 12
        if (cfg->lpc_order > zf::lpc::kMaxOrder || cfg->channels != 2 || !cfg->stereo_decorrelation || cfg->bit_depth == 32)
            return ZF_ERR_UNSUPPORTED;
    }
    int rc = zf_device_check(cfg->device_id);
    if (rc) return rc;
    ZF_CUDA(cudaSetDevice(cfg->device_id));
    zf_encoder *e = new (std::nothrow) zf_encoder();
    if (!e) return ZF_ERR_NOMEM;
    e->cfg = *cfg;
    e->stereo = cfg->channels == 2 && cfg->stereo_decorrelation;
    { const char *tr = getenv("ZF_TRACE"); e->trace = tr && tr[0] == '1'; }
    { const char *nt = getenv("ZF_NO_TAPER"); e->no_taper = nt && nt[0] == '1'; }  // A/B of the batch plan (development aid)
    e->frame_pcm_bytes = (size_t)cfg->block_size * cfg->channels * (cfg->bit_depth / 8);
    e->frame_dev_bytes = (size_t)cfg->block_size * cfg->channels * container_bytes(*cfg);
    e->max_frame_bytes = max_frame_bytes_of(cfg);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device_id) != cudaSuccess) { delete e; return ZF_ERR_CUDA; }
    e->sm_count = prop.multiProcessorCount;
    rc = setup_kernels(e);
    for (int i = 0; i < kSlots && !rc; i++) rc = slot_init(e, e->slot[i]);
    if (!rc && (cudaStreamCreateWithFlags(&e->s_up, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&e->s_down, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&e->ev_dev, cudaEventDisableTiming) != cudaSuccess))
        rc = ZF_ERR_CUDA;
    if (!rc) {
        std::vector<uint16_t> pw(kPow8Len);
        uint32_t v = 1;
        for (int k = 0; k < kPow8Len; k++) {
            pw[k] = (uint16_t)v;
            for (int b = 0; b < 8; b++) v = (v & 0x8000u) ? (((v << 1) ^ 0x8005u) & 0xffffu) : ((v << 1) & 0xffffu);
        }
        cudaError_t ce = cudaMalloc(&e->d_pow8, sizeof(uint16_t) * kPow8Len);
        if (ce == cudaSuccess) ce = cudaMemcpy(e->d_pow8, pw.data(), sizeof(uint16_t) * kPow8Len, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) {
            snprintf(g_cuda_err, sizeof g_cuda_err, "pow8 table: %s", cudaGetErrorString(ce));
            rc = ZF_ERR_CUDA;
        }
    }
    if (!rc && cfg->lpc_order) {
        std::vector<uint16_t> w;
        lpc_window(cfg->block_size, w);
        cudaError_t ce = cudaMalloc(&e->d_win, sizeof(uint16_t) * zf::kMaxBlock);
        if (ce == cudaSuccess) ce = cudaMemcpy(e->d_win, w.data(), sizeof(uint16_t) * w.size(), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) rc = ZF_ERR_CUDA;
    }
    if (rc) {
        zf_encoder_destroy(e);
        return rc;
    }
    *out = e;
    return ZF_OK;
}

void zf_encoder_destroy(zf_encoder *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device_id);
    if (e->s_up) cudaStreamSynchronize(e->s_up);
    if (e->s_down) cudaStreamSynchronize(e->s_down);
    for (int i = 0; i < kSlots; i++) slot_free(e->slot[i]);
    if (e->s_up) cudaStreamDestroy(e->s_up);
    if (e->s_down) cudaStreamDestroy(e->s_down);
    if (e->ev_dev) cudaEventDestroy(e->ev_dev);
    cudaFree(e->d_pow8);
    cudaFree(e->d_win);
    for (cudaEvent_t ev : e->tr_ev) cudaEventDestroy(ev);
    delete e;
}

int zf_host_alloc(size_t bytes, void **out) {
    if (!out) return ZF_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return ZF_ERR_NO_DEVICE;
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return ZF_ERR_NOMEM;
    }
    *out = p;
    return ZF_OK;
}

void zf_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int zf_encode_submit(zf_encoder *e, const uint8_t *pcm, uint64_t samples_per_channel, uint64_t first_frame_number) {
    if (!e || (!pcm && samples_per_channel)) return ZF_ERR_INVALID_ARG;
    if (e->slot[0].busy || e->slot[0].draining) return ZF_ERR_BUSY;
    ZF_CUDA(cudaSetDevice(e->cfg.device_id));
    int rc = slot_submit(e, e->slot[0], pcm, samples_per_channel, first_frame_number);
    if (rc) return rc;
    // The header's contract: `pcm` may be reused as soon as submit returns.  Pageable memory has been copied to the pinned
    // staging buffer already; pinned / managed memory is read by the copy engine itself, so wait for that upload (only
    // the upload: the kernels and the download stay asynchronous).
    if (samples_per_channel && is_pinned_or_device_visible(pcm)) ZF_CUDA(cudaEventSynchronize(e->slot[0].ev_up));
    return ZF_OK;
}

int zf_encode_collect(zf_encoder *e, uint8_t *out, size_t out_cap, size_t *out_len, uint32_t *frame_sizes,
                      uint32_t frame_sizes_cap, uint32_t *n_frames) {
    if (!e || !out) return ZF_ERR_INVALID_ARG;
    ZF_CUDA(cudaSetDevice(e->cfg.device_id));
    return slot_collect(e, e->slot[0], out, out_cap, out_len, frame_sizes, frame_sizes_cap, n_frames);
}

int zf_encode_pcm(zf_encoder *e, const uint8_t *pcm, uint64_t samples_per_channel, uint64_t first_frame_number, uint8_t *out,
                  size_t out_cap, size_t *out_len, uint32_t *frame_sizes, uint32_t frame_sizes_cap, uint32_t *n_frames) {
    if (!e || !out || (!pcm && samples_per_channel)) return ZF_ERR_INVALID_ARG;
    for (int i = 0; i < kSlots; i++)
        if (e->slot[i].busy || e->slot[i].draining) return ZF_ERR_BUSY;
    ZF_CUDA(cudaSetDevice(e->cfg.device_id));
    const uint32_t bs = e->cfg.block_size;
    const uint64_t frames = (samples_per_channel + bs - 1) / bs;
    if (n_frames) *n_frames = (uint32_t)frames;
    if (out_len) *out_len = 0;
    if (frames > frame_sizes_cap && frame_sizes) return ZF_ERR_OUT_TOO_SMALL;
    const uint64_t per = e->cfg.max_frames_per_batch;
    // Batch plan: a ramp of growing batches (256, 512, ... frames), full batches while more than one batch is left, then
    // halves down to kMinTailBatch frames.  The downloads are as long a chain as the uploads (both links run near 7 ms for
    // the 10-minute stream when they share the bus), so the first download must not wait for a full batch to be uploaded
    // and encoded; and what follows the last upload -- that batch's kernels and its download -- is overlapped with
    // nothing, so it is kept short.
    std::vector<uint64_t> first;  // first frame of every batch, then the frame count
    {
        const uint64_t kMinTailBatch = 256;
        const bool taper = !e->no_taper;
        uint64_t f = 0, ramp = taper ? kMinTailBatch : per;
        while (f < frames) {
            const uint64_t left = frames - f;
            uint64_t n = std::min(left, std::min(per, ramp));
            if (taper && left <= per && left > 2 * kMinTailBatch && n > (left + 1) / 2) n = (left + 1) / 2;
            first.push_back(f);
            f += n;
            ramp *= 2;
        }
        first.push_back(frames);
    }
    const uint64_t nbatch = first.size() - 1;
    const size_t ic_bytes = (size_t)e->cfg.channels * (e->cfg.bit_depth / 8);
    size_t pos = 0;
    float ms = 0.f;
    int launches = 0;
    auto fail = [&](int rc) {
        for (int i = 0; i < kSlots; i++) e->slot[i].busy = e->slot[i].draining = false;
        cudaDeviceSynchronize();
        return rc;
    };
    if (e->trace) {
        for (cudaEvent_t ev : e->tr_ev) cudaEventDestroy(ev);
        e->tr_ev.assign(nbatch * kTracePoints, nullptr);
        for (cudaEvent_t &ev : e->tr_ev) cudaEventCreate(&ev);
        e->tr_host.assign(nbatch * 3, 0.0);
        e->tr_h0 = std::chrono::steady_clock::now();
    }
    // pipeline over the slots: batch b is queued (upload, kernels), batch b-1's kernels are awaited and its download is
    // queued, and the arrival of the oldest batch in flight frees the slot for batch b+1
    constexpr uint64_t kLag = kSlots - 1;
    for (uint64_t b = 0; b < nbatch + kLag; b++) {
        if (b < nbatch) {
            const uint64_t f0 = first[b];
            const uint64_t s0 = f0 * bs;
            const uint64_t ns = std::min<uint64_t>((first[b + 1] - f0) * bs, samples_per_channel - s0);
            e->tr_batch = b;
            int rc = slot_submit(e, e->slot[b % kSlots], pcm + s0 * ic_bytes, ns, first_frame_number + f0);
            if (rc) return fail(rc);
            tr_host(e, b, 0);
            launches += e->launches_last;
        }
        if (b >= 1 && b - 1 < nbatch) {
            const uint64_t f0 = first[b - 1];
            size_t got = 0;
            uint32_t nf = 0;
            e->tr_batch = b - 1;
            int rc = slot_fetch(e, e->slot[(b - 1) % kSlots], out + pos, out_cap - pos, &got,
                                frame_sizes ? frame_sizes + f0 : nullptr, frame_sizes ? (uint32_t)(frames - f0) : 0xffffffffu, &nf);
            if (rc) return fail(rc);
            pos += got;
            ms += e->kernel_ms_last;
            tr_host(e, b - 1, 1);
        }
        if (b >= kLag) {
            int rc = slot_finish(e, e->slot[(b - kLag) % kSlots]);
            if (rc) return fail(rc);
            tr_host(e, b - kLag, 2);
        }
    }
    if (e->trace && nbatch) {
        static const char *names[kTracePoints] = {"up0", "up1", "kern1", "small1", "down0", "down1"};
        fprintf(stderr, "[zf trace] %llu batches; device times relative to batch 0 up0, host times relative to the call (ms)\n",
                (unsigned long long)nbatch);
        for (uint64_t b = 0; b < nbatch; b++) {
            fprintf(stderr, "[zf trace] batch %2llu:", (unsigned long long)b);
            for (int k = 0; k < kTracePoints; k++) {
                float t = 0.f;
                cudaEventElapsedTime(&t, e->tr_ev[0], e->tr_ev[b * kTracePoints + k]);
                fprintf(stderr, " %s %.3f", names[k], t);
            }
            fprintf(stderr, " | host submit %.3f fetch %.3f finish %.3f\n", e->tr_host[b * 3], e->tr_host[b * 3 + 1],
                    e->tr_host[b * 3 + 2]);
        }
    }
    e->kernel_ms_last = ms;
    e->launches_last = launches;
    if (out_len) *out_len = pos;
    return ZF_OK;
}

int zf_encode_device(zf_encoder *e, const void *d_pcm, uint64_t samples_per_channel, uint64_t first_frame_number, void *d_out,
                     size_t out_cap, uint32_t *d_frame_sizes, uint64_t *d_total_bytes, void *stream) {
    if (!e || !d_out || !d_frame_sizes || !d_total_bytes || (!d_pcm && samples_per_channel)) return ZF_ERR_INVALID_ARG;
    ZF_CUDA(cudaSetDevice(e->cfg.device_id));
    Slot &sl = e->slot[0];
    // the batch uses slot 0's tickets, status word, look-back descriptors and last-frame scratch
    if (sl.busy || sl.draining) return ZF_ERR_BUSY;
    cudaStream_t s = stream ? (cudaStream_t)stream : sl.stream;
    // ... and so did the previous device-resident batch, which may have been enqueued on another stream: order them
    if (e->dev_pending) ZF_CUDA(cudaStreamWaitEvent(s, e->ev_dev, 0));
    ZF_CUDA(cudaEventRecord(sl.ev_start, s));
    int launches = 0;
    int rc = launch_batch(e, sl, (const uint8_t *)d_pcm, samples_per_channel, first_frame_number, (uint8_t *)d_out, out_cap,
                          d_frame_sizes, (unsigned long long *)d_total_bytes, s, &launches);
    if (rc) return rc;
    ZF_CUDA(cudaEventRecord(sl.ev_stop, s));
    ZF_CUDA(cudaEventRecord(e->ev_dev, s));
    e->dev_pending = true;
    e->launches_last = launches;
    return ZF_OK;
}

int zf_encode_device_status(zf_encoder *e, uint32_t *flags) {
    if (!e || !flags) return ZF_ERR_INVALID_ARG;
    ZF_CUDA(cudaSetDevice(e->cfg.device_id));
    if (e->dev_pending) {
        ZF_CUDA(cudaEventSynchronize(e->ev_dev));
        e->dev_pending = false;
    }
    unsigned int st = 0;
    ZF_CUDA(cudaMemcpy(&st, e->slot[0].d_ctl + 2, sizeof st, cudaMemcpyDeviceToHost));
    *flags = (st & zf::kStatusOutOverflow ? ZF_STATUS_OUT_OVERFLOW : 0u) | (st & zf::kStatusBitOverflow ? ZF_STATUS_INTERNAL : 0u);
    return ZF_OK;
}

int zf_last_batch_stats(zf_encoder *e, float *kernel_ms, uint32_t *launches) {
    if (!e) return ZF_ERR_INVALID_ARG;
    if (kernel_ms) *kernel_ms = e->kernel_ms_last;
    if (launches) *launches = (uint32_t)e->launches_last;
    return ZF_OK;
}

int zf_kernel_times(zf_encoder *e, float *ms, uint32_t cap, uint32_t *n) {
    if (!e || !n) return ZF_ERR_INVALID_ARG;
    ZF_CUDA(cudaSetDevice(e->cfg.device_id));
    uint32_t got = 0;
    for (int si = 0; si < kSlots; si++) {
        Slot &sl = e->slot[si];
        const uint32_t have = std::min<uint32_t>(sl.kev_count, kRing);
        for (uint32_t k = 0; k < have; k++) {
            const uint32_t ring = (sl.kev_count - have + k) % kRing;
            float v = 0.f;
            ZF_CUDA(cudaEventSynchronize(sl.kev[2 * ring + 1]));
            ZF_CUDA(cudaEventElapsedTime(&v, sl.kev[2 * ring], sl.kev[2 * ring + 1]));
            if (ms && got < cap) ms[got] = v;
            got++;
        }
        sl.kev_count = 0;
    }
    *n = got < cap ? got : cap;
    return ZF_OK;
}

int zf_write_frame(zf_encoder *e, const int32_t *const *planes, uint32_t samples_count, uint64_t frame_number, uint8_t *out,
                   size_t out_cap, size_t *frame_len) {
    if (!e || !planes || !out || samples_count == 0) return ZF_ERR_INVALID_ARG;  // assert encoder.zig:235
    if (samples_count > e->cfg.block_size) return ZF_ERR_INVALID_ARG;
    // planar sign-extended i32 (Encoder.samples) -> the packed interleaved layout the kernels unpack
    const unsigned ch = e->cfg.channels, nb = e->cfg.bit_depth / 8;
    std::vector<uint8_t> pcm((size_t)samples_count * ch * nb);
    size_t o = 0;
    for (uint32_t i = 0; i < samples_count; i++)
        for (unsigned c = 0; c < ch; c++) {
            const uint32_t v = (uint32_t)planes[c][i];
            for (unsigned k = 0; k < nb; k++) pcm[o++] = (uint8_t)(v >> (8 * k));
        }
    uint32_t size = 0, nf = 0;
    size_t len = 0;
    int rc = zf_encode_pcm(e, pcm.data(), samples_count, frame_number, out, out_cap, &len, &size, 1, &nf);
    if (rc) return rc;
    if (frame_len) *frame_len = len;
    return ZF_OK;
}

}  // extern "C"

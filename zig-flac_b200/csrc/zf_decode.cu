// zf_decode.cu -- C ABI of the decoder (include/zigflac_b200.h, "decoder" section): handle, whole-stream decode from
// host or device memory, file-to-file.  Kernels: zf_kernel_decode.cuh.  No CPU fallback: every decode entry fails
// without an sm_100 device.
//
// Flow of one call:  metadata blocks (host)  ->  stream to the device  ->  header scan kernel  ->  hits back to the
// host, chained by frame number into the frame table  ->  per batch of frames, on alternating streams: one thread per
// frame parses, CRC-16, restore + interleave  ->  PCM back (overlapping the next batch's kernels).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/zigflac_b200.h"
#include "zf_decode_host.h"
#include "zf_kernel_decode.cuh"

void zf_internal_set_error(const char *msg);  // zf_capi.cu

namespace {

#define ZFD_CUDA(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess) {                                                                           \
            char buf__[256];                                                                                \
            snprintf(buf__, sizeof buf__, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            zf_internal_set_error(buf__);                                                                   \
            cudaGetLastError();                                                                             \
            return ZF_ERR_CUDA;                                                                             \
        }                                                                                                   \
    } while (0)

using zf::dec::Cand;
using zf::dec::FrameRec;
using zf::dec::StreamParams;

constexpr int kSlots = 2;
constexpr size_t kWorkBytesPerSlot = 2048ull << 20;  // frames of one batch (their bit parse is latency-bound: fewer, larger batches are faster)
constexpr uint32_t kMaxBatchFrames = 65536;

struct DSlot {
    cudaStream_t stream = nullptr;
    void *d_work = nullptr;
    size_t work_cap = 0;
    uint8_t *d_pcm = nullptr;  // batch output when the caller's buffer is host memory
    size_t pcm_cap = 0;
    FrameRec *h_rec = nullptr;  // pinned
    size_t rec_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t frames = 0, first = 0;
    bool busy = false;
};

template <typename T>
int grow(T *&p, size_t &cap, size_t need) {
    if (need <= cap) return ZF_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = need + need / 8 + 256;
    ZFD_CUDA(cudaMalloc(&p, want * sizeof(T)));
    cap = want;
    return ZF_OK;
}

}  // namespace

struct zf_decoder {
    int device = 0;
    int sm_count = 148;
    zf::dec::CrcPowers crc_pows;  // x^(8 * 2^i) mod x^16 + x^15 + x^2 + 1
    uint32_t warps_per_sm = 12;  // frames kernel: warps per SM before the lanes of a warp are filled (ZF_DEC_WARPS_PER_SM)
    cudaStream_t stream = nullptr;  // upload, scan, tables
    cudaEvent_t ev_ready = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
    uint8_t *d_stream = nullptr;
    size_t stream_cap = 0;
    Cand *d_cand = nullptr;
    size_t cand_cap = 0;
    uint32_t *d_count = nullptr;
    uint32_t *h_count = nullptr;  // pinned
    Cand *h_cand = nullptr;       // pinned: the scan kernel's hits
    size_t h_cand_cap = 0;
    unsigned long long *h_tab = nullptr;  // pinned: frame starts (n + 1), then first samples (n)
    size_t h_tab_cap = 0;
    unsigned long long *d_fpos = nullptr, *d_first = nullptr;
    size_t fpos_cap = 0, first_cap = 0;
    FrameRec *d_rec = nullptr;
    size_t rec_cap = 0;
    DSlot slot[kSlots];
};

namespace {

void fill_info(zf_decode_info *info, const zf::dec::HostStreamInfo &si) {
    if (!info) return;
    info->sample_rate = si.sample_rate;
    info->channels = si.channels;
    info->bit_depth = si.bits;
    info->min_block_size = si.min_block;
    info->max_block_size = si.max_block;
    info->streaminfo_samples = si.total_samples;
    info->md5_status = -1;
    memcpy(info->md5, si.md5, 16);
}

int metadata_status(int rc) {
    return rc == 0 ? ZF_OK : rc == -2 ? ZF_ERR_FLAC_TRUNCATED : ZF_ERR_FLAC_NOT_FLAC;
}

// waits for a slot's batch and looks at its frame records
int finish_slot(zf_decoder *d, DSlot &sl, zf_decode_info *info, float &kernel_ms, int &status) {
    if (!sl.busy) return ZF_OK;
    ZFD_CUDA(cudaStreamSynchronize(sl.stream));
    sl.busy = false;
    float ms = 0;
    ZFD_CUDA(cudaEventElapsedTime(&ms, sl.ev0, sl.ev1));
    kernel_ms += ms;
    if (status == ZF_OK) {
        for (uint32_t i = 0; i < sl.frames; i++) {
            if (sl.h_rec[i].status != zf::dec::kOk) {
                status = ZF_ERR_FLAC_FRAME;
                if (info) {
                    info->bad_frame = (uint64_t)sl.first + i;
                    info->bad_frame_status = sl.h_rec[i].status;
                }
                break;
            }
        }
    }
    (void)d;
    return ZF_OK;
}

template <typename ST>
void launch_batch(zf_decoder *d, DSlot &sl, uint32_t first, uint32_t nb, const StreamParams &sp, uint8_t *pcm_base,
                  unsigned long long pcm_cap) {
    // lanes per warp: enough warps for a few per scheduler first (zf_dec_frames_kernel)
    uint32_t lpw = 1;
    while (lpw < 32u && (nb + lpw - 1u) / lpw > (uint32_t)d->sm_count * d->warps_per_sm) lpw *= 2u;
    zf::dec::zf_dec_frames_kernel<ST><<<(nb + lpw - 1u) / lpw, 32, 0, sl.stream>>>(d->d_stream, d->d_fpos + first, nb, lpw, sp,
                                                                                 static_cast<ST *>(sl.d_work), d->d_rec + first);
    zf::dec::zf_dec_crc16_kernel<<<(nb + 3u) / 4u, 128, 0, sl.stream>>>(d->d_stream, d->d_fpos + first, nb, d->crc_pows, d->d_rec + first);
    zf::dec::zf_dec_output_kernel<ST><<<nb, 256, 0, sl.stream>>>(static_cast<const ST *>(sl.d_work), d->d_rec + first,
                                                                   d->d_first + first, sp, pcm_base, pcm_cap);
}

// flac_host XOR d_flac; out_host XOR d_out (both NULL: sizes only).
int decode_core(zf_decoder *d, const uint8_t *flac_host, const uint8_t *d_flac, size_t len, uint8_t *out_host, uint8_t *d_out,
                size_t out_cap, size_t *out_len, uint32_t flags, zf_decode_info *info) {
    if (!d || (!flac_host && !d_flac)) return ZF_ERR_INVALID_ARG;
    zf_decode_info local;
    memset(&local, 0, sizeof local);
    local.struct_size = sizeof local;
    local.md5_status = -1;
    if (out_len) *out_len = 0;
    ZFD_CUDA(cudaSetDevice(d->device));

    // ---- metadata (host) ----
    zf::dec::HostStreamInfo si;
    memset(&si, 0, sizeof si);
    std::vector<uint8_t> head;
    if (flac_host) {
        const int rc = metadata_status(zf::dec::parse_metadata(flac_host, len, si));
        if (rc) return rc;
    } else {
        size_t take = std::min<size_t>(len, 4096);  // metadata is usually a few hundred bytes; more is fetched when it is not
        for (;;) {
            head.resize(take);
            ZFD_CUDA(cudaMemcpy(head.data(), d_flac, take, cudaMemcpyDeviceToHost));
            const int prc = zf::dec::parse_metadata(head.data(), take, si);
            if (prc == -2 && take < len) {
                take = std::min<size_t>(len, take * 8);
                continue;
            }
            const int rc = metadata_status(prc);
            if (rc) return rc;
            break;
        }
    }
    fill_info(&local, si);
    auto publish = [&]() {
        if (info) {
            const uint32_t sz = info->struct_size ? std::min<uint32_t>(info->struct_size, sizeof local) : sizeof local;
            memcpy(info, &local, sz);
            info->struct_size = sz;
        }
    };
    if (!(si.bits == 8 || si.bits == 16 || si.bits == 24 || si.bits == 32) || si.channels < 1 || si.channels > 8) {
        publish();
        return ZF_ERR_UNSUPPORTED;
    }
    if (si.first_frame_offset >= len) {  // a stream without frames
        publish();
        return si.total_samples ? ZF_ERR_FLAC_COUNT : ZF_OK;
    }
    StreamParams sp;
    sp.channels = si.channels;
    sp.bits = si.bits;
    sp.max_block = si.max_block ? si.max_block : 65535u;
    sp.sample_rate = si.sample_rate;

    // ---- stream to the device, zero bytes behind it (the bit reader looks ahead) ----
    if (grow(d->d_stream, d->stream_cap, len + zf::dec::kStreamPad)) return ZF_ERR_CUDA;
    ZFD_CUDA(cudaMemcpyAsync(d->d_stream, flac_host ? flac_host : d_flac, len, flac_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                             d->stream));
    ZFD_CUDA(cudaMemsetAsync(d->d_stream + len, 0, zf::dec::kStreamPad, d->stream));

    // ---- header scan ----
    const uint64_t expect_frames = si.total_samples && si.min_block ? si.total_samples / si.min_block + 2 : len / 64 + 2;
    const size_t cand_cap = (size_t)std::min<uint64_t>(std::max<uint64_t>(expect_frames * 2 + 4096, 8192), len / 6 + 16);
    if (grow(d->d_cand, d->cand_cap, cand_cap)) return ZF_ERR_CUDA;
    ZFD_CUDA(cudaMemsetAsync(d->d_count, 0, sizeof(uint32_t), d->stream));
    const unsigned long long begin = si.first_frame_offset;
    const unsigned long long words = (len + 3) / 4 - (begin >> 2);
    ZFD_CUDA(cudaEventRecord(d->ev_s0, d->stream));
    zf::dec::zf_dec_scan_kernel<<<(unsigned)((words + 255) / 256), 256, 0, d->stream>>>(d->d_stream, begin, len, sp, d->d_cand,
                                                                                       (uint32_t)cand_cap, d->d_count);
    ZFD_CUDA(cudaEventRecord(d->ev_s1, d->stream));
    local.launches = 1;
    // the hit count and (in the same breath: one round trip) the first hits; page-locked staging kept in the handle
    const size_t first_take = std::min<size_t>(cand_cap, 32768);
    if (d->h_cand_cap < cand_cap) {
        if (d->h_cand) cudaFreeHost(d->h_cand);
        d->h_cand = nullptr;
        d->h_cand_cap = 0;
        ZFD_CUDA(cudaMallocHost(&d->h_cand, (cand_cap + cand_cap / 8 + 256) * sizeof(Cand)));
        d->h_cand_cap = cand_cap + cand_cap / 8 + 256;
    }
    ZFD_CUDA(cudaMemcpyAsync(d->h_count, d->d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
    ZFD_CUDA(cudaMemcpyAsync(d->h_cand, d->d_cand, first_take * sizeof(Cand), cudaMemcpyDeviceToHost, d->stream));
    ZFD_CUDA(cudaStreamSynchronize(d->stream));
    ZFD_CUDA(cudaGetLastError());
    const uint32_t n_cand = *d->h_count;
    if (n_cand > cand_cap) {
        zf_internal_set_error("decoder: more frame-header candidates than the stream can hold frames");
        publish();
        return ZF_ERR_FLAC_FRAME;
    }
    if (n_cand > first_take) {
        ZFD_CUDA(cudaMemcpyAsync(d->h_cand + first_take, d->d_cand + first_take, (n_cand - first_take) * sizeof(Cand),
                                 cudaMemcpyDeviceToHost, d->stream));
        ZFD_CUDA(cudaStreamSynchronize(d->stream));
    }
    const Cand *cand_raw = d->h_cand;
    std::vector<zf::dec::HostCand> cand(n_cand);
    for (uint32_t i = 0; i < n_cand; i++) {
        cand[i].pos = cand_raw[i].pos;
        cand[i].number = cand_raw[i].number;
        cand[i].block_size = cand_raw[i].block_size;
        cand[i].variable = cand_raw[i].variable;
    }
    float kernel_ms = 0;
    {
        float ms = 0;
        ZFD_CUDA(cudaEventElapsedTime(&ms, d->ev_s0, d->ev_s1));
        kernel_ms += ms;
    }
    std::vector<uint64_t> fpos, first_sample;
    uint64_t total = 0;
    int status = ZF_OK;
    for (int attempt = 0;; attempt++) {
    status = ZF_OK;
    local.bad_frame = 0;
    local.bad_frame_status = 0;
    const int crc = zf::dec::chain_frames(cand, si.first_frame_offset, len, fpos, first_sample, total);
    if (crc == -2) { publish(); return ZF_ERR_UNSUPPORTED; }
    if (crc) { publish(); return ZF_ERR_FLAC_TRUNCATED; }
    const uint64_t n_frames = first_sample.size();
    if (n_frames > 0xfffffff0ull) { publish(); return ZF_ERR_UNSUPPORTED; }
    const uint32_t bytes = si.bits / 8u;
    local.n_frames = n_frames;
    local.samples_per_channel = total;
    local.pcm_bytes = total * si.channels * bytes;
    if (out_len) *out_len = (size_t)local.pcm_bytes;
    if (!out_host && !d_out) {
        publish();
        return ZF_OK;
    }
    if (out_cap < local.pcm_bytes) {
        // a header image chained as a frame can also inflate the total: only the sizes-only call (above) trusts it
        publish();
        return ZF_ERR_OUT_TOO_SMALL;
    }

    // ---- frame table to the device ----
    if (grow(d->d_fpos, d->fpos_cap, (size_t)n_frames + 1) || grow(d->d_first, d->first_cap, (size_t)n_frames) ||
        grow(d->d_rec, d->rec_cap, (size_t)n_frames))
        return ZF_ERR_CUDA;
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "");
    ZFD_CUDA(cudaStreamSynchronize(d->stream));  // (a retry: the staging below may still be in flight)
    if (d->h_tab_cap < 2 * n_frames + 1) {
        if (d->h_tab) cudaFreeHost(d->h_tab);
        d->h_tab = nullptr;
        d->h_tab_cap = 0;
        const size_t want = 2 * (size_t)n_frames + 1 + n_frames / 4 + 256;
        ZFD_CUDA(cudaMallocHost(&d->h_tab, want * 8));
        d->h_tab_cap = want;
    }
    memcpy(d->h_tab, fpos.data(), (n_frames + 1) * 8);
    memcpy(d->h_tab + n_frames + 1, first_sample.data(), n_frames * 8);
    ZFD_CUDA(cudaMemcpyAsync(d->d_fpos, d->h_tab, (n_frames + 1) * 8, cudaMemcpyHostToDevice, d->stream));
    ZFD_CUDA(cudaMemcpyAsync(d->d_first, d->h_tab + n_frames + 1, n_frames * 8, cudaMemcpyHostToDevice, d->stream));
    ZFD_CUDA(cudaEventRecord(d->ev_ready, d->stream));  // the batches' streams wait for it; the host does not

    // ---- batches ----
    const bool wide = si.bits == 32;
    const size_t plane = (size_t)si.channels * (((size_t)sp.max_block + 7) & ~(size_t)7) * (wide ? 8 : 4);  // zf::dec::plane_stride
    uint32_t per_batch = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(kWorkBytesPerSlot / plane, 32), kMaxBatchFrames);
    if (n_frames <= per_batch) per_batch = (uint32_t)n_frames;                                   // one batch
    else if (n_frames < 2ull * per_batch) per_batch = (uint32_t)((n_frames + 1) / 2);          // two even ones
    int which = 0;
    for (uint64_t f0 = 0; f0 < n_frames && status == ZF_OK; f0 += per_batch, which ^= 1) {
        DSlot &sl = d->slot[which];
        int rc = finish_slot(d, sl, &local, kernel_ms, status);
        if (rc) return rc;
        if (status != ZF_OK) break;
        const uint32_t nb = (uint32_t)std::min<uint64_t>(per_batch, n_frames - f0);
        const uint64_t s0 = first_sample[f0], s1 = f0 + nb < n_frames ? first_sample[f0 + nb] : total;
        const uint64_t byte0 = s0 * si.channels * bytes, nbytes = (s1 - s0) * si.channels * bytes;
        {
            uint8_t *w = static_cast<uint8_t *>(sl.d_work);
            if (grow(w, sl.work_cap, plane * nb)) return ZF_ERR_CUDA;
            sl.d_work = w;
        }
        if (nb > sl.rec_cap) {
            if (sl.h_rec) cudaFreeHost(sl.h_rec);
            sl.h_rec = nullptr;
            sl.rec_cap = 0;
            ZFD_CUDA(cudaMallocHost(&sl.h_rec, (size_t)nb * sizeof(FrameRec)));
            sl.rec_cap = nb;
        }
        uint8_t *base;  // where sample 0 of the STREAM would lie
        if (d_out) {
            base = d_out;
        } else {
            if (grow(sl.d_pcm, sl.pcm_cap, (size_t)nbytes + 16)) return ZF_ERR_CUDA;
            base = sl.d_pcm - byte0;
        }
        ZFD_CUDA(cudaStreamWaitEvent(sl.stream, d->ev_ready, 0));
        ZFD_CUDA(cudaEventRecord(sl.ev0, sl.stream));
        if (wide) launch_batch<long long>(d, sl, (uint32_t)f0, nb, sp, base, byte0 + nbytes);
        else launch_batch<int32_t>(d, sl, (uint32_t)f0, nb, sp, base, byte0 + nbytes);
        ZFD_CUDA(cudaEventRecord(sl.ev1, sl.stream));
        ZFD_CUDA(cudaGetLastError());
        local.launches += 3;
        ZFD_CUDA(cudaMemcpyAsync(sl.h_rec, d->d_rec + f0, (size_t)nb * sizeof(FrameRec), cudaMemcpyDeviceToHost, sl.stream));
        if (out_host && nbytes) ZFD_CUDA(cudaMemcpyAsync(out_host + byte0, sl.d_pcm, nbytes, cudaMemcpyDeviceToHost, sl.stream));
        sl.frames = nb;
        sl.first = (uint32_t)f0;
        sl.busy = true;
    }
    for (int s = 0; s < kSlots; s++) {
        // in submission order: the older batch first
        DSlot &sl = d->slot[(which + s) & 1];
        const int rc = finish_slot(d, sl, &local, kernel_ms, status);
        if (rc) return rc;
    }
    if (status == ZF_ERR_FLAC_FRAME && attempt < 8 && zf::dec::drop_suspect(cand, fpos, local.bad_frame)) continue;
    break;
    }  // attempts
    local.kernel_ms = kernel_ms;
    if (status == ZF_OK && si.total_samples && si.total_samples != total) status = ZF_ERR_FLAC_COUNT;
    if (status == ZF_OK && out_host && (flags & (ZF_DECODE_CHECK_MD5 | ZF_DECODE_REQUIRE_MD5))) {
        static const uint8_t zero[16] = {0};
        if (memcmp(si.md5, zero, 16) != 0) {
            zf_md5 m;
            uint8_t digest[16];
            zf_md5_init(&m);
            zf_md5_update(&m, out_host, (size_t)local.pcm_bytes);
            zf_md5_final(&m, digest);
            local.md5_status = memcmp(digest, si.md5, 16) == 0 ? 1 : 0;
            if (!local.md5_status && (flags & ZF_DECODE_REQUIRE_MD5)) status = ZF_ERR_FLAC_MD5;
        }
    }
    publish();
    return status;
}

}  // namespace

extern "C" {

int zf_decoder_create(int device_id, zf_decoder **out) {
    if (!out) return ZF_ERR_INVALID_ARG;
    *out = nullptr;
    const int rc = zf_device_check(device_id);
    if (rc) return rc;
    ZFD_CUDA(cudaSetDevice(device_id));
    zf_decoder *d = new (std::nothrow) zf_decoder();
    if (!d) return ZF_ERR_NOMEM;
    d->device = device_id;
    cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device_id);
    {
        auto mulmod = [](uint32_t a, uint32_t b) {
            uint32_t r = 0;
            for (int i = 15; i >= 0; i--) {
                r = (r & 0x8000u) ? ((r << 1) ^ 0x8005u) & 0xffffu : (r << 1);
                if ((b >> i) & 1u) r ^= a;
            }
            return r;
        };
        uint32_t v = 0x100u;  // x^8
        for (int i = 0; i < 32; i++) {
            d->crc_pows.pw[i] = (uint16_t)v;
            v = mulmod(v, v);
        }
    }
    if (const char *e = getenv("ZF_DEC_WARPS_PER_SM")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 64) d->warps_per_sm = (uint32_t)v;
    }
    auto fail = [&](int code) {
        zf_decoder_destroy(d);
        return code;
    };
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&d->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&d->ev_s0) != cudaSuccess || cudaEventCreate(&d->ev_s1) != cudaSuccess ||
        cudaMalloc(&d->d_count, 64) != cudaSuccess || cudaMallocHost(&d->h_count, 64) != cudaSuccess)
        return fail(ZF_ERR_CUDA);
    for (int s = 0; s < kSlots; s++) {
        DSlot &sl = d->slot[s];
        if (cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&sl.ev0) != cudaSuccess ||
            cudaEventCreate(&sl.ev1) != cudaSuccess)
            return fail(ZF_ERR_CUDA);
    }
    *out = d;
    return ZF_OK;
}

void zf_decoder_destroy(zf_decoder *d) {
    if (!d) return;
    cudaSetDevice(d->device);
    for (int s = 0; s < kSlots; s++) {
        DSlot &sl = d->slot[s];
        if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
        if (sl.d_work) cudaFree(sl.d_work);
        if (sl.d_pcm) cudaFree(sl.d_pcm);
        if (sl.h_rec) cudaFreeHost(sl.h_rec);
        if (sl.ev0) cudaEventDestroy(sl.ev0);
        if (sl.ev1) cudaEventDestroy(sl.ev1);
    }
    if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
    if (d->ev_ready) cudaEventDestroy(d->ev_ready);
    if (d->ev_s0) cudaEventDestroy(d->ev_s0);
    if (d->ev_s1) cudaEventDestroy(d->ev_s1);
    if (d->d_stream) cudaFree(d->d_stream);
    if (d->d_cand) cudaFree(d->d_cand);
    if (d->d_count) cudaFree(d->d_count);
    if (d->h_count) cudaFreeHost(d->h_count);
    if (d->h_cand) cudaFreeHost(d->h_cand);
    if (d->h_tab) cudaFreeHost(d->h_tab);
    if (d->d_fpos) cudaFree(d->d_fpos);
    if (d->d_first) cudaFree(d->d_first);
    if (d->d_rec) cudaFree(d->d_rec);
    cudaGetLastError();
    delete d;
}

int zf_flac_stream_info(const uint8_t *flac, size_t len, zf_decode_info *info) {
    if (!flac || !info) return ZF_ERR_INVALID_ARG;
    zf::dec::HostStreamInfo si;
    memset(&si, 0, sizeof si);
    const int rc = metadata_status(zf::dec::parse_metadata(flac, len, si));
    if (rc) return rc;
    zf_decode_info local;
    memset(&local, 0, sizeof local);
    fill_info(&local, si);
    local.samples_per_channel = si.total_samples;
    local.pcm_bytes = si.total_samples * si.channels * (si.bits / 8u);
    const uint32_t sz = info->struct_size ? std::min<uint32_t>(info->struct_size, sizeof local) : sizeof local;
    local.struct_size = sz;
    memcpy(info, &local, sz);
    return ZF_OK;
}

int zf_decode_flac(zf_decoder *dec, const uint8_t *flac, size_t len, uint8_t *pcm, size_t pcm_cap, size_t *pcm_len, uint32_t flags,
                   zf_decode_info *info) {
    if (!dec || !flac) return ZF_ERR_INVALID_ARG;
    return decode_core(dec, flac, nullptr, len, pcm, nullptr, pcm_cap, pcm_len, flags, info);
}

int zf_decode_flac_device(zf_decoder *dec, const void *d_flac, size_t len, void *d_pcm, size_t pcm_cap, size_t *pcm_len,
                          zf_decode_info *info) {
    if (!dec || !d_flac) return ZF_ERR_INVALID_ARG;
    return decode_core(dec, nullptr, static_cast<const uint8_t *>(d_flac), len, nullptr, static_cast<uint8_t *>(d_pcm), pcm_cap,
                       pcm_len, 0, info);
}

int zf_decode_flac_memory(const uint8_t *flac, size_t len, int device_id, uint32_t flags, uint8_t **pcm, size_t *pcm_len,
                          zf_decode_info *info) {
    if (!flac || !pcm || !pcm_len) return ZF_ERR_INVALID_ARG;
    *pcm = nullptr;
    *pcm_len = 0;
    zf_decoder *d = nullptr;
    int rc = zf_decoder_create(device_id, &d);
    if (rc) return rc;
    zf_decode_info local;
    memset(&local, 0, sizeof local);
    local.struct_size = sizeof local;
    size_t need = 0;
    rc = decode_core(d, flac, nullptr, len, nullptr, nullptr, 0, &need, 0, &local);  // sizes: metadata + scan
    uint8_t *buf = nullptr;
    if (rc == ZF_OK) {
        buf = static_cast<uint8_t *>(malloc(need ? need : 1));
        if (!buf) rc = ZF_ERR_NOMEM;
    }
    if (rc == ZF_OK) rc = decode_core(d, flac, nullptr, len, buf, nullptr, need, pcm_len, flags, &local);
    zf_decoder_destroy(d);
    if (info) {
        const uint32_t sz = info->struct_size ? std::min<uint32_t>(info->struct_size, sizeof local) : sizeof local;
        memcpy(info, &local, sz);
        info->struct_size = sz;
    }
    if (rc) {
        free(buf);
        return rc;
    }
    *pcm = buf;
    return ZF_OK;
}

int zf_decode_flac_file(const char *in_path, const char *out_path, int device_id, uint32_t flags) {
    if (!in_path || !out_path) return ZF_ERR_INVALID_ARG;
    FILE *f = fopen(in_path, "rb");
    if (!f) return ZF_ERR_IO;
    std::vector<uint8_t> flac;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        if (sz > 0) flac.resize((size_t)sz);
        fseek(f, 0, SEEK_SET);
    }
    const size_t got = flac.empty() ? 0 : fread(flac.data(), 1, flac.size(), f);
    fclose(f);
    if (got != flac.size()) return ZF_ERR_IO;
    uint8_t *pcm = nullptr;
    size_t pcm_len = 0;
    zf_decode_info info;
    memset(&info, 0, sizeof info);
    info.struct_size = sizeof info;
    const int rc = zf_decode_flac_memory(flac.data(), flac.size(), device_id, flags, &pcm, &pcm_len, &info);
    if (rc) return rc;
    if (info.bit_depth == 8)
        for (size_t i = 0; i < pcm_len; i++) pcm[i] = (uint8_t)(pcm[i] + 128u);  // WAV stores 8-bit samples unsigned
    FILE *o = fopen(out_path, "wb");
    if (!o) { free(pcm); return ZF_ERR_IO; }
    uint8_t hdr[44];
    zf::dec::wav_header(hdr, info.channels, info.bit_depth, info.sample_rate, pcm_len);
    const bool ok = fwrite(hdr, 1, 44, o) == 44 && (pcm_len == 0 || fwrite(pcm, 1, pcm_len, o) == pcm_len);
    const bool closed = fclose(o) == 0;
    free(pcm);
    return ok && closed ? ZF_OK : ZF_ERR_IO;
}

static int read_file(const char *path, std::vector<uint8_t> &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return ZF_ERR_IO;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        if (sz > 0) out.resize((size_t)sz);
        fseek(f, 0, SEEK_SET);
    }
    const size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
    fclose(f);
    return got == out.size() ? ZF_OK : ZF_ERR_IO;
}

int zf_verify_flac_file(const char *wav_path, const char *flac_path, int device_id) {
    if (!wav_path || !flac_path) return ZF_ERR_INVALID_ARG;
    std::vector<uint8_t> wav, flac;
    int rc = read_file(wav_path, wav);
    if (rc == ZF_OK) rc = read_file(flac_path, flac);
    if (rc) return rc;
    zf_wav_format fmt;
    rc = zf_wav_parse(wav.data(), wav.size(), &fmt);
    if (rc) return rc;
    uint8_t *pcm = nullptr;
    size_t pcm_len = 0;
    zf_decode_info info;
    memset(&info, 0, sizeof info);
    info.struct_size = sizeof info;
    // 8-bit streams carry the MD5 of the raw file bytes (upstream's reader): compared below instead
    rc = zf_decode_flac_memory(flac.data(), flac.size(), device_id, fmt.bit_depth == 8 ? 0u : ZF_DECODE_REQUIRE_MD5, &pcm, &pcm_len, &info);
    if (rc) return rc;
    const size_t want = (size_t)fmt.samples_count * fmt.channels * fmt.bytes_per_sample;
    const uint8_t *data = wav.data() + fmt.data_offset;
    bool same = pcm_len == want && info.channels == fmt.channels && info.bit_depth == fmt.bit_depth;
    if (same && fmt.bit_depth == 8) {
        std::vector<uint8_t> state((size_t)info.max_block_size * fmt.channels);
        std::vector<int8_t> conv(want);
        zf_wav8_state_init(state.data(), state.size());
        zf_wav8_to_samples(data, fmt.samples_count, fmt.channels, info.max_block_size, 0, state.data(), conv.data());
        same = memcmp(conv.data(), pcm, want) == 0;
    } else if (same) {
        same = memcmp(data, pcm, want) == 0;
    }
    free(pcm);
    return same ? ZF_OK : ZF_ERR_FLAC_MD5;
}

}  // extern "C"

// zf_driver.cpp -- whole-file driver above the C ABI: the host side of `flac in.wav out.flac`.
//
// Mirrors src/cli/wav2flac.zig:10-97 (main + encode): parse the WAV, reserve the 42-byte header, write the vendor block,
// read the samples buffer by buffer (wav2flac.zig:66-97 reads `buffer_size` frames at a time; here a chunk is up to 2048
// frames), encode, then seek back and patch STREAMINFO with MD5, min/max frame size and sample count.
//
// What changes against the reference: the chunks flow through a pipeline of host threads --
//   reader   fills page-locked chunk buffers from the file (or from memory); one-byte samples are converted here
//            (wav_reader.zig:71-88, serial by nature)
//   md5      hashes the raw bytes of every chunk in stream order (wav_reader.zig:66; serial, so it has a thread of its own)
//   encoder  one per GPU: chunk k goes to device k mod G, each with its own handle (frames are independent: no
//            inter-GPU communication)
//   writer   (the calling thread) appends the chunks' frames in order and replays StreamInfo.updateFrameSize frame by
//            frame, because that function is order dependent (metadata.zig:35-40)
// so that file I/O, the hash and the GPUs overlap, and memory use is a few chunks, not the file.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/zigflac_b200.h"

extern "C" {
int zf_md5x_init(void *ctx, int use_openssl);
void zf_md5x_update(void *ctx, int ossl, const uint8_t *data, size_t len);
void zf_md5x_final(void *ctx, int ossl, uint8_t digest[16]);
}

namespace {
constexpr uint16_t kFrameSize = 4096;  // option.frame_size, build.zig:13
constexpr size_t kPrefix = 42 + 31;    // "fLaC" + STREAMINFO block + VORBIS_COMMENT block
constexpr uint32_t kChunkFrames = 2048;

// where the PCM comes from: a FILE positioned at the first sample, or memory
struct Source {
    FILE *file = nullptr;
    const uint8_t *mem = nullptr;
    size_t mem_left = 0;
    // reads up to n bytes; returns the count (short at the end of the stream)
    size_t read(uint8_t *dst, size_t n) {
        if (file) return fread(dst, 1, n, file);
        const size_t k = std::min(n, mem_left);
        memcpy(dst, mem, k);
        mem += k;
        mem_left -= k;
        return k;
    }
};

// where the FLAC goes: a FILE (seekable) or a growing malloc'd buffer
struct Sink {
    FILE *file = nullptr;
    uint8_t *buf = nullptr;
    size_t len = 0, cap = 0;
    bool append(const uint8_t *p, size_t n) {
        if (file) return fwrite(p, 1, n, file) == n;
        if (len + n > cap) {
            size_t nc = std::max(cap * 2, len + n + (1u << 20));
            uint8_t *nb = (uint8_t *)realloc(buf, nc);
            if (!nb) return false;
            buf = nb;
            cap = nc;
        }
        memcpy(buf + len, p, n);
        len += n;
        return true;
    }
    bool patch_front(const uint8_t *p, size_t n) {  // wav2flac.zig:60-62: seekTo(0), writeHeader
        if (file) return fseek(file, 0, SEEK_SET) == 0 && fwrite(p, 1, n, file) == n;
        memcpy(buf, p, n);
        return true;
    }
};

// Encoder handles and page-locked chunk buffers are kept between whole-file calls of one process: creating a handle
// (device buffers, streams, kernel attributes) and page-locking a few hundred MB cost far more than encoding a short
// file.  Whole-file calls are serialised on `call`; zf_driver_release_cache() gives everything back.
struct Cache {
    std::mutex call, mu;
    struct Enc { zf_config cfg; zf_encoder *h; bool used; };
    struct Buf { uint8_t *p; size_t cap; bool used; };
    std::vector<Enc> encs;
    std::vector<Buf> bufs;

    int encoder(const zf_config &cfg, zf_encoder **out) {
        {
            std::lock_guard<std::mutex> lk(mu);
            for (Enc &e : encs)
                if (!e.used && memcmp(&e.cfg, &cfg, sizeof cfg) == 0) { e.used = true; *out = e.h; return ZF_OK; }
        }
        zf_encoder *h = nullptr;
        const int rc = zf_encoder_create(&cfg, &h);
        if (rc) return rc;
        std::lock_guard<std::mutex> lk(mu);
        if (encs.size() >= 16) {  // bounded: drop an idle one
            for (size_t i = 0; i < encs.size(); i++)
                if (!encs[i].used) { zf_encoder_destroy(encs[i].h); encs.erase(encs.begin() + (long)i); break; }
        }
        encs.push_back({cfg, h, true});
        *out = h;
        return ZF_OK;
    }
    void put_encoder(zf_encoder *h) {
        std::lock_guard<std::mutex> lk(mu);
        for (Enc &e : encs)
            if (e.h == h) e.used = false;
    }
    int buffer(size_t bytes, uint8_t **out) {
        {
            std::lock_guard<std::mutex> lk(mu);
            Buf *best = nullptr;
            for (Buf &b : bufs)
                if (!b.used && b.cap >= bytes && (!best || b.cap < best->cap)) best = &b;
            if (best) { best->used = true; *out = best->p; return ZF_OK; }
        }
        void *p = nullptr;
        const int rc = zf_host_alloc(bytes, &p);
        if (rc) return rc;
        std::lock_guard<std::mutex> lk(mu);
        bufs.push_back({(uint8_t *)p, bytes, true});
        *out = (uint8_t *)p;
        return ZF_OK;
    }
    void put_buffer(uint8_t *p) {
        std::lock_guard<std::mutex> lk(mu);
        for (Buf &b : bufs)
            if (b.p == p) b.used = false;
    }
    void release() {
        std::lock_guard<std::mutex> lk(mu);
        for (Enc &e : encs) zf_encoder_destroy(e.h);
        for (Buf &b : bufs) zf_host_free(b.p);
        encs.clear();
        bufs.clear();
    }
};
Cache g_cache;

struct Chunk {
    uint8_t *pcm = nullptr;  // page-locked
    uint8_t *out = nullptr;  // page-locked
    size_t bytes = 0, out_len = 0;
    uint64_t first_frame = 0, samples = 0;
    uint32_t frames = 0;
    std::vector<uint32_t> sizes;
    long long index = -1;  // which chunk of the stream the slot holds
    bool filled = false, hashed = false, encoded = false;
};

struct Pipe {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Chunk> ring;
    int rc = ZF_OK;          // first failure; everybody stops
    long long n_chunks = -1; // known once the reader has hit the end of the stream
    uint64_t samples_read = 0;
    bool incomplete = false;

    void fail(int code) {
        std::lock_guard<std::mutex> lk(mu);
        if (rc == ZF_OK) rc = code;
        cv.notify_all();
    }
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int encode_stream(Source &src, Sink &sink, const zf_wav_format &fmt, uint64_t want_samples, const int *devices, int n_devices) {
    std::lock_guard<std::mutex> one_call(g_cache.call);
    const bool trace = getenv("ZF_TRACE") != nullptr;  // development aid: where the wall time of a whole-file call goes
    const double t_begin = now_s();
    double t_read = 0, t_hash = 0, t_enc = 0, t_write = 0, t_reader_done = 0, t_hasher_done = 0;
    const size_t ic_bytes = (size_t)fmt.channels * fmt.bytes_per_sample;
    const uint64_t want_frames = (want_samples + kFrameSize - 1) / kFrameSize;
    int one = 0;
    if (n_devices <= 0 || !devices) { devices = &one; n_devices = 1; }
    // chunk size: up to 2048 frames, smaller when the stream is short so that every device gets work and small files
    // do not pay for large page-locked buffers
    uint32_t chunk_frames = (uint32_t)std::min<uint64_t>(kChunkFrames, std::max<uint64_t>(1, (want_frames + n_devices - 1) / n_devices));
    const long long max_chunks = (long long)((want_frames + chunk_frames - 1) / chunk_frames);
    const int ring_n = (int)std::min<long long>(std::max<long long>(max_chunks, 1), 2LL * n_devices + 2);

    zf_config cfg;
    zf_config_default(&cfg, (uint8_t)fmt.channels, (uint8_t)fmt.bit_depth, fmt.sample_rate);  // wav2flac.zig:38-42
    cfg.max_frames_per_batch = std::min<uint32_t>(chunk_frames, 2048);
    const size_t chunk_bytes = (size_t)chunk_frames * kFrameSize * ic_bytes;
    const size_t out_cap = zf_max_batch_bytes(&cfg, chunk_frames);

    Pipe pipe;
    pipe.ring.resize((size_t)ring_n);
    auto release = [&] {
        for (Chunk &c : pipe.ring) { if (c.pcm) g_cache.put_buffer(c.pcm); if (c.out) g_cache.put_buffer(c.out); }
    };
    // the ring's page-locked buffers are taken (from the cache, or page-locked now) by the reader thread as it first fills
    // each slot: page-locking runs at a GB/s or two, and this way it overlaps the MD5 of the chunks already read
    for (Chunk &c : pipe.ring) c.sizes.resize(chunk_frames);

    // ---- reader (+ MD5 of one-byte streams: their raw bytes are converted in place, so they are hashed first) ----
    const bool one_byte = fmt.bytes_per_sample == 1;
    const char *md5_env = getenv("ZF_MD5");
    alignas(16) unsigned char md5_ctx[128];
    const int ossl = zf_md5x_init(md5_ctx, md5_env && !strcmp(md5_env, "openssl"));
    std::thread reader([&] {
        std::vector<uint8_t> state;
        if (one_byte) { state.resize((size_t)kFrameSize * fmt.channels); zf_wav8_state_init(state.data(), state.size()); }
        uint64_t left = want_samples, pos = 0;
        for (long long k = 0;; k++) {
            Chunk &c = pipe.ring[(size_t)(k % ring_n)];
            {
                std::unique_lock<std::mutex> lk(pipe.mu);
                pipe.cv.wait(lk, [&] { return pipe.rc != ZF_OK || c.index < 0; });
                if (pipe.rc != ZF_OK) return;
            }
            if (!c.pcm) {
                int rc = g_cache.buffer(chunk_bytes, &c.pcm);
                if (!rc) rc = g_cache.buffer(out_cap, &c.out);
                if (rc) { pipe.fail(rc); return; }
            }
            const uint64_t ask = std::min<uint64_t>(left, (uint64_t)chunk_frames * kFrameSize);
            const double tr0 = now_s();
            size_t got = ask ? src.read(c.pcm, (size_t)ask * ic_bytes) : 0;
            t_read += now_s() - tr0;
            if (got % ic_bytes) {  // StreamError.IncompleteStream, wav_reader.zig:52-53
                std::lock_guard<std::mutex> lk(pipe.mu);
                pipe.incomplete = true;
                got -= got % ic_bytes;
            }
            const uint64_t ns = got / ic_bytes;
            if (one_byte && ns) {
                zf_md5x_update(md5_ctx, ossl, c.pcm, got);
                zf_wav8_to_samples(c.pcm, ns, fmt.channels, kFrameSize, pos, state.data(), (int8_t *)c.pcm);
            }
            std::lock_guard<std::mutex> lk(pipe.mu);
            if (ns == 0) {  // fillSamples returned 0: end of the stream (wav2flac.zig:78-80)
                pipe.n_chunks = k;
                pipe.cv.notify_all();
                return;
            }
            c.index = k;
            c.bytes = got;
            c.samples = ns;
            c.first_frame = pos / kFrameSize;
            c.frames = (uint32_t)((ns + kFrameSize - 1) / kFrameSize);
            c.filled = true;
            c.hashed = one_byte;
            c.encoded = false;
            pos += ns;
            left -= ns;
            pipe.samples_read = pos;
            const bool last = ns < (uint64_t)chunk_frames * kFrameSize || left == 0;
            if (last) pipe.n_chunks = k + 1;
            pipe.cv.notify_all();
            if (last) return;
        }
    });

    // ---- MD5 over the raw bytes, in stream order (wav_reader.zig:66) ----
    std::thread hasher([&] {
        if (one_byte) return;
        for (long long k = 0;; k++) {
            Chunk &c = pipe.ring[(size_t)(k % ring_n)];
            {
                std::unique_lock<std::mutex> lk(pipe.mu);
                pipe.cv.wait(lk, [&] { return pipe.rc != ZF_OK || (c.index == k && c.filled) || (pipe.n_chunks >= 0 && k >= pipe.n_chunks); });
                if (pipe.rc != ZF_OK || (pipe.n_chunks >= 0 && k >= pipe.n_chunks)) return;
            }
            const double th0 = now_s();
            zf_md5x_update(md5_ctx, ossl, c.pcm, c.bytes);
            t_hash += now_s() - th0;
            t_hasher_done = now_s() - t_begin;
            std::lock_guard<std::mutex> lk(pipe.mu);
            c.hashed = true;
            pipe.cv.notify_all();
        }
    });

    // ---- one encoder thread per device: chunks g, g + G, g + 2 G, ... ----
    std::vector<std::thread> workers;
    for (int g = 0; g < n_devices; g++) {
        workers.emplace_back([&, g] {
            zf_config dc = cfg;
            dc.device_id = devices[g];
            zf_encoder *enc = nullptr;
            int rc = g_cache.encoder(dc, &enc);
            if (rc) { pipe.fail(rc); return; }
            for (long long k = g;; k += n_devices) {
                Chunk &c = pipe.ring[(size_t)(k % ring_n)];
                {
                    std::unique_lock<std::mutex> lk(pipe.mu);
                    pipe.cv.wait(lk, [&] { return pipe.rc != ZF_OK || (c.index == k && c.filled) || (pipe.n_chunks >= 0 && k >= pipe.n_chunks); });
                    if (pipe.rc != ZF_OK || (pipe.n_chunks >= 0 && k >= pipe.n_chunks)) break;
                }
                uint32_t nf = 0;
                const double te0 = now_s();
                rc = zf_encode_pcm(enc, c.pcm, c.samples, c.first_frame, c.out, out_cap, &c.out_len, c.sizes.data(),
                                   (uint32_t)c.sizes.size(), &nf);
                if (g == 0) t_enc += now_s() - te0;
                if (rc) { pipe.fail(rc); break; }
                std::lock_guard<std::mutex> lk(pipe.mu);
                c.encoded = true;
                pipe.cv.notify_all();
            }
            g_cache.put_encoder(enc);
        });
    }

    // ---- writer: this thread ----
    zf_streaminfo si;  // wav_reader.zig:95-109
    zf_streaminfo_init(&si);
    si.sample_rate = fmt.sample_rate;
    si.channels = (uint8_t)fmt.channels;
    si.bit_depth = (uint8_t)fmt.bit_depth;
    si.interchannel_samples = fmt.samples_count;  // from the header, even if the file is short (Q15)
    si.min_block_size = kFrameSize;
    si.max_block_size = kFrameSize;
    uint8_t prefix[kPrefix];
    memset(prefix, 0, sizeof prefix);           // skipHeader, encoder.zig:177-185
    zf_write_vorbis_comment(1, prefix + 42);    // wav2flac.zig:48
    int rc = sink.append(prefix, kPrefix) ? ZF_OK : ZF_ERR_IO;
    if (rc) pipe.fail(rc);
    for (long long k = 0; rc == ZF_OK; k++) {
        Chunk &c = pipe.ring[(size_t)(k % ring_n)];
        {
            std::unique_lock<std::mutex> lk(pipe.mu);
            pipe.cv.wait(lk, [&] { return pipe.rc != ZF_OK || (c.index == k && c.encoded && c.hashed) || (pipe.n_chunks >= 0 && k >= pipe.n_chunks); });
            if (pipe.rc != ZF_OK) { rc = pipe.rc; break; }
            if (pipe.n_chunks >= 0 && k >= pipe.n_chunks) break;
        }
        const double tw0 = now_s();
        if (!sink.append(c.out, c.out_len)) { rc = ZF_ERR_IO; pipe.fail(rc); break; }
        t_write += now_s() - tw0;
        for (uint32_t f = 0; f < c.frames; f++) zf_streaminfo_update_frame_size(&si, c.sizes[f]);  // wav2flac.zig:95
        std::lock_guard<std::mutex> lk(pipe.mu);
        c.index = -1;
        c.filled = c.hashed = c.encoded = false;
        pipe.cv.notify_all();
    }
    reader.join();
    t_reader_done = now_s() - t_begin;
    hasher.join();
    for (auto &w : workers) w.join();
    if (trace)
        fprintf(stderr, "zf trace: total %.1f ms | read %.1f (reader done at %.1f) | md5 %.1f (last chunk hashed at %.1f) | encode(dev0) %.1f | write %.1f | chunks of %u frames, ring %d\n",
                (now_s() - t_begin) * 1e3, t_read * 1e3, t_reader_done * 1e3, t_hash * 1e3, t_hasher_done * 1e3, t_enc * 1e3, t_write * 1e3,
                chunk_frames, ring_n);
    if (rc == ZF_OK && pipe.rc != ZF_OK) rc = pipe.rc;
    if (rc == ZF_OK && pipe.incomplete) rc = ZF_ERR_WAV_INCOMPLETE;
    if (rc == ZF_OK) {
        zf_md5x_final(md5_ctx, ossl, si.md5);        // finalizeStreamInfoMd5, encoder.zig:168-170
        zf_write_stream_header(&si, 0, prefix);      // wav2flac.zig:60-61 (last_metadata = false)
        if (!sink.patch_front(prefix, 42)) rc = ZF_ERR_IO;
    }
    release();
    return rc;
}

// WavReader.flacStreaminfo, wav_reader.zig:95-109 -> exit code 2 (wav2flac.zig:24-27), plus what the frame writer can carry
int format_check(const zf_wav_format &fmt) {
    if (fmt.bit_depth < 4 || fmt.bit_depth > 32 || fmt.channels == 0 || fmt.channels > 8 || fmt.sample_rate >= (1u << 20))
        return 2;
    // 4/12/20-bit and containers wider than the sample hit `unreachable` / stale-memory paths upstream
    // (frame_writer.zig:223-232): nothing to be compatible with, so they are refused.
    if ((fmt.bit_depth != 8 && fmt.bit_depth != 16 && fmt.bit_depth != 24 && fmt.bit_depth != 32) ||
        fmt.bytes_per_sample * 8u != fmt.bit_depth)
        return 2;
    return ZF_OK;
}
}  // namespace

extern "C" {

int zf_encode_wav_memory(const uint8_t *wav, size_t wav_len, uint8_t **flac, size_t *flac_len, const int *devices,
                         int n_devices) {
    if (!wav || !flac || !flac_len) return ZF_ERR_INVALID_ARG;
    *flac = nullptr;
    *flac_len = 0;
    zf_wav_format fmt;
    int rc = zf_wav_parse(wav, wav_len, &fmt);
    if (rc) return rc;
    rc = format_check(fmt);
    if (rc) return rc;
    Source src;
    src.mem = wav + fmt.data_offset;
    src.mem_left = wav_len - (size_t)fmt.data_offset;
    Sink sink;
    // room for the usual outcome up front (lossless audio seldom falls below half or rises above the input): growing the
    // buffer by doubling copies the stream a second time
    sink.cap = kPrefix + (src.mem_left / 4) * 3 + (1u << 20);
    sink.buf = (uint8_t *)malloc(sink.cap);
    if (!sink.buf) return ZF_ERR_NOMEM;
    rc = encode_stream(src, sink, fmt, fmt.samples_count, devices, n_devices);
    if (rc) { free(sink.buf); return rc; }
    *flac = sink.buf;
    *flac_len = sink.len;
    return ZF_OK;
}

int zf_encode_wav_file(const char *in_path, const char *out_path, const int *devices, int n_devices) {
    if (!in_path || !out_path) return ZF_ERR_INVALID_ARG;
    FILE *in = fopen(in_path, "rb");
    if (!in) return ZF_ERR_IO;
    // the header: read a prefix of the file and parse it; a file whose `data` chunk lies behind more than that gets more
    std::vector<uint8_t> head;
    zf_wav_format fmt;
    int rc = ZF_ERR_WAV_EOF;
    for (size_t want = 1u << 16;; want *= 4) {
        head.resize(want);
        if (fseek(in, 0, SEEK_SET) != 0) { fclose(in); return ZF_ERR_IO; }
        const size_t got = fread(head.data(), 1, want, in);
        rc = zf_wav_parse(head.data(), got, &fmt);
        const bool truncated = rc == ZF_ERR_WAV_EOF || rc == ZF_ERR_WAV_NO_DATA;
        if (!truncated || got < want || want >= (1u << 30)) break;
    }
    if (rc == ZF_OK) rc = format_check(fmt);
    if (rc) { fclose(in); return rc; }
    if (fseek(in, (long)fmt.data_offset, SEEK_SET) != 0) { fclose(in); return ZF_ERR_IO; }
    FILE *out = fopen(out_path, "wb");
    if (!out) { fclose(in); return ZF_ERR_IO; }
    Source src;
    src.file = in;
    Sink sink;
    sink.file = out;
    rc = encode_stream(src, sink, fmt, fmt.samples_count, devices, n_devices);
    fclose(in);
    if (fclose(out) != 0 && rc == ZF_OK) rc = ZF_ERR_IO;
    return rc;
}

void zf_free(void *p) { free(p); }

void zf_driver_release_cache(void) {
    std::lock_guard<std::mutex> one_call(g_cache.call);
    g_cache.release();
}

}  // extern "C"

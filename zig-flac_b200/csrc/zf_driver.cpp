// zf_driver.cpp -- whole-file driver above the C ABI: the host side of `flac in.wav out.flac`.
//
// Mirrors src/cli/wav2flac.zig:10-97 (main + encode): parse the WAV, reserve the 42-byte header,
// write the vendor block, encode every frame, then back-patch STREAMINFO with MD5, min/max frame
// size and sample count.  What changes against the reference: frames are encoded on the GPU(s) in
// batches; with several devices the stream is cut into contiguous frame ranges, one host thread
// and one encoder handle per device, no inter-GPU communication; MD5 (serial) runs on its own host
// thread concurrently with the GPUs; min/max frame size is replayed in frame order afterwards
// because StreamInfo.updateFrameSize is order dependent (metadata.zig:35-40).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/zigflac_b200.h"

namespace {
constexpr uint16_t kFrameSize = 4096;  // option.frame_size, build.zig:13
constexpr size_t kPrefix = 42 + 31;    // "fLaC" + STREAMINFO block + VORBIS_COMMENT block

struct Shard {
    int device = 0;
    uint64_t first_frame = 0, frames = 0, samples = 0;
    const uint8_t *pcm = nullptr;
    std::vector<uint8_t> out;
    std::vector<uint32_t> sizes;
    size_t out_len = 0;
    int rc = ZF_OK;
};

void run_shard(Shard *sh, const zf_wav_format *fmt) {
    zf_config cfg;
    zf_config_default(&cfg, (uint8_t)fmt->channels, (uint8_t)fmt->bit_depth, fmt->sample_rate);  // wav2flac.zig:38-42
    cfg.device_id = sh->device;
    cfg.max_frames_per_batch = 2048;
    zf_encoder *enc = nullptr;
    sh->rc = zf_encoder_create(&cfg, &enc);
    if (sh->rc) return;
    sh->out.resize(zf_max_batch_bytes(&cfg, (uint32_t)sh->frames));
    sh->sizes.resize(sh->frames ? sh->frames : 1);
    uint32_t nf = 0;
    sh->rc = zf_encode_pcm(enc, sh->pcm, sh->samples, sh->first_frame, sh->out.data(), sh->out.size(), &sh->out_len,
                           sh->sizes.data(), (uint32_t)sh->sizes.size(), &nf);
    zf_encoder_destroy(enc);
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
    // WavReader.flacStreaminfo, wav_reader.zig:95-109 -> exit code 2 (wav2flac.zig:24-27)
    if (fmt.bit_depth < 4 || fmt.bit_depth > 32 || fmt.channels == 0 || fmt.channels > 8 || fmt.sample_rate >= (1u << 20))
        return 2;
    // 4/8/12/20-bit and containers wider than the sample hit `unreachable` / stale-memory paths upstream
    // (frame_writer.zig:223-232, wav_reader.zig:71-78): nothing to be compatible with, so they are refused.
    if ((fmt.bit_depth != 16 && fmt.bit_depth != 24 && fmt.bit_depth != 32) || fmt.bytes_per_sample * 8u != fmt.bit_depth)
        return 2;
    const size_t ic_bytes = (size_t)fmt.channels * fmt.bytes_per_sample;
    uint64_t samples = fmt.samples_count;
    const size_t avail = wav_len - (size_t)fmt.data_offset;
    if ((uint64_t)(avail / ic_bytes) < samples) {  // truncated file: fillSamples stops at end of stream
        if (avail % ic_bytes) return ZF_ERR_WAV_INCOMPLETE;  // wav_reader.zig:52-53
        samples = avail / ic_bytes;
    }
    const uint8_t *pcm = wav + fmt.data_offset;
    const uint64_t frames = (samples + kFrameSize - 1) / kFrameSize;

    zf_md5 md5;  // wav_reader.zig:66 hashes the raw data bytes; serial, so it gets its own host thread
    zf_md5_init(&md5);
    uint8_t digest[16];
    std::thread md5_thread([&] {
        zf_md5_update(&md5, pcm, (size_t)samples * ic_bytes);
        zf_md5_final(&md5, digest);
    });

    int one = 0;
    if (n_devices <= 0 || !devices) { devices = &one; n_devices = 1; }
    std::vector<Shard> shards((size_t)n_devices);
    const uint64_t per = (frames + (uint64_t)n_devices - 1) / (uint64_t)n_devices;
    for (int g = 0; g < n_devices; g++) {
        Shard &sh = shards[(size_t)g];
        sh.device = devices[g];
        sh.first_frame = std::min<uint64_t>(per * (uint64_t)g, frames);
        const uint64_t last = std::min<uint64_t>(sh.first_frame + per, frames);
        sh.frames = last - sh.first_frame;
        const uint64_t s0 = sh.first_frame * kFrameSize;
        const uint64_t s1 = std::min<uint64_t>(last * kFrameSize, samples);
        sh.samples = s1 > s0 ? s1 - s0 : 0;
        sh.pcm = pcm + s0 * ic_bytes;
    }
    std::vector<std::thread> workers;
    for (int g = 0; g < n_devices; g++)
        if (shards[(size_t)g].frames) workers.emplace_back(run_shard, &shards[(size_t)g], &fmt);
    for (auto &w : workers) w.join();
    md5_thread.join();
    size_t body = 0;
    for (auto &sh : shards) {
        if (sh.rc) return sh.rc;
        body += sh.out_len;
    }
    uint8_t *buf = (uint8_t *)malloc(kPrefix + body);
    if (!buf) return ZF_ERR_NOMEM;

    zf_streaminfo si;  // wav_reader.zig:95-109
    zf_streaminfo_init(&si);
    si.sample_rate = fmt.sample_rate;
    si.channels = (uint8_t)fmt.channels;
    si.bit_depth = (uint8_t)fmt.bit_depth;
    si.interchannel_samples = fmt.samples_count;  // from the header, even if the file is short (Q15)
    si.min_block_size = kFrameSize;
    si.max_block_size = kFrameSize;
    size_t pos = kPrefix;
    for (auto &sh : shards) {  // host-side ordered concatenation; sizes replayed in frame order
        if (sh.out_len) memcpy(buf + pos, sh.out.data(), sh.out_len);
        pos += sh.out_len;
        for (uint64_t f = 0; f < sh.frames; f++) zf_streaminfo_update_frame_size(&si, sh.sizes[f]);  // wav2flac.zig:95
    }
    memcpy(si.md5, digest, 16);              // finalizeStreamInfoMd5, encoder.zig:168-170
    zf_write_stream_header(&si, 0, buf);     // wav2flac.zig:60-61 (last_metadata = false)
    zf_write_vorbis_comment(1, buf + 42);    // wav2flac.zig:48
    *flac = buf;
    *flac_len = pos;
    return ZF_OK;
}

int zf_encode_wav_file(const char *in_path, const char *out_path, const int *devices, int n_devices) {
    if (!in_path || !out_path) return ZF_ERR_INVALID_ARG;
    FILE *in = fopen(in_path, "rb");
    if (!in) return ZF_ERR_IO;
    if (fseek(in, 0, SEEK_END) != 0) { fclose(in); return ZF_ERR_IO; }
    const long len = ftell(in);
    if (len < 0 || fseek(in, 0, SEEK_SET) != 0) { fclose(in); return ZF_ERR_IO; }
    std::vector<uint8_t> wav((size_t)len);
    const size_t got = len ? fread(wav.data(), 1, (size_t)len, in) : 0;
    fclose(in);
    if (got != (size_t)len) return ZF_ERR_IO;
    uint8_t *flac = nullptr;
    size_t flac_len = 0;
    int rc = zf_encode_wav_memory(wav.data(), wav.size(), &flac, &flac_len, devices, n_devices);
    if (rc) return rc;
    FILE *out = fopen(out_path, "wb");
    if (!out) { free(flac); return ZF_ERR_IO; }
    const size_t put = fwrite(flac, 1, flac_len, out);
    const int crc = fclose(out);
    free(flac);
    return (put == flac_len && crc == 0) ? ZF_OK : ZF_ERR_IO;
}

void zf_free(void *p) { free(p); }

}  // extern "C"

// zf_decode_host.h -- host side of the decoder: metadata blocks, the frame table from the scan kernel's hits, the WAV
// header of a decoded file.  Plain C++ (shared by zf_decode.cu and the CPU test harness tests/kernel_emu/emu_decode.cpp).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace zf {
namespace dec {

struct HostStreamInfo {
    uint32_t min_block, max_block, min_frame, max_frame, sample_rate, channels, bits;
    uint64_t total_samples;
    uint8_t md5[16];
    uint64_t first_frame_offset;
};

// "fLaC", then metadata blocks up to the one flagged last (RFC 9639 section 8); STREAMINFO must be present.
// Returns 0, or -1 (not a FLAC stream) / -2 (truncated) / -3 (malformed STREAMINFO).
inline int parse_metadata(const uint8_t *f, uint64_t len, HostStreamInfo &si) {
    if (len < 42 || memcmp(f, "fLaC", 4) != 0) return -1;
    uint64_t pos = 4;
    bool seen = false;
    for (;;) {
        if (pos + 4 > len) return -2;
        const bool last = (f[pos] & 0x80u) != 0;
        const uint32_t type = f[pos] & 0x7fu;
        const uint64_t blen = ((uint64_t)f[pos + 1] << 16) | ((uint64_t)f[pos + 2] << 8) | f[pos + 3];
        pos += 4;
        if (pos + blen > len) return -2;
        if (type == 0) {
            if (blen != 34) return -3;
            const uint8_t *s = f + pos;
            si.min_block = ((uint32_t)s[0] << 8) | s[1];
            si.max_block = ((uint32_t)s[2] << 8) | s[3];
            si.min_frame = ((uint32_t)s[4] << 16) | ((uint32_t)s[5] << 8) | s[6];
            si.max_frame = ((uint32_t)s[7] << 16) | ((uint32_t)s[8] << 8) | s[9];
            si.sample_rate = ((uint32_t)s[10] << 12) | ((uint32_t)s[11] << 4) | (s[12] >> 4);
            si.channels = ((s[12] >> 1) & 7u) + 1u;
            si.bits = (((uint32_t)(s[12] & 1u) << 4) | (s[13] >> 4)) + 1u;
            si.total_samples = ((uint64_t)(s[13] & 0xFu) << 32) | ((uint64_t)s[14] << 24) | ((uint64_t)s[15] << 16) |
                               ((uint64_t)s[16] << 8) | s[17];
            memcpy(si.md5, s + 18, 16);
            seen = true;
        }
        pos += blen;
        if (last) break;
    }
    if (!seen) return -3;
    si.first_frame_offset = pos;
    return 0;
}

struct HostCand {
    uint64_t pos, number;
    uint32_t block_size, variable;
};

// Chains the scan kernel's hits into the frame table: the first frame sits at first_frame_offset, frame k carries
// number first + k (fixed-blocksize streams), and a hit whose number is not the expected one is a sync pattern inside
// frame data.  fpos gets n + 1 entries (the last is the end of the stream), first_sample the running sample count.
// Returns 0, -1 (no frame at the first offset), -2 (variable-blocksize stream).
inline int chain_frames(std::vector<HostCand> &cand, uint64_t first_frame_offset, uint64_t stream_len, std::vector<uint64_t> &fpos,
                        std::vector<uint64_t> &first_sample, uint64_t &total_samples) {
    // by position: the hits arrive in the order of the scan kernel's atomic counter.  Sorting 64-bit keys (position, index)
    // and gathering is several times faster than sorting the records themselves.
    if (cand.size() < (1u << 24) && (cand.empty() || cand.back().pos < (1ull << 40))) {
        bool fits = true;
        std::vector<uint64_t> key(cand.size());
        for (size_t k = 0; k < cand.size(); k++) {
            fits = fits && cand[k].pos < (1ull << 40);
            key[k] = (cand[k].pos << 24) | k;
        }
        if (fits) {
            std::sort(key.begin(), key.end());
            std::vector<HostCand> sorted(cand.size());
            for (size_t k = 0; k < cand.size(); k++) sorted[k] = cand[key[k] & 0xffffffu];
            cand.swap(sorted);
        } else {
            std::sort(cand.begin(), cand.end(), [](const HostCand &a, const HostCand &b) { return a.pos < b.pos; });
        }
    } else {
        std::sort(cand.begin(), cand.end(), [](const HostCand &a, const HostCand &b) { return a.pos < b.pos; });
    }
    fpos.clear();
    first_sample.clear();
    total_samples = 0;
    size_t i = 0;
    while (i < cand.size() && cand[i].pos < first_frame_offset) i++;
    if (i == cand.size() || cand[i].pos != first_frame_offset) return -1;
    if (cand[i].variable) return -2;
    uint64_t expect = cand[i].number;
    for (; i < cand.size(); i++) {
        if (cand[i].variable || cand[i].number != expect) continue;
        fpos.push_back(cand[i].pos);
        first_sample.push_back(total_samples);
        total_samples += cand[i].block_size;
        expect++;
    }
    fpos.push_back(stream_len);
    return 0;
}

// A frame failed although every frame header in the table is a valid one: the header chained as frame bad + 1 may be
// a header IMAGE inside frame `bad`'s data that happens to carry the expected number (possible in crafted streams;
// in natural data the odds are below 2^-40 per stream).  Drops that hit so that the chain picks the next one with the
// same number; false when there is nothing to drop.
inline bool drop_suspect(std::vector<HostCand> &cand, const std::vector<uint64_t> &fpos, uint64_t bad) {
    if (bad + 2 >= fpos.size()) return false;  // fpos = n frame starts + the stream's end: frame `bad` is the last one
    const uint64_t pos = fpos[bad + 1];
    for (size_t i = 0; i < cand.size(); i++) {
        if (cand[i].pos == pos) {
            cand.erase(cand.begin() + (long)i);
            return true;
        }
    }
    return false;
}

// canonical 44-byte PCM WAV header (WAVE_FORMAT_EXTENSIBLE is not needed by any reader for <= 2 channels; more channels
// get the same header, as most tools accept it)
inline size_t wav_header(uint8_t *h, uint32_t channels, uint32_t bits, uint32_t sample_rate, uint64_t data_len) {
    const uint32_t bytes = bits / 8u, dl = data_len > 0xffffffd0ull ? 0xffffffd0u : (uint32_t)data_len;
    auto u32 = [&](size_t o, uint32_t v) { h[o] = (uint8_t)v; h[o + 1] = (uint8_t)(v >> 8); h[o + 2] = (uint8_t)(v >> 16); h[o + 3] = (uint8_t)(v >> 24); };
    auto u16 = [&](size_t o, uint32_t v) { h[o] = (uint8_t)v; h[o + 1] = (uint8_t)(v >> 8); };
    memcpy(h, "RIFF", 4); u32(4, 36u + dl); memcpy(h + 8, "WAVEfmt ", 8); u32(16, 16); u16(20, 1); u16(22, channels);
    u32(24, sample_rate); u32(28, sample_rate * channels * bytes); u16(32, channels * bytes); u16(34, bits);
    memcpy(h + 36, "data", 4); u32(40, dl);
    return 44;
}

}  // namespace dec
}  // namespace zf

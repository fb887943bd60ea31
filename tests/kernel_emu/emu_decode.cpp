// emu_decode.cpp -- TEST-ONLY harness: runs the decoder's device functions (zig-flac_b200/csrc/zf_kernel_decode.cuh) on
// the CPU, thread by thread, in the order zf_decode.cu launches the kernels.  Not linked into the product library; the
// product decoder fails without a CUDA device.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define ZF_HOST_EMU 1
#include "cuda_emu.h"

namespace emu {  // the decoder's device functions use no barriers or collectives
emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
uint64_t g_xchg[64][32];
void block_barrier() {}
void warp_barrier() {}
void run_block(void (*)(void *), void *, int) {}
}  // namespace emu

#include "../../zig-flac_b200/csrc/zf_kernel_decode.cuh"
#include "../../zig-flac_b200/csrc/zf_decode_host.h"

using namespace zf::dec;

template <typename ST>
static int run(const uint8_t *s, const HostStreamInfo &si, const StreamParams &sp, const std::vector<uint64_t> &fpos,
               const std::vector<uint64_t> &first, uint8_t *pcm, size_t cap, uint32_t *bad) {
    const size_t n = first.size();
    std::vector<ST> work((size_t)sp.channels * plane_stride(sp.max_block));
    ST ring[8 * 32];
    uint16_t tab[4][256];
    for (uint32_t b = 0; b < 256; b++) crc16_build_tables(tab, b);
    const uint32_t bytes = si.bits / 8u, stride = si.channels * bytes;
    for (size_t f = 0; f < n; f++) {
        FrameRec rec;
        decode_frame<ST>(s, fpos[f], fpos[f + 1], sp, work.data(), ring + (f & 31), rec);
        if (rec.status == kOk) {  // zf_dec_crc16_kernel: 64 chunks
            const uint64_t begin = fpos[f], len = fpos[f + 1] - begin, chunk = (len + 63) / 64;
            const uint32_t xc = crc16_xpow8(chunk);
            uint32_t c = 0;
            uint64_t done = 0;
            for (uint32_t k = 0; k < 64 && done < len; k++) {
                const uint64_t m = len - done < chunk ? len - done : chunk;
                c = crc16_mulmod(c, m == chunk ? xc : crc16_xpow8(m)) ^ crc16_span(s + begin + done, m, tab);
                done += m;
            }
            if (c != 0) rec.status = kErrCrc16;
        }
        if (rec.status == kOk) {  // zf_dec_output_kernel
            for (uint32_t i = 0; i < rec.block_size; i++) {
                long long v[8];
                if (!restore_sample<ST>(work.data(), sp.max_block, i, rec, sp.channels, sp.bits, v)) rec.status = kErrRange;
                const uint64_t off = (first[f] + i) * stride;
                if (off + stride > cap) return -100;
                for (uint32_t c = 0; c < sp.channels; c++)
                    for (uint32_t k = 0; k < bytes; k++) pcm[off + c * bytes + k] = (uint8_t)((unsigned long long)v[c] >> (8 * k));
            }
        }
        if (rec.status != kOk) {
            bad[0] = (uint32_t)f;
            bad[1] = rec.status;
            return -34;
        }
    }
    return 0;
}

// returns the PCM byte count or a negative status (the product's ZF_ERR_* numbers); info = {channels, bits, rate, frames}
extern "C" long long emu_decode_flac(const uint8_t *flac, size_t len, uint8_t *pcm, size_t cap, uint32_t *info, uint32_t *bad) {
    HostStreamInfo si;
    memset(&si, 0, sizeof si);
    const int mrc = parse_metadata(flac, len, si);
    if (mrc) return mrc == -2 ? -33 : -32;
    info[0] = si.channels; info[1] = si.bits; info[2] = si.sample_rate; info[3] = 0;
    if (!(si.bits == 8 || si.bits == 16 || si.bits == 24 || si.bits == 32)) return -2;
    std::vector<uint32_t> buf((len + kStreamPad + 3) / 4, 0u);  // 4-byte aligned, zero bytes behind the stream
    uint8_t *s = reinterpret_cast<uint8_t *>(buf.data());
    memcpy(s, flac, len);
    StreamParams sp;
    sp.channels = si.channels; sp.bits = si.bits; sp.max_block = si.max_block ? si.max_block : 65535u; sp.sample_rate = si.sample_rate;
    std::vector<HostCand> cand;
    for (uint64_t pos = si.first_frame_offset; pos + 6 <= len; pos++) {  // zf_dec_scan_kernel
        FrameHdr h;
        if (s[pos] != 0xFF || (s[pos + 1] & 0xFE) != 0xF8) continue;
        if (!parse_header(s + pos, len - pos, sp, h)) continue;
        cand.push_back(HostCand{pos, h.number, h.block_size, h.variable});
    }
    std::vector<uint64_t> fpos, first;
    uint64_t total = 0, need = 0;
    for (int attempt = 0;; attempt++) {  // as decode_core in zf_decode.cu
        const int crc = chain_frames(cand, si.first_frame_offset, len, fpos, first, total);
        if (crc == -2) return -2;
        if (crc) return si.first_frame_offset >= len ? 0 : -33;
        info[3] = (uint32_t)first.size();
        need = total * si.channels * (si.bits / 8u);
        if (need > cap) return -6;
        const int rc = si.bits == 32 ? run<long long>(s, si, sp, fpos, first, pcm, cap, bad) : run<int32_t>(s, si, sp, fpos, first, pcm, cap, bad);
        if (rc == -34 && attempt < 8 && drop_suspect(cand, fpos, bad[0])) continue;
        if (rc) return rc;
        break;
    }
    if (si.total_samples && si.total_samples != total) return -35;
    return (long long)need;
}

#ifdef ZF_EMU_DECODE_MAIN
// Stand-alone form for AddressSanitizer runs (tests/test_decode_emu.py::test_decoder_logic_memory_safety): streams from
// files named on the command line, output buffer of exactly the size STREAMINFO announces (or 1 MB), work planes and the
// padded stream copy of exactly the sizes the product allocates -- an access outside them stops the program.
int main(int argc, char **argv) {
    int worst = 0;
    for (int a = 1; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb");
        if (!f) return 2;
        std::vector<uint8_t> data;
        uint8_t tmp[65536];
        size_t got;
        while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) data.insert(data.end(), tmp, tmp + got);
        fclose(f);
        HostStreamInfo si;
        memset(&si, 0, sizeof si);
        size_t cap = 1 << 20;
        if (parse_metadata(data.data(), data.size(), si) == 0 && si.total_samples)
            cap = (size_t)si.total_samples * si.channels * ((si.bits + 7) / 8);
        std::vector<uint8_t> pcm(cap ? cap : 1);
        uint32_t info[4] = {0, 0, 0, 0}, bad[2] = {0, 0};
        const long long rc = emu_decode_flac(data.data(), data.size(), pcm.data(), cap, info, bad);
        printf("%s %lld\n", argv[a], rc);
        if (rc < 0) worst = 1;
    }
    return worst ? 1 : 0;
}
#endif

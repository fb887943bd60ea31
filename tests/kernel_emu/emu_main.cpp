// emu_main.cpp -- TEST-ONLY harness: runs the encode kernels' source on the CPU under cuda_emu.h.
// Not linked into the product library (see cuda_emu.h).
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <vector>

#define ZF_HOST_EMU 1
#include "cuda_emu.h"

namespace zf { alignas(128) unsigned char zf_smem[256 * 1024]; namespace v3 { alignas(128) unsigned char zf_smem[256 * 1024]; } namespace lpc { alignas(128) unsigned char zf_smem[256 * 1024]; } }

#include "../../zig-flac_b200/csrc/zf_kernel.cuh"
#include "../../zig-flac_b200/csrc/zf_kernel_indep.cuh"
#include "../../zig-flac_b200/csrc/zf_kernel_v3.cuh"
#include "../../zig-flac_b200/csrc/zf_kernel_lpc.cuh"

namespace emu {
emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
uint64_t g_xchg[64][32];

enum { RUN = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
    ucontext_t ctx;
    char *stack;
    int state;
};
static std::vector<Fiber> g_fibers;
static ucontext_t g_sched;
static int g_cur = 0;
static void (*g_fn)(void *);
static void *g_arg;
static const size_t kStack = 512 * 1024;

static void yield_as(int st) {
    g_fibers[g_cur].state = st;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
void block_barrier() { yield_as(WAIT_BLOCK); }
void warp_barrier() { yield_as(WAIT_WARP); }

static void trampoline() {
    g_fn(g_arg);
    g_fibers[g_cur].state = DONE;
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}

void run_block(void (*fn)(void *), void *arg, int nthreads) {
    g_fn = fn;
    g_arg = arg;
    g_fibers.resize(nthreads);
    for (int t = 0; t < nthreads; t++) {
        Fiber &f = g_fibers[t];
        f.stack = (char *)malloc(kStack);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &g_sched;
        makecontext(&f.ctx, trampoline, 0);
        f.state = RUN;
    }
    for (;;) {
        bool ran = false;
        int done = 0;
        for (int t = 0; t < nthreads; t++) {
            if (g_fibers[t].state == RUN) {
                g_cur = t;
                g_threadIdx.x = (unsigned)t;
                swapcontext(&g_sched, &g_fibers[t].ctx);
                ran = true;
            }
            if (g_fibers[t].state == DONE) done++;
        }
        if (done == nthreads) break;
        bool released = false;
        // warp barriers
        for (int w = 0; w * 32 < nthreads; w++) {
            int waiting = 0, alive = 0;
            for (int l = 0; l < 32 && w * 32 + l < nthreads; l++) {
                const int st = g_fibers[w * 32 + l].state;
                if (st != DONE) alive++;
                if (st == WAIT_WARP) waiting++;
            }
            if (alive && waiting == alive) {
                for (int l = 0; l < 32 && w * 32 + l < nthreads; l++)
                    if (g_fibers[w * 32 + l].state == WAIT_WARP) g_fibers[w * 32 + l].state = RUN;
                released = true;
            }
        }
        // block barrier
        int waiting = 0, alive = 0;
        for (int t = 0; t < nthreads; t++) {
            if (g_fibers[t].state != DONE) alive++;
            if (g_fibers[t].state == WAIT_BLOCK) waiting++;
        }
        if (alive && waiting == alive) {
            for (int t = 0; t < nthreads; t++)
                if (g_fibers[t].state == WAIT_BLOCK) g_fibers[t].state = RUN;
            released = true;
        }
        if (!ran && !released) {
            fprintf(stderr, "cuda_emu: deadlock (divergent barrier)\n");
            abort();
        }
    }
    for (int t = 0; t < nthreads; t++) free(g_fibers[t].stack);
}
}  // namespace emu

static zf::FrameJob g_job;
static int g_bytes, g_full, g_indep, g_v3;
static int g_allow_v3 = 1;
static unsigned g_bit_depth = 0;  // 0 = 8 x container bytes
static unsigned g_lpc_order = 0;
static unsigned g_exact = 0;
static std::vector<uint16_t> g_win;
static void make_window(uint32_t n) {  // zf-LPC v1 window, as zf_capi.cu computes it
    g_win.assign(n, 16384);
    if (n <= 1) return;
    const unsigned long long den = (unsigned long long)(n - 1) * (n - 1);
    for (uint32_t i = 0; i < n; i++) {
        const long long d = 2ll * i - (long long)(n - 1);
        g_win[i] = (uint16_t)(16384u - (uint32_t)((((unsigned long long)(d * d)) << 14) / den));
    }
}
static unsigned long long g_v3_frames = 0;

static void kernel_entry(void *) {
    if (g_lpc_order && !g_indep) {
        if (g_bytes == 2) zf::lpc::zf_encode_stereo_lpc_kernel<2>(g_job);
        else zf::lpc::zf_encode_stereo_lpc_kernel<3>(g_job);
        return;
    }
    if (g_exact && !g_indep) {
        if (g_bytes == 2) zf::zf_encode_stereo_kernel<2, false, true>(g_job);
        else if (g_bytes == 3) zf::zf_encode_stereo_kernel<3, false, true>(g_job);
        else zf::zf_encode_stereo_kernel<4, false, true>(g_job);
        return;
    }
    if (g_indep) {
        if (g_bytes == 2) zf::zf_encode_indep_kernel<2>(g_job);
        else if (g_bytes == 3) zf::zf_encode_indep_kernel<3>(g_job);
        else zf::zf_encode_indep_kernel<4>(g_job);
        return;
    }
    if (g_v3) {
        if (g_bytes == 2) zf::v3::zf_encode_stereo_v3_kernel<2>(g_job);
        else if (g_bytes == 3) zf::v3::zf_encode_stereo_v3_kernel<3>(g_job);
        else zf::v3::zf_encode_stereo_v3_kernel<4>(g_job);
        return;
    }
    if (g_bytes == 2) zf::zf_encode_stereo_kernel<2, false>(g_job);
    else if (g_bytes == 3) zf::zf_encode_stereo_kernel<3, false>(g_job);
    else zf::zf_encode_stereo_kernel<4, false>(g_job);
}

static uint16_t *make_pow8(int n) {
    uint16_t *p = (uint16_t *)malloc(sizeof(uint16_t) * n);
    uint32_t v = 1;
    for (int k = 0; k < n; k++) {
        p[k] = (uint16_t)v;
        for (int b = 0; b < 8; b++) v = (v & 0x8000u) ? (((v << 1) ^ 0x8005u) & 0xffffu) : ((v << 1) & 0xffffu);
    }
    return p;
}

extern "C" {

// Emulated stereo encode of a PCM buffer: full-frame launch + (optional) partial last frame launch,
// exactly like the host side of the product does it.  Returns total bytes or -1 on error status.
long long emu_encode(const uint8_t *pcm, unsigned long long samples, int bytes_per_sample, unsigned channels,
                     int stereo_decorrelation, unsigned block_size, unsigned sample_rate,
                     unsigned long long first_frame_number, unsigned max_rice_order, unsigned max_rice_param,
                     uint8_t *out, unsigned long long out_cap, uint32_t *frame_sizes, uint32_t *n_frames_out) {
    g_indep = !(channels == 2 && stereo_decorrelation);
    const unsigned long long frames = (samples + block_size - 1) / block_size;
    const unsigned long long full = samples / block_size;
    const unsigned tail = (unsigned)(samples - full * block_size);
    std::vector<unsigned long long> desc(frames + 1, 0);
    unsigned ticket = 0, status = 0;
    unsigned long long total = 0;
    static uint16_t *pow8 = make_pow8(1 << 18);
    emu::g_blockDim.x = zf::kThreads;
    emu::g_gridDim.x = 1;
    emu::g_blockIdx.x = 0;
    zf::FrameJob j;
    memset(&j, 0, sizeof j);
    j.out = out; j.out_cap = out_cap; j.frame_sizes = frame_sizes; j.desc = desc.data();
    j.ticket = &ticket; j.status = &status; j.total_bytes = &total; j.pow8 = pow8;
    j.batch_frames = (uint32_t)frames; j.first_frame_number = first_frame_number;
    j.frame_stride = block_size * channels * bytes_per_sample; j.sample_rate = sample_rate; j.channels = channels;
    j.max_rice_order = max_rice_order; j.max_rice_param = max_rice_param; j.use_tma = 1;
    j.bit_depth = g_bit_depth;
    j.lpc_order = g_lpc_order;
    g_bytes = bytes_per_sample;
    if (full) {
        j.pcm = pcm; j.n_frames = (uint32_t)full; j.frame_base = 0; j.block_size = block_size;
        g_full = (block_size == (unsigned)zf::kMaxBlock) && !g_indep && g_bit_depth == 0 && !g_lpc_order && !g_exact;
        if (g_lpc_order) { make_window(block_size); j.lpc_window = g_win.data(); }
        g_v3 = g_full && g_allow_v3 && g_bit_depth == 0 && (bytes_per_sample != 4 || max_rice_param >= 20);
        g_job = j;
        ticket = 0;
        if (g_v3) g_v3_frames += full;
        emu::g_blockDim.x = g_v3 ? zf::v3::kT : zf::kThreads;
        emu::run_block(kernel_entry, nullptr, g_v3 ? zf::v3::kT : zf::kThreads);
        emu::g_blockDim.x = zf::kThreads;
        g_v3 = 0;
    }
    if (tail) {
        j.pcm = pcm + full * j.frame_stride; j.n_frames = 1; j.frame_base = (uint32_t)full; j.block_size = tail;
        if (g_lpc_order) { make_window(tail); j.lpc_window = g_win.data(); }
        g_full = 0;
        g_job = j;
        ticket = 0;
        emu::run_block(kernel_entry, nullptr, zf::kThreads);
    }
    if (n_frames_out) *n_frames_out = (uint32_t)frames;
    if (status) return -(long long)status;
    return (long long)total;
}

void emu_allow_v3(int on) { g_allow_v3 = on; }
void emu_set_bit_depth(unsigned d) { g_bit_depth = d; }
void emu_set_lpc_order(unsigned o) { g_lpc_order = o; }
void emu_set_exact_rice(unsigned on) { g_exact = on; }
unsigned long long emu_v3_frames(void) { return g_v3_frames; }
unsigned long long emu_v3_wide_frames(void) { return zf::v3::g_emu_wide_frames; }
unsigned long long emu_v3_narrow_frames(void) { return zf::v3::g_emu_narrow_frames; }

void emu_best_param_nw(unsigned long long S, unsigned B, unsigned n, unsigned P, unsigned *choice, unsigned long long *cost) {
    uint32_t c, k;
    zf::v3::best_param_nw(S, B, n, P, c, k);
    *choice = c;
    *cost = k;
}

// exhaustive-ish check hook for the closed-form parameter search
void emu_best_param(unsigned long long S, unsigned B, unsigned n, unsigned P, unsigned *choice, unsigned long long *cost) {
    uint32_t c;
    unsigned long long k;
    zf::best_param(S, B, n, P, c, k);
    *choice = c;
    *cost = k;
}
}

// cuda_emu.h -- TEST-ONLY host emulation of the handful of CUDA primitives the encode kernels use.
//
// tests/kernel_emu compiles zig-flac_b200/csrc/zf_kernel.cuh with -DZF_HOST_EMU into a separate
// test library so the kernel's integer logic can be exercised against the oracle on a machine
// without a GPU.  Each CUDA thread is a ucontext fiber; __syncthreads and warp collectives are
// cooperative barriers.  It is NOT a CPU fallback: libzigflac_b200.so never contains this code and
// every product entry point fails without a CUDA device.
#pragma once
#include <stdint.h>
#include <string.h>

#define ZF_DEVICE inline
#define __global__
#define __host__
#define __device__
#define __launch_bounds__(...)
#define __shared__
#define __align__(x) __attribute__((aligned(x)))

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct ulonglong2 { unsigned long long x, y; };
struct emu_dim3 { unsigned x, y, z; };

namespace emu {
extern emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
void block_barrier();
void warp_barrier();
extern uint64_t g_xchg[64][32];
void run_block(void (*fn)(void *), void *arg, int nthreads);
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }

static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
}
static inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t sh) {  // shift clamped to 32
    if (sh >= 32u) return lo;
    return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31u;
    return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }

template <typename T>
static inline T emu_exchange(T v, int src_lane_rel_mode, int arg) {
    const int tid = (int)emu::g_threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::g_xchg[warp][lane] = raw;
    emu::warp_barrier();
    int src = lane;
    if (src_lane_rel_mode == 0) src = arg & 31;                      // shfl
    else if (src_lane_rel_mode == 1) src = lane - arg >= 0 ? lane - arg : lane;  // shfl_up
    else if (src_lane_rel_mode == 2) src = lane + arg < 32 ? lane + arg : lane;  // shfl_down
    else src = lane ^ arg;                                           // shfl_xor
    uint64_t got = emu::g_xchg[warp][src];
    emu::warp_barrier();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, 0, src); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_exchange(v, 1, d); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, 2, d); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, 3, m); }

static inline uint32_t emu_warp_fold(uint32_t v, int op) {
    const int tid = (int)emu::g_threadIdx.x, warp = tid >> 5, lane = tid & 31;
    emu::g_xchg[warp][lane] = v;
    emu::warp_barrier();
    uint32_t r = (op == 3) ? 0u : 0u;
    for (int l = 0; l < 32; l++) {
        const uint32_t x = (uint32_t)emu::g_xchg[warp][l];
        if (op == 0) r += x;
        else if (op == 1) r |= x;
        else if (op == 2) r ^= x;
        else if (op == 3) r = x > r ? x : r;
        else r = (l == 0 || x < r) ? x : r;
    }
    emu::warp_barrier();
    return r;
}
static inline uint32_t __ballot_sync(unsigned, int pred) {
    const int lane = (int)emu::g_threadIdx.x & 31;
    return emu_warp_fold(pred ? (1u << lane) : 0u, 1);
}

static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }

static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }

namespace zf {
static inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    const uint64_t both = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int k = 0; k < 4; k++) {
        const uint32_t s = (sel >> (4 * k)) & 0xf;
        uint32_t byte = (uint32_t)(both >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        d |= byte << (8 * k);
    }
    return d;
}
static inline void mbar_init(unsigned long long *, uint32_t) {}
static inline void fence_mbar_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_expect_tx(unsigned long long *, uint32_t) {}
static inline void mbar_wait(unsigned long long *, uint32_t) {}
static inline void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *) { memcpy(dst, src, bytes); }
static inline void tma_load_1d_elect(void *dst, const void *src, uint32_t bytes, unsigned long long *) {
    if (((int)emu::g_threadIdx.x & 31) == 0) memcpy(dst, src, bytes);
}
static inline void cp_async16_cg(void *dst, const void *src) { memcpy(dst, src, 16); }
static inline void cp_async_commit() {}
static inline void pdl_launch_dependents() {}
static inline void pdl_wait_primary() {}
static inline void cp_async_wait_all() {}
static inline void st_relaxed_gpu(unsigned long long *p, unsigned long long v) { *p = v; }
static inline unsigned long long ld_relaxed_gpu(const unsigned long long *p) { return *p; }
static inline uint32_t reduce_add(uint32_t v) { return emu_warp_fold(v, 0); }
static inline uint32_t reduce_or(uint32_t v) { return emu_warp_fold(v, 1); }
static inline uint32_t reduce_xor(uint32_t v) { return emu_warp_fold(v, 2); }
static inline uint32_t reduce_max(uint32_t v) { return emu_warp_fold(v, 3); }
static inline uint32_t reduce_min(uint32_t v) { return emu_warp_fold(v, 4); }
}  // namespace zf

// zf_dev.h -- thin device-intrinsic layer for the encode kernels.
//
// On the product build (nvcc, sm_100a) these are the raw CUDA / PTX primitives.  When ZF_HOST_EMU is
// defined (tests/kernel_emu only -- a fiber-based harness that runs the SAME kernel source on the CPU
// to debug integer logic without a GPU; never part of libzigflac_b200.so) they come from cuda_emu.h.
#pragma once
#include <stdint.h>

#ifdef ZF_HOST_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>

#define ZF_DEVICE __device__ __forceinline__

namespace zf {

// prmt.b32 in default mode: selector nibble bit 3 replicates the sign of the selected byte.
ZF_DEVICE uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

ZF_DEVICE uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D TMA bulk copy (global -> shared), sm_90+/sm_100a ---------------------------------
ZF_DEVICE void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
ZF_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
ZF_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
ZF_DEVICE void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
ZF_DEVICE void mbar_wait(unsigned long long *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(phase)
        : "memory");
}
// bytes must be a multiple of 16, both addresses 16-byte aligned
ZF_DEVICE void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// The same issued by one elected lane of a converged warp (all 32 lanes call it with the same arguments): the addresses
// stay in uniform registers, so a loop of such copies costs a few uniform-datapath instructions per copy instead of a
// per-lane serialisation loop.
ZF_DEVICE void tma_load_1d_elect(void *smem_dst, const void *gmem_src, uint32_t bytes, unsigned long long *bar) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
        "}\n" ::"r"(smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

// ---- Ampere-style asynchronous copy, L2 only (.cg): used to fetch look-back descriptors without holding registers
ZF_DEVICE void cp_async16_cg(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
ZF_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
ZF_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- programmatic dependent launch: lets the next kernel in the stream start while this one is still running
ZF_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// A kernel launched as a programmatic dependent executes this once before it ends, so that its own completion implies that
// of the kernel in front of it (whose output the kernel after it reads); without such a launch it returns at once.
ZF_DEVICE void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- look-back descriptors: one 64-bit word carries flag + value, so relaxed gpu-scope accesses suffice
ZF_DEVICE void st_relaxed_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
ZF_DEVICE unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

ZF_DEVICE uint32_t reduce_add(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }
ZF_DEVICE uint32_t reduce_or(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
ZF_DEVICE uint32_t reduce_xor(uint32_t v) { return __reduce_xor_sync(0xffffffffu, v); }
ZF_DEVICE uint32_t reduce_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
ZF_DEVICE uint32_t reduce_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }

}  // namespace zf
#endif

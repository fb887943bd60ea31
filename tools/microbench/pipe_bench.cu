// pipe_bench.cu -- integer/FP pipe issue-rate probe for sm_100a (development tool, not part of the library).
// Prints warp-instructions per cycle per SM for a set of instruction mixes, to size the encode kernels'
// instruction budget (which ops share the ALU pipe, which go to the FMA pipe, what SHFL/REDUX/LDS/ATOMS cost).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[256 * 4];
    uint32_t v[CHAINS], w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { v[i] = seed * (threadIdx.x + i + 1); w[i] = seed ^ (i * 77 + threadIdx.x); }
    sm[threadIdx.x] = seed;
    __syncthreads();
    float f[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) f[i] = (float)v[i];
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
            if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 3) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 5) asm volatile("abs.s32 %0, %0;" : "+r"(v[i]));
            if (OP == 6) asm volatile("max.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
            if (OP == 7) asm volatile("vabsdiff.s32.s32.s32.add %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 8) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(__uint_as_float(w[i])));
            if (OP == 9) asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+f"(f[i]) : "f"(__uint_as_float(w[i])));
            if (OP == 10) asm volatile("clz.b32 %0, %0;" : "+r"(v[i]));
            if (OP == 11) asm volatile("popc.b32 %0, %0;" : "+r"(v[i]));
            if (OP == 12) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 13) { asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i])); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(seed), "r"(seed)); }
            if (OP == 14) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x + i) & 255]))); v[i] += r; }
            if (OP == 15) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x * 4 + i) & 1023])), "r"(v[i]) : "memory");
            if (OP == 16) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1);
            if (OP == 17) v[i] = __reduce_add_sync(0xffffffffu, v[i]);
            if (OP == 18) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(v[i]));
            if (OP == 19) asm volatile("sub.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
            if (OP == 20) { asm volatile("abs.s32 %0, %1;" : "=r"(w[i]) : "r"(v[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i])); }
            if (OP == 21) asm volatile("bfe.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            if (OP == 22) asm volatile("shr.s32 %0, %0, 31;" : "+r"(v[i]));
            if (OP == 23) asm volatile("shl.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
            if (OP == 24) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(w[i]), "r"(seed)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(__uint_as_float(seed))); }
            if (OP == 25) { uint32_t r; asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x * 4 + i) & 1023])), "r"(v[i]) : "memory"); v[i] ^= r; }
            if (OP == 26) asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x + i * 32) & 1023])), "r"(v[i]) : "memory");
            if (OP == 27) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(*(unsigned long long *)&v[i & ~1]) : "r"(w[i]), "r"(seed));
            if (OP == 28) asm volatile("vabsdiff.s32.s32.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc += v[i] + w[i] + (uint32_t)f[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

template <int OP>
void run(const char *name, int per_iter, uint32_t *out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int grid = p.multiProcessorCount * 4;  // 32 warps per SM
    k<OP><<<grid, 256>>>(out, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<OP><<<grid, 256>>>(out, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double warp_inst = (double)grid * 8 * ITERS * CHAINS * per_iter;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SM (at %d MHz nominal)  err=%s\n", name, ms, warp_inst / cycles / p.multiProcessorCount, clk / 1000,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *out; cudaMalloc(&out, 4096);
    run<0>("IADD", 1, out); run<19>("ISUB", 1, out); run<1>("IMAD", 1, out); run<2>("LOP3", 1, out); run<3>("SHF var", 1, out);
    run<22>("SHR imm", 1, out); run<23>("SHL var", 1, out); run<4>("PRMT", 1, out); run<5>("IABS", 1, out); run<6>("IMNMX", 1, out);
    run<7>("VABSDIFF.add", 1, out); run<28>("VABSDIFF", 1, out); run<8>("FADD", 1, out); run<9>("FADD |x|", 1, out); run<10>("CLZ/FLO", 1, out);
    run<11>("POPC", 1, out); run<12>("SETP+SELP", 2, out); run<13>("IADD+IMAD pair", 2, out); run<20>("IABS+IADD pair", 2, out);
    run<24>("LOP3+FADD pair", 2, out); run<21>("BFE", 1, out); run<18>("I2F", 1, out); run<27>("IMAD.WIDE", 1, out);
    run<14>("LDS", 1, out); run<26>("STS", 1, out); run<15>("RED.shared.or", 1, out); run<25>("ATOMS.or ret", 1, out);
    run<16>("SHFL", 1, out); run<17>("REDUX.add", 1, out);
    return 0;
}

// pipe_bench2.cu -- second probe: dependent-chain-free forms of IABS / max-relu / 3-input ops / 64-bit ops /
// mixed pipes, to cost the abs-sum and bit-writer inner loops (development tool).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 1024
#define CH 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[2048];
    uint32_t v[CH], w[CH], acc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { v[i] = seed * (threadIdx.x + i + 1); w[i] = seed ^ (i * 77 + threadIdx.x); acc[i] = 0; }
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = seed;
    __syncthreads();
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            uint32_t t0, t1;
            if (OP == 0) {  // e = v - w ; acc += |e|  (3 ops)
                asm volatile("sub.s32 %0, %1, %2;" : "=r"(t0) : "r"(v[i]), "r"(w[i]));
                asm volatile("abs.s32 %0, %1;" : "=r"(t1) : "r"(t0));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(t1));
                v[i] = t0;
            }
            if (OP == 1) {  // e = v - w ; acc += max(e, 0)
                asm volatile("sub.s32 %0, %1, %2;" : "=r"(t0) : "r"(v[i]), "r"(w[i]));
                asm volatile("max.s32 %0, %1, 0;" : "=r"(t1) : "r"(t0));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(t1));
                v[i] = t0;
            }
            if (OP == 2) {  // abs only, fresh input each time
                asm volatile("abs.s32 %0, %1;" : "=r"(t1) : "r"(v[i]));
                asm volatile("add.u32 %0, %1, %2;" : "=r"(v[i]) : "r"(t1), "r"(w[i]));
            }
            if (OP == 3) {  // max relu only + add
                asm volatile("max.s32 %0, %1, 0;" : "=r"(t1) : "r"(v[i]));
                asm volatile("add.u32 %0, %1, %2;" : "=r"(v[i]) : "r"(t1), "r"(w[i]));
            }
            if (OP == 4) {  // 64-bit add
                unsigned long long a = ((unsigned long long)v[i] << 32) | acc[i], b = ((unsigned long long)w[i] << 32) | seed;
                asm volatile("add.u64 %0, %0, %1;" : "+l"(a) : "l"(b));
                v[i] = (uint32_t)(a >> 32); acc[i] = (uint32_t)a;
            }
            if (OP == 5) {  // 64-bit shl var
                unsigned long long a = ((unsigned long long)v[i] << 32) | acc[i];
                asm volatile("shl.b64 %0, %0, %1;" : "+l"(a) : "r"(w[i] & 63));
                v[i] = (uint32_t)(a >> 32); acc[i] = (uint32_t)a | 1;
            }
            if (OP == 6) {  // xor-sign zigzag: (v<<1) ^ (v>>31)
                asm volatile("{.reg .s32 s; shr.s32 s, %1, 31; shl.b32 %0, %1, 1; xor.b32 %0, %0, s;}" : "=r"(t0) : "r"(v[i]));
                v[i] = t0 + w[i];
            }
            if (OP == 7) {  // LDS.128
                uint32_t a, b, c, d;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sbase + ((threadIdx.x * 16 + i * 4096 + (v[i] & 16)) & 8191)));
                v[i] += a + b + c + d;
            }
            if (OP == 8) {  // LDS.32 random address (bank conflicts)
                uint32_t a;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sbase + ((v[i] * 4) & 1023)));
                v[i] = v[i] * 5 + a + 1;
            }
            if (OP == 9) {  // FLO
                asm volatile("bfind.u32 %0, %1;" : "=r"(t0) : "r"(v[i]));
                v[i] = t0 + w[i];
            }
            if (OP == 10) {  // setp + predicated add
                asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; @p add.u32 %0, %0, %2;}" : "+r"(v[i]) : "r"(w[i]), "r"(seed));
            }
            if (OP == 11) {  // 3-input add
                asm volatile("{add.u32 %0, %0, %1; add.u32 %0, %0, %2;}" : "+r"(v[i]) : "r"(w[i]), "r"(acc[(i + 1) % CH]));
            }
            if (OP == 12) {  // min+max of same value (pass-2 pattern)
                asm volatile("min.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
                asm volatile("max.s32 %0, %0, %1;" : "+r"(acc[i]) : "r"(w[i]));
                w[i] += seed;
            }
            if (OP == 13) {  // STS.32 conflict-free
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + (threadIdx.x * 4 + i * 1024) % 8192), "r"(v[i]) : "memory");
                v[i] += 1;
            }
            if (OP == 14) {  // predicated-off atomics: only lane 0 active
                if ((threadIdx.x & 31) == 0) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sbase + (threadIdx.x * 4 + i * 1024) % 8192), "r"(v[i]) : "memory");
                v[i] += 1;
            }
            if (OP == 15) {  // REDUX.OR
                v[i] = __reduce_or_sync(0xffffffffu, v[i]) + w[i];
            }
            if (OP == 16) {  // SHFL + 3 IADD (does SHFL overlap ALU work?)
                t0 = __shfl_xor_sync(0xffffffffu, v[i], 1);
                asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(t0));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(seed));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
            }
            if (OP == 17) {  // IMAD with small constant (x*3 + y)
                asm volatile("mad.lo.s32 %0, %1, 3, %0;" : "+r"(v[i]) : "r"(w[i]));
            }
            if (OP == 18) {  // lop3 OR of 3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xfe;" : "+r"(v[i]) : "r"(w[i]), "r"(acc[(i + 1) % CH]));
                w[i] += seed;
            }
            if (OP == 19) {  // funnel shift
                v[i] = __funnelshift_l(v[i], w[i], acc[i] + i) + 1;
            }
            if (OP == 20) {  // I2F + FADD|x|
                float f;
                asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v[i]));
                float g = __uint_as_float(acc[i]);
                asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+f"(g) : "f"(f));
                acc[i] = __float_as_uint(g);
                v[i] += w[i];
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) r += v[i] + w[i] + acc[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}

template <int OP>
void run(const char *name, int per_iter, uint32_t *out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int grid = p.multiProcessorCount * 4;
    k<OP><<<grid, 256>>>(out, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<OP><<<grid, 256>>>(out, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double groups = (double)grid * 8 * ITERS * CH;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms  %6.3f SM-clk per warp-group (%d nominal instr)  err=%s\n", name, ms, cycles * p.multiProcessorCount / groups, per_iter,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *out; cudaMalloc(&out, 4096);
    run<0>("sub+abs+add", 3, out); run<1>("sub+max0+add", 3, out); run<2>("abs+add", 2, out); run<3>("max0+add", 2, out);
    run<4>("add.u64", 2, out); run<5>("shl.b64 var", 2, out); run<6>("zigzag(shr,shl,xor)+add", 4, out); run<7>("LDS.128", 1, out);
    run<8>("LDS.32 random (conflicts)", 1, out); run<9>("FLO+add", 2, out); run<10>("setp+@p add", 2, out); run<11>("add3", 1, out);
    run<12>("min+max+add", 3, out); run<13>("STS.32", 1, out); run<14>("RED.or lane0 only", 1, out); run<15>("REDUX.OR+add", 2, out);
    run<16>("SHFL+3 add", 4, out); run<17>("IMAD x*3+y", 1, out); run<18>("LOP3 or3 + add", 2, out); run<19>("SHF funnel + add", 2, out);
    run<20>("I2F+FADD|x|+add", 3, out);
    return 0;
}

// chain_bench.cu -- what does pass 1 (fixed.bestOrder's difference chains, sum |delta^k x| for k = 0..4 of the four
// candidate channels) cost per inter-channel sample in int32, fp32 and fp64 arithmetic?  Development tool:
// 444 CTAs x 256 threads (the v3 kernel's residency), samples from shared memory, 16 samples per thread per round.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/chain_bench tools/microbench/chain_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ROUNDS 256

template <typename V>
struct Ch {
    V xp, e1p, e2p, e3p, s0, s1, s2, s3, s4;
    __device__ void init() { xp = e1p = e2p = e3p = 0; s0 = s1 = s2 = s3 = s4 = 0; }
};

__device__ __forceinline__ void step(Ch<int32_t> &c, int32_t x) {
    const int32_t e1 = x - c.xp, e2 = e1 - c.e1p, e3 = e2 - c.e2p, e4 = e3 - c.e3p;
    c.xp = x; c.e1p = e1; c.e2p = e2; c.e3p = e3;
    c.s0 += abs(x); c.s1 += abs(e1); c.s2 += abs(e2); c.s3 += abs(e3); c.s4 += abs(e4);
}
__device__ __forceinline__ void step(Ch<float> &c, float x) {
    const float e1 = x - c.xp, e2 = e1 - c.e1p, e3 = e2 - c.e2p, e4 = e3 - c.e3p;
    c.xp = x; c.e1p = e1; c.e2p = e2; c.e3p = e3;
    c.s0 += fabsf(x); c.s1 += fabsf(e1); c.s2 += fabsf(e2); c.s3 += fabsf(e3); c.s4 += fabsf(e4);
}
__device__ __forceinline__ void step(Ch<double> &c, double x) {
    const double e1 = x - c.xp, e2 = e1 - c.e1p, e3 = e2 - c.e2p, e4 = e3 - c.e3p;
    c.xp = x; c.e1p = e1; c.e2p = e2; c.e3p = e3;
    c.s0 += fabs(x); c.s1 += fabs(e1); c.s2 += fabs(e2); c.s3 += fabs(e3); c.s4 += fabs(e4);
}

// MODE 0: int32, four chains side by side (the kernel today)
// MODE 1: fp64, two chains per trip, two trips
// MODE 2: fp32, four chains side by side
// MODE 3: fp64 with the range predicate (32-bit PCM: |delta^k| >= 2^31 anywhere)
// MODE 4: int32 sums of order 0 + fp64 orders 1..4, two chains per trip
template <int MODE>
__global__ void __launch_bounds__(256, 3) k(uint32_t *out, uint32_t seed) {
    __shared__ uint4 raw[256 * 4 + 4];
    for (int i = threadIdx.x; i < 256 * 4 + 4; i += 256) raw[i] = make_uint4(seed * i, seed ^ i, seed + i, i);
    __syncthreads();
    const int t = threadIdx.x;
    uint32_t res = 0;
    for (int r = 0; r < ROUNDS; r++) {
        if (MODE == 0) {
            Ch<int32_t> c0, c1, c2, c3;
            c0.init(); c1.init(); c2.init(); c3.init();
            uint32_t orv = 0;
#pragma unroll 1
            for (int g = 0; g < 4; g++) {
                const uint4 v = raw[4 * t + g + (r & 1)];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int32_t L = (int32_t)__byte_perm(w[q], 0, 0x9910), R = (int32_t)__byte_perm(w[q], 0, 0xBB32);
                    const int32_t M = (L + R) >> 1, S = L - R;
                    orv |= (uint32_t)L | (uint32_t)R;
                    step(c0, L); step(c1, R); step(c2, M); step(c3, S);
                }
            }
            res += orv + c0.s0 + c0.s1 + c0.s2 + c0.s3 + c0.s4 + c1.s0 + c1.s1 + c1.s2 + c1.s3 + c1.s4 + c2.s0 + c2.s1 + c2.s2 +
                   c2.s3 + c2.s4 + c3.s0 + c3.s1 + c3.s2 + c3.s3 + c3.s4;
        } else if (MODE == 2) {
            Ch<float> c0, c1, c2, c3;
            c0.init(); c1.init(); c2.init(); c3.init();
            uint32_t orv = 0;
#pragma unroll 1
            for (int g = 0; g < 4; g++) {
                const uint4 v = raw[4 * t + g + (r & 1)];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int32_t L = (int32_t)__byte_perm(w[q], 0, 0x9910), R = (int32_t)__byte_perm(w[q], 0, 0xBB32);
                    const int32_t M = (L + R) >> 1;
                    orv |= (uint32_t)L | (uint32_t)R;
                    const float fl = (float)L, fr = (float)R;
                    step(c0, fl); step(c1, fr); step(c2, (float)M); step(c3, fl - fr);
                }
            }
            const float s = c0.s0 + c0.s1 + c0.s2 + c0.s3 + c0.s4 + c1.s0 + c1.s1 + c1.s2 + c1.s3 + c1.s4 + c2.s0 + c2.s1 + c2.s2 +
                            c2.s3 + c2.s4 + c3.s0 + c3.s1 + c3.s2 + c3.s3 + c3.s4;
            res += orv + (uint32_t)s;
        } else {
            double tot = 0;
            uint32_t orv = 0, isum = 0;
            bool big = false;
#pragma unroll 1
            for (int trip = 0; trip < 2; trip++) {
                Ch<double> ca, cb;
                ca.init(); cb.init();
#pragma unroll 1
                for (int g = 0; g < 4; g++) {
                    const uint4 v = raw[4 * t + g + (r & 1)];
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int32_t L = (int32_t)__byte_perm(w[q], 0, 0x9910), R = (int32_t)__byte_perm(w[q], 0, 0xBB32);
                        const int32_t xa = trip ? (L + R) >> 1 : L, xb = trip ? L - R : R;
                        orv |= (uint32_t)xa | (uint32_t)xb;
                        if (MODE == 4) isum += (uint32_t)abs(xa) + (uint32_t)abs(xb);
                        const double da = (double)xa, db = (double)xb;
                        if (MODE == 3) {
                            const double a1 = da - ca.xp, a2 = a1 - ca.e1p, a3 = a2 - ca.e2p, a4 = a3 - ca.e3p;
                            const double b1 = db - cb.xp, b2 = b1 - cb.e1p, b3 = b2 - cb.e2p, b4 = b3 - cb.e3p;
                            big = big || fabs(a1) >= 2147483648.0 || fabs(a2) >= 2147483648.0 || fabs(a3) >= 2147483648.0 ||
                                  fabs(a4) >= 2147483648.0 || fabs(b1) >= 2147483648.0 || fabs(b2) >= 2147483648.0 ||
                                  fabs(b3) >= 2147483648.0 || fabs(b4) >= 2147483648.0;
                        }
                        if (MODE == 4) {
                            const double a1 = da - ca.xp, a2 = a1 - ca.e1p, a3 = a2 - ca.e2p, a4 = a3 - ca.e3p;
                            const double b1 = db - cb.xp, b2 = b1 - cb.e1p, b3 = b2 - cb.e2p, b4 = b3 - cb.e3p;
                            ca.xp = da; ca.e1p = a1; ca.e2p = a2; ca.e3p = a3;
                            cb.xp = db; cb.e1p = b1; cb.e2p = b2; cb.e3p = b3;
                            ca.s1 += fabs(a1); ca.s2 += fabs(a2); ca.s3 += fabs(a3); ca.s4 += fabs(a4);
                            cb.s1 += fabs(b1); cb.s2 += fabs(b2); cb.s3 += fabs(b3); cb.s4 += fabs(b4);
                        } else {
                            step(ca, da); step(cb, db);
                        }
                    }
                }
                tot += ca.s0 + ca.s1 + ca.s2 + ca.s3 + ca.s4 + cb.s0 + cb.s1 + cb.s2 + cb.s3 + cb.s4;
            }
            res += orv + isum + (uint32_t)(long long)tot + (big ? 1u : 0u);
        }
    }
    out[blockIdx.x * 256 + t] = res;
}

template <int MODE>
void run(const char *name, uint32_t *d) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<444, 256>>>(d, 12345u);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int i = 0; i < 5; i++) {
        cudaEventRecord(a);
        k<MODE><<<444, 256>>>(d, 12345u + i);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    // per SM: 3 CTAs x 8 warps x ROUNDS x 16 warp-samples
    const double warp_samples_per_sm = 3.0 * 8 * ROUNDS * 16;
    printf("%-44s %8.3f ms  %7.2f SM-clk per warp-sample (1965 MHz)  err=%s\n", name, best,
           best * 1e-3 * 1.965e9 / warp_samples_per_sm, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *d;
    cudaMalloc(&d, 444 * 256 * 4);
    run<0>("int32 x4 chains", d);
    run<1>("fp64 2 chains x 2 trips", d);
    run<2>("fp32 x4 chains", d);
    run<3>("fp64 2x2 + range predicate", d);
    run<4>("fp64 orders 1..4 + int order 0, 2x2", d);
    return 0;
}

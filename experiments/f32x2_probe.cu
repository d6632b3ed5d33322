// Throughput probe: scalar FFMA vs packed FFMA2 / FMUL2 / FADD2 (fp32x2) and MUFU.RSQ on sm_100a.
// Each warp runs 8 independent dependency chains so latency is hidden; reports lane-results per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void probe(float *out, long long *cyc, int iters) {
    float s[8]; u64 p[8];
    for (int k = 0; k < 8; ++k) { s[k] = threadIdx.x * 1e-3f + k; p[k] = ((u64)__float_as_uint(s[k]) << 32) | __float_as_uint(s[k] + 0.5f); }
    const float c = 1.0001f; const u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(s[k]) : "f"(c));
            if (MODE == 1) p[k] = fma2(p[k], c2, c2);
            if (MODE == 2) p[k] = mul2(p[k], c2);
            if (MODE == 3) p[k] = add2(p[k], c2);
            if (MODE == 4) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[k]));
            if (MODE == 5) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[k]) : "f"(c));
        }
    }
    long long t1 = clock64();
    float acc = 0; for (int k = 0; k < 8; ++k) acc += s[k] + __uint_as_float((unsigned)p[k]) + __uint_as_float((unsigned)(p[k] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const char *names[] = {"FFMA", "FFMA2", "FMUL2", "FADD2", "MUFU.RSQ", "FADD"};
    const int iters = 4096;
    for (int mode = 0; mode < 6; ++mode) for (int threads = 128; threads <= 1024; threads *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
            case 0: probe<0><<<148, threads>>>(out, cyc, iters); break;
            case 1: probe<1><<<148, threads>>>(out, cyc, iters); break;
            case 2: probe<2><<<148, threads>>>(out, cyc, iters); break;
            case 3: probe<3><<<148, threads>>>(out, cyc, iters); break;
            case 4: probe<4><<<148, threads>>>(out, cyc, iters); break;
            case 5: probe<5><<<148, threads>>>(out, cyc, iters); break;
            }
            cudaDeviceSynchronize();
        }
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double instr = (double)iters * 8 * (threads / 32);
        double lanes = instr * 32 * ((mode >= 1 && mode <= 3) ? 2 : 1);
        printf("%-9s threads=%4d  warp-instr/clk/SM %.2f  lane-results/clk/SM %.1f\n", names[mode], threads, instr / h[0], lanes / h[0]);
    }
    return 0;
}

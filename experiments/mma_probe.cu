// Micro-probe for sm_100a tcgen05 behaviour that the kernel designs in csrc/ depend on:
//   (A) cycles per tcgen05.mma for the operand forms / shapes used (SS vs TS, K- vs MN-major B, N),
//       thread-side issue cost, tcgen05.ld cost;
//   (B) whether a SW128 K-major A operand may start at an arbitrary ROW offset inside a larger
//       TMA-style tile (descriptor base_offset semantics) — the sliding-window temporal conv needs it.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../computer-vision-system-for-analyzing-golfer-action_b200/csrc \
//              experiments/mma_probe.cu -o experiments/_build/mma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "gcn_fused.cuh"

namespace gs {
void set_error(const char *, ...) {}
const char *kernel_name(int) { return ""; }
}  // namespace gs

using namespace gs;
using namespace gs::tc;
using namespace gs::gcn;

__device__ __forceinline__ uint64_t desc_kmajor_bo(uint32_t smem_addr, uint32_t row_bytes, uint32_t base_off) {
    return make_kmajor_desc(smem_addr, row_bytes) | ((uint64_t)(base_off & 7) << 49);
}

struct Result {
    long long cyc[32][4];   // [test][0: issue cycles, 1: total cycles, 2: reps]
    int mism[8][4];         // [shift idx][variant]
    int chain[8];           // f16 TMEM-chain variants: mismatches (or -1 = not run)
};

__global__ void __launch_bounds__(256, 1) probe_kernel(Result *out, int reps, int vmask) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *sA = smem;               // 384 rows x 128 B
    unsigned char *sB = smem + 384 * 128;   // 256 rows x 128 B
    // A[r][k] = (r*64+k) % 251 ; B[n][k] = (n == k)   (both K-major SW128: chunk ^= row & 7)
    for (int e = threadIdx.x; e < 384 * 64; e += blockDim.x) {
        const int r = e / 64, k = e % 64;
        const int chunk = k / 8, within = k % 8;
        __nv_bfloat16 *p = reinterpret_cast<__nv_bfloat16 *>(sA + r * 128 + ((chunk ^ (r & 7)) << 4)) + within;
        *p = __float2bfloat16_rn((float)((r * 64 + k) % 251));
    }
    for (int e = threadIdx.x; e < 256 * 64; e += blockDim.x) {
        const int n = e / 64, k = e % 64;
        const int chunk = k / 8, within = k % 8;
        __nv_bfloat16 *p = reinterpret_cast<__nv_bfloat16 *>(sB + n * 128 + ((chunk ^ (n & 7)) << 4)) + within;
        *p = __float2bfloat16_rn(n == k ? 1.f : 0.f);
    }
    unsigned char *sA16 = sB + 256 * 128;   // 128 rows x 128 B, f16 copy of A rows 0..127
    unsigned char *sB16 = sA16 + 128 * 128; // 64 rows x 128 B, f16 identity
    for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
        const int r = e / 64, k = e % 64;
        __half *p = reinterpret_cast<__half *>(sA16 + r * 128 + (((k / 8) ^ (r & 7)) << 4)) + (k % 8);
        *p = __float2half_rn((float)((r * 64 + k) % 251));
        if (r < 64) {
            __half *q = reinterpret_cast<__half *>(sB16 + r * 128 + (((k / 8) ^ (r & 7)) << 4)) + (k % 8);
            *q = __float2half_rn(r == k ? 1.f : 0.f);
        }
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;

    // ---------------- (A) rates ----------------
    for (int test = 0; test < 9; ++test) {
        if (threadIdx.x == 32) {
            const uint32_t a = smem_u32(sA), b = smem_u32(sB);
            long long t0 = clock64(), t1 = 0;
            for (int i = 0; i < reps; ++i) {
                const uint32_t acc = (uint32_t)(i > 0);
                switch (test) {
                    case 0: umma_bf16(tmem, make_kmajor_desc(a, 128), make_kmajor_desc(b, 128), make_idesc_bf16(256), acc); break;
                    case 1: umma_bf16(tmem, make_kmajor_desc(a, 128), make_kmajor_desc(b, 128), make_idesc_bf16(128), acc); break;
                    case 2: umma_bf16(tmem, make_kmajor_desc(a, 128), make_kmajor_desc(b, 128), make_idesc_bf16(64), acc); break;
                    case 3: umma_bf16(tmem, make_kmajor_desc(a, 32), make_kmajor_desc(b, 32), make_idesc_bf16(16), acc); break;
                    case 4: umma_bf16_ts(tmem, tmem + 320, make_mnmajor_desc(a), make_idesc_agg(), acc); break;
                    case 5: umma_bf16_ts(tmem, tmem + 320, make_kmajor_desc(b, 128), make_idesc_bf16(64), acc); break;
                    case 6: umma_bf16_ts(tmem, tmem + 320, make_kmajor_desc(b, 128), make_idesc_bf16(256), acc); break;
                    case 7: umma_bf16(tmem, make_kmajor_desc(a, 128), make_mnmajor_desc(b),
                                      make_idesc_agg(), acc); break;
                    case 8: umma_bf16(tmem, make_kmajor_desc(a, 128), make_kmajor_desc(b, 128), make_idesc_bf16(32), acc); break;
                }
            }
            t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, phase);
            const long long t2 = clock64();
            out->cyc[test][0] = t1 - t0;
            out->cyc[test][1] = t2 - t0;
            out->cyc[test][2] = reps;
        }
        phase ^= 1;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // same MMAs issued with warp-uniform control flow: the whole warp runs the loop, one elected lane issues
    for (int test = 0; test < 8; ++test) {
        if (warp == 1) {
            const uint32_t a = smem_u32(sA), b = smem_u32(sB);
            const uint64_t da = make_kmajor_desc(a, 128), db = make_kmajor_desc(b, 128), dmn = make_mnmajor_desc(a);
            const uint32_t id256 = make_idesc_bf16(256), id64 = make_idesc_bf16(64), idagg = make_idesc_agg();
            __syncwarp();
            long long t0 = clock64(), t1 = 0;
            for (int i = 0; i < reps; ++i) {
                const uint32_t acc = (uint32_t)(i > 0);
                if (elect_one()) {
                    if (test == 0) umma_bf16(tmem, da, db, id256, acc);
                    else if (test == 1) umma_bf16(tmem, da, db, id64, acc);
                    else if (test == 2) umma_bf16_ts(tmem, tmem + 320, dmn, idagg, acc);
                    else if (test == 3) umma_bf16_ts(tmem, tmem + 320, db, id256, acc);
                    else if (test == 4) umma_bf16(tmem + (i & 1) * 64, da, db, id64, (uint32_t)(i > 1));
                    else if (test == 5) umma_bf16(tmem + (i & 3) * 64, da, db, id64, (uint32_t)(i > 3));
                    else if (test == 6) umma_bf16(tmem + (i & 1) * 256, da, db, id256, (uint32_t)(i > 1));
                    else umma_bf16(tmem + (i & 7) * 16, da, db, make_idesc_bf16(16), (uint32_t)(i > 7));
                }
                __syncwarp();
            }
            t1 = clock64();
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, phase);
            const long long t2 = clock64();
            if (lane == 0) {
                out->cyc[11 + test][0] = t1 - t0;
                out->cyc[11 + test][1] = t2 - t0;
                out->cyc[11 + test][2] = reps;
            }
        }
        phase ^= 1;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // two / four issuing warps at once (different accumulators): is the limit per issuing warp or per SM?
    for (int test = 0; test < 12; ++test) {
        const int nw = test == 0 ? 2 : (test >= 3 ? 1 : 4);                 // warps 0..nw-1 issue
        const uint32_t n = test == 2 ? 16u : (test == 4 ? 256u : (test == 5 ? 128u : ((test == 6 || test >= 9) ? 16u : 64u)));
        // tests 7..11: A start row offset inside a SW128 tile and B layout, as the sliding-window conv uses them
        const uint32_t arow = (test == 7 || test == 10) ? 1u : (test == 8 || test == 11 ? 8u : 0u);
        const bool b_sw32 = test >= 9;
        __shared__ uint64_t bars2[4];
        if (threadIdx.x == 0) {
            for (int i = 0; i < 4; ++i) mbar_init(&bars2[i], 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (warp < nw) {
            const uint32_t a = smem_u32(sA), b = smem_u32(sB);
            const uint64_t da = make_kmajor_desc(a + arow * 128u, 128), db = make_kmajor_desc(b, b_sw32 ? 32 : 128);
            const uint32_t id = make_idesc_bf16(n);
            __syncwarp();
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                if (elect_one()) umma_bf16(tmem + warp * 64, da, db, id, (uint32_t)(i > 0));
                __syncwarp();
            }
            const long long t1 = clock64();
            if (elect_one()) umma_commit(&bars2[warp]);
            __syncwarp();
            mbar_wait(&bars2[warp], 0);
            const long long t2 = clock64();
            if (lane == 0 && warp == 0) {
                out->cyc[19 + test][0] = t1 - t0;
                out->cyc[19 + test][1] = t2 - t0;
                out->cyc[19 + test][2] = (long long)reps * nw;
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // tcgen05.ld cost: warps 4-7 read 64 columns `reps` times
    if (warp >= 4) {
        uint32_t v[32];
        const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
        __syncwarp();
        const long long t0 = clock64();
        uint32_t sink = 0;
        for (int i = 0; i < reps; ++i) {
            tmem_ld32(tmem + lane_base + (i & 1) * 32, v);
            tmem_ld_wait();
            sink += v[i & 31];
        }
        const long long t1 = clock64();
        if (lane == 0 && warp == 4) {
            out->cyc[9][0] = t1 - t0;
            out->cyc[9][1] = sink;
            out->cyc[9][2] = reps;
        }
    }
    __syncthreads();
    // mbarrier try_wait on an already-completed phase: cost per wait
    if (threadIdx.x == 32) {
        const long long t0 = clock64();
        for (int i = 0; i < 16; ++i) mbar_wait(&bar, phase ^ 1);
        out->cyc[10][0] = clock64() - t0;
        out->cyc[10][2] = 16;
    }
    __syncthreads();

    // ---------------- (B) row-shifted SW128 K-major A operand ----------------
    const int shifts[8] = {0, 8, 1, 3, 17, 34, 51, 68};
    for (int si = 0; si < 8; ++si) {
        for (int variant = 0; variant < 3; ++variant) {
            const int s = shifts[si];
            if (threadIdx.x == 32) {
                const uint32_t a = smem_u32(sA) + (uint32_t)s * 128, b = smem_u32(sB);
                const uint32_t bo = variant == 0 ? 0u : (variant == 1 ? (uint32_t)(s & 7) : (uint32_t)((8 - (s & 7)) & 7));
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem, desc_kmajor_bo(a, 128, bo) + (uint64_t)(2 * k), make_kmajor_desc(b, 128) + (uint64_t)(2 * k),
                              make_idesc_bf16(64), (uint32_t)(k > 0));
                umma_commit(&bar);
                mbar_wait(&bar, phase);
            }
            phase ^= 1;
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (warp >= 4) {
                uint32_t v[64];
                const int m = (warp - 4) * 32 + lane;
                const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
                tmem_ld32(tmem + lane_base, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tmem_ld32(tmem + lane_base + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                tmem_ld_wait();
                int bad = 0;
                for (int n = 0; n < 64; ++n) {
                    const float want = (float)(((m + s) * 64 + n) % 251);
                    if (__uint_as_float(v[n]) != want) ++bad;
                }
                if (bad) atomicAdd(&out->mism[si][variant], bad);
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
    }
    // ---------------- (C) f16 accumulator of MMA-a re-used in place as the TMEM A operand of MMA-b ----------------
    // variant bit0: operands of MMA-a are f16 (else bf16); bit1: B of MMA-b is f16 (else bf16)
    for (int variant = 0; variant < 4; ++variant) {
        if (!((vmask >> variant) & 1)) {
            if (threadIdx.x == 0) out->chain[variant] = -1;
            continue;
        }
        const bool a16 = variant & 1, b16 = variant & 2;
        if (warp == 1) {
            const uint32_t fa = a16 ? 0u : 1u, fb = b16 ? 0u : 1u;
            // idesc: c_format [4,6) (0 = f16, 1 = f32), a_format [7,10), b_format [10,13), N>>3 at 17, M>>4 at 24
            const uint32_t id_a = (0u << 4) | (fa << 7) | (fa << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t id_b = (1u << 4) | (0u << 7) | (fb << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t da = make_kmajor_desc(smem_u32(a16 ? sA16 : sA), 128);
            const uint64_t db1 = make_kmajor_desc(smem_u32(a16 ? sB16 : sB), 128);
            const uint64_t db2 = make_kmajor_desc(smem_u32(b16 ? sB16 : sB), 128);
            if (elect_one()) {
                for (int k = 0; k < 4; ++k) umma_bf16(tmem + 256, da + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), id_a, (uint32_t)(k > 0));
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase);
            tc_fence_after();
            if (elect_one()) {
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ts(tmem, tmem + 256 + (uint32_t)(k * 8), db2 + (uint64_t)(2 * k), id_b, (uint32_t)(k > 0));
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase ^ 1);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (warp >= 4) {
            uint32_t v[64];
            const int m = (warp - 4) * 32 + lane;
            const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
            tmem_ld32(tmem + lane_base, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld32(tmem + lane_base + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tmem_ld_wait();
            int bad = 0;
            for (int n = 0; n < 64; ++n)
                if (__uint_as_float(v[n]) != (float)((m * 64 + n) % 251)) ++bad;
            if (bad) atomicAdd(&out->chain[variant], bad);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 2) tmem_dealloc(tmem, 512);
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 64;
    const int grid = argc > 2 ? atoi(argv[2]) : 1;
    const int vmask = argc > 3 ? atoi(argv[3]) : 0;
    Result *d;
    cudaMalloc(&d, sizeof(Result));
    cudaMemset(d, 0, sizeof(Result));
    const int smem = 384 * 128 + 256 * 128 + 128 * 128 + 64 * 128 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<grid, 256, smem>>>(d, reps, vmask);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    Result h;
    cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char *names[31] = {"SS K/K  N=256", "SS K/K  N=128", "SS K/K  N=64", "SS K/K  N=16 (SW32)", "TS A=tmem, B MN-major N=64",
                             "TS A=tmem, B K-major N=64", "TS A=tmem, B K-major N=256", "SS A K-major, B MN-major N=64",
                             "SS K/K  N=32", "tcgen05.ld 32x32b.x32 + wait", "mbar_wait (already complete)",
                             "elect: SS K/K N=256", "elect: SS K/K N=64", "elect: TS B MN-major N=64", "elect: TS B K-major N=256",
                             "elect: SS N=64, 2 accumulators", "elect: SS N=64, 4 accumulators", "elect: SS N=256, 2 accumulators",
                             "elect: SS N=16, 8 accumulators", "2 issuing warps N=64 (per MMA)", "4 issuing warps N=64 (per MMA)",
                             "4 issuing warps N=16 (per MMA)", "1 warp clean loop N=64", "1 warp clean loop N=256",
                             "1 warp clean loop N=128", "1 warp clean loop N=16",
                             "clean N=64, A start +1 row", "clean N=64, A start +8 rows", "clean N=16 A SW128 B SW32",
                             "clean N=16 A +1 row, B SW32", "clean N=16 A +8 rows, B SW32"};
    for (int t = 0; t < 31; ++t)
        printf("%-34s issue %8.1f cyc/op   total %8.1f cyc/op   (reps %lld)\n", names[t],
               (double)h.cyc[t][0] / h.cyc[t][2], (double)h.cyc[t][1] / h.cyc[t][2], h.cyc[t][2]);
    const int shifts[8] = {0, 8, 1, 3, 17, 34, 51, 68};
    for (int si = 0; si < 8; ++si)
        printf("row shift %2d: mismatches base_offset=0: %5d   =s%%8: %5d   =(8-s)%%8: %5d\n", shifts[si], h.mism[si][0],
               h.mism[si][1], h.mism[si][2]);
    for (int v = 0; v < 4; ++v)
        printf("f16 TMEM chain, MMA-a operands %s, MMA-b B %s: mismatches %d\n", (v & 1) ? "f16 " : "bf16", (v & 2) ? "f16 " : "bf16",
               h.chain[v]);
    return 0;
}

// Kernels shared by the fp32 and bf16 segmentation paths: clip-global pooling
// statistics, SE channel attention, ST-joint attention, head.  All reductions are
// fixed-order (no float atomics) so results are run-to-run deterministic.
//
// Stages replaced: /root/reference/README.md:31-32 (channel attention), 33-34
// (ST-joint attention), 17-18 (per-frame phase logits).  Math: oracle/segnet.py
// ChannelAttention / STJointAttention / SegNet.forward.
//
// Deferred gating: a block's pre-attention output U is stored once; its gates
//   gT[b,t,c] = s[b,c] * a_t[b,t,c]     (SE gate folded into the frame gate)
//   gV[b,v,c] = a_v[b,v,c]
// are applied by whichever kernel reads U next (next block's aggregation, head).
#pragma once
#include <cuda_fp16.h>
#include <string.h>
#include <algorithm>

#include "common.cuh"

namespace gs {

template <typename T> __device__ __forceinline__ float ld_act(const T *p);
template <> __device__ __forceinline__ float ld_act<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float ld_act<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_act(T *p, float v);
template <> __device__ __forceinline__ void st_act<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
    *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float hardswishf(float x) {
    return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) / 6.0f;
}

constexpr int V17 = 17;

// ---- gate-on-load + adjacency contraction --------------------------------------
// Xg[row, c]          = gated input (block 0: skel*scale+shift; else U*gT*gV)
// XA[row(w), p*Cin+c] = sum_v A[p,w,v] * Xg[frame, v, c]
// one work item = (frame, channel); A staged in shared memory.
template <typename TIN, typename TOUT, int P>
__global__ void __launch_bounds__(256)
aggregate_kernel(const TIN *__restrict__ Uin, const float *__restrict__ gT, const float *__restrict__ gV,
                 const float *__restrict__ in_scale, const float *__restrict__ in_shift,
                 const float *__restrict__ A, int T, int Cin, size_t nframes, TOUT *__restrict__ Xg,
                 TOUT *__restrict__ XA) {
    __shared__ float sA[P * V17 * V17];
    for (int k = threadIdx.x; k < P * V17 * V17; k += blockDim.x) sA[k] = A[k];
    __syncthreads();
    const size_t items = nframes * Cin;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < items;
         it += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(it % Cin);
        const size_t f = it / Cin;           // frame index = b*T + t
        const size_t b = f / T;
        float x[V17];
        const TIN *src = Uin + (f * V17) * Cin + c;
        if (in_scale) {
#pragma unroll
            for (int v = 0; v < V17; ++v)
                x[v] = ld_act(src + (size_t)v * Cin) * in_scale[v * Cin + c] + in_shift[v * Cin + c];
        } else {
            const float gt = gT[f * Cin + c];
            const float *gv = gV + (b * V17) * Cin + c;
#pragma unroll
            for (int v = 0; v < V17; ++v) x[v] = ld_act(src + (size_t)v * Cin) * gt * gv[(size_t)v * Cin];
        }
        TOUT *xg = Xg + (f * V17) * Cin + c;
#pragma unroll
        for (int v = 0; v < V17; ++v) st_act(xg + (size_t)v * Cin, x[v]);
        TOUT *xa = XA + (f * V17) * (size_t)(P * Cin) + c;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            for (int w = 0; w < V17; ++w) {
                const float *ar = sA + (p * V17 + w) * V17;
                float acc = 0.f;
#pragma unroll
                for (int v = 0; v < V17; ++v) acc += ar[v] * x[v];
                st_act(xa + (size_t)w * (P * Cin) + p * Cin, acc);
            }
        }
    }
}

constexpr int kStatChunk = 30;   // frames per partial-sum chunk of PV

// U [B,T,V,C] -> PT[b,t,c] = sum_v U ; PVpart[b,chunk,v,c] = sum_{t in chunk} U.
// grid (ceil(T/kStatChunk), B, ceil(C/blockDim)), thread = channel.
template <typename TU, int V>
__global__ void __launch_bounds__(128)
stats_kernel(const TU *__restrict__ U, int T, int C, float *__restrict__ PT, float *__restrict__ PVpart) {
    const int c = blockIdx.z * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
    const int t0 = chunk * kStatChunk;
    const int t1 = min(T, t0 + kStatChunk);
    float pv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) pv[v] = 0.f;
    for (int t = t0; t < t1; ++t) {
        const TU *row = U + (((size_t)b * T + t) * V) * C + c;
        float pt = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float x = ld_act(row + (size_t)v * C);
            pt += x;
            pv[v] += x;
        }
        PT[((size_t)b * T + t) * C + c] = pt;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) PVpart[(((size_t)b * nchunk + chunk) * V + v) * C + c] = pv[v];
}

// ---- SE channel attention (README.md:31-32) ---------------------------------------
// grid B, block 1024 = G groups x C channels (G = min(1024/C, V)).  Group g folds the
// per-chunk partial sums of joints v = g, g+G, ... into PV[b,v,c]; the clip mean over (T,V)
// is the fixed-order sum of those; then FC -> ReLU -> FC -> sigmoid with K split over groups.
// smem floats: G*C (partials) + C (mean) + cs (hidden).
template <int V>
__global__ void __launch_bounds__(1024)
se_kernel(const float *__restrict__ PVpart, int T, int C, int cs, int nchunk, int G,
          const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
          const float *__restrict__ b2, float *__restrict__ seS, float *__restrict__ PV) {
    extern __shared__ float sm[];
    float *part = sm, *m = sm + G * C, *hid = m + C;
    const int b = blockIdx.x;
    const int c = threadIdx.x % C, g = threadIdx.x / C;
    if (g < G) {
        float acc = 0.f;
        for (int v = g; v < V; v += G) {
            float s = 0.f;
            for (int k = 0; k < nchunk; ++k) s += PVpart[(((size_t)b * nchunk + k) * V + v) * C + c];
            PV[((size_t)b * V + v) * C + c] = s;
            acc += s;
        }
        part[g * C + c] = acc;
    }
    __syncthreads();
    if (g == 0) {
        float acc = 0.f;
        for (int k = 0; k < G; ++k) acc += part[k * C + c];
        m[c] = acc / (float)(T * V);
    }
    __syncthreads();
    // FC1: hid[j] = relu(b1[j] + sum_k m[k] W1[k,j]); K split over P1 parts, fixed-order fold
    const int P1 = max(1, min((int)blockDim.x / cs, 32));
    {
        const int j = threadIdx.x % cs, pt = threadIdx.x / cs;
        if (pt < P1) {
            float acc = 0.f;
            for (int k = pt; k < C; k += P1) acc += m[k] * W1[k * cs + j];
            part[pt * cs + j] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x < cs) {
        float acc = b1[threadIdx.x];
        for (int p = 0; p < P1; ++p) acc += part[p * cs + threadIdx.x];
        hid[threadIdx.x] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    // FC2: K = cs split over the G groups in contiguous ranges (a 64-deep chain of dependent L2 loads in one
    // group was most of this kernel's latency), partials folded in group order: still a fixed summation order
    if (g < G) {
        const int per = (cs + G - 1) / G, k0 = g * per, k1 = min(cs, k0 + per);
        float acc = 0.f;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) acc += hid[k] * __ldg(W2 + k * C + c);
        part[g * C + c] = acc;
    }
    __syncthreads();
    if (g == 0) {
        float acc = b2[c];
        for (int k = 0; k < G; ++k) acc += part[k * C + c];
        seS[(size_t)b * C + c] = sigmoidf_acc(acc);
    }
}

// ---- ST-joint attention (README.md:33-34) --------------------------------------------
// Position tiles of kStjPos: blockIdx.x < ntT covers frames, the rest joints, so one CTA
// uses one output matrix (Wt or Wv).  Register-tiled SIMT GEMMs: a work unit is one output
// column x 8 positions (two broadcast LDS.128 + one coalesced weight load per 8 FMAs).
// smem floats: C*(kStjPos+4) pooled (transposed, padded) + cj*kStjPos hidden.
constexpr int kStjPos = 32;
constexpr int kStjLd = kStjPos + 4;

template <int V>
__global__ void __launch_bounds__(256)
stj_kernel(const float *__restrict__ PT, const float *__restrict__ PV, const float *__restrict__ seS, int T,
           int C, int cj, int ntT, const float *__restrict__ W, const float *__restrict__ bW,
           const float *__restrict__ Wt, const float *__restrict__ bt, const float *__restrict__ Wv,
           const float *__restrict__ bv, float *__restrict__ gT, float *__restrict__ gV) {
    extern __shared__ __align__(16) float sm[];
    float *pooled = sm;                    // [C][kStjLd]   pooled[c][q]
    float *att = sm + (size_t)C * kStjLd;  // [cj][kStjPos] att[k][q]
    const int b = blockIdx.y;
    const bool is_t = (int)blockIdx.x < ntT;
    const int p0 = (is_t ? blockIdx.x : blockIdx.x - ntT) * kStjPos;
    const int np = min(kStjPos, (is_t ? T : V) - p0);
    const float inv = is_t ? 1.0f / (float)V : 1.0f / (float)T;
    const float *src = is_t ? PT + ((size_t)b * T + p0) * C : PV + ((size_t)b * V + p0) * C;
    for (int e = threadIdx.x; e < kStjPos * C; e += blockDim.x) {
        const int q = e / C, c = e - q * C;
        // oracle order: (sum / count) scaled by the SE gate
        pooled[c * kStjLd + q] = q < np ? seS[(size_t)b * C + c] * (src[(size_t)q * C + c] * inv) : 0.f;
    }
    __syncthreads();
    // phase 1: att[k][q] = hswish(bW[k] + sum_i pooled[i][q] W[i][k]); unit = 2 k x 8 positions
    for (int u = threadIdx.x; u < (cj / 2) * (kStjPos / 8); u += blockDim.x) {
        const int k = (u % (cj / 2)) * 2, q8 = (u / (cj / 2)) * 8;
        float acc[2][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.f;
        const float *pp = pooled + q8;
#pragma unroll 4
        for (int i = 0; i < C; ++i) {
            const float2 w = __ldg(reinterpret_cast<const float2 *>(W + (size_t)i * cj + k));
            const float4 a = *reinterpret_cast<const float4 *>(pp + i * kStjLd);
            const float4 d = *reinterpret_cast<const float4 *>(pp + i * kStjLd + 4);
            const float p[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                acc[0][e] += p[e] * w.x;
                acc[1][e] += p[e] * w.y;
            }
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const float bk = bW[k + kk];
            float4 o0, o1;
            o0.x = hardswishf(acc[kk][0] + bk); o0.y = hardswishf(acc[kk][1] + bk);
            o0.z = hardswishf(acc[kk][2] + bk); o0.w = hardswishf(acc[kk][3] + bk);
            o1.x = hardswishf(acc[kk][4] + bk); o1.y = hardswishf(acc[kk][5] + bk);
            o1.z = hardswishf(acc[kk][6] + bk); o1.w = hardswishf(acc[kk][7] + bk);
            *reinterpret_cast<float4 *>(att + (k + kk) * kStjPos + q8) = o0;
            *reinterpret_cast<float4 *>(att + (k + kk) * kStjPos + q8 + 4) = o1;
        }
    }
    __syncthreads();
    // phase 2: gate[q][c] = sigmoid(bo[c] + sum_k att[k][q] Wo[k][c]); unit = 4 channels x 8 positions
    const float *Wo = is_t ? Wt : Wv;
    const float *bo = is_t ? bt : bv;
    float *dst = is_t ? gT + ((size_t)b * T + p0) * C : gV + ((size_t)b * V + p0) * C;
    for (int u = threadIdx.x; u < (C / 4) * (kStjPos / 8); u += blockDim.x) {
        const int c = (u % (C / 4)) * 4, q8 = (u / (C / 4)) * 8;
        if (q8 >= np) continue;
        float acc[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
        const float *ap = att + q8;
#pragma unroll 2
        for (int k = 0; k < cj; ++k) {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(Wo + (size_t)k * C + c));
            const float4 a = *reinterpret_cast<const float4 *>(ap + k * kStjPos);
            const float4 d = *reinterpret_cast<const float4 *>(ap + k * kStjPos + 4);
            const float p[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                acc[0][e] += p[e] * w.x;
                acc[1][e] += p[e] * w.y;
                acc[2][e] += p[e] * w.z;
                acc[3][e] += p[e] * w.w;
            }
        }
        const float4 bc = __ldg(reinterpret_cast<const float4 *>(bo + c));
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (is_t) s4 = __ldg(reinterpret_cast<const float4 *>(seS + (size_t)b * C + c));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (q8 + e >= np) break;
            float4 o;
            o.x = s4.x * sigmoidf_acc(acc[0][e] + bc.x);
            o.y = s4.y * sigmoidf_acc(acc[1][e] + bc.y);
            o.z = s4.z * sigmoidf_acc(acc[2][e] + bc.z);
            o.w = s4.w * sigmoidf_acc(acc[3][e] + bc.w);
            *reinterpret_cast<float4 *>(dst + (size_t)(q8 + e) * C + c) = o;
        }
    }
}

// ---- ST-joint attention on warp-level tensor cores (bf16 path) ----------------------------------
// Same math as stj_kernel with both small GEMMs as mma.sync m16n8k16 FP16 (fp32 accumulate; fp16 has
// TF32's 10-bit mantissa, twice the K per instruction and half the shared-memory bytes; operands are
// clamped to the fp16 range when they are staged).  The
// launch is PERSISTENT: one CTA per SM keeps the first-layer matrix W and ONE second-layer matrix
// (Wt for the frame CTAs, Wv for the joint CTAs) resident in shared memory in B-fragment order and
// walks 64-position items, so the weights are read from L2 once per CTA instead of once per k-step
// and warp (the per-item form was bound by those loads: 168 us per C=256 launch for 5 GFLOP and
// 160 MB).  While an item is in the MMAs the next item's pooled rows are already in flight to
// registers.  Items: frame item = 64 frames of one clip; joint item = the 17 joints of 3 clips.
// 16 warps = 4 (16-row slabs) x 4 (column groups).  Weights are rounded to fp16 once at context
// creation, activations when they are staged.  The fp32 parity path keeps stj_kernel (exact fp32).
constexpr int kStjTcPos = 64;
constexpr int kStjThreads = 512;
constexpr int kStjClipsPerJointItem = 3;

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// A / B registers hold two fp16 each (consecutive k); D as for the TF32 form
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {     // (a -> low half), clamped to +-65504
    const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
    return *reinterpret_cast<const uint32_t *>(&h);
}
// 1 / (1 + 2^(-x log2 e)) on the SFU approximations (rel. error ~1e-6, the bf16 path's bar is 1e-2)
__device__ __forceinline__ float sigmoidf_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// Q = cj / 16 with C = 64 * Q (ST-joint reduction 4)
template <int Q>
struct StjCfg {
    static constexpr int C = 64 * Q, CJ = 16 * Q;
    static constexpr int WN1 = (CJ / 8 < 4) ? CJ / 8 : 4;   // column groups that work in the first GEMM
    static constexpr int NT1 = CJ / 8 / WN1;                // n-tiles per warp, first GEMM
    static constexpr int NT2 = 2 * Q;                       // n-tiles per warp, second GEMM (C/4 columns)
    static constexpr int LDP = C / 2 + 4, LDA = CJ / 2 + 4; // row strides in 32-bit words (2 fp16 each); +4: conflict-free A-fragment loads
    static constexpr int NPF = kStjTcPos * C / 4 / kStjThreads;   // float4 per thread and item
    static constexpr size_t smem_bytes =
        ((size_t)C * CJ + (size_t)kStjTcPos * (LDP + LDA) + CJ + 2 * C) * sizeof(float);
};

template <int V, int Q>
__global__ void __launch_bounds__(kStjThreads, Q <= 2 ? 2 : 1)
stj_tc_kernel(const float *__restrict__ PT, const float *__restrict__ PV, const float *__restrict__ seS, int B, int T,
              int ntT, int ctasT, const float *__restrict__ P1, const float *__restrict__ bW,
              const float *__restrict__ P2t, const float *__restrict__ bt, const float *__restrict__ P2v,
              const float *__restrict__ bv, float *__restrict__ gT, float *__restrict__ gV) {
    using Cf = StjCfg<Q>;
    constexpr int C = Cf::C, cj = Cf::CJ, NT1 = Cf::NT1, NT2 = Cf::NT2, WN1 = Cf::WN1, ldp = Cf::LDP, lda = Cf::LDA;
    constexpr int NPF = Cf::NPF, C4 = C / 4, KC = kStjClipsPerJointItem;
    extern __shared__ __align__(16) float sm[];
    float *w1 = sm;                                               // [C/16][WN1][NT1*2 words][32 lanes] fragments (2 fp16 per word)
    float *w2 = w1 + C * cj / 2;                                  // [cj/16][4][NT2/2 x 4 words][32 lanes]
    uint32_t *sp = reinterpret_cast<uint32_t *>(w2 + cj * C / 2); // pooled [64][ldp] fp16 pairs
    uint32_t *sa = sp + kStjTcPos * ldp;                          // hidden [64][lda] fp16 pairs
    float *s_b1 = reinterpret_cast<float *>(sa + kStjTcPos * lda);   // [cj] first-layer bias
    float *s_b2 = s_b1 + cj;                                      // [C]  second-layer bias
    float *s_se = s_b2 + C;                                       // [C]  SE gate of the item's clip (frame items)
    const bool is_t = (int)blockIdx.x < ctasT;
    const int first = is_t ? blockIdx.x : blockIdx.x - ctasT;
    const int stride = is_t ? ctasT : (int)gridDim.x - ctasT;
    const int nitems = is_t ? B * ntT : (B + KC - 1) / KC;
    const float inv = is_t ? 1.0f / (float)V : 1.0f / (float)T;
    {
        const float4 *g1 = reinterpret_cast<const float4 *>(P1);
        const float4 *g2 = reinterpret_cast<const float4 *>(is_t ? P2t : P2v);
        for (int e = threadIdx.x; e < C * cj / 8; e += kStjThreads) {
            reinterpret_cast<float4 *>(w1)[e] = __ldg(g1 + e);
            reinterpret_cast<float4 *>(w2)[e] = __ldg(g2 + e);
        }
        const float *bo = is_t ? bt : bv;
        for (int e = threadIdx.x; e < C; e += kStjThreads) s_b2[e] = bo[e];
        for (int e = threadIdx.x; e < cj; e += kStjThreads) s_b1[e] = bW[e];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 3, wn = warp >> 2, gr = lane >> 2, tg = lane & 3;
    const int r0 = wm * 16 + gr, r1 = r0 + 8;
    // staging map: float4 e = tid + k*512 -> row e / C4, channels 4*(e % C4) (the same channels for every k)
    const int srow0 = threadIdx.x / C4, sc = (threadIdx.x % C4) * 4;
    constexpr int srow_step = kStjThreads / C4;

    // An item is 64 rows.  Frame item: rows = frames p0.. of clip `clip0`.  Joint item: row r = joint r % 17
    // of clip clip0 + r / 17 (3 clips).  row_src returns the pooled row or nullptr for padding rows.
    auto decode = [&](int item, int &clip0, int &p0) {
        if (is_t) {
            clip0 = item / ntT;
            p0 = (item - clip0 * ntT) * kStjTcPos;
        } else {
            clip0 = item * KC;
            p0 = 0;
        }
    };
    auto row_index = [&](int clip0, int p0, int row) -> long long {   // row of PT/gT or PV/gV, -1: padding
        if (is_t) return p0 + row < T ? (long long)clip0 * T + p0 + row : -1;
        const int q = row / V;
        return (q < KC && clip0 + q < B) ? (long long)(clip0 + q) * V + (row - q * V) : -1;
    };
    float4 pre[NPF], pse[KC];
    auto prefetch = [&](int item) {
        int clip0, p0;
        decode(item, clip0, p0);
#pragma unroll
        for (int k = 0; k < NPF; ++k) {
            const long long ri = row_index(clip0, p0, srow0 + k * srow_step);
            pre[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ri >= 0) pre[k] = __ldg(reinterpret_cast<const float4 *>((is_t ? PT : PV) + ri * C + sc));
        }
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            pse[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((q == 0 || !is_t) && clip0 + q < B)
                pse[q] = __ldg(reinterpret_cast<const float4 *>(seS + (size_t)(clip0 + q) * C + sc));
        }
    };
    if (first < nitems) prefetch(first);
    for (int item = first; item < nitems; item += stride) {
        int clip0, p0;
        decode(item, clip0, p0);
        // stage the prefetched rows: oracle order (sum / count) scaled by the SE gate, rounded to fp16
#pragma unroll
        for (int k = 0; k < NPF; ++k) {
            const int row = srow0 + k * srow_step;
            float4 s4 = pse[0];
            if (!is_t) {
                const int q = row / V;
                s4 = q == 0 ? pse[0] : (q == 1 ? pse[1] : pse[2]);
            }
            const float4 v4 = pre[k];
            *reinterpret_cast<uint2 *>(sp + row * ldp + sc / 2) =
                make_uint2(pack_half2_sat(s4.x * (v4.x * inv), s4.y * (v4.y * inv)),
                           pack_half2_sat(s4.z * (v4.z * inv), s4.w * (v4.w * inv)));
        }
        if (srow0 == 0) *reinterpret_cast<float4 *>(s_se + sc) = pse[0];
        __syncthreads();
        if (item + stride < nitems) prefetch(item + stride);
        if (wn < WN1) {   // hidden = hswish(pooled . W + bW): this warp's 16 rows x NT1 n-tiles
            float acc[NT1][4];
#pragma unroll
            for (int t = 0; t < NT1; ++t)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
#pragma unroll 4
            for (int ks = 0; ks < C / 16; ++ks) {
                const int k0 = ks * 8;                 // in 32-bit words: 16 fp16 per k-step
                const uint32_t a0 = sp[r0 * ldp + k0 + tg], a1 = sp[r1 * ldp + k0 + tg];
                const uint32_t a2 = sp[r0 * ldp + k0 + tg + 4], a3 = sp[r1 * ldp + k0 + tg + 4];
                float bf[NT1 * 2];
                const float *pk = w1 + ((size_t)(ks * WN1 + wn) * 32 + lane) * (NT1 * 2);
                if constexpr (NT1 == 1) {
                    const float2 v = *reinterpret_cast<const float2 *>(pk);
                    bf[0] = v.x; bf[1] = v.y;
                } else {
                    const float4 v = *reinterpret_cast<const float4 *>(pk);
                    bf[0] = v.x; bf[1] = v.y; bf[2] = v.z; bf[3] = v.w;
                }
#pragma unroll
                for (int t = 0; t < NT1; ++t)
                    mma_f16(acc[t], a0, a1, a2, a3, __float_as_uint(bf[2 * t]), __float_as_uint(bf[2 * t + 1]));
            }
#pragma unroll
            for (int t = 0; t < NT1; ++t) {
                const int col = wn * (NT1 * 8) + t * 8 + 2 * tg;
                const float b0 = s_b1[col], b1 = s_b1[col + 1];
                sa[r0 * lda + col / 2] = pack_half2_sat(hardswishf(acc[t][0] + b0), hardswishf(acc[t][1] + b1));
                sa[r1 * lda + col / 2] = pack_half2_sat(hardswishf(acc[t][2] + b0), hardswishf(acc[t][3] + b1));
            }
        }
        __syncthreads();
        {   // gate = sigmoid(hidden . W2 + b): this warp's 16 rows x C/4 columns
            float acc[NT2][4];
#pragma unroll
            for (int t = 0; t < NT2; ++t)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < cj / 16; ++ks) {
                const int k0 = ks * 8;
                const uint32_t a0 = sa[r0 * lda + k0 + tg], a1 = sa[r1 * lda + k0 + tg];
                const uint32_t a2 = sa[r0 * lda + k0 + tg + 4], a3 = sa[r1 * lda + k0 + tg + 4];
                const float4 *pk = reinterpret_cast<const float4 *>(w2) + (size_t)(ks * 4 + wn) * (NT2 / 2) * 32 + lane;
#pragma unroll
                for (int i = 0; i < NT2 / 2; ++i) {
                    const float4 v = pk[i * 32];
                    mma_f16(acc[2 * i], a0, a1, a2, a3, __float_as_uint(v.x), __float_as_uint(v.y));
                    mma_f16(acc[2 * i + 1], a0, a1, a2, a3, __float_as_uint(v.z), __float_as_uint(v.w));
                }
            }
            const long long ri0 = row_index(clip0, p0, r0), ri1 = row_index(clip0, p0, r1);
            float *dst = is_t ? gT : gV;
#pragma unroll
            for (int t = 0; t < NT2; ++t) {
                const int col = wn * (NT2 * 8) + t * 8 + 2 * tg;
                const float2 bc = *reinterpret_cast<const float2 *>(s_b2 + col);
                float2 s2 = make_float2(1.f, 1.f);
                if (is_t) s2 = *reinterpret_cast<const float2 *>(s_se + col);
                if (ri0 >= 0)
                    *reinterpret_cast<float2 *>(dst + ri0 * C + col) =
                        make_float2(s2.x * sigmoidf_fast(acc[t][0] + bc.x), s2.y * sigmoidf_fast(acc[t][1] + bc.y));
                if (ri1 >= 0)
                    *reinterpret_cast<float2 *>(dst + ri1 * C + col) =
                        make_float2(s2.x * sigmoidf_fast(acc[t][2] + bc.x), s2.y * sigmoidf_fast(acc[t][3] + bc.y));
            }
        }
        __syncthreads();   // pooled / hidden / s_se are rewritten by the next item
    }
}

// Host: re-pack W [K][N] (row-major) into mma.sync m16n8k16 fp16 B-fragment order for `wn` column groups of
// N/wn columns (nt = N/wn/8 n-tiles each).  A lane's registers for one 16-deep k-step are
//   frag[t*2 + h] = {W[k][col], W[k+1][col]} as two fp16 (k in the low half),
//   k = ks*16 + 2*(lane&3) + 8h,  col = g*(N/wn) + t*8 + (lane>>2),  t < nt, h < 2,
// stored in chunks of `cw` words (cw = 2 when nt == 1, else 4) with the 32 lanes of a chunk contiguous:
//   out[(((ks*wn + g) * (2*nt/cw) + chunk) * 32 + lane) * cw + e] = frag[chunk*cw + e]
// so every warp-wide fragment load is one conflict-free 8- or 16-byte access per lane.  `out` carries
// the 32-bit words in float storage (K*N/2 of them).
inline void pack_stj_fragments(const float *W, int K, int N, int wn, std::vector<float> &out) {
    const int nt = N / wn / 8, cw = nt == 1 ? 2 : 4, nchunk = 2 * nt / cw;
    out.assign((size_t)K * N / 2, 0.f);
    for (int ks = 0; ks < K / 16; ++ks)
        for (int g = 0; g < wn; ++g)
            for (int lane = 0; lane < 32; ++lane)
                for (int t = 0; t < nt; ++t)
                    for (int h = 0; h < 2; ++h) {
                        const int f = t * 2 + h, chunk = f / cw, e = f % cw;
                        const int k = ks * 16 + 2 * (lane & 3) + 8 * h;
                        const int col = g * (N / wn) + t * 8 + (lane >> 2);
                        const __half lo = __float2half_rn(W[(size_t)k * N + col]);
                        const __half hi = __float2half_rn(W[(size_t)(k + 1) * N + col]);
                        const uint32_t word = (uint32_t)(*reinterpret_cast<const unsigned short *>(&lo)) |
                                              ((uint32_t)(*reinterpret_cast<const unsigned short *>(&hi)) << 16);
                        memcpy(&out[((((size_t)ks * wn + g) * nchunk + chunk) * 32 + lane) * cw + e], &word, 4);
                    }
}

// ---- head (README.md:17-18) -----------------------------------------------------------
// logits[b,t,k] = sum_c (gT[b,t,c] * sum_v U[b,t,v,c]*gV[b,v,c] / V) * Wh[c,k] + bh[k]
// gV[b] and the transposed head weights WhT [K][C] are staged in shared memory once per CTA of
// kHeadFrames frames of one clip.  C % 8 == 0, K <= 32.
constexpr int kHeadFrames = 32;

template <typename TU> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    typedef uint4 Raw;     // 8 channels as loaded: kept packed in registers until they are used
    static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16 *p) {
        return __ldg(reinterpret_cast<const uint4 *>(p));
    }
    static __device__ __forceinline__ void unpack(const Raw &r, float (&f)[8]) {
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 t = __bfloat1622float2(h[e]);
            f[2 * e] = t.x;
            f[2 * e + 1] = t.y;
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[8]) {
        const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 t = __bfloat1622float2(h[e]);
            f[2 * e] = t.x;
            f[2 * e + 1] = t.y;
        }
    }
};
template <> struct Vec8<float> {
    struct Raw { float4 a, d; };
    static __device__ __forceinline__ Raw load_raw(const float *p) {
        Raw r;
        r.a = __ldg(reinterpret_cast<const float4 *>(p));
        r.d = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        return r;
    }
    static __device__ __forceinline__ void unpack(const Raw &r, float (&f)[8]) {
        f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.d.x; f[5] = r.d.y; f[6] = r.d.z; f[7] = r.d.w;
    }
    static __device__ __forceinline__ void load(const float *p, float (&f)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
        const float4 d = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = d.x; f[5] = d.y; f[6] = d.z; f[7] = d.w;
    }
};

// Generic form (any C % 8 == 0, K <= 32; the fp32 parity path): loads as the compiler schedules them.
// A warp owns 4 frames: lane = (frame f = lane / 8, channel group g = lane % 8); a lane walks the 8-channel
// vectors g, g+8, ... of its frame, folds them straight into K logit partials, and the partials are reduced
// over the 8 lanes of the frame (3 shuffle steps for 4 frames at once).  kHeadFrames = 32 frames per CTA.
template <typename TU, int V>
__global__ void __launch_bounds__(256)
head_kernel(const TU *__restrict__ U, const float *__restrict__ gT, const float *__restrict__ gV, int T, int C,
            int K, const float *__restrict__ WhT, const float *__restrict__ bh, float *__restrict__ logits,
            uint8_t *__restrict__ labels) {
    extern __shared__ __align__(16) float sm[];     // gV[b]: [V][C], then WhT: [K][C]
    float *sgv = sm, *swh = sm + V * C;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x * 4; e < V * C; e += blockDim.x * 4)
        *reinterpret_cast<float4 *>(sgv + e) = __ldg(reinterpret_cast<const float4 *>(gV + (size_t)b * V * C + e));
    for (int e = threadIdx.x * 4; e < K * C; e += blockDim.x * 4)
        *reinterpret_cast<float4 *>(swh + e) = __ldg(reinterpret_cast<const float4 *>(WhT + e));
    __syncthreads();
    const int f = lane >> 3, g = lane & 7;
    const int t = blockIdx.x * kHeadFrames + warp * 4 + f;
    const bool tv = t < T;
    const size_t bt = (size_t)b * T + (tv ? t : T - 1);
    float part[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) part[k] = 0.f;
    for (int c0 = g * 8; c0 < C; c0 += 64) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const TU *row = U + (bt * V) * C + c0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float x[8];
            Vec8<TU>::load(row + (size_t)v * C, x);
            const float4 g0 = *reinterpret_cast<const float4 *>(sgv + v * C + c0);
            const float4 g1 = *reinterpret_cast<const float4 *>(sgv + v * C + c0 + 4);
            acc[0] += x[0] * g0.x; acc[1] += x[1] * g0.y; acc[2] += x[2] * g0.z; acc[3] += x[3] * g0.w;
            acc[4] += x[4] * g1.x; acc[5] += x[5] * g1.y; acc[6] += x[6] * g1.z; acc[7] += x[7] * g1.w;
        }
        float gt[8];
        Vec8<float>::load(gT + bt * C + c0, gt);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = acc[e] * gt[e] * (1.0f / (float)V);
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (k < K) {
                const float4 w0 = *reinterpret_cast<const float4 *>(swh + k * C + c0);
                const float4 w1 = *reinterpret_cast<const float4 *>(swh + k * C + c0 + 4);
                part[k] += acc[0] * w0.x + acc[1] * w0.y + acc[2] * w0.z + acc[3] * w0.w + acc[4] * w1.x +
                           acc[5] * w1.y + acc[6] * w1.z + acc[7] * w1.w;
            }
    }
    // reduce over the 8 channel-group lanes of each frame; lane g == 0 of a frame ends with the totals
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k < K) {
            float x = part[k];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 2);
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            x += bh[k];
            if (g == 0 && tv) logits[bt * K + k] = x;
            if (k == 0 || x > best) { best = x; arg = k; }      // first arg-max: ties -> lowest class index
        }
    if (labels && g == 0 && tv) labels[bt] = (uint8_t)arg;
}

// Streaming form (bf16 path, C % 64 == 0, K <= 16).
// A warp owns 4 frames: lane = (frame f = lane / 8, channel group g = lane % 8); a lane walks the 8-channel
// vectors g, g+8, ... of its frame, folds them straight into K logit partials, and the partials are reduced
// over the 8 lanes of the frame (3 shuffle steps for 4 frames at once).  kHeadFrames = 32 frames per CTA.
// The launch is HBM-bound (one pass over U), so what matters is bytes in flight: the 17 joint vectors of
// an 8-channel group are loaded as one batch of 16-byte loads before any of them is used (the 80-register
// build issued them 4-5 at a time: 3.3 TB/s), and the gate / weight rows sit in shared memory in a
// lane-interleaved order so that the 8 lanes of a frame read 128 contiguous bytes per LDS.128
// (channel c of a 64-channel group lives at float (c%8/4)*32 + (c/8)*4 + c%4; the plain order was a
// 2-way bank conflict on every gate load).
constexpr int kHeadMaxK = 16;

__device__ __forceinline__ int head_perm(int c) {    // position of channel c inside its 64-channel group
    const int w = c & 63;
    return (c & ~63) + ((w & 7) >> 2) * 32 + (w >> 3) * 4 + (w & 3);
}

template <typename TU, int V>
__global__ void __launch_bounds__(256, 2)
head_stream_kernel(const TU *__restrict__ U, const float *__restrict__ gT, const float *__restrict__ gV, int T, int C,
                   int K, const float *__restrict__ WhT, const float *__restrict__ bh, float *__restrict__ logits,
                   uint8_t *__restrict__ labels) {
    extern __shared__ __align__(16) float sm[];     // gV[b]: [V][C], then WhT: [K][C], both lane-interleaved
    float *sgv = sm, *swh = sm + V * C;
    // last clips first: the producing kernel wrote them last, so they are the ones still in L2
    const int b = gridDim.y - 1 - blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = lane >> 3, g = lane & 7;
    const int t = blockIdx.x * kHeadFrames + warp * 4 + f;
    const bool tv = t < T;
    const size_t bt = (size_t)b * T + (tv ? t : T - 1);
    const TU *row = U + (bt * V) * C + g * 8;
    const float *gtrow = gT + bt * C + g * 8;
    // the first batch of activation loads goes out before the gate / weight rows are staged, so HBM is
    // busy while this CTA fills its shared memory
    typename Vec8<float>::Raw gtr = Vec8<float>::load_raw(gtrow);
    typename Vec8<TU>::Raw xr[V];
#pragma unroll
    for (int v = 0; v < V; ++v) xr[v] = Vec8<TU>::load_raw(row + (size_t)v * C);
    for (int e = threadIdx.x * 4; e < V * C; e += blockDim.x * 4) {
        const int r = e / C, c = e - r * C;
        *reinterpret_cast<float4 *>(sgv + r * C + head_perm(c)) =
            __ldg(reinterpret_cast<const float4 *>(gV + (size_t)b * V * C + e));
    }
    for (int e = threadIdx.x * 4; e < K * C; e += blockDim.x * 4) {
        const int r = e / C, c = e - r * C;
        *reinterpret_cast<float4 *>(swh + r * C + head_perm(c)) = __ldg(reinterpret_cast<const float4 *>(WhT + e));
    }
    __syncthreads();
    float part[kHeadMaxK];
#pragma unroll
    for (int k = 0; k < kHeadMaxK; ++k) part[k] = 0.f;
    for (int cg = 0; cg < C; cg += 64) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const float *gbase = sgv + cg + g * 4;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 g0 = *reinterpret_cast<const float4 *>(gbase + v * C);
            const float4 g1 = *reinterpret_cast<const float4 *>(gbase + v * C + 32);
            float x[8];
            Vec8<TU>::unpack(xr[v], x);
            acc[0] += x[0] * g0.x; acc[1] += x[1] * g0.y; acc[2] += x[2] * g0.z; acc[3] += x[3] * g0.w;
            acc[4] += x[4] * g1.x; acc[5] += x[5] * g1.y; acc[6] += x[6] * g1.z; acc[7] += x[7] * g1.w;
        }
        {
            float gt[8];
            Vec8<float>::unpack(gtr, gt);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = acc[e] * gt[e] * (1.0f / (float)V);
        }
        if (cg + 64 < C) {                       // next batch in flight while the logit partials are folded
            gtr = Vec8<float>::load_raw(gtrow + cg + 64);
#pragma unroll
            for (int v = 0; v < V; ++v) xr[v] = Vec8<TU>::load_raw(row + (size_t)v * C + cg + 64);
        }
        const float *wbase = swh + cg + g * 4;
#pragma unroll
        for (int k = 0; k < kHeadMaxK; ++k)
            if (k < K) {
                const float4 w0 = *reinterpret_cast<const float4 *>(wbase + k * C);
                const float4 w1 = *reinterpret_cast<const float4 *>(wbase + k * C + 32);
                part[k] += acc[0] * w0.x + acc[1] * w0.y + acc[2] * w0.z + acc[3] * w0.w + acc[4] * w1.x +
                           acc[5] * w1.y + acc[6] * w1.z + acc[7] * w1.w;
            }
    }
    // reduce over the 8 channel-group lanes of each frame; lane g == 0 of a frame ends with the totals
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < kHeadMaxK; ++k)
        if (k < K) {
            float x = part[k];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 2);
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            x += bh[k];
            if (g == 0 && tv) logits[bt * K + k] = x;
            if (k == 0 || x > best) { best = x; arg = k; }      // first arg-max: ties -> lowest class index
        }
    if (labels && g == 0 && tv) labels[bt] = (uint8_t)arg;
}

// out[b,t,v,c] = U * gT * gV  (fp32; debug / parity hook)
template <typename TU>
__global__ void __launch_bounds__(256)
features_kernel(const TU *__restrict__ U, const float *__restrict__ gT, const float *__restrict__ gV, int T, int V,
                int C, size_t total, float *__restrict__ out) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const size_t row = e / C;
        const int v = (int)(row % V);
        const size_t bt = row / V;
        const size_t b = bt / T;
        out[e] = ld_act(U + e) * gT[bt * C + c] * gV[(b * V + v) * C + c];
    }
}

// Host-side launcher for the attention tail of one block (stats -> SE -> ST-joint).
// `stj_packed` (bf16 path): fragment-ordered fp16 copies of {W, Wt, Wv} for stj_tc_kernel.
// `fused_nchunk` > 0: PT and that many PVpart partials per clip were already written by the producing
// kernel's epilogue (tcn_fused.cuh); 0: run stats_kernel here.
template <typename TU>
int launch_attention(Ctx *ctx, const BlockParams &bp, const TU *U, int B, int T, cudaStream_t st,
                     int fused_nchunk = 0, const float *const *stj_packed = nullptr) {
    constexpr int V = 17;
    const int C = bp.c;
    const int nchunk = fused_nchunk > 0 ? fused_nchunk : cdiv(T, kStatChunk);
    if (fused_nchunk == 0) {
        dim3 grid(nchunk, B, cdiv(C, 128));
        {
            LaunchScope ls(ctx, K_STATS, st, 2.0 * B * T * V * C, (double)B * T * V * C * sizeof(TU));
            stats_kernel<TU, V><<<grid, 128, 0, st>>>(U, T, C, ctx->PT, ctx->PVpart);
        }
        GS_KERNEL_CHECK();
    }
    {
        const int G = std::max(1, std::min(1024 / C, V));
        const size_t smem = ((size_t)std::max(G * C, 1024) + C + bp.cs) * sizeof(float);
        {
            LaunchScope ls(ctx, K_SE, st, 4.0 * B * C * bp.cs, (double)B * nchunk * V * C * 4);
            se_kernel<V><<<B, 1024, smem, st>>>(ctx->PVpart, T, C, bp.cs, nchunk, G, bp.seW1, bp.seb1, bp.seW2,
                                                bp.seb2, ctx->seS, ctx->PV);
        }
        GS_KERNEL_CHECK();
    }
    if (stj_packed && C == 4 * bp.cj && (C == 64 || C == 128 || C == 256)) {
        const int ntT = cdiv(T, kStjTcPos);
        const int itemsT = B * ntT, itemsV = cdiv(B, kStjClipsPerJointItem);
        // two CTAs per SM at C <= 128 (17 / 40 KB of shared memory, 64 registers): an item is a chain of short phases
        // between CTA barriers, and a second CTA fills the gaps
        int grid = std::min(ctx->sm_count * (C <= 128 ? 2 : 1), itemsT + itemsV);
        int ctasV = (int)((double)grid * itemsV / (itemsT + itemsV) + 0.5);
        ctasV = std::max(1, std::min(ctasV, grid - 1));
        if (grid < 2) { grid = 2; ctasV = 1; }
        const size_t smem = C == 64 ? StjCfg<1>::smem_bytes : (C == 128 ? StjCfg<2>::smem_bytes : StjCfg<4>::smem_bytes);
        auto kern = C == 64 ? stj_tc_kernel<V, 1> : (C == 128 ? stj_tc_kernel<V, 2> : stj_tc_kernel<V, 4>);
        if (smem > 48 * 1024) GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            LaunchScope ls(ctx, K_STJ, st, 4.0 * B * (T + V) * C * bp.cj, 2.0 * B * (T + V) * C * 4);
            kern<<<grid, kStjThreads, smem, st>>>(ctx->PT, ctx->PV, ctx->seS, B, T, ntT, grid - ctasV, stj_packed[0], bp.jb,
                                                  stj_packed[1], bp.jbt, stj_packed[2], bp.jbv, ctx->gT, ctx->gV);
        }
        GS_KERNEL_CHECK();
    } else {
        const int ntT = cdiv(T, kStjPos);
        dim3 grid(ntT + cdiv(V, kStjPos), B);
        const size_t smem = ((size_t)C * kStjLd + (size_t)bp.cj * kStjPos) * sizeof(float);
        if (smem > 48 * 1024)
            GS_CUDA(cudaFuncSetAttribute(stj_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            LaunchScope ls(ctx, K_STJ, st, 4.0 * B * (T + V) * C * bp.cj, 2.0 * B * (T + V) * C * 4);
            stj_kernel<V><<<grid, 256, smem, st>>>(ctx->PT, ctx->PV, ctx->seS, T, C, bp.cj, ntT, bp.jW, bp.jb,
                                                   bp.jWt, bp.jbt, bp.jWv, bp.jbv, ctx->gT, ctx->gV);
        }
        GS_KERNEL_CHECK();
    }
    return GS_OK;
}

template <typename TU>
int launch_head(Ctx *ctx, const TU *U, int B, int T, int C, float *logits, uint8_t *labels, cudaStream_t st) {
    constexpr int V = 17;
    const int K = ctx->cfg.num_classes;
    const size_t smem = (size_t)(V + K) * C * sizeof(float);
    if (sizeof(TU) == 2 && K <= kHeadMaxK && C % 64 == 0) {
        if (smem > 48 * 1024)
            GS_CUDA(cudaFuncSetAttribute(head_stream_kernel<TU, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            LaunchScope ls(ctx, K_HEAD, st, 2.0 * B * T * C * (V + K), (double)B * T * V * C * sizeof(TU));
            dim3 grid(cdiv(T, kHeadFrames), B);
            head_stream_kernel<TU, V><<<grid, 256, smem, st>>>(U, ctx->gT, ctx->gV, T, C, K, ctx->headWT, ctx->headb,
                                                               logits, labels);
        }
        GS_KERNEL_CHECK();
        return GS_OK;
    }
    if (smem > 48 * 1024)
        GS_CUDA(cudaFuncSetAttribute(head_kernel<TU, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        LaunchScope ls(ctx, K_HEAD, st, 2.0 * B * T * C * (V + K), (double)B * T * V * C * sizeof(TU));
        dim3 grid(cdiv(T, kHeadFrames), B);
        head_kernel<TU, V><<<grid, 256, smem, st>>>(U, ctx->gT, ctx->gV, T, C, K, ctx->headWT, ctx->headb, logits,
                                                    labels);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

template <typename TU>
int launch_features(Ctx *ctx, const TU *U, int B, int T, int C, float *out, cudaStream_t st) {
    const size_t total = (size_t)B * T * 17 * C;
    int grid = (int)((total + 255) / 256 < (size_t)ctx->sm_count * 16 ? (total + 255) / 256
                                                                       : (size_t)ctx->sm_count * 16);
    {
        LaunchScope ls(ctx, K_FEAT, st);
        features_kernel<TU><<<grid, 256, 0, st>>>(U, ctx->gT, ctx->gV, T, 17, C, total, out);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace gs

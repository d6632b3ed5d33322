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

// grid B, block = C threads (C <= 1024).  smem: C + cs floats.
template <int V>
__global__ void __launch_bounds__(1024)
se_kernel(const float *__restrict__ PT, const float *__restrict__ PVpart, int T, int C, int cs, int nchunk,
          const float *__restrict__ W1, const float *__restrict__ b1, const float *__restrict__ W2,
          const float *__restrict__ b2, float *__restrict__ seS, float *__restrict__ PV) {
    extern __shared__ float sm[];
    float *m = sm, *hid = sm + C;
    const int b = blockIdx.x, c = threadIdx.x;
    if (c < C) {
        float acc = 0.f;
        for (int t = 0; t < T; ++t) acc += PT[((size_t)b * T + t) * C + c];
        m[c] = acc / (float)(T * V);
        for (int v = 0; v < V; ++v) {
            float s = 0.f;
            for (int k = 0; k < nchunk; ++k) s += PVpart[(((size_t)b * nchunk + k) * V + v) * C + c];
            PV[((size_t)b * V + v) * C + c] = s;
        }
    }
    __syncthreads();
    if (c < cs) {
        float acc = b1[c];
        for (int k = 0; k < C; ++k) acc += m[k] * W1[k * cs + c];
        hid[c] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (c < C) {
        float acc = b2[c];
        for (int k = 0; k < cs; ++k) acc += hid[k] * W2[k * C + c];
        seS[(size_t)b * C + c] = sigmoidf_acc(acc);
    }
}

constexpr int kStjPos = 8;   // positions (frames or joints) per CTA

// grid (ceil((T+V)/kStjPos), B), block = C threads.  smem: kStjPos*(C+cj) floats.
template <int V>
__global__ void __launch_bounds__(1024)
stj_kernel(const float *__restrict__ PT, const float *__restrict__ PV, const float *__restrict__ seS, int T,
           int C, int cj, const float *__restrict__ W, const float *__restrict__ bW,
           const float *__restrict__ Wt, const float *__restrict__ bt, const float *__restrict__ Wv,
           const float *__restrict__ bv, float *__restrict__ gT, float *__restrict__ gV) {
    extern __shared__ float sm[];
    float *pooled = sm;                 // [kStjPos][C]
    float *att = sm + kStjPos * C;      // [kStjPos][cj]
    const int b = blockIdx.y, c = threadIdx.x;
    const int p0 = blockIdx.x * kStjPos;
    const int np = min(kStjPos, T + V - p0);
    const float s = c < C ? seS[(size_t)b * C + c] : 0.f;
    if (c < C) {
        for (int q = 0; q < np; ++q) {
            const int pos = p0 + q;
            float x;
            if (pos < T) x = PT[((size_t)b * T + pos) * C + c] / (float)V;
            else x = PV[((size_t)b * V + (pos - T)) * C + c] / (float)T;
            pooled[q * C + c] = s * x;
        }
    }
    __syncthreads();
    for (int e = c; e < np * cj; e += blockDim.x) {
        const int q = e / cj, k = e % cj;
        float acc = bW[k];
        for (int i = 0; i < C; ++i) acc += pooled[q * C + i] * W[i * cj + k];
        att[q * cj + k] = hardswishf(acc);
    }
    __syncthreads();
    if (c < C) {
        for (int q = 0; q < np; ++q) {
            const int pos = p0 + q;
            const bool is_t = pos < T;
            const float *Wo = is_t ? Wt : Wv;
            float acc = is_t ? bt[c] : bv[c];
            for (int k = 0; k < cj; ++k) acc += att[q * cj + k] * Wo[k * C + c];
            const float g = sigmoidf_acc(acc);
            if (is_t) gT[((size_t)b * T + pos) * C + c] = s * g;
            else gV[((size_t)b * V + (pos - T)) * C + c] = g;
        }
    }
}

// logits[b,t,k] = sum_c (gT[b,t,c] * sum_v U[b,t,v,c]*gV[b,v,c] / V) * Wh[c,k] + bh[k]
// one CTA per frame, C threads; K <= 32.  smem: (C/32+1)*K floats
template <typename TU, int V>
__global__ void __launch_bounds__(1024)
head_kernel(const TU *__restrict__ U, const float *__restrict__ gT, const float *__restrict__ gV, int T, int C,
            int K, const float *__restrict__ Wh, const float *__restrict__ bh, float *__restrict__ logits,
            uint8_t *__restrict__ labels) {
    extern __shared__ float sm[];       // [nwarps][K]
    const int bt = blockIdx.x;
    const int b = bt / T;
    const int c = threadIdx.x;
    const int lane = c & 31, warp = c >> 5, nwarps = (blockDim.x + 31) >> 5;
    float pooled = 0.f;
    if (c < C) {
        const TU *row = U + ((size_t)bt * V) * C + c;
        const float *gv = gV + ((size_t)b * V) * C + c;
        float acc = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) acc += ld_act(row + (size_t)v * C) * gv[(size_t)v * C];
        pooled = acc * gT[(size_t)bt * C + c] / (float)V;
    }
    for (int k = 0; k < K; ++k) {
        float x = c < C ? pooled * Wh[c * K + k] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[warp * K + k] = x;
    }
    __syncthreads();
    if (c == 0) {
        float best = 0.f;
        int arg = 0;
        for (int k = 0; k < K; ++k) {
            float acc = bh[k];
            for (int w = 0; w < nwarps; ++w) acc += sm[w * K + k];
            logits[(size_t)bt * K + k] = acc;
            if (k == 0 || acc > best) { best = acc; arg = k; }
        }
        if (labels) labels[bt] = (uint8_t)arg;
    }
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
template <typename TU>
int launch_attention(Ctx *ctx, const BlockParams &bp, const TU *U, int B, int T, cudaStream_t st) {
    constexpr int V = 17;
    const int C = bp.c;
    const int nchunk = cdiv(T, kStatChunk);
    {
        dim3 grid(nchunk, B, cdiv(C, 128));
        {
            LaunchScope ls(ctx, K_STATS, st, 2.0 * B * T * V * C, (double)B * T * V * C * sizeof(TU));
            stats_kernel<TU, V><<<grid, 128, 0, st>>>(U, T, C, ctx->PT, ctx->PVpart);
        }
        GS_KERNEL_CHECK();
    }
    {
        const int threads = ((C + 31) / 32) * 32;
        {
            LaunchScope ls(ctx, K_SE, st, 4.0 * B * C * bp.cs, (double)B * T * C * 4);
            se_kernel<V><<<B, threads, (C + bp.cs) * sizeof(float), st>>>(
                ctx->PT, ctx->PVpart, T, C, bp.cs, nchunk, bp.seW1, bp.seb1, bp.seW2, bp.seb2, ctx->seS,
                ctx->PV);
        }
        GS_KERNEL_CHECK();
    }
    {
        const int threads = ((C + 31) / 32) * 32;
        dim3 grid(cdiv(T + V, kStjPos), B);
        {
            LaunchScope ls(ctx, K_STJ, st, 4.0 * B * (T + V) * C * bp.cj, 2.0 * B * (T + V) * C * 4);
            stj_kernel<V><<<grid, threads, kStjPos * (C + bp.cj) * sizeof(float), st>>>(
                ctx->PT, ctx->PV, ctx->seS, T, C, bp.cj, bp.jW, bp.jb, bp.jWt, bp.jbt, bp.jWv, bp.jbv, ctx->gT,
                ctx->gV);
        }
        GS_KERNEL_CHECK();
    }
    return GS_OK;
}

template <typename TU>
int launch_head(Ctx *ctx, const TU *U, int B, int T, int C, float *logits, uint8_t *labels, cudaStream_t st) {
    constexpr int V = 17;
    const int K = ctx->cfg.num_classes;
    const int threads = ((C + 31) / 32) * 32;
    {
        LaunchScope ls(ctx, K_HEAD, st, 2.0 * B * T * C * (V + K), (double)B * T * V * C * sizeof(TU));
        head_kernel<TU, V><<<B * T, threads, (threads / 32) * K * sizeof(float), st>>>(
            U, ctx->gT, ctx->gV, T, C, K, ctx->headW, ctx->headb, logits, labels);
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

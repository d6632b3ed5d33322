// fp32 parity path of the segmentation network: fp32 storage, fp32 CUDA-core math.
// This is the path held to 1e-5 against oracle/segnet.py (north_star "fp32 path");
// the throughput path is segment_bf16.cu.
//
// Stages replaced: /root/reference/README.md:27-28 (graph convolution), 29-30
// (multi-branch temporal convolution); attention + head in segment_common.cuh.
#include "segment_common.cuh"

namespace gs {

namespace {

// ---- SIMT row GEMM: Out[M,N] = act(In[M,K] @ W[K,N] + bias) ------------------------
// 64x64 tile, BK=16, 256 threads, 4x4 outputs per thread; any M,N,K (guarded).
template <bool RELU>
__global__ void __launch_bounds__(256)
gemm_rows_kernel(const float *__restrict__ In, const float *__restrict__ W, const float *__restrict__ bias,
                 float *__restrict__ Out, size_t M, int N, int K) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float sIn[BK][BM + 4];
    __shared__ float sW[BK][BN + 4];
    const size_t m0 = (size_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        for (int e = threadIdx.x; e < BM * BK; e += 256) {
            const int r = e / BK, k = e % BK;
            const size_t m = m0 + r;
            sIn[k][r] = (m < M && k0 + k < K) ? In[m * K + k0 + k] : 0.f;
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            const int k = e / BN, nn = e % BN;
            sW[k][nn] = (k0 + k < K && n0 + nn < N) ? W[(size_t)(k0 + k) * N + n0 + nn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sIn[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = sW[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn >= N) continue;
            float v = acc[i][j] + bias[nn];
            if (RELU) v = fmaxf(v, 0.f);
            Out[m * N + nn] = v;
        }
    }
}

// ---- multi-branch dilated temporal conv + residual + ReLU -----------------------
// U[row, r*cr+co] = relu( sum_{j<3, ci<cr} H[row + (j-1)*d_r*V, r*cr+ci] * W2[r,j,ci,co]
//                         + b2 + Res[row, r*cr+co] ),   zero outside [0,T) of the clip.
// one thread per output element; a CTA covers 256/C... rows x all C channels.
struct Dil { int d[GS_MAX_BRANCHES]; };

__global__ void __launch_bounds__(256)
tconv_kernel(const float *__restrict__ H, const float *__restrict__ Res, const float *__restrict__ W2,
             const float *__restrict__ b2, float *__restrict__ U, int T, int C, int cr, Dil dil,
             size_t total) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        const int co_full = (int)(e % C);
        const size_t row = e / C;
        const int r = co_full / cr, co = co_full % cr;
        const int t = (int)((row / V17) % T);
        const int d = dil.d[r];
        float acc = b2[co_full];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int ts = t + (j - 1) * d;
            if (ts < 0 || ts >= T) continue;
            const float *h = H + ((ptrdiff_t)row + (ptrdiff_t)(j - 1) * d * V17) * C + r * cr;
            const float *w = W2 + ((size_t)(r * 3 + j) * cr) * cr + co;
            for (int ci = 0; ci < cr; ++ci) acc += h[ci] * w[(size_t)ci * cr];
        }
        acc += Res[e];
        U[e] = fmaxf(acc, 0.f);
    }
}

int grid_for(const Ctx *ctx, size_t total, int threads) {
    size_t g = (total + threads - 1) / threads;
    const size_t cap = (size_t)ctx->sm_count * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <bool RELU>
int launch_gemm(Ctx *ctx, int kid, const float *In, const float *W, const float *bias, float *Out, size_t M,
                int N, int K, cudaStream_t st) {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
    {
        LaunchScope ls(ctx, kid, st, 2.0 * M * N * K, 4.0 * M * (N + K));
        gemm_rows_kernel<RELU><<<grid, 256, 0, st>>>(In, W, bias, Out, M, N, K);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace

int segment_fp32_forward(Ctx *ctx, const float *skel, float *logits, uint8_t *labels, int B, int T,
                         int upto_block, float *feat_out, cudaStream_t st) {
    const size_t nframes = (size_t)B * T;
    const size_t rows = nframes * V17;
    const int nb = ctx->cfg.num_blocks;
    const int last = upto_block >= 0 ? upto_block : nb - 1;
    float *X = (float *)ctx->bufX, *XA = (float *)ctx->bufXA, *Y = (float *)ctx->bufY;
    float *H = (float *)ctx->bufH, *R = (float *)ctx->bufR;
    Dil dil;
    for (int r = 0; r < GS_MAX_BRANCHES; ++r) dil.d[r] = ctx->cfg.dilations[r];
    const float *Uprev = nullptr;
    int rc;
    for (int i = 0; i <= last; ++i) {
        ctx->cur_block = i;
        const BlockParams &bp = ctx->blocks[i];
        float *U = (float *)ctx->bufU[i & 1];
        const size_t items = nframes * bp.cin;
        {
            LaunchScope ls(ctx, K_AGG, st, 2.0 * rows * V17 * 3 * bp.cin, 4.0 * rows * bp.cin * 5);
            if (i == 0) {
                aggregate_kernel<float, float, 3><<<grid_for(ctx, items, 256), 256, 0, st>>>(
                    skel, nullptr, nullptr, ctx->in_scale, ctx->in_shift, bp.A, T, bp.cin, nframes, X, XA);
            } else {
                aggregate_kernel<float, float, 3><<<grid_for(ctx, items, 256), 256, 0, st>>>(
                    Uprev, ctx->gT, ctx->gV, nullptr, nullptr, bp.A, T, bp.cin, nframes, X, XA);
            }
        }
        GS_KERNEL_CHECK();
        if ((rc = launch_gemm<true>(ctx, K_GEMM_GCN, XA, bp.Wg, bp.bg, Y, rows, bp.c, 3 * bp.cin, st))) return rc;
        if ((rc = launch_gemm<true>(ctx, K_GEMM_TCN1, Y, bp.W1, bp.b1, H, rows, bp.c, bp.c, st))) return rc;
        const float *res = X;
        if (bp.has_res) {
            if ((rc = launch_gemm<false>(ctx, K_GEMM_RES, X, bp.Wr, bp.br, R, rows, bp.c, bp.cin, st))) return rc;
            res = R;
        }
        const size_t total = rows * bp.c;
        {
            LaunchScope ls(ctx, K_TCONV, st, 2.0 * total * 3 * bp.cr, 4.0 * total * 3);
            tconv_kernel<<<grid_for(ctx, total, 256), 256, 0, st>>>(H, res, bp.W2, bp.b2, U, T, bp.c, bp.cr, dil,
                                                                    total);
        }
        GS_KERNEL_CHECK();
        if ((rc = launch_attention<float>(ctx, bp, U, B, T, st))) return rc;
        Uprev = U;
    }
    ctx->cur_block = GS_MAX_BLOCKS;
    const int C = ctx->blocks[last].c;
    if (feat_out) return launch_features<float>(ctx, Uprev, B, T, C, feat_out, st);
    return launch_head<float>(ctx, Uprev, B, T, C, logits, labels, st);
}

}  // namespace gs

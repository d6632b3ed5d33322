// Learned alignment embedding (/root/reference/README.md:44-47: the alignment model is trained).
// Contract: oracle/embed.py (AlignEmbedConfig v0: per-frame MLP 34 -> 128 -> 128 over the (x, y) of the 17 joints,
// cost c[i,j] = sqrt(max(|fa_i|^2 + |fb_j|^2 - 2 fa_i . fb_j, 0)), then the DP / tie-break / backtrack of align.cu).
//
//   embed_encoder_kernel   frames -> F [rows, 128] bf16 (the GEMM operand) + |F|^2 fp32 of the ROUNDED values
//                          (CUDA cores, fp32: 20.7 kFMA per frame, weights in shared memory, 4 frames x 8 outputs per thread)
//   embed_cost_kernel      one CTA per (pair, 128-row tile, 128-column tile): D = Fa_tile . Fb_tile^T on tcgen05
//                          (K = 128: 8 MMAs of M = N = 128, operands by TMA, accumulator in TMEM), epilogue
//                          sqrt(max(na + nb - 2 D, 0)) staged through the (now dead) operand buffers so that rows leave
//                          as 512-byte coalesced segments -> cost matrix [N, ra, rb] fp32
//   dtw_costmat_kernel     align.cu: the pipelined DP sweep reading that matrix; dtw_backtrack_kernel
// This is the one place where `align` is tensor-core work.  Parity policy: oracle/embed.py header (cost within 1e-2
// of the fp32 oracle / 1e-3 of its bf16 emulation; DP and path bit-exact on the GPU's own cost matrix).
#include "umma.cuh"

namespace gs {

constexpr int kEmbIn = 34, kEmbHidden = 128, kEmbDim = 128;

struct EmbedPath {
    float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr;    // device fp32
    __nv_bfloat16 *F = nullptr;       // [N*(Ta+Tb) (+128 rows of slack), 128] embeddings of a then b
    float *norm = nullptr;            // [N*(Ta+Tb)]
    float *cm = nullptr;              // [N, ra, rb] cost matrices (when the caller does not take them)
    size_t f_elems = 0, cm_floats = 0;
};

namespace {

using namespace tc;

constexpr int kEncFrames = 128;       // frames per CTA pass of the encoder
constexpr int kEncThreads = 512;      // 32 frame groups x 16 output groups: 4 frames x 8 outputs per thread

// smem: W1 [34][128], b1[128], W2 [128][128], b2[128], x [128][34], h [128][128 + 4]: 166 KB, 16 warps per SM (the
// 8-warp form ran at 28 % of the FMA pipe: shared-memory latency with nothing to hide it)
__global__ void __launch_bounds__(kEncThreads)
embed_encoder_kernel(const float *__restrict__ frames, int V, int Cc, size_t nframes, const float *__restrict__ W1,
                     const float *__restrict__ b1, const float *__restrict__ W2, const float *__restrict__ b2,
                     __nv_bfloat16 *__restrict__ F, float *__restrict__ norm) {
    extern __shared__ __align__(16) float sm[];
    float *sW1 = sm, *sb1 = sW1 + kEmbIn * kEmbHidden, *sW2 = sb1 + kEmbHidden, *sb2 = sW2 + kEmbHidden * kEmbDim;
    float *sx = sb2 + kEmbDim, *sh = sx + kEncFrames * kEmbIn;
    constexpr int ldh = kEmbHidden + 4;
    for (int e = threadIdx.x; e < kEmbIn * kEmbHidden; e += kEncThreads) sW1[e] = W1[e];
    for (int e = threadIdx.x; e < kEmbHidden * kEmbDim; e += kEncThreads) sW2[e] = W2[e];
    if (threadIdx.x < kEmbHidden) {
        sb1[threadIdx.x] = b1[threadIdx.x];
        sb2[threadIdx.x] = b2[threadIdx.x];
    }
    const int fg = threadIdx.x >> 4, og = threadIdx.x & 15;       // 4 frames x 8 outputs per thread
    for (size_t f0 = (size_t)blockIdx.x * kEncFrames; f0 < nframes; f0 += (size_t)gridDim.x * kEncFrames) {
        __syncthreads();
        const int nf = (int)(nframes - f0 < (size_t)kEncFrames ? nframes - f0 : (size_t)kEncFrames);
        for (int e = threadIdx.x; e < kEncFrames * kEmbIn; e += kEncThreads) {
            const int f = e / kEmbIn, k = e - f * kEmbIn;          // k = joint * 2 + (x | y)
            sx[e] = f < nf ? frames[((f0 + f) * V + (k >> 1)) * Cc + (k & 1)] : 0.f;
        }
        __syncthreads();
        float acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[a][o] = sb1[og * 8 + o];
#pragma unroll 2
        for (int k = 0; k < kEmbIn; ++k) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sW1 + k * kEmbHidden + og * 8);
            const float4 w1 = *reinterpret_cast<const float4 *>(sW1 + k * kEmbHidden + og * 8 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float x = sx[(fg * 4 + a) * kEmbIn + k];
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[a][o] = fmaf(x, w[o], acc[a][o]);
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int o = 0; o < 8; ++o) sh[(fg * 4 + a) * ldh + og * 8 + o] = fmaxf(acc[a][o], 0.f);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[a][o] = sb2[og * 8 + o];
#pragma unroll 4
        for (int k = 0; k < kEmbHidden; ++k) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sW2 + k * kEmbDim + og * 8);
            const float4 w1 = *reinterpret_cast<const float4 *>(sW2 + k * kEmbDim + og * 8 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float x = sh[(fg * 4 + a) * ldh + k];
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[a][o] = fmaf(x, w[o], acc[a][o]);
            }
        }
        // round to bf16 (the GEMM operand) and take |f|^2 of the ROUNDED values: 16 threads share a frame
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int f = fg * 4 + a;
            uint32_t pk[4];
            float nn = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                pk[o] = pack_bf16(acc[a][2 * o], acc[a][2 * o + 1]);
                const float lo = bf16lo_to_f32(pk[o]), hi = bf16hi_to_f32(pk[o]);
                nn = fmaf(lo, lo, nn);
                nn = fmaf(hi, hi, nn);
            }
#pragma unroll
            for (int d = 8; d >= 1; d >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, d);
            if (f < nf) {
                *reinterpret_cast<uint4 *>(F + (f0 + f) * kEmbDim + og * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (og == 0) norm[f0 + f] = nn;
            }
        }
    }
}

// One CTA = one 128 x 128 tile of one pair's cost matrix.  192 threads: warp 0 TMA + MMA issue, warp 1 TMEM
// allocation, warps 2-5 epilogue (TMEM lane quarter = warp % 4).
constexpr int kCostThreads = 192;
constexpr uint32_t kCostSmem = 4 * 16384 + 2048 + 1024;      // 4 operand boxes, staging pad, alignment slack

__global__ void __launch_bounds__(kCostThreads)
embed_cost_kernel(const __grid_constant__ CUtensorMap mapF, const float *__restrict__ norm, int N, int ra, int rb,
                  size_t row_a0, size_t row_b0, float *__restrict__ cm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t full, done;
    __shared__ uint32_t tmem_slot;
    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int mtiles = (ra + 127) / 128, ntiles = (rb + 127) / 128;
    int id = blockIdx.x;
    const int nt = id % ntiles;
    id /= ntiles;
    const int mt = id % mtiles, n = id / mtiles;
    const int i0 = mt * 128, j0 = nt * 128;
    // rows of the embedding matrix: sequence "a" of the sweep (ra frames per pair) starts at row_a0, "b" at row_b0
    const size_t arow = row_a0 + (size_t)n * ra + i0, brow = row_b0 + (size_t)n * rb + j0;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapF);
        mbar_init(&full, 1);
        mbar_init(&done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&full, 4 * 16384);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(smem + kb * 16384, &mapF, &full, kb * 64, (int)arow);
                tma_load_2d(smem + 32768 + kb * 16384, &mapF, &full, kb * 64, (int)brow);
            }
        }
        __syncwarp();
        mbar_wait(&full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128u);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = make_kmajor_desc(smem_u32(smem + kb * 16384), 128);
            const uint64_t db = make_kmajor_desc(smem_u32(smem + 32768 + kb * 16384), 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (elect_one()) umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb > 0) | (k > 0)));
        }
        if (elect_one()) umma_commit(&done);
        __syncwarp();
    } else if (warp >= 2) {
        const int ew = warp & 3, r = ew * 32 + lane;
        const int i = i0 + r;
        const float na = i < ra ? norm[arow + r] : 0.f;
        mbar_wait(&done, 0);
        tc_fence_after();
        // the operand boxes are dead once the MMAs have completed: stage the fp32 tile there, row stride 129 floats
        float *stage = reinterpret_cast<float *>(smem);
        constexpr int lds = 129;
        for (int c = 0; c < 8; ++c) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(c * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int j = j0 + c * 16 + e;
                const float nb = j < rb ? __ldg(norm + brow + c * 16 + e) : 0.f;
                const float d2 = fmaf(-2.f, __uint_as_float(v[e]), na + nb);
                stage[r * lds + c * 16 + e] = sqrtf(fmaxf(d2, 0.f));
            }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // coalesced rows: warp w writes rows w, w+4, ...; a lane writes columns lane, lane+32, ...
        const int w = warp - 2;
        float *dst = cm + ((size_t)n * ra) * rb;
        for (int rr = w; rr < 128; rr += 4) {
            const int ii = i0 + rr;
            if (ii >= ra) break;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q * 32 + lane;
                if (j < rb) dst[(size_t)ii * rb + j] = stage[rr * lds + q * 32 + lane];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

template <typename T>
int grow_dev(Ctx *ctx, T **p, size_t *cap, size_t count) {
    if (count <= *cap) return GS_OK;
    if (*p) {
        GS_CUDA(cudaFree(*p));      // synchronises the device: nothing in flight still uses the old buffer
        ctx->ws_bytes -= *cap * sizeof(T);
        *p = nullptr;
        *cap = 0;
    }
    GS_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    ctx->ws_bytes += count * sizeof(T);
    *cap = count;
    return GS_OK;
}

}  // namespace

int align_embed_set_encoder(Ctx *ctx, const float *blob, size_t nfloats) {
    const size_t want = (size_t)kEmbIn * kEmbHidden + kEmbHidden + (size_t)kEmbHidden * kEmbDim + kEmbDim;
    if (!blob || nfloats != want) {
        set_error("align encoder blob holds %zu floats, AlignEmbedConfig v0 needs %zu (W1 [34,128], b1, W2 [128,128], b2)",
                  nfloats, want);
        return GS_ERR_INVALID;
    }
    if (!tc::get_encode_fn()) {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return GS_ERR_CUDA;
    }
    if (!ctx->embed) ctx->embed = new EmbedPath();
    EmbedPath *ep = ctx->embed;
    if (!ep->W1) {
        GS_CUDA(cudaMalloc((void **)&ep->W1, want * sizeof(float)));
        ctx->ws_bytes += want * sizeof(float);
        ep->b1 = ep->W1 + (size_t)kEmbIn * kEmbHidden;
        ep->W2 = ep->b1 + kEmbHidden;
        ep->b2 = ep->W2 + (size_t)kEmbHidden * kEmbDim;
    }
    GS_CUDA(cudaMemcpy(ep->W1, blob, want * sizeof(float), cudaMemcpyHostToDevice));
    return GS_OK;
}

void align_embed_destroy(Ctx *ctx) {
    EmbedPath *ep = ctx->embed;
    if (!ep) return;
    if (ep->W1) cudaFree(ep->W1);
    if (ep->F) cudaFree(ep->F);
    if (ep->norm) cudaFree(ep->norm);
    if (ep->cm) cudaFree(ep->cm);
    delete ep;
    ctx->embed = nullptr;
}

int align_embed_launch(Ctx *ctx, const float *a, const float *b, int N, int Ta, int Tb, int V, int Cc, float *cost,
                       int32_t *path, int32_t *plen, float *cost_matrix_out, cudaStream_t st) {
    EmbedPath *ep = ctx->embed;
    if (!ep || !ep->W1) {
        set_error("align_embed: no encoder set (gs_set_align_encoder)");
        return GS_ERR_INVALID;
    }
    if (V * 2 != kEmbIn) {
        set_error("align_embed: AlignEmbedConfig v0 takes V = 17 joints (34 inputs), got V = %d", V);
        return GS_ERR_UNSUPPORTED;
    }
    // the sweep wants the shorter sequence on the column axis: with Ta < Tb the cost matrix is built for the
    // exchanged sequences (rows = b frames) and the DP / backtrack run in their SWAP form (align.cu)
    const bool swap = Ta < Tb;
    const int ra = swap ? Tb : Ta, rb = swap ? Ta : Tb;
    const size_t rows_a = (size_t)N * Ta, rows_b = (size_t)N * Tb, rows = rows_a + rows_b;
    if (rows + 128 >= ((size_t)1 << 31)) {
        set_error("align_embed: %zu frames exceed the 2^31 rows of one tensor map", rows);
        return GS_ERR_UNSUPPORTED;
    }
    int rc;
    size_t fcap = ep->f_elems;
    if ((rc = grow_dev(ctx, &ep->F, &fcap, (rows + 128) * kEmbDim)) != GS_OK) return rc;    // +128 rows: tile overrun of the last pair
    if (fcap != ep->f_elems) {
        ep->f_elems = fcap;
        size_t ncap = 0;
        if (ep->norm) { cudaFree(ep->norm); ep->norm = nullptr; }
        if ((rc = grow_dev(ctx, &ep->norm, &ncap, rows + 128)) != GS_OK) return rc;
        GS_CUDA(cudaMemsetAsync(ep->F, 0, fcap * sizeof(__nv_bfloat16), st));
        GS_CUDA(cudaMemsetAsync(ep->norm, 0, (rows + 128) * sizeof(float), st));
    }
    float *cm = cost_matrix_out;
    if (!cm) {
        if ((rc = grow_dev(ctx, &ep->cm, &ep->cm_floats, (size_t)N * ra * rb)) != GS_OK) return rc;
        cm = ep->cm;
    }
    const size_t enc_smem = ((size_t)kEmbIn * kEmbHidden + kEmbHidden + (size_t)kEmbHidden * kEmbDim + kEmbDim +
                             (size_t)kEncFrames * kEmbIn + (size_t)kEncFrames * (kEmbHidden + 4)) * sizeof(float);
    if ((rc = ensure_dyn_smem(ctx, (const void *)embed_encoder_kernel, enc_smem)) != GS_OK) return rc;
    for (int which = 0; which < 2; ++which) {
        const size_t nfr = which == 0 ? rows_a : rows_b;
        int grid = (int)((nfr + kEncFrames - 1) / kEncFrames < (size_t)ctx->sm_count ? (nfr + kEncFrames - 1) / kEncFrames
                                                                                     : (size_t)ctx->sm_count);
        if (grid < 1) grid = 1;
        {
            LaunchScope ls(ctx, K_EMBED, st, 2.0 * nfr * (kEmbIn * kEmbHidden + kEmbHidden * kEmbDim), (double)nfr * (V * Cc * 4 + kEmbDim * 2));
            embed_encoder_kernel<<<grid, kEncThreads, enc_smem, st>>>(which == 0 ? a : b, V, Cc, nfr, ep->W1, ep->b1, ep->W2, ep->b2,
                                                              ep->F + (which == 0 ? 0 : rows_a * kEmbDim),
                                                              ep->norm + (which == 0 ? 0 : rows_a));
        }
        GS_KERNEL_CHECK();
    }
    CUtensorMap mapF;
    if ((rc = tc::make_weight_map(&mapF, ep->F, kEmbDim, (int)(rows + 128), 64, 128)) != GS_OK) return rc;
    if ((rc = ensure_dyn_smem(ctx, (const void *)embed_cost_kernel, kCostSmem)) != GS_OK) return rc;
    const int mtiles = (ra + 127) / 128, ntiles = (rb + 127) / 128;
    const long long ctas = (long long)N * mtiles * ntiles;
    if (ctas >= (1ll << 31)) {
        set_error("align_embed: too many cost tiles (%lld)", ctas);
        return GS_ERR_UNSUPPORTED;
    }
    {
        LaunchScope ls(ctx, K_EMBED_COST, st, 2.0 * N * (double)ra * rb * kEmbDim, (double)N * ((double)ra * rb * 4 + (ra + rb) * kEmbDim * 2.0));
        // rows of the sweep's "a" (row) sequence and "b" (column) sequence inside F
        embed_cost_kernel<<<(int)ctas, kCostThreads, kCostSmem, st>>>(mapF, ep->norm, N, ra, rb, swap ? rows_a : 0,
                                                                      swap ? 0 : rows_a, cm);
    }
    GS_KERNEL_CHECK();
    return dtw_costmat_launch(ctx, cm, N, ra, rb, swap, cost, path, plen, st);
}

}  // namespace gs

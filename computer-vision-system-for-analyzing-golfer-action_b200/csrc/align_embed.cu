// Learned alignment embedding (/root/reference/README.md:44-47: the alignment model is trained).
// Contract: oracle/embed.py (AlignEmbedConfig v0: per-frame MLP 34 -> 128 -> 128 over the (x, y) of the 17 joints,
// cost c[i,j] = sqrt(max(|fa_i|^2 + |fb_j|^2 - 2 fa_i . fb_j, 0)), then the DP / tie-break / backtrack of align.cu).
//
//   embed_encoder_tc_kernel  frames -> F [rows, 128] bf16 (the GEMM operand) + |F|^2 fp32 of the ROUNDED values
//                          (mma.sync m16n8k8 3xTF32 = fp32-accurate; both weight matrices as B fragments in shared memory)
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
    float *Wfrag = nullptr;           // W1 (K padded to 40) and W2 as 3xTF32 mma.sync B fragments (embed_encoder_tc_kernel)
    __nv_bfloat16 *F = nullptr;       // [N*(Ta+Tb) (+128 rows of slack), 128] embeddings of a then b
    float *norm = nullptr;            // [N*(Ta+Tb)]
    float *cm = nullptr;              // [N, ra, rb] cost matrices (when the caller does not take them)
    size_t f_elems = 0, cm_floats = 0;
};

namespace {

using namespace tc;

constexpr int kEncFrames = 128;       // frames per CTA pass of the encoder

// ---- the encoder on warp-level tensor cores: 3xTF32 (error-compensated: fp32 operands split into a TF32 head and a
// TF32 tail, a.b = a_hi.b_hi + a_hi.b_lo + a_lo.b_hi, the dropped term is 2^-22 relative), so the embeddings agree with
// an fp32 evaluation to summation order and the parity policy of oracle/embed.py holds unchanged.  The first form of
// this kernel was fp32 FMAs on CUDA cores (4 frames x 8 outputs per thread, weights in shared memory): 3.97 ms for the
// 2.46 M frames of 4096 pairs, 37 % of the FMA pipe, bound by its shared-memory operand loads (6 loads per 32 FMAs);
// this one: 2.14 ms.
// One CTA pass = 128 frames, 8 warps x 16 frames (one m16 tile per warp); both weight matrices live in shared memory as
// ready-made B fragments {hi(k=t), hi(k=t+4), lo(k=t), lo(k=t+4)} per (k-step, n-tile, lane): one conflict-free LDS.128
// per three MMAs.  The hidden layer never leaves registers: accumulator tile j of layer 1 IS k-tile j of layer 2, its
// A fragment is gathered inside each quad with 8 shuffles.
constexpr int kEncTcThreads = 256;
constexpr int kEncTcKs1 = 5;                     // K = 34 padded to 40
constexpr int kEncLdx = 44;                      // sx row stride: conflict-free A-fragment loads
constexpr size_t kEncFragFloats = (size_t)(kEncTcKs1 + kEmbHidden / 8) * (kEmbHidden / 8) * 32 * 4;   // 43008

__device__ __forceinline__ uint32_t tf32_of(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kEncTcThreads, 1)
embed_encoder_tc_kernel(const float *__restrict__ frames, int V, int Cc, size_t nframes, const float *__restrict__ Wfrag,
                        const float *__restrict__ b1, const float *__restrict__ b2, __nv_bfloat16 *__restrict__ F,
                        float *__restrict__ norm) {
    extern __shared__ __align__(16) float sm[];
    float4 *fr1 = reinterpret_cast<float4 *>(sm);                       // [5][16][32]
    float4 *fr2 = fr1 + kEncTcKs1 * 16 * 32;                            // [16][16][32]
    float *sx = reinterpret_cast<float *>(fr2 + 16 * 16 * 32);          // [128][kEncLdx]
    float *sb1 = sx + kEncFrames * kEncLdx, *sb2 = sb1 + kEmbHidden;
    for (int e = threadIdx.x; e < (int)(kEncFragFloats / 4); e += kEncTcThreads)
        fr1[e] = __ldg(reinterpret_cast<const float4 *>(Wfrag) + e);
    if (threadIdx.x < kEmbHidden) {
        sb1[threadIdx.x] = b1[threadIdx.x];
        sb2[threadIdx.x] = b2[threadIdx.x];
    }
    for (int e = threadIdx.x; e < kEncFrames * kEncLdx; e += kEncTcThreads) sx[e] = 0.f;    // K padding stays zero
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const int src_lo = (lane & ~3) | (t >> 1), src_hi = src_lo + 2;
    const bool odd = t & 1;
    for (size_t f0 = (size_t)blockIdx.x * kEncFrames; f0 < nframes; f0 += (size_t)gridDim.x * kEncFrames) {
        __syncthreads();
        const int nf = (int)(nframes - f0 < (size_t)kEncFrames ? nframes - f0 : (size_t)kEncFrames);
        for (int e = threadIdx.x; e < kEncFrames * kEmbIn; e += kEncTcThreads) {
            const int f = e / kEmbIn, k = e - f * kEmbIn;          // k = joint * 2 + (x | y)
            sx[f * kEncLdx + k] = f < nf ? frames[((f0 + f) * V + (k >> 1)) * Cc + (k & 1)] : 0.f;
        }
        __syncthreads();
        // ---- layer 1: h = relu(x . W1 + b1), 16 n-tiles of 8 hidden units
        float h[16][4];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const float2 bb = *reinterpret_cast<const float2 *>(sb1 + nt * 8 + 2 * t);
            h[nt][0] = h[nt][2] = bb.x;
            h[nt][1] = h[nt][3] = bb.y;
        }
#pragma unroll
        for (int ks = 0; ks < kEncTcKs1; ++ks) {
            const float xa[4] = {sx[r0 * kEncLdx + ks * 8 + t], sx[r1 * kEncLdx + ks * 8 + t],
                                 sx[r0 * kEncLdx + ks * 8 + t + 4], sx[r1 * kEncLdx + ks * 8 + t + 4]};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[i] = tf32_of(xa[i]);
                al[i] = tf32_of(xa[i] - __uint_as_float(ah[i]));
            }
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                const float4 b = fr1[(ks * 16 + nt) * 32 + lane];
                mma_tf32_16x8x8(h[nt], al, __float_as_uint(b.x), __float_as_uint(b.y));
                mma_tf32_16x8x8(h[nt], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                mma_tf32_16x8x8(h[nt], ah, __float_as_uint(b.x), __float_as_uint(b.y));
            }
        }
        // ---- layer 2: f = h . W2 + b2; accumulator tile ks of layer 1 is k-tile ks here
        float o[16][4];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const float2 bb = *reinterpret_cast<const float2 *>(sb2 + nt * 8 + 2 * t);
            o[nt][0] = o[nt][2] = bb.x;
            o[nt][1] = o[nt][3] = bb.y;
        }
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            // A fragment (g, t) (g+8, t) (g, t+4) (g+8, t+4) of relu(h tile ks): column c of a row sits in lane (g, c / 2),
            // element c & 1 (rows g) or 2 + (c & 1) (rows g + 8)
            float c[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) c[e] = fmaxf(h[ks][e], 0.f);
            const float l0 = __shfl_sync(0xffffffffu, c[0], src_lo), l1 = __shfl_sync(0xffffffffu, c[1], src_lo);
            const float l2 = __shfl_sync(0xffffffffu, c[2], src_lo), l3 = __shfl_sync(0xffffffffu, c[3], src_lo);
            const float u0 = __shfl_sync(0xffffffffu, c[0], src_hi), u1 = __shfl_sync(0xffffffffu, c[1], src_hi);
            const float u2 = __shfl_sync(0xffffffffu, c[2], src_hi), u3 = __shfl_sync(0xffffffffu, c[3], src_hi);
            const float xa[4] = {odd ? l1 : l0, odd ? l3 : l2, odd ? u1 : u0, odd ? u3 : u2};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[i] = tf32_of(xa[i]);
                al[i] = tf32_of(xa[i] - __uint_as_float(ah[i]));
            }
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                const float4 b = fr2[(ks * 16 + nt) * 32 + lane];
                mma_tf32_16x8x8(o[nt], al, __float_as_uint(b.x), __float_as_uint(b.y));
                mma_tf32_16x8x8(o[nt], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                mma_tf32_16x8x8(o[nt], ah, __float_as_uint(b.x), __float_as_uint(b.y));
            }
        }
        // ---- round to bf16 (the GEMM operand), |f|^2 of the ROUNDED values, store
        float n0 = 0.f, n1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const uint32_t p0 = pack_bf16(o[nt][0], o[nt][1]), p1 = pack_bf16(o[nt][2], o[nt][3]);
            n0 = fmaf(bf16lo_to_f32(p0), bf16lo_to_f32(p0), n0);
            n0 = fmaf(bf16hi_to_f32(p0), bf16hi_to_f32(p0), n0);
            n1 = fmaf(bf16lo_to_f32(p1), bf16lo_to_f32(p1), n1);
            n1 = fmaf(bf16hi_to_f32(p1), bf16hi_to_f32(p1), n1);
            if (r0 < nf) *reinterpret_cast<uint32_t *>(F + (f0 + r0) * kEmbDim + nt * 8 + 2 * t) = p0;
            if (r1 < nf) *reinterpret_cast<uint32_t *>(F + (f0 + r1) * kEmbDim + nt * 8 + 2 * t) = p1;
        }
        n0 += __shfl_xor_sync(0xffffffffu, n0, 1);
        n0 += __shfl_xor_sync(0xffffffffu, n0, 2);
        n1 += __shfl_xor_sync(0xffffffffu, n1, 1);
        n1 += __shfl_xor_sync(0xffffffffu, n1, 2);
        if (t == 0) {
            if (r0 < nf) norm[f0 + r0] = n0;
            if (r1 < nf) norm[f0 + r1] = n1;
        }
    }
}

// Host: W [K][128] (row-major, K rows used, padded with zeros to ksteps * 8) -> B fragments of mma.sync m16n8k8 (col-major B:
// b0 = W[ks*8 + t][nt*8 + g], b1 = W[ks*8 + t + 4][nt*8 + g]) split into TF32 head and tail: {hi0, hi1, lo0, lo1} per lane.
inline float tf32_round_host(float x) {          // cvt.rna.tf32.f32: nearest, ties away from zero, 10 mantissa bits kept
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & ~0x1fffu;
    float r;
    memcpy(&r, &u, 4);
    return r;
}
inline void pack_tf32x3_fragments(const float *W, int K, int ksteps, float *out) {
    for (int ks = 0; ks < ksteps; ++ks)
        for (int nt = 0; nt < kEmbHidden / 8; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3, n = nt * 8 + g;
                float *dst = out + ((size_t)(ks * 16 + nt) * 32 + lane) * 4;
                for (int h = 0; h < 2; ++h) {
                    const int k = ks * 8 + t + 4 * h;
                    const float w = k < K ? W[(size_t)k * kEmbHidden + n] : 0.f;
                    const float hi = tf32_round_host(w);
                    dst[h] = hi;
                    dst[2 + h] = tf32_round_host(w - hi);
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
        // the tile's 128 column norms once, in shared memory (every thread needs all of them: a global load per element
        // per thread was most of this kernel's epilogue)
        __shared__ float s_nb[128];
        {
            const int e = threadIdx.x - 64;
            s_nb[e] = j0 + e < rb ? norm[brow + e] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
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
                const float d2 = fmaf(-2.f, __uint_as_float(v[e]), na + s_nb[c * 16 + e]);
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
    std::vector<float> frag(kEncFragFloats);
    pack_tf32x3_fragments(blob, kEmbIn, kEncTcKs1, frag.data());
    pack_tf32x3_fragments(blob + (size_t)kEmbIn * kEmbHidden + kEmbHidden, kEmbHidden, kEmbHidden / 8,
                          frag.data() + (size_t)kEncTcKs1 * 16 * 32 * 4);
    if (!ep->Wfrag) {
        GS_CUDA(cudaMalloc((void **)&ep->Wfrag, kEncFragFloats * sizeof(float)));
        ctx->ws_bytes += kEncFragFloats * sizeof(float);
    }
    GS_CUDA(cudaMemcpy(ep->Wfrag, frag.data(), kEncFragFloats * sizeof(float), cudaMemcpyHostToDevice));
    return GS_OK;
}

void align_embed_destroy(Ctx *ctx) {
    EmbedPath *ep = ctx->embed;
    if (!ep) return;
    if (ep->W1) cudaFree(ep->W1);
    if (ep->Wfrag) cudaFree(ep->Wfrag);
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
    const size_t enc_smem = (kEncFragFloats + (size_t)kEncFrames * kEncLdx + 2 * kEmbHidden) * sizeof(float);
    if ((rc = ensure_dyn_smem(ctx, (const void *)embed_encoder_tc_kernel, enc_smem)) != GS_OK) return rc;
    for (int which = 0; which < 2; ++which) {
        const size_t nfr = which == 0 ? rows_a : rows_b;
        int grid = (int)((nfr + kEncFrames - 1) / kEncFrames < (size_t)ctx->sm_count ? (nfr + kEncFrames - 1) / kEncFrames
                                                                                     : (size_t)ctx->sm_count);
        if (grid < 1) grid = 1;
        {
            LaunchScope ls(ctx, K_EMBED, st, 2.0 * nfr * (kEmbIn * kEmbHidden + kEmbHidden * kEmbDim), (double)nfr * (V * Cc * 4 + kEmbDim * 2));
            embed_encoder_tc_kernel<<<grid, kEncTcThreads, enc_smem, st>>>(which == 0 ? a : b, V, Cc, nfr, ep->Wfrag, ep->b1, ep->b2,
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

// bf16 throughput path of the segmentation network: bf16 activations in HBM, every dense
// contraction on tcgen05 tensor cores with fp32 accumulation in TMEM.
//
// Block 0 (Cin = 3):  front_mma_kernel   input BN + adjacency on CUDA cores, the K=9 mix and the K=3
//                                        residual projection as one TF32 mma.sync GEMM -> Y (joint-major), R0
// Block i >= 1:       gcn_fused_kernel   gate-on-load, adjacency contraction + channel mix -> Y (joint-major)
//                                        (+ the gated input Xg for the residual)
// every block:        tcn_fused_kernel   branch 1x1 + dilated taps + residual + ReLU -> U, pooling sums
//                     se_kernel, stj_tc_kernel (segment_common.cuh) -> gates applied by the NEXT reader of U
// XA (3x the input) and H (the branch 1x1 output) never reach HBM.  Rounding points: weights, Xg, XA, Y,
// H, U (and block 0's residual projection) are rounded to bf16; everything else is fp32.
//
// Stages replaced: /root/reference/README.md:27-34.
#include "segment_common.cuh"
#include "umma.cuh"
#include "gcn_fused.cuh"
#include "tcn_fused.cuh"
#include <stdlib.h>

namespace gs {

struct BlockMaps {
    tf::Maps tf;                                                 // fused temporal kernel
    CUtensorMap f_x_load, f_xg_store, f_y_store, f_gt, f_gv;    // fused GCN kernel (7-frame tiles)
    CUtensorMap wg;
};

struct Bf16Path {
    std::vector<__nv_bfloat16 *> WgT, W1T, W2p, WrT;   // device bf16 weights, K contiguous
    std::vector<__nv_bfloat16 *> Bt1, Bt2;              // bias tiles [C][16] of the temporal kernel: b1; b2 (+ br)
    std::vector<float *> stjP[3];                       // ST-joint {W, Wt, Wv} in fp16 fragment order (stj_tc_kernel)
    std::vector<std::vector<float>> bgHost;             // host copies of the GCN biases: they travel as kernel parameters (gcn_fused.cuh)
    std::vector<std::vector<float>> b1Host, b2Host;     // temporal kernel at C = 256: b1 and b2 (+ br) as kernel parameters (tcn_fused.cuh)
    float *frontB = nullptr;                            // block 0: [16][2C] TF32 matrix of front_mma_kernel
    __nv_bfloat16 *ident64 = nullptr;                   // 64x64 identity: "projection" weights of the identity residual (tcn_fused.cuh)
    std::vector<BlockMaps> maps;
    int maps_T = -1;
    bool debug_xa = false;      // GOLFER_DEBUG_XA=1: the fused GCN kernel also dumps its XA chunks into bufXA (tests/test_gpu_kernels.py)
    unsigned long long *trace = nullptr;   // GOLFER_TRACE_GCN=1: [blocks][5 roles][6 tiles][64 events] clock64 (tools/trace_gcn.py)
    unsigned long long *trace_tcn = nullptr;   // GOLFER_TRACE_TCN=1: [blocks][4 roles][8 steps][16 events] (tools/trace_tcn.py)
};

namespace {

constexpr int kFrontFrames = 15;     // 255 rows: one A row per thread of the 256 (16 frames = 272 rows left a second pass for 16 of them)

// ---- block 0 with the two K = 9 / K = 3 mixes on warp-level tensor cores ---------------------------
// A CUDA-core form is bound by FMA issue (12 x 2C FMAs per row; 0.15 ms against a 0.054 ms
// write floor).  Here a row's inputs form one 16-wide A row
//     [ 9 aggregated inputs | 3 normalised raw inputs | 1 | 0 0 0 ]
// and BOTH outputs come from one [16 x 2C] matrix (host: pack_front_matrix)
//     columns [0,C)  : rows 0-8 = Wg, row 12 = bg          -> Y  = relu(.)
//     columns [C,2C) : rows 9-11 = Wr, row 12 = br         -> R0 (the block's residual projection)
// as mma.sync m16n8k8 TF32 with fp32 accumulation (inputs and weights rounded to TF32: 2^-11, below the
// bf16 rounding of the outputs).  One CTA = kFrontFrames frames = 16 m-tiles of 16 rows (255 used); a warp owns a
// 32-column group (its 16 B-fragment registers never change) and every (8 / groups)-th m-tile.
constexpr int kFrontLd = 20;     // A-row stride in floats: conflict-free fragment loads

// 4x4 transpose of 32-bit values across the 4 lanes of a quad (lane q holds row q on entry, column q on exit)
__device__ __forceinline__ void quad_transpose(uint32_t (&p)[4], int q) {
    const bool odd = q & 1, hi = q & 2;
    uint32_t s0 = odd ? p[0] : p[1], s1 = odd ? p[2] : p[3];
    s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
    s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    if (odd) { p[0] = s0; p[2] = s1; } else { p[1] = s0; p[3] = s1; }
    s0 = hi ? p[0] : p[2];
    s1 = hi ? p[1] : p[3];
    s0 = __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (hi) { p[0] = s0; p[1] = s1; } else { p[2] = s0; p[3] = s1; }
}

// Y is written JOINT-MAJOR [B,V,T,C] (tcn_fused.cuh reads frame windows per joint), R0 frame-major [B,T,V,C].
template <int CIN>
__global__ void __launch_bounds__(256)
front_mma_kernel(const float *__restrict__ skel, const float *__restrict__ in_scale, const float *__restrict__ in_shift,
                 const float *__restrict__ A, const float *__restrict__ Bm, int C, int T, size_t nframes,
                 __nv_bfloat16 *__restrict__ Y, __nv_bfloat16 *__restrict__ R0) {
    static_assert(CIN == 3, "A-row layout assumes 3 input channels");
    extern __shared__ __align__(16) float sm[];
    float *sA = sm;                                    // [3][17][17] adjacency
    float *sx = sA + 3 * V17 * V17 + 1;                // [rows][CIN] normalised input
    uint32_t *sam = reinterpret_cast<uint32_t *>(sx + kFrontFrames * V17 * CIN);   // [rows][kFrontLd] tf32 A rows
    constexpr int kRows = (kFrontFrames * V17 + 15) / 16 * 16;     // 256 = 16 m-tiles, row 255 is padding
    for (int k = threadIdx.x; k < 3 * V17 * V17; k += blockDim.x) sA[k] = A[k];
    const size_t f0 = (size_t)blockIdx.x * kFrontFrames;
    const int nf = (int)(nframes - f0 < (size_t)kFrontFrames ? nframes - f0 : (size_t)kFrontFrames);
    for (int e = threadIdx.x; e < nf * V17 * CIN; e += blockDim.x) {
        const int vc = e % (V17 * CIN);
        sx[e] = skel[f0 * V17 * CIN + e] * in_scale[vc] + in_shift[vc];
    }
    // this warp's B fragments: column group of 32 = 4 n-tiles, 2 k-steps
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gr = lane >> 2, tg = lane & 3;
    const int ngroups = 2 * C / 32;                    // 4 (C = 64) or 8 (C = 128)
    const int grp = warp % ngroups, mt0 = warp / ngroups, mstep = 8 / ngroups;
    uint32_t bf[2][4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int col = grp * 32 + t * 8 + gr;
            bf[ks][t][0] = __float_as_uint(__ldg(Bm + (size_t)(ks * 8 + tg) * 2 * C + col));
            bf[ks][t][1] = __float_as_uint(__ldg(Bm + (size_t)(ks * 8 + tg + 4) * 2 * C + col));
        }
    __syncthreads();
    // A rows: adjacency contraction of the row's frame, then the raw inputs, the bias one, zero padding
    for (int u = threadIdx.x; u < kRows; u += blockDim.x) {
        uint32_t *arow = sam + u * kFrontLd;
        if (u < nf * V17) {
            const int f = u / V17, w = u - f * V17;
            float acc[3 * CIN];
#pragma unroll
            for (int k = 0; k < 3 * CIN; ++k) acc[k] = 0.f;
            const float *xf = sx + f * V17 * CIN;
#pragma unroll
            for (int v = 0; v < V17; ++v) {
                float x[CIN];
#pragma unroll
                for (int c = 0; c < CIN; ++c) x[c] = xf[v * CIN + c];
#pragma unroll
                for (int pp = 0; pp < 3; ++pp) {
                    const float a = sA[(pp * V17 + w) * V17 + v];
#pragma unroll
                    for (int c = 0; c < CIN; ++c) acc[pp * CIN + c] += a * x[c];
                }
            }
#pragma unroll
            for (int k = 0; k < 3 * CIN; ++k) arow[k] = to_tf32(acc[k]);
#pragma unroll
            for (int c = 0; c < CIN; ++c) arow[3 * CIN + c] = to_tf32(sx[u * CIN + c]);
            arow[12] = __float_as_uint(1.0f);
            arow[13] = arow[14] = arow[15] = 0u;
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) arow[k] = 0u;
        }
    }
    __syncthreads();
    const bool is_y = grp * 32 < C;
    __nv_bfloat16 *dst = is_y ? Y : R0;
    const int cbase = is_y ? grp * 32 : grp * 32 - C;
    for (int mt = mt0; mt < kRows / 16; mt += mstep) {
        const int r0 = mt * 16 + gr, r1 = r0 + 8;
        float acc[4][4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a0 = sam[r0 * kFrontLd + ks * 8 + tg], a1 = sam[r1 * kFrontLd + ks * 8 + tg];
            const uint32_t a2 = sam[r0 * kFrontLd + ks * 8 + tg + 4], a3 = sam[r1 * kFrontLd + ks * 8 + tg + 4];
#pragma unroll
            for (int t = 0; t < 4; ++t) mma_tf32(acc[t], a0, a1, a2, a3, bf[ks][t][0], bf[ks][t][1]);
        }
        const bool v0 = r0 < nf * V17, v1 = r1 < nf * V17;
        // pack, then transpose 4x4 inside the quad so that lane tg ends with the 8 columns of n-tile tg:
        // one 16-byte store per lane and row (a quad writes 64 contiguous bytes) instead of four 4-byte ones
        uint32_t p0[4], p1[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float y0 = acc[t][0], y1 = acc[t][1], y2 = acc[t][2], y3 = acc[t][3];
            if (is_y) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
            p0[t] = tc::pack_bf16(y0, y1);
            p1[t] = tc::pack_bf16(y2, y3);
        }
        quad_transpose(p0, tg);
        quad_transpose(p1, tg);
        const int col = cbase + tg * 8;
        // destination row: frame-major (f*17 + v) for R0, joint-major ((b*17 + v)*T + t) for Y
        auto dst_row = [&](int r) -> size_t {
            const size_t f = f0 + (size_t)(r / V17);
            if (!is_y) return f * V17 + (size_t)(r % V17);
            const size_t b = f / (size_t)T, t = f - b * (size_t)T;
            return (b * V17 + (size_t)(r % V17)) * (size_t)T + t;
        };
        if (v0) *reinterpret_cast<uint4 *>(dst + dst_row(r0) * C + col) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
        if (v1) *reinterpret_cast<uint4 *>(dst + dst_row(r1) * C + col) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    }
}

// Host: the [16][2C] matrix of front_mma_kernel, rounded to TF32.
inline void pack_front_matrix(const float *Wg, const float *bg, const float *Wr, const float *br, int C,
                              std::vector<float> &out) {
    out.assign((size_t)16 * 2 * C, 0.f);
    for (int k = 0; k < 9; ++k)
        for (int n = 0; n < C; ++n) out[(size_t)k * 2 * C + n] = Wg[(size_t)k * C + n];
    for (int k = 0; k < 3; ++k)
        for (int n = 0; n < C; ++n) out[(size_t)(9 + k) * 2 * C + C + n] = Wr[(size_t)k * C + n];
    for (int n = 0; n < C; ++n) {
        out[(size_t)12 * 2 * C + n] = bg[n];
        out[(size_t)12 * 2 * C + C + n] = br[n];
    }
    for (float &x : out) {      // round to nearest TF32 (10 explicit mantissa bits)
        uint32_t u;
        memcpy(&u, &x, 4);
        u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
        memcpy(&x, &u, 4);
    }
}

int upload_bf16(Ctx *ctx, const std::vector<__nv_bfloat16> &h, __nv_bfloat16 **d) {
    GS_CUDA(cudaMalloc((void **)d, h.size() * 2));
    ctx->ws_bytes += h.size() * 2;
    GS_CUDA(cudaMemcpy(*d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return GS_OK;
}
int build_maps(Ctx *ctx, int T) {
    Bf16Path *bp = ctx->bf16;
    if (bp->maps_T == T) return GS_OK;
    const int rows = T * V17, batch = ctx->max_B;
    const int R = ctx->cfg.num_branches;
    bp->maps.assign(ctx->blocks.size(), BlockMaps{});
    int dmax = 1;
    for (int r = 0; r < R; ++r) dmax = dmax > ctx->cfg.dilations[r] ? dmax : ctx->cfg.dilations[r];
    const int nout = tf::kWin - 2 * dmax;
    int rc;
    for (size_t i = 0; i < ctx->blocks.size(); ++i) {
        const BlockParams &b = ctx->blocks[i];
        BlockMaps &m = bp->maps[i];
        const int C = b.c, cin = b.cin, cr = b.cr;
        const int crm = cr < 16 ? 16 : cr;      // MMA width of a branch tap (tcn_fused.cuh)
        const bool proj = (i > 0 && b.has_res);
        // temporal kernel: joint-major Y windows, (C,V,T,B) views of the [B,T,V,C] tensors
        if ((rc = tf::make_bvtc_map(&m.tf.y_win, ctx->bufY, C, T, batch, false, 64, tf::kWin))) return rc;
        if ((rc = tf::make_btvc_joint_map(&m.tf.out, ctx->bufU[i & 1], C, T, batch, 64, nout))) return rc;
        if ((rc = tc::make_weight_map(&m.tf.w1, bp->W1T[i], C, C, 64, 64))) return rc;
        if ((rc = tc::make_weight_map(&m.tf.w2, bp->W2p[i], crm, R * 3 * crm, crm, crm))) return rc;
        if ((rc = tc::make_weight_map(&m.tf.bt1, bp->Bt1[i], 16, C, 16, 64))) return rc;
        if ((rc = tc::make_weight_map(&m.tf.bt2, bp->Bt2[i], 16, C, 16, 64))) return rc;
        if (proj) {
            if ((rc = tf::make_btvc_joint_map(&m.tf.xg, ctx->bufX, cin, T, batch, 64, tf::kWin))) return rc;
            if ((rc = tc::make_weight_map(&m.tf.wr, bp->WrT[i], cin, C, 64, 64))) return rc;
        } else {
            // identity residual through the tensor core: K box q of the residual source (block 0: its projected input
            // R0, later blocks: their gated input Xg), Wr = I_64
            if ((rc = tf::make_btvc_joint_map(&m.tf.xg, i == 0 ? ctx->bufR : ctx->bufX, C, T, batch, 64, tf::kWin))) return rc;
            if ((rc = tc::make_weight_map(&m.tf.wr, bp->ident64, 64, 64, 64, 64))) return rc;
        }
        if (i > 0) {
            if ((rc = tc::make_weight_map(&m.wg, bp->WgT[i], 3 * cin, C, 64, C))) return rc;
            if ((rc = tc::make_act_map(&m.f_x_load, ctx->bufU[(i - 1) & 1], cin, rows, batch, 64, tc::kTileM))) return rc;
            if ((rc = tc::make_act_map(&m.f_xg_store, ctx->bufX, cin, rows, batch, 64, gcn::kRowsPerTile))) return rc;
            if ((rc = tf::make_bvtc_map(&m.f_y_store, ctx->bufY, C, T, batch, true, 64, gcn::kFramesPerTile))) return rc;
            if ((rc = tc::make_f32_map(&m.f_gt, ctx->gT, cin, (long long)batch * T, 64, 8))) return rc;
            if ((rc = tc::make_f32_map(&m.f_gv, ctx->gV, cin, (long long)batch * V17, 64, V17))) return rc;
        }
    }
    bp->maps_T = T;
    return GS_OK;
}

}  // namespace

int bf16_path_create(Ctx *ctx) {
    const gs_config &c = ctx->cfg;
    const int R = c.num_branches;
    if (ctx->blocks.empty() || !ctx->blocks[0].has_res || c.in_channels != 3 ||
        (ctx->blocks[0].c != 64 && ctx->blocks[0].c != 128)) {
        set_error("bf16 path expects 3 input channels and a first block of width 64 or 128");
        return GS_ERR_UNSUPPORTED;
    }
    int dmax = 1;
    for (int r = 0; r < R; ++r) dmax = dmax > c.dilations[r] ? dmax : c.dilations[r];
    for (size_t i = 0; i < ctx->blocks.size(); ++i) {
        const BlockParams &b = ctx->blocks[i];
        const bool ok = (b.c % 64 == 0) && b.c <= 256 && (b.cr == 8 || b.cr == 16 || b.cr == 32 || b.cr == 64) &&
                        (i == 0 || b.cin % 64 == 0) && b.c == 4 * b.cj && dmax <= 16;
        if (!ok) {
            set_error("bf16 tensor-core path needs widths in {64,128,256}, C/R in {8,16,32,64}, ST-joint reduction 4 "
                      "and dilations <= 16 (block %zu: cin=%d c=%d c/R=%d); use precision fp32 for this config", i, b.cin,
                      b.c, b.cr);
            return GS_ERR_UNSUPPORTED;
        }
    }
    if (!tc::get_encode_fn()) {
        set_error("cuTensorMapEncodeTiled entry point not found");
        return GS_ERR_CUDA;
    }
    Bf16Path *bp = new Bf16Path();
    ctx->bf16 = bp;
    // instrumentation only (no alternative code paths hang off the environment)
    if (const char *e = getenv("GOLFER_DEBUG_XA")) bp->debug_xa = (e[0] == '1');
    if (const char *e = getenv("GOLFER_TRACE_GCN")) {
        if (e[0] == '1') {
            const size_t n = (size_t)GS_MAX_BLOCKS * 5 * gcn::kTraceTiles * gcn::kTraceEv;
            GS_CUDA(cudaMalloc((void **)&bp->trace, n * 8));
            GS_CUDA(cudaMemset(bp->trace, 0, n * 8));
        }
    }
    if (const char *e = getenv("GOLFER_TRACE_TCN")) {
        if (e[0] == '1') {
            const size_t n = (size_t)GS_MAX_BLOCKS * 4 * tf::kTrSteps * tf::kTrEv;
            GS_CUDA(cudaMalloc((void **)&bp->trace_tcn, n * 8));
            GS_CUDA(cudaMemset(bp->trace_tcn, 0, n * 8));
        }
    }
    {
        std::vector<__nv_bfloat16> id((size_t)64 * 64, __float2bfloat16_rn(0.f));
        for (int k = 0; k < 64; ++k) id[(size_t)k * 64 + k] = __float2bfloat16_rn(1.f);
        int rc0;
        if ((rc0 = upload_bf16(ctx, id, &bp->ident64))) return rc0;
    }
    const size_t nb = ctx->blocks.size();
    bp->WgT.assign(nb, nullptr);
    bp->W1T.assign(nb, nullptr);
    bp->W2p.assign(nb, nullptr);
    bp->WrT.assign(nb, nullptr);
    bp->Bt1.assign(nb, nullptr);
    bp->Bt2.assign(nb, nullptr);
    for (auto &v : bp->stjP) v.assign(nb, nullptr);
    const float *hb = ctx->h_blob.data();
    auto host = [&](const float *dev_ptr) { return hb + (dev_ptr - ctx->d_blob); };
    int rc;
    for (size_t i = 0; i < nb; ++i) {
        const BlockParams &b = ctx->blocks[i];
        const int C = b.c, cin = b.cin, cr = b.cr;
        std::vector<__nv_bfloat16> h;
        if (i > 0) {   // WgT [C][3cin]
            const float *W = host(b.Wg);
            h.assign((size_t)C * 3 * cin, __nv_bfloat16());
            for (int k = 0; k < 3 * cin; ++k)
                for (int n = 0; n < C; ++n) h[(size_t)n * 3 * cin + k] = __float2bfloat16_rn(W[(size_t)k * C + n]);
            if ((rc = upload_bf16(ctx, h, &bp->WgT[i]))) return rc;
        }
        {   // W1T [C][C]
            const float *W = host(b.W1);
            h.assign((size_t)C * C, __nv_bfloat16());
            for (int k = 0; k < C; ++k)
                for (int n = 0; n < C; ++n) h[(size_t)n * C + k] = __float2bfloat16_rn(W[(size_t)k * C + n]);
            if ((rc = upload_bf16(ctx, h, &bp->W1T[i]))) return rc;
        }
        {   // W2p [(r*3+j)*crm + co'][ci'], crm = max(cr, 16): an 8-channel branch sits in its diagonal 8x8
            // block of a zeroed 16x16 box (the other branch of the pair owns the other diagonal block)
            const float *W = host(b.W2);
            const int crm = cr < 16 ? 16 : cr;
            h.assign((size_t)R * 3 * crm * crm, __float2bfloat16_rn(0.f));
            for (int rj = 0; rj < R * 3; ++rj) {
                const int off = ((rj / 3) * cr) % crm;
                for (int ci = 0; ci < cr; ++ci)
                    for (int co = 0; co < cr; ++co)
                        h[((size_t)rj * crm + off + co) * crm + off + ci] =
                            __float2bfloat16_rn(W[((size_t)rj * cr + ci) * cr + co]);
            }
            if ((rc = upload_bf16(ctx, h, &bp->W2p[i]))) return rc;
        }
        bp->bgHost.resize(nb);
        bp->bgHost[i].assign(host(b.bg), host(b.bg) + C);
        std::vector<float> bias(host(b.b2), host(b.b2) + C);
        if (i > 0 && b.has_res) {   // WrT [C][cin]; its bias joins b2
            const float *W = host(b.Wr);
            h.assign((size_t)C * cin, __nv_bfloat16());
            for (int k = 0; k < cin; ++k)
                for (int n = 0; n < C; ++n) h[(size_t)n * cin + k] = __float2bfloat16_rn(W[(size_t)k * C + n]);
            if ((rc = upload_bf16(ctx, h, &bp->WrT[i]))) return rc;
            const float *brh = host(b.br);
            for (int n = 0; n < C; ++n) bias[n] += brh[n];
        }
        {   // ST-joint matrices in mma.sync fragment order
            const int Q = C / 64;
            const float *srcs[3] = {host(b.jW), host(b.jWt), host(b.jWv)};
            for (int w = 0; w < 3; ++w) {
                std::vector<float> packed;
                if (w == 0) pack_stj_fragments(srcs[w], C, b.cj, Q == 1 ? 2 : 4, packed);
                else pack_stj_fragments(srcs[w], b.cj, C, 4, packed);
                GS_CUDA(cudaMalloc((void **)&bp->stjP[w][i], packed.size() * sizeof(float)));
                ctx->ws_bytes += packed.size() * sizeof(float);
                GS_CUDA(cudaMemcpy(bp->stjP[w][i], packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
            }
        }
        // bias tiles: the bias enters the accumulator through an all-ones MMA (tcn_fused.cuh); two bf16 terms
        auto bias_tile = [&](const float *bsrc, __nv_bfloat16 **dst) -> int {
            std::vector<__nv_bfloat16> t((size_t)C * 16, __float2bfloat16_rn(0.f));
            for (int n = 0; n < C; ++n) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(bsrc[n]);
                t[(size_t)n * 16] = hi;
                t[(size_t)n * 16 + 1] = __float2bfloat16_rn(bsrc[n] - __bfloat162float(hi));
            }
            return upload_bf16(ctx, t, dst);
        };
        if ((rc = bias_tile(host(b.b1), &bp->Bt1[i]))) return rc;
        if ((rc = bias_tile(bias.data(), &bp->Bt2[i]))) return rc;
        bp->b1Host.resize(nb);
        bp->b2Host.resize(nb);
        bp->b1Host[i].assign(host(b.b1), host(b.b1) + C);
        bp->b2Host[i] = bias;
        if (i == 0) {   // block 0 on warp-level tensor cores (front_mma_kernel)
            std::vector<float> bm;
            pack_front_matrix(host(b.Wg), host(b.bg), host(b.Wr), host(b.br), C, bm);
            GS_CUDA(cudaMalloc((void **)&bp->frontB, bm.size() * sizeof(float)));
            ctx->ws_bytes += bm.size() * sizeof(float);
            GS_CUDA(cudaMemcpy(bp->frontB, bm.data(), bm.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    return GS_OK;
}

void bf16_path_destroy(Ctx *ctx) {
    Bf16Path *bp = ctx->bf16;
    if (!bp) return;
    for (auto *v : {&bp->WgT, &bp->W1T, &bp->W2p, &bp->WrT, &bp->Bt1, &bp->Bt2})
        for (__nv_bfloat16 *p : *v)
            if (p) cudaFree(p);
    for (auto &v : bp->stjP)
        for (float *p : v)
            if (p) cudaFree(p);
    if (bp->frontB) cudaFree(bp->frontB);
    if (bp->ident64) cudaFree(bp->ident64);
    if (bp->trace) cudaFree(bp->trace);
    if (bp->trace_tcn) cudaFree(bp->trace_tcn);
    delete bp;
    ctx->bf16 = nullptr;
}

// Debug / parity hooks (tests/test_gpu_kernels.py, tools/trace_gcn.py): raw copies of the workspace tensors
// of the LAST block that ran.  "X" gated block input [B,T,V,Cin]; "XA" aggregated input (GOLFER_DEBUG_XA=1);
// "Y" GCN output, joint-major [B,V,T,C]; "U0"/"U1" block outputs before their gates [B,T,V,C] (ping-pong);
// "R" block 0's projected input; "gT" [B,T,C] / "gV" [B,V,C] fp32 gates; "PT" / "PV" pooling sums; "seS" SE gate.
int bf16_debug_read(Ctx *ctx, const char *name, void *host, size_t nbytes) {
    Bf16Path *bp = ctx->bf16;
    const void *src = nullptr;
    if (bp && !strcmp(name, "gcn_trace")) src = bp->trace;
    else if (bp && !strcmp(name, "tcn_trace")) src = bp->trace_tcn;
    else if (!strcmp(name, "XA")) src = ctx->bufXA;
    else if (!strcmp(name, "X")) src = ctx->bufX;
    else if (!strcmp(name, "Y")) src = ctx->bufY;
    else if (!strcmp(name, "R")) src = ctx->bufR;
    else if (!strcmp(name, "U0")) src = ctx->bufU[0];
    else if (!strcmp(name, "U1")) src = ctx->bufU[1];
    else if (!strcmp(name, "gT")) src = ctx->gT;
    else if (!strcmp(name, "gV")) src = ctx->gV;
    else if (!strcmp(name, "PT")) src = ctx->PT;
    else if (!strcmp(name, "PV")) src = ctx->PV;
    else if (!strcmp(name, "seS")) src = ctx->seS;
    if (!src) {
        set_error("debug buffer '%s' not available", name);
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaDeviceSynchronize());
    GS_CUDA(cudaMemcpy(host, src, nbytes, cudaMemcpyDeviceToHost));
    return GS_OK;
}

int segment_bf16_forward(Ctx *ctx, const float *skel, float *logits, uint8_t *labels, int B, int T,
                         int upto_block, float *feat_out, cudaStream_t st) {
    typedef __nv_bfloat16 bf;
    Bf16Path *bp = ctx->bf16;
    int rc;
    if ((rc = build_maps(ctx, T))) return rc;
    const size_t nframes = (size_t)B * T;
    const double rows = (double)nframes * V17;
    const int nb = ctx->cfg.num_blocks, R = ctx->cfg.num_branches;
    const int last = upto_block >= 0 ? upto_block : nb - 1;
    bf *XA = (bf *)ctx->bufXA, *Y = (bf *)ctx->bufY, *R0 = (bf *)ctx->bufR;
    const bf *Uprev = nullptr;
    for (int i = 0; i <= last; ++i) {
        ctx->cur_block = i;
        const BlockParams &b = ctx->blocks[i];
        const BlockMaps &m = bp->maps[i];
        const int C = b.c, cin = b.cin, cr = b.cr;
        bf *U = (bf *)ctx->bufU[i & 1];
        const bool proj = (i > 0 && b.has_res);
        if (i == 0) {
            const size_t smem = ((size_t)3 * V17 * V17 + 1 + (size_t)kFrontFrames * V17 * cin + (size_t)((kFrontFrames * V17 + 15) / 16 * 16) * kFrontLd) * sizeof(float);
            // one launch, or one per input chunk of gs_segment_host (whole clips: both output layouts are per clip)
            const int nch = ctx->front_nchunks > 0 ? ctx->front_nchunks : 1;
            for (int k = 0; k < nch; ++k) {
                const size_t b0 = ctx->front_nchunks > 0 ? (size_t)ctx->front_b0[k] : 0;
                const size_t nbk = ctx->front_nchunks > 0 ? (size_t)ctx->front_nb[k] : (size_t)B;
                if (ctx->front_nchunks > 0) GS_CUDA(cudaStreamWaitEvent(st, ctx->ev_front[k], 0));
                const size_t nfk = nbk * T;
                const size_t act0 = b0 * (size_t)T * V17 * C;
                {
                    LaunchScope ls(ctx, K_B_FRONT, st, 2.0 * nfk * V17 * (V17 * 3 * cin + 4 * cin * C),
                                   (double)nfk * V17 * (cin * 4 + 4.0 * C));
                    front_mma_kernel<3><<<cdiv(nfk, kFrontFrames), 256, smem, st>>>(
                        skel + b0 * (size_t)T * V17 * cin, ctx->in_scale, ctx->in_shift, b.A, bp->frontB, C, T, nfk,
                        Y + act0, R0 + act0);
                }
                GS_KERNEL_CHECK();
            }
            // host entry points: the staged input has been consumed (gs_segment_host_submit overwrites it for the next batch)
            if (ctx->front_nchunks > 0) GS_CUDA(cudaEventRecord(ctx->ev_pipe_front, st));
        } else {
            gcn::LaunchGcn L{};
            L.mapX = m.f_x_load;
            L.mapXg = m.f_xg_store;
            L.mapW = m.wg;
            L.mapY = m.f_y_store;
            L.mapGT = m.f_gt;
            L.mapGV = m.f_gv;
            gcn::Params &q = L.prm;
            q.Cin = cin;
            q.C = C;
            q.T = T;
            q.B = B;
            q.rows_per_clip = T * V17;
            q.mtiles = cdiv(T, gcn::kFramesPerTile);
            q.ntiles = B * q.mtiles;
            q.gT = ctx->gT;
            q.gV = ctx->gV;
            q.store_xg = 1;
            q.A = b.A;
            memcpy(q.biasv, bp->bgHost[i].data(), (size_t)C * sizeof(float));
            q.dbg_xa = bp->debug_xa ? XA : nullptr;      // sized for whole tiles in alloc_workspace when the switch is on
            q.trace = bp->trace ? bp->trace + (size_t)i * 5 * gcn::kTraceTiles * gcn::kTraceEv : nullptr;
            L.flops = 2.0 * rows * (V17 * 3.0 * cin + 3.0 * cin * C);
            L.bytes = 2.0 * rows * ((q.store_xg ? 2.0 : 1.0) * cin + C);
            if ((rc = gcn::launch(ctx, K_B_GEMM_GCN, L, st))) return rc;
        }
        {   // branch 1x1 + dilated taps (+ residual projection) + residual + ReLU + pooling sums (tcn_fused.cuh)
            tf::LaunchTf L{};
            L.maps = m.tf;
            tf::Params &q = L.prm;
            memset(&q, 0, sizeof(q));
            q.B = B; q.T = T; q.C = C; q.cr = cr; q.cin = cin;
            if (cr >= 64 && C <= 256) {
                memcpy(q.b1v, bp->b1Host[i].data(), (size_t)C * sizeof(float));
                memcpy(q.b2v, bp->b2Host[i].data(), (size_t)C * sizeof(float));
            }
            q.nbr = 64 / cr;
            q.res_q = proj ? 0 : 1;
            q.nkx = proj ? cin / 64 : 1;
            q.nky = C / 64;
            q.dmax = 1;
            for (int r = 0; r < R; ++r) {
                q.dil[r] = ctx->cfg.dilations[r];
                q.dmax = q.dmax > q.dil[r] ? q.dmax : q.dil[r];
            }
            q.nout = tf::kWin - 2 * q.dmax;
            q.ttiles = cdiv(T, q.nout);
            q.nboxes = C / 64;
            q.nq_items = B * q.ttiles;
            q.PT = ctx->PT;
            q.PVpart = ctx->PVpart;
            q.trace = bp->trace_tcn ? bp->trace_tcn + (size_t)i * 4 * tf::kTrSteps * tf::kTrEv : nullptr;
            L.flops = 2.0 * rows * ((double)C * C + 3.0 * cr * C + (proj ? (double)cin * C : 0.0));
            L.bytes = 2.0 * rows * (2.0 * C + (proj ? cin : C));
            if ((rc = tf::launch(ctx, K_B_TCONV, L, st))) return rc;
        }
        const float *stjp[3] = {bp->stjP[0][i], bp->stjP[1][i], bp->stjP[2][i]};
        const int dmax_all = [&] { int d = 1; for (int r = 0; r < R; ++r) d = d > ctx->cfg.dilations[r] ? d : ctx->cfg.dilations[r]; return d; }();
        if ((rc = launch_attention<bf>(ctx, b, U, B, T, st, cdiv(T, tf::kWin - 2 * dmax_all), stjp))) return rc;
        Uprev = U;
    }
    ctx->cur_block = GS_MAX_BLOCKS;
    const int C = ctx->blocks[last].c;
    if (feat_out) return launch_features<bf>(ctx, Uprev, B, T, C, feat_out, st);
    return launch_head<bf>(ctx, Uprev, B, T, C, logits, labels, st);
}

}  // namespace gs

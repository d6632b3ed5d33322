// Multi-branch dilated temporal convolution for sm_100a, sliding-window form.
//
//   U[b,t,v, r*cr+co] = relu( sum_{j<3, ci<cr} H[b, t+(j-1)d_r, v, r*cr+ci] * W2[r,j,ci,co] + b2
//                             + residual[b,t,v, r*cr+co] )          zero outside [0,T)
//   residual = the block's gated input (identity) or its 1x1 projection Xg . Wr (width change)
// Stage replaced: /root/reference/README.md:29-30 (Temporal Module - Multi-branch Temporal Convolution).
//
// Why not the generic tap GEMM of tc_gemm.cuh: with [B,T,V,C] rows a temporal shift is 17*d rows, so
// every tap needs its own 128-row box — the H tile crosses L2->SM three times and the launch is
// bound by TMA latency x bytes in flight (measured: 650 cycles per chunk whatever the box size).
// Here the producing 1x1 GEMM writes H JOINT-MAJOR ([B,V,T,C], a free transposition in its TMA
// store map), so one M tile is 128 consecutive FRAMES of one joint and a shift of d frames is d rows:
//   * one TMA box per step: the (128 + 2*dmax)-frame window of 64 channels (1.06x the tile, not 3x);
//     frames outside [0,T) are zero-filled by TMA = the conv's zero padding;
//   * the 3 taps of every branch in the box are tcgen05 MMAs whose SW128 K-major A descriptors start
//     at ROW offsets dmax+(j-1)d inside that one box (verified in experiments/mma_probe.cu: the swizzle
//     is a function of the absolute smem address, any row offset works with base_offset = 0);
//   * one pipeline step = one (joint, 64-channel box): one full/empty handshake for all its MMAs.
// A CTA owns one 64-channel box q (its tap weights stay resident) and walks (clip, frame-tile) items
// two at a time; epilogue group g owns item g of the pair, loops the 17 joints, and so can keep the
// frame-pooling sums PT (sum over joints) in registers; the joint-pooling partial sums PVpart (sum
// over the tile's frames) are column sums of each staged tile.  Both feed SE / ST-joint attention
// (segment_common.cuh), replacing a separate pass over U.
//
// Warp roles (640 threads, 1 CTA/SM): w0 TMA producer (window + projection-input boxes), w1 MMA issuer,
// w2 TMEM alloc then residual producer of group 1, w3 residual producer of group 0, w4-11 / w12-19 the
// two epilogue groups (8 warps: TMEM lanes = rows by w%4, column halves by (w-4)/4%2).
#pragma once
#include "tc_gemm.cuh"

namespace gs {
namespace tw {

using namespace tc;

constexpr int kTwThreads = 640;
constexpr int kTwGroup = 256;
constexpr int kFramesTile = 128;
constexpr int kTwMaxStages = 8;

struct Params {
    int B, T, C, cr, cin;
    int nbr;          // branches inside one 64-channel box (64 / cr)
    int proj;         // 1: residual = Xg . Wr (extra MMAs), 0: identity residual box added in the epilogue
    int nkx;          // projection K boxes (cin / 64)
    int dil[GS_MAX_BRANCHES];
    int dmax, wrows;  // window rows = 128 + 2*dmax
    int ttiles, nboxes, nq_items;   // frame tiles per clip, 64-channel boxes, items per box (B * ttiles)
    int rev;              // 1: walk the items from the last clip down (reads the tail of H, still in L2, first)
    int stages, eslots;   // A ring depth; staging slots per epilogue group (2 or 3)
    uint32_t a_span, xg_off, stage_bytes, w_off, wr_off, w_bytes, wr_bytes, out_off, scr_off, bar_off, total;
    const float *bias;    // [C]  (b2 + folded projection bias)
    float *PT;            // [B,T,C]
    float *PVpart;        // [B,ttiles,17,C]
    unsigned long long *trace;
};

struct Maps {
    CUtensorMap h_win;    // H joint-major [B,V,T,C]: box (64 ch, wrows frames, 1 joint, 1 clip)
    CUtensorMap xg;       // Xg [B,T,V,cin] seen as (C, V, T, B): box (64 ch, 1 joint, 128 frames, 1)  (projection input)
    CUtensorMap res;      // identity residual [B,T,V,C], same box shape
    CUtensorMap out;      // U [B,T,V,C], same box shape
    CUtensorMap w2;       // tap weights [(r*3+j)*cr + co][ci], box (cr, cr)
    CUtensorMap wr;       // projection weights [C][cin], box (64 k, 64 n)
};

constexpr int kTwTraceTiles = 8, kTwTraceEv = 32;
#define TW_TRACE(role, n, ev)                                                                              \
    do {                                                                                                   \
        if (prm.trace && blockIdx.x == 0 && (n) < kTwTraceTiles && (ev) < kTwTraceEv)                        \
            prm.trace[((role)*kTwTraceTiles + (n)) * kTwTraceEv + (ev)] = (unsigned long long)clock64();    \
    } while (0)

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// CR = channels per branch (8 / 16 / 32 / 64): compile-time so the tap loop of the MMA issuer is straight-line
// code (12 tcgen05.mma per step with loop-invariant operand offsets; 24 for CR = 8).  A bf16 MMA needs
// K = 16 and N % 16 == 0, so 8-channel branches (the R = 8 stress config at C = 64) run as 16-wide MMAs over
// the pair of branches that shares a 16-channel slice: the host packs each branch's 8x8 tap into its
// diagonal block of a zeroed 16x16 box, and the two branches accumulate into the same 16 columns.
template <int CR>
__global__ void __launch_bounds__(kTwThreads, 1)
tconv_window_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params prm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + prm.bar_off);
    uint64_t *empty = full + kTwMaxStages;
    uint64_t *tfull = empty + kTwMaxStages;      // [4 accumulator buffers]
    uint64_t *tempty = tfull + 4;
    uint64_t *wres = tempty + 4;
    uint64_t *res_full = wres + 1;               // [2 groups][3 slots]
    uint64_t *res_empty = res_full + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_empty + 6);
    float *sbias = reinterpret_cast<float *>(smem + prm.bar_off + 512);    // [64] (barriers + TMEM slot use < 512 B)

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int STAGES = prm.stages, ES = prm.eslots;
    const int q = blockIdx.x % prm.nboxes;            // this CTA's 64-channel box
    const int cta_in_box = blockIdx.x / prm.nboxes;
    const int ctas_per_box = gridDim.x / prm.nboxes;
    const int T = prm.T;
    constexpr int cr = CR, NBR = 64 / CR;
    constexpr int CRM = CR < 16 ? 16 : CR;          // MMA width (K and N) per branch tap

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&maps.h_win);
        tma_prefetch_desc(&maps.xg);
        tma_prefetch_desc(&maps.res);
        tma_prefetch_desc(&maps.out);
        tma_prefetch_desc(&maps.w2);
        tma_prefetch_desc(&maps.wr);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTwMaxStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], kTwGroup);
        }
        for (int s = 0; s < 6; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 1);
        }
        mbar_init(wres, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    if (threadIdx.x < 64) sbias[threadIdx.x] = prm.bias[q * 64 + threadIdx.x];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The step sequence every role walks: pairs of items (i0, i1 = i0 + ctas_per_box), joints v = 0..16,
    // group g = 0, 1 (skipped when the pair has no second item).
    const int first_item = cta_in_box;
    const int item_stride = 2 * ctas_per_box;

    if (warp == 0) {
        // ===== producer: one stage per step = H window of (joint v, box q) [+ projection-input boxes] =====
        if (elect_one()) {
            mbar_expect_tx(wres, prm.w_bytes + prm.wr_bytes);
            for (int rl = 0; rl < NBR; ++rl)
                for (int j = 0; j < 3; ++j)
                    tma_load_2d(smem + prm.w_off + (size_t)(rl * 3 + j) * (CRM * CRM * 2), &maps.w2, wres, 0,
                                ((q * NBR + rl) * 3 + j) * CRM);
            for (int kx = 0; kx < (prm.proj ? prm.nkx : 0); ++kx)
                tma_load_2d(smem + prm.wr_off + (size_t)kx * 8192, &maps.wr, wres, kx * 64, q * 64);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx = (uint32_t)prm.wrows * 128u + (prm.proj ? (uint32_t)prm.nkx * 16384u : 0u);
        for (int i0 = first_item; i0 < prm.nq_items; i0 += item_stride) {
            for (int v = 0; v < 17; ++v) {
                for (int g = 0; g < 2; ++g) {
                    const int item = i0 + g * ctas_per_box;
                    if (item >= prm.nq_items) continue;
                    const int itm = prm.rev ? prm.nq_items - 1 - item : item;
                    const int b = itm / prm.ttiles, t0 = (itm % prm.ttiles) * kFramesTile;
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *sa = smem + (size_t)stage * prm.stage_bytes;
                    if (elect_one()) {
                        TW_TRACE(0, v < 8 ? v : 7, g);
                        mbar_expect_tx(&full[stage], tx);
                        tma_load_4d(sa, &maps.h_win, &full[stage], q * 64, t0 - prm.dmax, v, b);
                        for (int kx = 0; kx < (prm.proj ? prm.nkx : 0); ++kx)
                            tma_load_4d(sa + prm.xg_off + (size_t)kx * 16384, &maps.xg, &full[stage], kx * 64, v, t0, b);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: per step 3 taps x branches-in-box (+ projection chunks) into one 64-column buffer =====
        int stage = 0;
        uint32_t phase = 0;
        uint32_t ng[2] = {0, 0};                 // steps issued per group
        constexpr uint32_t wrow_bytes = (uint32_t)CRM * 2;
        constexpr int ksteps = CRM / 16;
        const uint32_t idesc_tap = make_idesc_bf16((uint32_t)CRM);
        const uint32_t idesc_proj = make_idesc_bf16(64u);
        // Everything that does not depend on the stage is computed ONCE: per (branch, tap) the byte offset of
        // the A start row inside the window and the weight descriptor; per step an MMA costs one 64-bit add.
        // (Descriptors rebuilt next to every tcgen05.mma cost ~200 cycles each through the uniform datapath.)
        constexpr int ntap = 3 * NBR;
        uint32_t aoff[ntap], dcol[ntap];
        uint64_t dbv[ntap];
#pragma unroll
        for (int i = 0; i < ntap; ++i) {
            const int rl = i / 3, j = i % 3;
            const int d = prm.dil[q * NBR + rl];
            const int ch0 = (rl * CR / CRM) * CRM;      // first channel of the MMA slice this branch lives in
            aoff[i] = ((uint32_t)(prm.dmax + (j - 1) * d) * 128u + (uint32_t)(ch0 * 2)) >> 4;
            dcol[i] = (uint32_t)ch0;
            dbv[i] = make_kmajor_desc(smem_u32(smem + prm.w_off + (size_t)i * (CRM * CRM * 2)), wrow_bytes);
        }
        mbar_wait(wres, 0);
        for (int i0 = first_item; i0 < prm.nq_items; i0 += item_stride) {
            for (int v = 0; v < 17; ++v) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (i0 + g * ctas_per_box >= prm.nq_items) continue;
                    const uint32_t n = ng[g];
                    const uint32_t buf = 2u * (n & 1u) + (uint32_t)g;
                    if (lane == 0) TW_TRACE(1, (int)n, 6 + g);
                    mbar_wait(&tempty[buf], ((n >> 1) & 1u) ^ 1u);
                    if (lane == 0) TW_TRACE(1, (int)n, 2 + g);
                    mbar_wait(&full[stage], phase);
                    if (lane == 0) TW_TRACE(1, (int)n, 4 + g);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * prm.stage_bytes);
                    const uint32_t td = tmem_base + buf * 64u;
                    const uint64_t da0 = make_kmajor_desc(sa, 128);
                    // operands are warp-uniform values computed in uniform code; only the tcgen05 instruction itself
                    // sits under the elected lane (operands produced inside the elected branch live in vector
                    // registers and cost an R2UR round trip per MMA: 237 cycles each, measured)
                    if (prm.proj) {
                        for (int kx = 0; kx < prm.nkx; ++kx) {
                            const uint64_t da = make_kmajor_desc(sa + prm.xg_off + (uint32_t)kx * 16384u, 128);
                            const uint64_t db = make_kmajor_desc(smem_u32(smem + prm.wr_off + (size_t)kx * 8192), 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t accum = (uint32_t)((kx > 0) | (k > 0));
                                if (elect_one()) umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_proj, accum);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < ntap; ++i) {
                        {
                            const uint64_t da = da0 + (uint64_t)aoff[i];
                            const uint64_t db = dbv[i];
                            const uint32_t tdr = td + dcol[i];
                            // first tap of the first branch in a column slice overwrites, everything else accumulates
                            const uint32_t accum = (uint32_t)(prm.proj | ((i % 3) > 0) | ((((i / 3) * CR) % CRM) != 0));
#pragma unroll
                            for (int k = 0; k < ksteps; ++k) {
                                const uint32_t acc_k = accum | (uint32_t)(k > 0);
                                if (elect_one()) umma_bf16(tdr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_tap, acc_k);
                            }
                        }
                    }
                    if (elect_one()) {
                        TW_TRACE(1, (int)n, g);
                        umma_commit(&empty[stage]);
                        umma_commit(&tfull[buf]);
                    }
                    __syncwarp();
                    ++ng[g];
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===== residual producers (warp 3 -> group 0, warp 2 -> group 1): the identity-residual box of each
        // step straight into the group's staging slot =====
        if (!prm.proj) {
            const int g = 3 - warp;
            uint32_t ecnt = 0;
            for (int i0 = first_item + g * ctas_per_box; i0 < prm.nq_items; i0 += item_stride) {
                const int itm = prm.rev ? prm.nq_items - 1 - i0 : i0;
                const int b = itm / prm.ttiles, t0 = (itm % prm.ttiles) * kFramesTile;
                for (int v = 0; v < 17; ++v) {
                    const uint32_t sl = ecnt % (uint32_t)ES, ph = (ecnt / (uint32_t)ES) & 1u;
                    uint64_t *rf = &res_full[g * 3 + (int)sl];
                    mbar_wait(&res_empty[g * 3 + (int)sl], ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(rf, 16384u);
                        tma_load_4d(smem + prm.out_off + (size_t)(g * ES + (int)sl) * 16384, &maps.res, rf, q * 64, v, t0, b);
                    }
                    __syncwarp();
                    ++ecnt;
                }
            }
        }
    } else {
        // ===== epilogue group g: item g of every pair, joints 0..16 =====
        const int g = (warp - 4) >> 3;
        const int ew = (warp - 4) & 3;
        const int half = ((warp - 4) >> 2) & 1;
        const int gt = threadIdx.x - 128 - g * kTwGroup;     // 0..255
        const int r = ew * 32 + lane;                        // frame inside the tile
        const bool leader = (gt == 0);
        const int bar_id = 1 + g;
        unsigned char *sout = smem + prm.out_off + (size_t)(g * ES) * 16384;
        float *scr = reinterpret_cast<float *>(smem + prm.scr_off) + g * (2 * 8 * 64);    // [2][8 parts][64]
        const int c2 = gt & 31, part = gt >> 5;              // pooling: channel pair, row part (16 rows each)
        uint32_t n = 0;                                      // steps done by this group
        int pending = -1;
        // deferred finalisation of the previous step's joint-pooling partial sums
        int fin_b = -1, fin_tt = 0, fin_v = 0;
        auto finalize = [&](uint32_t step) {
            if (fin_b >= 0 && gt < 32) {
                const float *sp = scr + (step & 1u) * (8 * 64);
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const float2 t = *reinterpret_cast<const float2 *>(sp + p * 64 + 2 * gt);
                    acc.x += t.x;
                    acc.y += t.y;
                }
                *reinterpret_cast<float2 *>(prm.PVpart + (((size_t)fin_b * prm.ttiles + fin_tt) * 17 + fin_v) * prm.C + q * 64 +
                                            2 * gt) = acc;
            }
        };
        for (int i0 = first_item + g * ctas_per_box; i0 < prm.nq_items; i0 += item_stride) {
            const int itm = prm.rev ? prm.nq_items - 1 - i0 : i0;
            const int b = itm / prm.ttiles, tt = itm % prm.ttiles, t0 = tt * kFramesTile;
            const int nvalid = min(kFramesTile, T - t0);
            float pt[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) pt[e] = 0.f;
            for (int v = 0; v < 17; ++v, ++n) {
                const uint32_t buf = 2u * (n & 1u) + (uint32_t)g;
                const uint32_t es = n % (uint32_t)ES, eph = (n / (uint32_t)ES) & 1u;
                unsigned char *box = sout + (size_t)es * 16384;
                if (leader) TW_TRACE(3 + g, (int)n, 5);
                mbar_wait(&tfull[buf], (n >> 1) & 1u);
                if (leader) TW_TRACE(3 + g, (int)n, 6);
                tc_fence_after();
                // a warp whose 32 frames all lie past the end of the clip (last frame tile: T = 300 leaves 44 of
                // 128 rows) skips the accumulator read and the row math; it still takes part in every handshake
                const bool rows_live = ew * 32 < nvalid;
                uint32_t acc[32];
                if (rows_live) tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 64u + (uint32_t)(half * 32), acc);
                if (!prm.proj) mbar_wait(&res_full[g * 3 + (int)es], eph);   // slot is ours and holds the residual box
                if (rows_live) tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&tempty[buf]);
                if (leader) TW_TRACE(3 + g, (int)n, 0);
                unsigned char *rowp = box + (size_t)r * 128;
                const float *bq = sbias + half * 32;
                if (rows_live)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(bq + cc * 8);
                    const float4 b1 = *reinterpret_cast<const float4 *>(bq + cc * 8 + 4);
                    float f[8];
                    f[0] = __uint_as_float(acc[cc * 8 + 0]) + b0.x; f[1] = __uint_as_float(acc[cc * 8 + 1]) + b0.y;
                    f[2] = __uint_as_float(acc[cc * 8 + 2]) + b0.z; f[3] = __uint_as_float(acc[cc * 8 + 3]) + b0.w;
                    f[4] = __uint_as_float(acc[cc * 8 + 4]) + b1.x; f[5] = __uint_as_float(acc[cc * 8 + 5]) + b1.y;
                    f[6] = __uint_as_float(acc[cc * 8 + 6]) + b1.z; f[7] = __uint_as_float(acc[cc * 8 + 7]) + b1.w;
                    uint4 *sp = reinterpret_cast<uint4 *>(rowp + (((half * 4 + cc) ^ (r & 7)) << 4));
                    if (!prm.proj) {
                        const uint4 rr = *sp;
                        const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 t = __bfloat1622float2(rp[e]);
                            f[2 * e] += t.x;
                            f[2 * e + 1] += t.y;
                        }
                    }
                    uint4 packed;
                    __nv_bfloat162 *pp = reinterpret_cast<__nv_bfloat162 *>(&packed);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        f[2 * e] = fmaxf(f[2 * e], 0.f);
                        f[2 * e + 1] = fmaxf(f[2 * e + 1], 0.f);
                        pp[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                        // frame pooling (sum over joints) of the fp32 values, as the oracle pools (the stored
                        // copy is their bf16 rounding; re-expanding it cost 2 more instructions per pair)
                        pt[cc * 8 + 2 * e] += f[2 * e];
                        pt[cc * 8 + 2 * e + 1] += f[2 * e + 1];
                    }
                    *sp = packed;
                }
                fence_proxy_async_smem();
                // projection blocks have no slot producer: the store issued a step ago (the only one pending)
                // must have read its slot before the NEXT step overwrites it; that step starts after this barrier
                if (leader && prm.proj) tma_store_wait_read0();
                if (leader) TW_TRACE(3 + g, (int)n, 1);
                asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
                if (leader) {
                    TW_TRACE(3 + g, (int)n, 2);
                    tma_store_4d(&maps.out, box, q * 64, v, t0, b);
                    tma_store_commit();
                    if (!prm.proj) {
                        if (pending >= 0) {
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            mbar_arrive(&res_empty[g * 3 + pending]);
                        }
                        pending = (int)es;
                    }
                }
                // joint pooling: column sums of the staged tile over its valid frames, 8 row parts -> scratch;
                // the 8-way fold of the PREVIOUS step's scratch is published by this step's barrier
                finalize(n + 1);
                {
                    float2 a2 = make_float2(0.f, 0.f);
                    const uint32_t coff = (uint32_t)(c2 & 3) * 4u;
                    const int cchunk = c2 >> 2;
                    float2 t[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int row = part * 16 + k;
                        t[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(
                            box + (size_t)row * 128 + ((cchunk ^ (row & 7)) << 4) + coff));
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (part * 16 + k < nvalid) {
                            a2.x += t[k].x;
                            a2.y += t[k].y;
                        }
                    *reinterpret_cast<float2 *>(scr + (n & 1u) * (8 * 64) + part * 64 + 2 * c2) = a2;
                }
                if (leader) TW_TRACE(3 + g, (int)n, 3);
                fin_b = b;
                fin_tt = tt;
                fin_v = v;
            }
            // frame-pooling sums of this item: 32 channels of frame r, 128 contiguous bytes per thread
            if (r < nvalid) {
                float4 *dst = reinterpret_cast<float4 *>(prm.PT + ((size_t)b * T + (size_t)(t0 + r)) * prm.C + q * 64 + half * 32);
#pragma unroll
                for (int e = 0; e < 8; ++e) dst[e] = make_float4(pt[4 * e], pt[4 * e + 1], pt[4 * e + 2], pt[4 * e + 3]);
            }
        }
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        finalize(n + 1);
        if (leader) {
            tma_store_wait_read0();
            tma_store_wait_all0();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// bf16 [B][T][V][width] activation seen as (C, V, T, B): box (box_w channels, 1 joint, box_t frames, 1 clip);
// the box lands in shared memory as box_t rows of box_w channels (128 B rows, SW128).
inline int make_btvc_joint_map(CUtensorMap *m, const void *base, int width, int T, int B, int box_w, int box_t) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GS_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)width, 17, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)width * 2, (cuuint64_t)width * 2 * 17, (cuuint64_t)width * 2 * 17 * (cuuint64_t)T};
    cuuint32_t box[4] = {(cuuint32_t)box_w, 1, (cuuint32_t)box_t, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_w * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(btvc joint box %dx%d) failed: %d", box_w, box_t, (int)r);
        return GS_ERR_CUDA;
    }
    return GS_OK;
}

// bf16 joint-major H [B][V][T][width]:
//   as (C, T, V, B), box (box_w, box_t, 1, 1): the tconv window load;
//   as (C, V, T, B), box (64, 17, 7, 1): the 1x1 GEMM's store of a 7-frame (119-row, frame-major) tile.
inline int make_bvtc_map(CUtensorMap *m, const void *base, int width, int T, int B, bool frame_major_box, int box_w,
                         int box_t) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GS_ERR_CUDA;
    }
    const cuuint64_t sT = (cuuint64_t)width * 2, sV = sT * (cuuint64_t)T, sB = sV * 17;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    dims[0] = (cuuint64_t)width;
    box[0] = (cuuint32_t)box_w;
    if (frame_major_box) {   // (C, V, T, B)
        dims[1] = 17; dims[2] = (cuuint64_t)T; dims[3] = (cuuint64_t)B;
        strides[0] = sV; strides[1] = sT; strides[2] = sB;
        box[1] = 17; box[2] = (cuuint32_t)box_t; box[3] = 1;
    } else {                 // (C, T, V, B)
        dims[1] = (cuuint64_t)T; dims[2] = 17; dims[3] = (cuuint64_t)B;
        strides[0] = sT; strides[1] = sV; strides[2] = sB;
        box[1] = (cuuint32_t)box_t; box[2] = 1; box[3] = 1;
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_w * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(bvtc box %dx%d, frame-major %d) failed: %d", box_w, box_t, (int)frame_major_box, (int)r);
        return GS_ERR_CUDA;
    }
    return GS_OK;
}

inline bool plan(Params &p) {
    const uint32_t limit = 227u * 1024u;
    p.a_span = ((uint32_t)p.wrows * 128u + 1023u) & ~1023u;
    p.xg_off = p.a_span;
    p.stage_bytes = p.a_span + (p.proj ? (uint32_t)p.nkx * 16384u : 0u);
    const int crm = p.cr < 16 ? 16 : p.cr;
    p.w_bytes = (uint32_t)(p.nbr * 3 * crm * crm * 2);
    p.wr_bytes = p.proj ? (uint32_t)p.nkx * 8192u : 0u;
    const uint32_t wspan = ((p.w_bytes + 1023u) & ~1023u) + p.wr_bytes;
    // three staging slots per group give the residual box two steps of lookahead (its TMA latency is about
    // one step); projection blocks have no residual box and stay with two
    for (int es = p.proj ? 2 : 3; es >= 2; --es) {
        const uint32_t fixed = wspan + 2u * es * 16384u + 8192u /*pooling scratch: 2 groups x [2][8][64] floats*/ +
                               1024u /*barriers + bias*/ + 1024u /*slack*/;
        for (int st = kTwMaxStages; st >= (es == 3 ? 4 : 2); --st) {
            if (p.stage_bytes * st + fixed <= limit) {
                p.stages = st;
                p.eslots = es;
                p.w_off = p.stage_bytes * st;
                p.wr_off = p.w_off + ((p.w_bytes + 1023u) & ~1023u);
                p.out_off = p.w_off + wspan;
                p.scr_off = p.out_off + 2u * es * 16384u;
                p.bar_off = p.scr_off + 8192u;
                p.total = p.bar_off + 1024u + 1024u;
                return true;
            }
        }
    }
    return false;
}

struct LaunchTw {
    Maps maps;
    Params prm;
    double flops = 0, bytes = 0;
};

inline int launch(Ctx *ctx, int kid, LaunchTw &L, cudaStream_t st) {
    if (!plan(L.prm)) {
        set_error("tconv_window: shared memory plan does not fit (C=%d cin=%d dmax=%d)", L.prm.C, L.prm.cin, L.prm.dmax);
        return GS_ERR_UNSUPPORTED;
    }
    int grid = (ctx->sm_count / L.prm.nboxes) * L.prm.nboxes;
    const int need = L.prm.nq_items * L.prm.nboxes;
    if (grid > need) grid = need;
    if (grid < 1) return GS_OK;
    auto kern = L.prm.cr == 8 ? tconv_window_kernel<8>
                              : (L.prm.cr == 16 ? tconv_window_kernel<16>
                                                : (L.prm.cr == 32 ? tconv_window_kernel<32> : tconv_window_kernel<64>));
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.prm.total));
    {
        LaunchScope ls(ctx, kid, st, L.flops, L.bytes);
        kern<<<grid, kTwThreads, L.prm.total, st>>>(L.maps, L.prm);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace tw
}  // namespace gs

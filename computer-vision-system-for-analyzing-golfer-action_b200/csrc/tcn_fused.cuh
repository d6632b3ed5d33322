// Multi-branch temporal convolution for sm_100a, fused: branch 1x1 + dilated taps + residual in ONE kernel.
//
//   H[b,t,v, :]       = relu(Y[b,t,v, :] . W1 + b1)                         zero outside [0,T)
//   U[b,t,v, r*cr+co] = relu( sum_{j<3, ci<cr} H[b, t+(j-1)d_r, v, r*cr+ci] * W2[r,j,ci,co] + b2
//                             + residual[b,t,v, r*cr+co] )
//   residual = the block's gated input (identity) or its 1x1 projection Xg . Wr (width change)
// Stage replaced: /root/reference/README.md:29-30 (Temporal Module - Multi-branch Temporal Convolution).
//
// H never reaches HBM (round 1 ran the 1x1 as its own GEMM: one write and one read of a [rows, C]
// tensor per block, 28 % of the step's DRAM traffic).  The graph-conv kernel stores Y JOINT-MAJOR
// ([B,V,T,C], a free transposition in its TMA store map), so one M tile is 128 consecutive FRAMES of one
// joint and a temporal shift of d frames is d rows:
//   * a step = (clip, frame tile, joint) for the CTA's 64-channel output box q.  Its input is the
//     128-frame window [t0-dmax, t0+128-dmax) of Y, all C channels, as C/64 TMA boxes that stream
//     through a ring; frames outside [0,T) are zero-filled by TMA;
//   * MMA A (tcgen05, SS): Hacc[128 x 64] = Ywin[128 x C] . W1[:, box q]      (W1 slice resident in smem)
//   * the step's epilogue group turns Hacc into the bf16, 128B-swizzled K-major box Hbox (+b1, ReLU,
//     rows whose frame lies outside [0,T) forced to zero = the conv's zero padding);
//   * MMA B: the 3 taps of every branch in the box are MMAs whose A descriptors start at ROW offsets
//     dmax+(j-1)d inside Hbox (the swizzle is a function of the absolute smem address: any row offset
//     works), accumulator row m = frame t0+m; rows m >= 128-2*dmax read past the window and are
//     discarded, so a tile yields nout = 128-2*dmax output frames;
//   * epilogue: +b2, + residual box (TMA-loaded into the staging slot) or the projection MMAs' sum,
//     ReLU, bf16, TMA store; frame-pooling sums PT stay in registers over the 17 joints of an item,
//     joint-pooling partials PVpart are column sums over each warp's 32 rows, reduced by shuffles from the
//     fp32 registers (SE / ST-joint attention).
// A CTA owns one output box (its W1 slice, tap and projection weights stay resident) and walks its
// (clip, frame-tile) items joint by joint.  The 16 epilogue warps run convert(step n), then
// epilogue(step n-1): MMA B of step n and MMA A of step n+1 run under the epilogue of step n-1.
//
// Warp roles (640 threads, 1 CTA/SM): w0 TMA producer (Y windows, projection-input boxes), w1 TMEM alloc + 1x1
// issuer (MMA A, up to two steps ahead: two Hacc buffers), w2-17 epilogue warps (TMEM lanes = rows by w%4, 16 of
// the box's 64 columns by (w-2)/4: 16 pooling sums + a 16-column accumulator slice per thread keep the row
// math inside 96 registers), w18 store warp (TMA stores of the staged tiles, identity-residual box loads two
// steps ahead), w19 tap issuer (MMA B).  No CTA barrier in the steady state: every hand-over is an mbarrier,
// so the epilogue warps drift apart and hide each other's latencies, and the two issuers never wait behind
// each other's hand-shakes (one issuer thread paced the loop at its own serial waits: traced).
// TMEM (256 columns): [0,128) two 64-column U accumulators, [128,256) two Hacc buffers.
#pragma once
#include "umma.cuh"

namespace gs {
namespace tf {

using namespace tc;

constexpr int kTfThreads = 640;      // producer + 1x1 issuer + 16 epilogue warps + store warp + tap issuer
constexpr int kTfEpi = 512;
constexpr int kWin = 128;            // window rows = UMMA M
constexpr int kTfMaxSlots = 8;
constexpr uint32_t kBoxBytes = 16384;   // 128 rows x 64 bf16

struct Params {
    int B, T, C, cr, cin;
    int nbr;          // branches inside one 64-channel box (64 / cr)
    int proj;         // 1: residual = Xg . Wr (extra MMAs), 0: identity residual box added in the epilogue
    int nkx;          // projection K boxes (cin / 64)
    int nky;          // 1x1 K boxes (C / 64)
    int dil[GS_MAX_BRANCHES];
    int dmax, nout;   // largest dilation; output frames per tile = 128 - 2*dmax
    int ttiles, nboxes, nq_items;   // frame tiles per clip, 64-channel boxes, items per box (B * ttiles)
    int slots, eslots, xslots;   // Y ring depth (boxes); staging slots (2 or 3); projection-input ring depth (0, 2..4)
    uint32_t w1_off, w2_off, wr_off, w1_bytes, w2_bytes, wr_bytes, hbox_off, hbox_span, out_off, xring_off, bar_off, total;
    const float *bias1;   // [C]  b1 (folded BN of the 1x1)
    const float *bias;    // [C]  b2 (+ folded projection bias)
    float *PT;            // [B,T,C]
    float *PVpart;        // [B,ttiles*4,17,C]: partial sums per (frame tile, 32-row quarter)
    unsigned long long *trace;   // optional clock64 trace of CTA 0 (GOLFER_TRACE_TCN=1, tools/trace_tcn.py)
};

// trace layout: [role 0..3][step 0..kTrSteps)[event 0..kTrEv): roles 0 = first epilogue warp (warp 2), 1 = last epilogue
// warp, 2 = 1x1 issuer, 3 = tap issuer; steps kTrFirst.. of CTA 0
constexpr int kTrSteps = 8, kTrEv = 16, kTrFirst = 20;
#define TF_TRACE(role, step, ev)                                                                          \
    do {                                                                                                  \
        if (prm.trace && blockIdx.x == 0 && (int)(step) >= kTrFirst && (int)(step) < kTrFirst + kTrSteps)  \
            prm.trace[((role)*kTrSteps + ((int)(step)-kTrFirst)) * kTrEv + (ev)] = (unsigned long long)clock64(); \
    } while (0)

struct Maps {
    CUtensorMap y_win;    // Y joint-major [B,V,T,C]: box (64 ch, 128 frames, 1 joint, 1 clip)
    CUtensorMap xg;       // Xg [B,T,V,cin] seen as (C, V, T, B): box (64 ch, 1 joint, 128 frames, 1)  (projection input)
    CUtensorMap res;      // identity residual [B,T,V,C]: box (64 ch, 1 joint, nout frames, 1)
    CUtensorMap out;      // U [B,T,V,C], same box shape
    CUtensorMap w1;       // 1x1 weights W1T [C n][C k]: box (64 k, 64 n)
    CUtensorMap w2;       // tap weights [(r*3+j)*cr + co][ci], box (cr, cr)
    CUtensorMap wr;       // projection weights [C][cin], box (64 k, 64 n)
};

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
__global__ void __launch_bounds__(kTfThreads, 1)
tcn_fused_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params prm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + prm.bar_off);
    uint64_t *empty = full + kTfMaxSlots;
    uint64_t *tfull = empty + kTfMaxSlots;       // [2 U accumulators]
    uint64_t *tempty = tfull + 2;
    uint64_t *hfull = tempty + 2;                // [2] Hacc buffer written by MMA A
    uint64_t *hempty = hfull + 2;                // [2] Hacc buffer read by the epilogue warps
    uint64_t *hready = hempty + 2;               // Hbox written and visible to the async proxy
    uint64_t *wres = hready + 1;
    uint64_t *res_full = wres + 1;               // [3 staging slots] residual box landed (identity blocks)
    uint64_t *out_ready = res_full + 3;          // [3] staged output tile complete (all 512 epilogue threads)
    uint64_t *slot_free = out_ready + 3;         // [3] the slot's store has read it (projection blocks)
    uint64_t *xfull = slot_free + 3;             // [4] projection-input ring
    uint64_t *xempty = xfull + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(xempty + 4);
    float *sbias = reinterpret_cast<float *>(smem + prm.bar_off + 512);    // [64] b2, then [64] b1

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int SLOTS = prm.slots, ES = prm.eslots;
    const int q = blockIdx.x % prm.nboxes;            // this CTA's 64-channel box
    const int cta_in_box = blockIdx.x / prm.nboxes;
    const int ctas_per_box = gridDim.x / prm.nboxes;
    const int T = prm.T, nout = prm.nout;
    constexpr int NBR = 64 / CR;
    constexpr int CRM = CR < 16 ? 16 : CR;          // MMA width (K and N) per branch tap

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&maps.y_win);
        tma_prefetch_desc(&maps.xg);
        tma_prefetch_desc(&maps.res);
        tma_prefetch_desc(&maps.out);
        tma_prefetch_desc(&maps.w1);
        tma_prefetch_desc(&maps.w2);
        tma_prefetch_desc(&maps.wr);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTfMaxSlots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], kTfEpi);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&hfull[s], 1);
            mbar_init(&hempty[s], kTfEpi);
        }
        mbar_init(hready, kTfEpi);
        for (int s = 0; s < 3; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&out_ready[s], kTfEpi);
            mbar_init(&slot_free[s], 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&xfull[s], 1);
            mbar_init(&xempty[s], 1);
        }
        mbar_init(wres, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    if (threadIdx.x < 64) sbias[threadIdx.x] = prm.bias[q * 64 + threadIdx.x];
    else if (threadIdx.x < 128) sbias[threadIdx.x] = prm.bias1[q * 64 + threadIdx.x - 64];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColH = 128;                  // TMEM: U accumulators at columns [0,64) [64,128), Hacc buffers at [128,192) [192,256)

    // This CTA's items: cta_in_box, + ctas_per_box, ...; step s = (item s/17, joint s%17).
    const int nitems = cta_in_box < prm.nq_items ? (prm.nq_items - cta_in_box + ctas_per_box - 1) / ctas_per_box : 0;
    const int nsteps = nitems * 17;

    if (warp == 0) {
        // ===== producer: weights once, then per step the C/64 Y-window boxes (and, projection blocks, the cin/64
        // input boxes of the PREVIOUS step), in the order the MMA issuer consumes them: A(0) [A(1) B(0)] [A(2) B(1)] ...
        if (elect_one()) {
            mbar_expect_tx(wres, prm.w1_bytes + prm.w2_bytes + prm.wr_bytes);
            for (int kb = 0; kb < prm.nky; ++kb)
                tma_load_2d(smem + prm.w1_off + (size_t)kb * 8192, &maps.w1, wres, kb * 64, q * 64);
            for (int rl = 0; rl < NBR; ++rl)
                for (int j = 0; j < 3; ++j)
                    tma_load_2d(smem + prm.w2_off + (size_t)(rl * 3 + j) * (CRM * CRM * 2), &maps.w2, wres, 0,
                                ((q * NBR + rl) * 3 + j) * CRM);
            for (int kx = 0; kx < (prm.proj ? prm.nkx : 0); ++kx)
                tma_load_2d(smem + prm.wr_off + (size_t)kx * 8192, &maps.wr, wres, kx * 64, q * 64);
        }
        __syncwarp();
        int slot = 0, xslot = 0;
        uint32_t phase = 0, xphase = 0;
        // Y-window boxes of step st into the main ring; projection-input boxes into their own small ring (each
        // ring has one consumer: the 1x1 issuer / the tap issuer)
        auto load_y = [&](int st) {
            const int item = cta_in_box + (st / 17) * ctas_per_box, v = st % 17;
            const int b = item / prm.ttiles, t0 = (item % prm.ttiles) * nout;
            for (int kb = 0; kb < prm.nky; ++kb) {
                mbar_wait(&empty[slot], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[slot], kBoxBytes);
                    tma_load_4d(smem + (size_t)slot * kBoxBytes, &maps.y_win, &full[slot], kb * 64, t0 - prm.dmax, v, b);
                }
                __syncwarp();
                if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
        };
        auto load_x = [&](int st) {
            const int item = cta_in_box + (st / 17) * ctas_per_box, v = st % 17;
            const int b = item / prm.ttiles, t0 = (item % prm.ttiles) * nout;
            for (int kx = 0; kx < prm.nkx; ++kx) {
                mbar_wait(&xempty[xslot], xphase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&xfull[xslot], kBoxBytes);
                    tma_load_4d(smem + prm.xring_off + (size_t)xslot * kBoxBytes, &maps.xg, &xfull[xslot], kx * 64, v, t0, b);
                }
                __syncwarp();
                if (++xslot == prm.xslots) { xslot = 0; xphase ^= 1; }
            }
        };
        // the 1x1 runs up to two steps ahead of the taps: Y(st+1) before X(st)
        if (nsteps > 0) load_y(0);
        for (int st = 0; st < nsteps; ++st) {
            if (st + 1 < nsteps) load_y(st + 1);
            if (prm.proj) load_x(st);
        }
    } else if (warp == 1) {
        // ===== 1x1 issuer: A(n): Hacc[n & 1] = Ywin . W1[:, box q].  Runs ahead of the epilogue warps by up to two
        // steps (two Hacc buffers); its own warp, so it never waits behind the tap issuer's hand-shakes. =====
        int slot = 0;
        uint32_t phase = 0;
        const uint32_t idesc_64 = make_idesc_bf16(64u);
        mbar_wait(wres, 0);
        for (int st = 0; st < nsteps; ++st) {
            const uint32_t n = (uint32_t)st, hb = n & 1u;
            if (lane == 0) TF_TRACE(2, n, 0);
            mbar_wait(&hempty[hb], ((n >> 1) & 1u) ^ 1u);
            if (lane == 0) TF_TRACE(2, n, 1);
            tc_fence_after();
            const uint32_t th = tmem_base + kColH + hb * 64u;
            for (int kb = 0; kb < prm.nky; ++kb) {
                mbar_wait(&full[slot], phase);
                tc_fence_after();
                const uint64_t da = make_kmajor_desc(smem_u32(smem + (size_t)slot * kBoxBytes), 128);
                const uint64_t db = make_kmajor_desc(smem_u32(smem + prm.w1_off + (size_t)kb * 8192), 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t accum = (uint32_t)((kb > 0) | (k > 0));
                    if (elect_one()) umma_bf16(th, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_64, accum);
                }
                if (elect_one()) umma_commit(&empty[slot]);
                __syncwarp();
                if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&hfull[hb]);
            __syncwarp();
            if (lane == 0) TF_TRACE(2, n, 2);
        }
    } else if (warp == 19) {
        // ===== tap issuer: B(n) = projection chunks + 3 taps x branches-in-box into one 64-column U accumulator =====
        int xslot = 0;
        uint32_t xphase = 0;
        constexpr uint32_t wrow_bytes = (uint32_t)CRM * 2;
        constexpr int ksteps = CRM / 16;
        const uint32_t idesc_tap = make_idesc_bf16((uint32_t)CRM);
        const uint32_t idesc_64 = make_idesc_bf16(64u);
        // Everything that does not depend on the step is computed ONCE: per (branch, tap) the byte offset of
        // the A start row inside Hbox and the weight descriptor; per step an MMA costs one 64-bit add.
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
            dbv[i] = make_kmajor_desc(smem_u32(smem + prm.w2_off + (size_t)i * (CRM * CRM * 2)), wrow_bytes);
        }
        const uint64_t dh = make_kmajor_desc(smem_u32(smem + prm.hbox_off), 128);
        mbar_wait(wres, 0);
        for (int st = 0; st < nsteps; ++st) {
            const uint32_t n = (uint32_t)st, buf = n & 1u;
            if (lane == 0) TF_TRACE(3, n, 4);
            mbar_wait(&tempty[buf], ((n >> 1) & 1u) ^ 1u);
            if (lane == 0) TF_TRACE(3, n, 5);
            tc_fence_after();
            const uint32_t td = tmem_base + buf * 64u;
            if (prm.proj) {
                for (int kx = 0; kx < prm.nkx; ++kx) {
                    mbar_wait(&xfull[xslot], xphase);
                    tc_fence_after();
                    const uint64_t da = make_kmajor_desc(smem_u32(smem + prm.xring_off + (size_t)xslot * kBoxBytes), 128);
                    const uint64_t db = make_kmajor_desc(smem_u32(smem + prm.wr_off + (size_t)kx * 8192), 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t accum = (uint32_t)((kx > 0) | (k > 0));
                        if (elect_one()) umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_64, accum);
                    }
                    if (elect_one()) umma_commit(&xempty[xslot]);
                    __syncwarp();
                    if (++xslot == prm.xslots) { xslot = 0; xphase ^= 1; }
                }
            }
            mbar_wait(hready, n & 1u);
            if (lane == 0) TF_TRACE(3, n, 6);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < ntap; ++i) {
                const uint64_t da = dh + (uint64_t)aoff[i];
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
            if (elect_one()) umma_commit(&tfull[buf]);     // also: Hbox may be rewritten
            __syncwarp();
            if (lane == 0) TF_TRACE(3, n, 7);
        }
    } else if (warp == 18) {
        // ===== store warp: per step, once all 512 epilogue threads have staged the tile: TMA store; once the
        // PREVIOUS step's store has read its slot: that slot gets the identity-residual box of the step that uses it
        // next (ES-1 steps ahead), or (projection blocks) is handed back through slot_free. =====
        unsigned char *sout = smem + prm.out_off;
        auto load_residual = [&](int st) {
            const int item = cta_in_box + (st / 17) * ctas_per_box;
            const int sl = st % ES;
            mbar_expect_tx(&res_full[sl], (uint32_t)nout * 128u);
            tma_load_4d(sout + (size_t)sl * kBoxBytes, &maps.res, &res_full[sl], q * 64, st % 17, (item % prm.ttiles) * nout,
                        item / prm.ttiles);
        };
        if (!prm.proj && elect_one())
            for (int st = 0; st < ES && st < nsteps; ++st) load_residual(st);
        __syncwarp();
        int sl = 0, v = 0, item = cta_in_box;
        uint32_t ph = 0;
        int b = item / prm.ttiles, t0 = (item % prm.ttiles) * nout;
        for (int m = 0; m < nsteps; ++m) {
            mbar_wait(&out_ready[sl], ph);
            if (elect_one()) {
                tma_store_4d(&maps.out, sout + (size_t)sl * kBoxBytes, q * 64, v, t0, b);
                tma_store_commit();
                if (m > 0) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // store(m-1) has read its slot
                    if (prm.proj) mbar_arrive(&slot_free[(m - 1) % ES]);
                    else if (m - 1 + ES < nsteps) load_residual(m - 1 + ES);
                }
            }
            __syncwarp();
            if (++sl == ES) { sl = 0; ph ^= 1u; }
            if (++v == 17) {
                v = 0;
                item += ctas_per_box;
                b = item / prm.ttiles;
                t0 = (item - b * prm.ttiles) * nout;
            }
        }
        if (elect_one()) {
            tma_store_wait_read0();
            tma_store_wait_all0();
        }
        __syncwarp();
    } else {
        // ===== epilogue warps: per iteration convert(n) and epilogue(n-1).  Warp w owns TMEM lanes (rows)
        // 32*(w%4).. and columns 16*cq.., cq = (w-2)/4.  The 16 warps never meet at a CTA barrier: they talk to the
        // MMA issuer and the store warp through mbarriers only, so their latencies overlap instead of adding up
        // behind the slowest warp.  Everything that does not change per step is hoisted: shared addresses as
        // 32-bit values, step coordinates advanced incrementally (a division per ITEM, none per step), both
        // accumulator slices fetched up front.
        const int ew = warp & 3;                             // TMEM lane quarter this warp may access
        const int cq = (warp - 2) >> 2;                      // 16-column quarter of the box
        const int r = ew * 32 + lane;                        // row inside the tile
        const uint32_t t_h = tmem_base + ((uint32_t)(ew * 32) << 16) + kColH + (uint32_t)(cq * 16);
        const uint32_t t_u = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(cq * 16);
        // this thread's two 16-byte chunks of a 128-byte box row (SW128: chunk index XOR row & 7)
        const uint32_t row_off0 = (uint32_t)r * 128u + (uint32_t)(((cq * 2) ^ (r & 7)) << 4);
        const uint32_t row_off1 = (uint32_t)r * 128u + (uint32_t)(((cq * 2 + 1) ^ (r & 7)) << 4);
        const uint32_t hb0 = smem_u32(smem + prm.hbox_off) + row_off0, hb1 = smem_u32(smem + prm.hbox_off) + row_off1;
        const uint32_t out_base = smem_u32(smem + prm.out_off);
        const uint32_t a_hfull = smem_u32(hfull), a_hempty = smem_u32(hempty), a_hready = smem_u32(hready);
        const uint32_t a_tfull = smem_u32(tfull), a_tempty = smem_u32(tempty);
        const uint32_t a_slot = smem_u32(prm.proj ? slot_free : res_full), a_outready = smem_u32(out_ready);
        float bias2[16];                                     // b2 of this thread's 16 channels (b1 is re-read: registers)
#pragma unroll
        for (int e = 0; e < 16; ++e) bias2[e] = sbias[cq * 16 + e];
        const float4 *bq1 = reinterpret_cast<const float4 *>(sbias + 64 + cq * 16);
        // joint pooling: after the transposing butterfly below, lane l holds the column sum of channel
        // ((l >> 1) & 15 bit-reversed into (bit4, bit3, bit2, bit1) -> 8, 4, 2, 1) over the warp's 32 rows
        const int pv_ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        float pt[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pt[e] = 0.f;
        // step coordinates: convert side (step n) and epilogue side (step n-1), advanced incrementally
        int cv = 0, c_item = cta_in_box;                     // convert: joint, item
        bool c_inside = false;
        int e_v = 0, e_item = cta_in_box, e_b = 0, e_tt = 0, e_t0 = 0, e_nvalid = 0;
        uint32_t e_slot = 0, e_ph = 0;                       // staging slot of the epilogue step and its phase
        auto item_coords = [&](int item, int &b, int &tt) { b = item / prm.ttiles; tt = item - b * prm.ttiles; };
        {
            int b0, tt0;
            item_coords(c_item, b0, tt0);
            const int frame = tt0 * nout - prm.dmax + r;
            c_inside = frame >= 0 && frame < T;
            e_b = b0; e_tt = tt0; e_t0 = tt0 * nout; e_nvalid = min(nout, T - e_t0);
        }
        // Iteration n: convert(n) and epilogue(n-2).  The epilogue lags TWO steps: MMA B of step n-1 starts only when
        // the slowest warp has converted step n-1, so an epilogue of step n-1 here would make every warp wait for
        // the slowest one at every step (traced: 15 % of all stall samples on that wait); step n-2's accumulator
        // was completed an iteration ago.
        for (uint32_t n = 0; n < (uint32_t)nsteps + 2u; ++n) {
            const bool do_cv = n < (uint32_t)nsteps, do_ep = n >= 2;
            const uint32_t m = n - 2;
            const uint32_t buf = m & 1u;
            const bool rows_live = do_ep && ew * 32 < e_nvalid;
            uint32_t acch[16], accu[16];
            const int trole = (warp == 2 && lane == 0) ? 0 : ((warp == 17 && lane == 0) ? 1 : -1);
            if (trole >= 0) TF_TRACE(trole, n, 0);
            if (do_ep) mbar_wait_u32(a_tfull + buf * 8u, (m >> 1) & 1u);
            if (do_cv) mbar_wait_u32(a_hfull + (n & 1u) * 8u, (n >> 1) & 1u);
            if (trole >= 0) TF_TRACE(trole, n, 2);
            tc_fence_after();
            if (do_cv) tmem_ld16(t_h + (n & 1u) * 64u, acch);
            if (rows_live) tmem_ld16(t_u + buf * 64u, accu);
            tmem_ld_wait();
            if (trole >= 0) TF_TRACE(trole, n, 3);
            tc_fence_before();
            if (do_cv) mbar_arrive_u32(a_hempty + (n & 1u) * 8u);
            if (do_ep) mbar_arrive_u32(a_tempty + buf * 8u);
            // Hbox is rewritten below: the taps of step n-1 must have finished reading it (their accumulator is complete)
            if (do_cv && n > 0) mbar_wait_u32(a_tfull + ((n - 1) & 1u) * 8u, ((n - 1) >> 1) & 1u);
            if (do_cv) {
                // ---- convert(n): Hacc -> +b1, ReLU, zero outside [0,T) -> bf16 K-major Hbox
                float bias1[16];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 t = bq1[e];
                    bias1[4 * e] = t.x; bias1[4 * e + 1] = t.y; bias1[4 * e + 2] = t.z; bias1[4 * e + 3] = t.w;
                }
                uint4 p0, p1;
                p0.x = pack_bf16(fmaxf(__uint_as_float(acch[0]) + bias1[0], 0.f), fmaxf(__uint_as_float(acch[1]) + bias1[1], 0.f));
                p0.y = pack_bf16(fmaxf(__uint_as_float(acch[2]) + bias1[2], 0.f), fmaxf(__uint_as_float(acch[3]) + bias1[3], 0.f));
                p0.z = pack_bf16(fmaxf(__uint_as_float(acch[4]) + bias1[4], 0.f), fmaxf(__uint_as_float(acch[5]) + bias1[5], 0.f));
                p0.w = pack_bf16(fmaxf(__uint_as_float(acch[6]) + bias1[6], 0.f), fmaxf(__uint_as_float(acch[7]) + bias1[7], 0.f));
                p1.x = pack_bf16(fmaxf(__uint_as_float(acch[8]) + bias1[8], 0.f), fmaxf(__uint_as_float(acch[9]) + bias1[9], 0.f));
                p1.y = pack_bf16(fmaxf(__uint_as_float(acch[10]) + bias1[10], 0.f), fmaxf(__uint_as_float(acch[11]) + bias1[11], 0.f));
                p1.z = pack_bf16(fmaxf(__uint_as_float(acch[12]) + bias1[12], 0.f), fmaxf(__uint_as_float(acch[13]) + bias1[13], 0.f));
                p1.w = pack_bf16(fmaxf(__uint_as_float(acch[14]) + bias1[14], 0.f), fmaxf(__uint_as_float(acch[15]) + bias1[15], 0.f));
                if (!c_inside) p0 = p1 = make_uint4(0u, 0u, 0u, 0u);
                st_shared_v4(hb0, p0);
                st_shared_v4(hb1, p1);
                fence_proxy_async_smem();
                mbar_arrive_u32(a_hready);
                if (trole >= 0) TF_TRACE(trole, n, 5);
                if (++cv == 17) {          // next convert step starts a new item: frame validity of this row
                    cv = 0;
                    c_item += ctas_per_box;
                    int b0, tt0;
                    item_coords(c_item, b0, tt0);
                    const int frame = tt0 * nout - prm.dmax + r;
                    c_inside = frame >= 0 && frame < T;
                }
            }
            if (!do_ep) continue;
            // ---- epilogue(n-2)
            const uint32_t box_u32 = out_base + e_slot * kBoxBytes;
            // identity blocks: the slot holds the residual box (which also means its previous store has drained);
            // projection blocks: the slot's previous store has drained (first use of a slot passes at once)
            mbar_wait_u32(a_slot + e_slot * 8u, prm.proj ? (e_ph ^ 1u) : e_ph);
            if (trole >= 0) TF_TRACE(trole, n, 6);
            float f[16];
            if (rows_live) {
                // a warp whose 32 rows all lie past the end of the tile / clip skips the row math; it still takes
                // part in every handshake
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(accu[e]) + bias2[e];
                if (!prm.proj) {
                    const uint4 r0 = ld_shared_v4(box_u32 + row_off0), r1 = ld_shared_v4(box_u32 + row_off1);
                    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        f[2 * e] += bf16lo_to_f32(rw[e]);
                        f[2 * e + 1] += bf16hi_to_f32(rw[e]);
                    }
                }
                const bool row_valid = r < e_nvalid;         // rows past the tile / clip: stored nowhere, pooled as zero
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    f[e] = row_valid ? fmaxf(f[e], 0.f) : 0.f;
                    // frame pooling (sum over joints) of the fp32 values, as the oracle pools (the stored copy is
                    // their bf16 rounding)
                    pt[e] += f[e];
                }
                st_shared_v4(box_u32 + row_off0, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
                st_shared_v4(box_u32 + row_off1, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) f[e] = 0.f;
            }
            fence_proxy_async_smem();
            mbar_arrive_u32(a_outready + e_slot * 8u);
            if (trole >= 0) TF_TRACE(trole, n, 8);
            // ---- joint pooling: column sums over this warp's 32 rows straight from the registers (fp32, fixed
            // order).  Transposing butterfly: each exchange halves the channels a lane keeps and doubles the rows
            // they cover: 16 shuffles + 16 adds for 16 channels x 32 rows.
            {
                float g8[8], g4[4], g2[2], g1;
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float send = b4 ? f[i] : f[i + 8], keep = b4 ? f[i + 8] : f[i];
                    g8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float send = b3 ? g8[i] : g8[i + 4], keep = b3 ? g8[i + 4] : g8[i];
                    g4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float send = b2 ? g4[i] : g4[i + 2], keep = b2 ? g4[i + 2] : g4[i];
                    g2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                {
                    const float send = b1 ? g2[0] : g2[1], keep = b1 ? g2[1] : g2[0];
                    g1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
                g1 += __shfl_xor_sync(0xffffffffu, g1, 1);
                // partial of (clip, frame tile, row quarter ew): PVpart[b][tt*4 + ew][v][c]
                if (!(lane & 1))
                    prm.PVpart[((((size_t)e_b * prm.ttiles + e_tt) * 4 + ew) * 17 + e_v) * prm.C + q * 64 + cq * 16 + pv_ch] = g1;
            }
            if (trole >= 0) TF_TRACE(trole, n, 12);
            if (++e_slot == (uint32_t)ES) { e_slot = 0; e_ph ^= 1u; }
            if (++e_v == 17) {
                // frame-pooling sums of this item: 16 channels of frame r, 64 contiguous bytes per thread
                if (r < e_nvalid) {
                    float4 *dst = reinterpret_cast<float4 *>(prm.PT + ((size_t)e_b * T + (size_t)(e_t0 + r)) * prm.C + q * 64 + cq * 16);
#pragma unroll
                    for (int e = 0; e < 4; ++e) dst[e] = make_float4(pt[4 * e], pt[4 * e + 1], pt[4 * e + 2], pt[4 * e + 3]);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) pt[e] = 0.f;
                e_v = 0;
                e_item += ctas_per_box;
                item_coords(e_item, e_b, e_tt);
                e_t0 = e_tt * nout;
                e_nvalid = min(nout, T - e_t0);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
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

// bf16 joint-major Y [B][V][T][width]:
//   as (C, T, V, B), box (box_w, box_t, 1, 1): the window load of this kernel;
//   as (C, V, T, B), box (64, 17, 7, 1): the graph-conv kernel's store of a 7-frame (119-row, frame-major) tile.
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

// Shared-memory plan: resident weights, Hbox (128 + 2*dmax rows: the taps of the discarded
// accumulator rows read past row 127), staging slots, pooling scratch, barriers; the rest is the box ring.
inline bool plan(Params &p) {
    const uint32_t limit = 227u * 1024u;
    const int crm = p.cr < 16 ? 16 : p.cr;
    p.w1_bytes = (uint32_t)p.nky * 8192u;
    p.w2_bytes = (uint32_t)(p.nbr * 3 * crm * crm * 2);
    p.wr_bytes = p.proj ? (uint32_t)p.nkx * 8192u : 0u;
    p.hbox_span = (((uint32_t)(kWin + 2 * p.dmax) * 128u) + 1023u) & ~1023u;
    const uint32_t wspan = p.w1_bytes + ((p.w2_bytes + 1023u) & ~1023u) + p.wr_bytes;
    // identity blocks: three staging slots give the residual box two steps of lookahead (its TMA latency is
    // about one step); projection blocks have no residual box and take two
    const int es_hi = p.proj ? 2 : 3, es_lo = 2;
    for (int min_slots = 4; min_slots >= 2; --min_slots) {
        for (int es = es_hi; es >= es_lo; --es) {
            const int xs = p.proj ? (p.nkx >= 2 ? p.nkx : 2) : 0;       // one step of projection-input boxes (at least 2)
            const uint32_t fixed = wspan + p.hbox_span + (uint32_t)(es + xs) * kBoxBytes + 1024u /*barriers + biases*/ +
                                   1024u /*alignment slack*/;
            if (fixed + (uint32_t)min_slots * kBoxBytes > limit) continue;
            int st = (int)((limit - fixed) / kBoxBytes);
            if (st > kTfMaxSlots) st = kTfMaxSlots;
            p.slots = st;
            p.eslots = es;
            p.w1_off = (uint32_t)st * kBoxBytes;
            p.w2_off = p.w1_off + p.w1_bytes;
            p.wr_off = p.w2_off + ((p.w2_bytes + 1023u) & ~1023u);
            p.hbox_off = p.w1_off + wspan;
            p.out_off = p.hbox_off + p.hbox_span;
            p.xslots = xs;
            p.xring_off = p.out_off + (uint32_t)es * kBoxBytes;
            p.bar_off = p.xring_off + (uint32_t)xs * kBoxBytes;
            p.total = p.bar_off + 1024u + 1024u;
            return true;
        }
    }
    return false;
}

struct LaunchTf {
    Maps maps;
    Params prm;
    double flops = 0, bytes = 0;
};

inline int launch(Ctx *ctx, int kid, LaunchTf &L, cudaStream_t st) {
    if (!plan(L.prm)) {
        set_error("tcn_fused: shared memory plan does not fit (C=%d cin=%d dmax=%d)", L.prm.C, L.prm.cin, L.prm.dmax);
        return GS_ERR_UNSUPPORTED;
    }
    int grid = (ctx->sm_count / L.prm.nboxes) * L.prm.nboxes;
    const int need = L.prm.nq_items * L.prm.nboxes;
    if (grid > need) grid = need;
    if (grid < 1) return GS_OK;
    typedef void (*Kern)(const Maps, const Params);
    const Kern kern = L.prm.cr == 8 ? (Kern)tcn_fused_kernel<8>
                                    : (L.prm.cr == 16 ? (Kern)tcn_fused_kernel<16>
                                                      : (L.prm.cr == 32 ? (Kern)tcn_fused_kernel<32> : (Kern)tcn_fused_kernel<64>));
    int rc = ensure_dyn_smem(ctx, (const void *)kern, L.prm.total);
    if (rc != GS_OK) return rc;
    {
        LaunchScope ls(ctx, kid, st, L.flops, L.bytes);
        kern<<<grid, kTfThreads, L.prm.total, st>>>(L.maps, L.prm);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace tf
}  // namespace gs

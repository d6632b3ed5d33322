// Multi-branch temporal convolution for sm_100a, fused: branch 1x1 + dilated taps + residual in ONE kernel.
//
//   H[b,t,v, :]       = relu(Y[b,t,v, :] . W1 + b1)                         zero outside [0,T)
//   U[b,t,v, r*cr+co] = relu( sum_{j<3, ci<cr} H[b, t+(j-1)d_r, v, r*cr+ci] * W2[r,j,ci,co] + b2
//                             + residual[b,t,v, r*cr+co] )
//   residual = the block's gated input (identity) or its 1x1 projection Xg . Wr (width change); BOTH run through
//   the tensor core into the U accumulator: the identity as Xg[:, box q] . I_64 (1.0 * x is exact in the fp32
//   accumulator).  Adding the residual box in the epilogue instead (a TMA load per step into the staging slot, an
//   unpack and 16 adds per thread) measured 7-20 % slower per launch: the epilogue warps are the bound.
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
//   * both biases enter through the tensor core as well: one extra K = 16 MMA per accumulator whose A operand is a
//     resident all-ones tile and whose B operand holds the bias split into two bf16 terms (hi + lo, 2^-17 relative)
//     in its first two K columns: 32 FADDs per thread and step less in the warps that bound the kernel;
//   * the step's epilogue group turns Hacc into the bf16, 128B-swizzled K-major box Hbox (ReLU inside the
//     bf16x2 conversion, rows whose frame lies outside [0,T) forced to zero = the conv's zero padding);
//   * MMA B: the 3 taps of every branch in the box are MMAs whose A descriptors start at ROW offsets
//     dmax+(j-1)d inside Hbox (the swizzle is a function of the absolute smem address: any row offset
//     works), accumulator row m = frame t0+m; rows m >= 128-2*dmax read past the window and are
//     discarded, so a tile yields nout = 128-2*dmax output frames;
//   * epilogue: ReLU (bias and residual are already in the accumulator), bf16, TMA store; frame-pooling sums PT stay in registers over the 17 joints of an item,
//     joint-pooling partials PVpart are column sums of each staged tile (SE / ST-joint attention).
// A CTA owns one output box (its W1 slice, tap and projection weights stay resident) and walks its
// (clip, frame-tile) items joint by joint.  The 16 epilogue warps run convert(step n), then
// epilogue(step n-1): MMA B of step n and MMA A of step n+1 run under the epilogue of step n-1.
//
// Warp roles (576 threads, 1 CTA/SM): w0 TMA producer (Y windows, projection-input boxes), w1 TMEM alloc +
// MMA issuer, w2-17 epilogue warps (TMEM lanes = rows by w%4, 16 of the box's 64 columns by (w-2)/4: 16
// pooling sums + a 16-column accumulator slice per thread keep the row math inside 96 registers).  The
// leader epilogue thread issues the TMA stores.
// TMEM (256 columns): [0,128) two 64-column U accumulators, [128,256) two 64-column Hacc buffers.  With ONE Hacc the
// issuer's work sat inside the loop that paced the kernel (trace at C = 256: convert(n) done -> 12 tap MMAs issued
// 365 cycles -> the 17 MMAs of A(n+1) issued 900 cycles, 52 per MMA -> completion 350 -> TMEM load + convert(n+1)
// 750 = the 2600-cycle step); with two, A(n+2) is issued behind B(n) and is complete long before convert(n+2) asks.
//
// Measured and kept out (all parity-green, experiments/tcn_fused_decoupled.cuh.txt): hand-overs through
// mbarriers only (no CTA barrier per step, a store warp, shuffle-reduced pooling), two issuer warps with a
// double-buffered Hacc: 2.16 ms over the six launches of a 256-clip step against 2.18 ms for this simpler
// lock-step form, i.e. the same.  The chain convert -> taps -> epilogue makes all 16 warps wait for the
// slowest one at every step whichever primitive carries the hand-over; ncu: issue slots 57 % busy, no pipe
// above 35 %, top stall long-scoreboard (TMEM loads, mbarrier waits).
#pragma once
#include "umma.cuh"

namespace gs {
namespace tf {

using namespace tc;

constexpr int kTfThreads = 576;      // producer + issuer + 16 epilogue warps
constexpr int kTfEpi = 512;
constexpr int kWin = 128;            // window rows = UMMA M
constexpr int kTfMaxSlots = 8;
constexpr uint32_t kBoxBytes = 16384;   // 128 rows x 64 bf16

struct Params {
    int B, T, C, cr, cin;
    int nbr;          // branches inside one 64-channel box (64 / cr)
    int res_q;        // 0: residual = Xg . Wr (cin / 64 K boxes); 1: identity residual: the only K box is channel box q
                      //    of the block input and Wr is the 64x64 identity
    int nkx;          // residual K boxes (cin / 64; 1 with res_q)
    int nky;          // 1x1 K boxes (C / 64)
    int dil[GS_MAX_BRANCHES];
    int dmax, nout;   // largest dilation; output frames per tile = 128 - 2*dmax
    int ttiles, nboxes, nq_items;   // frame tiles per clip, 64-channel boxes, items per box (B * ttiles)
    int slots, eslots;   // ring depth (boxes); staging slots (2)
    uint32_t w1_off, w2_off, wr_off, w1_bytes, w2_bytes, wr_bytes, bias_off, hbox_off, hbox_span, out_off, scr_off, bar_off, total;
    float *PT;            // [B,T,C]
    float *PVpart;        // [B,ttiles,17,C]
    float b1v[256], b2v[256];   // CR = 64 (C = 256) only: the two biases by value, added in registers from the constant bank
    unsigned long long *trace;   // optional clock64 trace of CTA 0 (GOLFER_TRACE_TCN=1, tools/trace_tcn.py)
};

// trace layout: [role 0..3][step 0..kTrSteps)[event 0..kTrEv): roles 0 = epilogue leader (warp 2), 1 = last epilogue
// warp, 2 = MMA issuer; steps kTrFirst.. of CTA 0
constexpr int kTrSteps = 8, kTrEv = 16, kTrFirst = 20;
// compiled in only with -DGOLFER_TCN_TRACE (tools/trace_tcn.py builds its own copy of the library): the trace
// points cost ~10 % of the kernel even when the buffer pointer is null
#ifdef GOLFER_TCN_TRACE
#define TF_TRACE(role, step, ev)                                                                          \
    do {                                                                                                  \
        if (prm.trace && blockIdx.x == 0 && (int)(step) >= kTrFirst && (int)(step) < kTrFirst + kTrSteps)  \
            prm.trace[((role)*kTrSteps + ((int)(step)-kTrFirst)) * kTrEv + (ev)] = (unsigned long long)clock64(); \
    } while (0)
#else
#define TF_TRACE(role, step, ev) do { } while (0)
#endif

struct Maps {
    CUtensorMap y_win;    // Y joint-major [B,V,T,C]: box (64 ch, 128 frames, 1 joint, 1 clip)
    CUtensorMap xg;       // residual source Xg [B,T,V,cin] seen as (C, V, T, B): box (64 ch, 1 joint, 128 frames, 1)
    CUtensorMap out;      // U [B,T,V,C], same box shape
    CUtensorMap w1;       // 1x1 weights W1T [C n][C k]: box (64 k, 64 n)
    CUtensorMap w2;       // tap weights [(r*3+j)*cr + co][ci], box (cr, cr)
    CUtensorMap wr;       // projection weights [C][cin] (or the 64x64 identity), box (64 k, 64 n)
    CUtensorMap bt1, bt2; // bias tiles [C][16]: column 0 = bf16(b), column 1 = bf16(b - column 0), rest 0; box (16 k, 64 n)
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// L2 prefetch of a box (no shared-memory slot, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *m, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wres + 1);

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int SLOTS = prm.slots, ES = prm.eslots;
    const int q = blockIdx.x % prm.nboxes;            // this CTA's 64-channel box
    const int cta_in_box = blockIdx.x / prm.nboxes;
    const int ctas_per_box = gridDim.x / prm.nboxes;
    const int T = prm.T, nout = prm.nout;
    constexpr int NBR = 64 / CR;
    constexpr int CRM = CR < 16 ? 16 : CR;          // MMA width (K and N) per branch tap
    constexpr bool kBiasRegs = (CR >= 64);          // C = 256: biases added in registers instead of through the tensor core

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&maps.y_win);
        tma_prefetch_desc(&maps.xg);
        tma_prefetch_desc(&maps.out);
        tma_prefetch_desc(&maps.w1);
        tma_prefetch_desc(&maps.w2);
        tma_prefetch_desc(&maps.wr);
        tma_prefetch_desc(&maps.bt1);
        tma_prefetch_desc(&maps.bt2);
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
        mbar_init(wres, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    // the all-ones A tile of the bias MMAs: [128 rows][16 k] bf16, 32-byte rows (invariant under the 32B swizzle)
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 256)
        *reinterpret_cast<uint4 *>(smem + prm.bias_off + (size_t)(threadIdx.x - 64) * 16) =
            make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColH = 128;                  // TMEM: U accumulators at columns [0,64) [64,128), Hacc at [128,192) [192,256)

    // This CTA's items: cta_in_box, + ctas_per_box, ...; step s = (item s/17, joint s%17).
    const int nitems = cta_in_box < prm.nq_items ? (prm.nq_items - cta_in_box + ctas_per_box - 1) / ctas_per_box : 0;
    const int nsteps = nitems * 17;

    if (warp == 0) {
        // ===== producer: weights once, then the C/64 Y-window boxes two steps ahead and the residual-input boxes of
        // this step, in the order the MMA issuer consumes them: A(0) A(1) [B(0) A(2)] [B(1) A(3)] ...
        if (elect_one()) {
            mbar_expect_tx(wres, prm.w1_bytes + prm.w2_bytes + prm.wr_bytes + 4096u);
            tma_load_2d(smem + prm.bias_off + 4096, &maps.bt1, wres, 0, q * 64);
            tma_load_2d(smem + prm.bias_off + 6144, &maps.bt2, wres, 0, q * 64);
            for (int kb = 0; kb < prm.nky; ++kb)
                tma_load_2d(smem + prm.w1_off + (size_t)kb * 8192, &maps.w1, wres, kb * 64, q * 64);
            for (int rl = 0; rl < NBR; ++rl)
                for (int j = 0; j < 3; ++j)
                    tma_load_2d(smem + prm.w2_off + (size_t)(rl * 3 + j) * (CRM * CRM * 2), &maps.w2, wres, 0,
                                ((q * NBR + rl) * 3 + j) * CRM);
            for (int kx = 0; kx < prm.nkx; ++kx)
                tma_load_2d(smem + prm.wr_off + (size_t)kx * 8192, &maps.wr, wres, kx * 64, prm.res_q ? 0 : q * 64);
        }
        __syncwarp();
        int slot = 0;
        uint32_t phase = 0;
        // boxes of step st: which = 0 the Y-window boxes, which = 1 the projection-input boxes
        auto load_boxes = [&](int st, int which) {
            const int item = cta_in_box + (st / 17) * ctas_per_box, v = st % 17;
            const int b = item / prm.ttiles, t0 = (item % prm.ttiles) * nout;
            const int nb = which == 0 ? prm.nky : prm.nkx;
            for (int kb = 0; kb < nb; ++kb) {
                mbar_wait(&empty[slot], phase ^ 1);
                unsigned char *sa = smem + (size_t)slot * kBoxBytes;
                if (elect_one()) {
                    mbar_expect_tx(&full[slot], kBoxBytes);
                    if (which == 0) tma_load_4d(sa, &maps.y_win, &full[slot], kb * 64, t0 - prm.dmax, v, b);
                    else tma_load_4d(sa, &maps.xg, &full[slot], prm.res_q ? q * 64 : kb * 64, v, t0, b);
                }
                __syncwarp();
                if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
        };
        // The ring holds 5-8 boxes (80-128 KB in flight) and the Y windows come from DRAM (the graph-conv kernel wrote
        // 170-670 MB just before): measured, the C = 256 launches slow down 17 % with 4 slots instead of 6, i.e. they
        // are bound by bytes in flight x latency.  When the ring cannot hold one step's boxes (the 128 -> 256 block: 6
        // boxes, 5 slots) each CTA asks L2 for ITS share of the boxes kPfAhead steps ahead (Y box q of the window - the
        // nboxes CTAs of an item cover the window between them - and its residual box), so the ring's loads find them
        // in L2: 0.61 -> 0.56 ms on that launch.  With a deeper ring the prefetches only cost (+3 %): off.
        auto prefetch_boxes = [&](int st) {
            const int item = cta_in_box + (st / 17) * ctas_per_box, v = st % 17;
            const int b = item / prm.ttiles, t0 = (item % prm.ttiles) * nout;
            if (elect_one()) {
                tma_prefetch_4d(&maps.y_win, q * 64, t0 - prm.dmax, v, b);
                if (prm.res_q) tma_prefetch_4d(&maps.xg, q * 64, v, t0, b);
                else if (q < prm.nkx) tma_prefetch_4d(&maps.xg, q * 64, v, t0, b);
            }
            __syncwarp();
        };
        constexpr int kPfAhead = 6;
        const bool pf = prm.nky + prm.nkx >= SLOTS;
        for (int st = 0; pf && st < kPfAhead && st < nsteps; ++st) prefetch_boxes(st);
        if (nsteps > 0) load_boxes(0, 0);
        if (nsteps > 1) load_boxes(1, 0);
        for (int st = 0; st < nsteps; ++st) {
            if (pf && st + kPfAhead < nsteps) prefetch_boxes(st + kPfAhead);
            load_boxes(st, 1);
            if (st + 2 < nsteps) load_boxes(st + 2, 0);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: A(n) = bias + the 1x1 into Hacc[n & 1], B(n) = bias + residual chunks + 3 taps x
        // branches-in-box into one 64-column U accumulator.  Order A(0) A(1) [B(0) A(2)] [B(1) A(3)] ... =====
        int slot = 0;
        uint32_t phase = 0;
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
        const uint64_t d_ring = make_kmajor_desc(smem_u32(smem), 128);                  // ring slot 0; + slot * (16384 >> 4)
        const uint64_t d_w1 = make_kmajor_desc(smem_u32(smem + prm.w1_off), 128);       // + kb * (8192 >> 4)
        const uint64_t d_wr = make_kmajor_desc(smem_u32(smem + prm.wr_off), 128);
        const uint64_t d_ones = make_kmajor_desc(smem_u32(smem + prm.bias_off), 32);
        const uint64_t d_b1 = make_kmajor_desc(smem_u32(smem + prm.bias_off + 4096), 32);
        const uint64_t d_b2 = make_kmajor_desc(smem_u32(smem + prm.bias_off + 6144), 32);
        mbar_wait(wres, 0);
        auto issue_a = [&](uint32_t n) {
            if (lane == 0) TF_TRACE(2, n, 0);
            const uint32_t hb = n & 1u;
            const uint32_t th = tmem_base + kColH + hb * 64u;
            mbar_wait(&hempty[hb], ((n >> 1) & 1u) ^ 1u);
            if (lane == 0) TF_TRACE(2, n, 1);
            tc_fence_after();
            if (!kBiasRegs && elect_one()) umma_bf16(th, d_ones, d_b1, idesc_64, 0u);     // Hacc = b1
            for (int kb = 0; kb < prm.nky; ++kb) {
                mbar_wait(&full[slot], phase);
                tc_fence_after();
                const uint64_t da = d_ring + (uint64_t)((uint32_t)slot * (kBoxBytes >> 4));
                const uint64_t db = d_w1 + (uint64_t)((uint32_t)kb * (8192u >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(th, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_64,
                                  kBiasRegs ? (uint32_t)((kb > 0) | (k > 0)) : 1u);
                    umma_commit(&empty[slot]);
                }
                __syncwarp();
                if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&hfull[hb]);
            __syncwarp();
            if (lane == 0) TF_TRACE(2, n, 2);
        };
        auto issue_b = [&](uint32_t n) {
            const uint32_t buf = n & 1u;
            if (lane == 0) TF_TRACE(2, n, 4);
            mbar_wait(&tempty[buf], ((n >> 1) & 1u) ^ 1u);
            if (lane == 0) TF_TRACE(2, n, 5);
            tc_fence_after();
            const uint32_t td = tmem_base + buf * 64u;
            if (!kBiasRegs && elect_one()) umma_bf16(td, d_ones, d_b2, idesc_64, 0u);     // U = b2 (+ projection bias)
            for (int kx = 0; kx < prm.nkx; ++kx) {
                mbar_wait(&full[slot], phase);
                tc_fence_after();
                const uint64_t da = d_ring + (uint64_t)((uint32_t)slot * (kBoxBytes >> 4));
                const uint64_t db = d_wr + (uint64_t)((uint32_t)kx * (8192u >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_64,
                                  kBiasRegs ? (uint32_t)((kx > 0) | (k > 0)) : 1u);
                    umma_commit(&empty[slot]);
                }
                __syncwarp();
                if (++slot == SLOTS) { slot = 0; phase ^= 1; }
            }
            mbar_wait(hready, n & 1u);
            if (lane == 0) TF_TRACE(2, n, 6);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < ntap; ++i) {
                const uint64_t da = dh + (uint64_t)aoff[i];
                const uint64_t db = dbv[i];
                const uint32_t tdr = td + dcol[i];
                // the bias and residual MMAs above started the accumulator: every tap accumulates
#pragma unroll
                for (int k = 0; k < ksteps; ++k)
                    if (elect_one()) umma_bf16(tdr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_tap, 1u);
            }
            if (elect_one()) umma_commit(&tfull[buf]);     // also: Hbox may be rewritten
            __syncwarp();
            if (lane == 0) TF_TRACE(2, n, 7);
        };
        if (nsteps > 0) issue_a(0u);
        if (nsteps > 1) issue_a(1u);
        for (int st = 0; st < nsteps; ++st) {
            issue_b((uint32_t)st);
            if (st + 2 < nsteps) issue_a((uint32_t)st + 2u);
        }
    } else {
        // ===== epilogue warps: per iteration convert(n) and epilogue(n-1).  Warp w owns TMEM lanes (rows)
        // 32*(w%4).. and columns 16*cq.., cq = (w-2)/4.  The loop is bound by instruction issue and by the chain
        // convert -> taps -> epilogue, so everything that does not change per step is hoisted: shared addresses as
        // 32-bit values, step coordinates advanced incrementally (a division per ITEM, none per step), both
        // accumulator slices fetched up front.
        const int ew = warp & 3;                             // TMEM lane quarter this warp may access
        const int cq = (warp - 2) >> 2;                      // 16-column quarter of the box
        const float *b1q = prm.b1v + q * 64 + cq * 16, *b2q = prm.b2v + q * 64 + cq * 16;   // kBiasRegs: constant bank
        const int gt = threadIdx.x - 64;                     // 0..511
        const int r = ew * 32 + lane;                        // row inside the tile
        const bool leader = (gt == 0);
        const uint32_t t_h = tmem_base + ((uint32_t)(ew * 32) << 16) + kColH + (uint32_t)(cq * 16);
        const uint32_t t_u = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(cq * 16);
        // this thread's two 16-byte chunks of a 128-byte box row (SW128: chunk index XOR row & 7)
        const uint32_t row_off0 = (uint32_t)r * 128u + (uint32_t)(((cq * 2) ^ (r & 7)) << 4);
        const uint32_t row_off1 = (uint32_t)r * 128u + (uint32_t)(((cq * 2 + 1) ^ (r & 7)) << 4);
        const uint32_t hb0 = smem_u32(smem + prm.hbox_off) + row_off0, hb1 = smem_u32(smem + prm.hbox_off) + row_off1;
        const uint32_t out_base = smem_u32(smem + prm.out_off);
        const uint32_t a_hfull = smem_u32(hfull), a_hempty = smem_u32(hempty), a_hready = smem_u32(hready);   // [2], [2], [1]
        const uint32_t a_tfull = smem_u32(tfull), a_tempty = smem_u32(tempty);
        unsigned char *sout = smem + prm.out_off;
        float *scr = reinterpret_cast<float *>(smem + prm.scr_off);    // [2][16 parts][64]
        const int c2 = gt & 31, part = gt >> 5;              // pooling: channel pair, row part (8 rows each)
        // pooling reads: 8 rows of this thread's channel pair; the swizzled chunk depends on row & 7 = k only
        const uint32_t pool_off = (uint32_t)part * 1024u + (uint32_t)(c2 & 3) * 4u;
        // deferred finalisation of the previous step's joint-pooling partial sums
        int fin_b = -1, fin_tt = 0, fin_v = 0;
        auto finalize = [&](uint32_t step) {
            if (fin_b >= 0 && gt < 32) {
                const float *sp = scr + (step & 1u) * (16 * 64);
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const float2 t = *reinterpret_cast<const float2 *>(sp + p * 64 + 2 * gt);
                    acc.x += t.x;
                    acc.y += t.y;
                }
                *reinterpret_cast<float2 *>(prm.PVpart + (((size_t)fin_b * prm.ttiles + fin_tt) * 17 + fin_v) * prm.C + q * 64 +
                                            2 * gt) = acc;
            }
        };
        float pt[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pt[e] = 0.f;
        // step coordinates: convert side (step n) and epilogue side (step n-1), advanced incrementally
        int cv = 0, c_item = cta_in_box;                     // convert: joint, item
        bool c_inside = false;
        int e_v = 0, e_item = cta_in_box, e_b = 0, e_tt = 0, e_t0 = 0, e_nvalid = 0;
        uint32_t e_slot = 0;                                 // staging slot of the epilogue step
        auto item_coords = [&](int item, int &b, int &tt) { b = item / prm.ttiles; tt = item - b * prm.ttiles; };
        {
            int b0, tt0;
            item_coords(c_item, b0, tt0);
            const int frame = tt0 * nout - prm.dmax + r;
            c_inside = frame >= 0 && frame < T;
            e_b = b0; e_tt = tt0; e_t0 = tt0 * nout; e_nvalid = min(nout, T - e_t0);
        }
        for (uint32_t n = 0; n <= (uint32_t)nsteps; ++n) {
            const bool do_cv = n < (uint32_t)nsteps, do_ep = n > 0;
            const uint32_t m = n - 1;
            const uint32_t buf = m & 1u;
            const bool rows_live = do_ep && ew * 32 < e_nvalid;
            uint32_t acch[16], accu[16];
            const int trole = gt == 0 ? 0 : (gt == 480 ? 1 : -1);
            if (trole >= 0) TF_TRACE(trole, n, 0);
            if (do_ep) {
                // U accumulator of step n-1 complete => its MMAs have also finished reading Hbox
                mbar_wait_u32(a_tfull + buf * 8u, (m >> 1) & 1u);
            }
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
            if (do_cv) {
                // ---- convert(n): Hacc (b1 included) -> ReLU + bf16 in one conversion, zero outside [0,T) -> K-major Hbox
                uint4 p0, p1;
                if (kBiasRegs) {
                    // C = 256: the kernel sits at the shared-memory pipe, and the bias MMAs read 12 KB of operands per
                    // step through it; here the biases come from the constant bank (warp-uniform index: no register,
                    // no shared-memory load) and cost 32 FADDs per thread and step in warps that have the slack
#pragma unroll
                    for (int e = 0; e < 16; ++e) acch[e] = __float_as_uint(__uint_as_float(acch[e]) + b1q[e]);
                }
                p0.x = pack_bf16_relu(__uint_as_float(acch[0]), __uint_as_float(acch[1]));
                p0.y = pack_bf16_relu(__uint_as_float(acch[2]), __uint_as_float(acch[3]));
                p0.z = pack_bf16_relu(__uint_as_float(acch[4]), __uint_as_float(acch[5]));
                p0.w = pack_bf16_relu(__uint_as_float(acch[6]), __uint_as_float(acch[7]));
                p1.x = pack_bf16_relu(__uint_as_float(acch[8]), __uint_as_float(acch[9]));
                p1.y = pack_bf16_relu(__uint_as_float(acch[10]), __uint_as_float(acch[11]));
                p1.z = pack_bf16_relu(__uint_as_float(acch[12]), __uint_as_float(acch[13]));
                p1.w = pack_bf16_relu(__uint_as_float(acch[14]), __uint_as_float(acch[15]));
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
            // ---- epilogue(n-1)
            const uint32_t box_u32 = out_base + e_slot * kBoxBytes;
            if (trole >= 0) TF_TRACE(trole, n, 6);
            if (rows_live) {
                // a warp whose 32 rows all lie past the end of the tile / clip skips the row math; it still takes
                // part in every handshake
                float f[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    f[e] = fmaxf(kBiasRegs ? __uint_as_float(accu[e]) + b2q[e] : __uint_as_float(accu[e]), 0.f);
                    // frame pooling (sum over joints) of the fp32 values, as the oracle pools (the stored copy is
                    // their bf16 rounding)
                    pt[e] += f[e];
                }
                st_shared_v4(box_u32 + row_off0, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
                st_shared_v4(box_u32 + row_off1, make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15])));
            }
            fence_proxy_async_smem();
            // two staging slots: the store issued a step ago (the only one pending) must have read its slot before the
            // NEXT step overwrites it; that step starts after this barrier
            if (leader) tma_store_wait_read0();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (trole >= 0) TF_TRACE(trole, n, 9);
            if (leader) {
                tma_store_4d(&maps.out, sout + (size_t)e_slot * kBoxBytes, q * 64, e_v, e_t0, e_b);
                tma_store_commit();
            }
            // joint pooling: column sums of the staged tile over its valid frames, 16 row parts -> scratch;
            // the 16-way fold of the PREVIOUS step's scratch is published by this step's barrier
            finalize(m + 1);
            {
                float2 a2 = make_float2(0.f, 0.f);
                const int cchunk = c2 >> 2;
                const int nrows = e_nvalid - part * 8;       // valid rows among this thread's 8
                if (nrows > 0) {
                    uint32_t w[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[k]) : "r"(box_u32 + pool_off + (uint32_t)k * 128u + (uint32_t)((cchunk ^ k) << 4)));
                    if (nrows >= 8) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            a2.x += bf16lo_to_f32(w[k]);
                            a2.y += bf16hi_to_f32(w[k]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < nrows) {
                                a2.x += bf16lo_to_f32(w[k]);
                                a2.y += bf16hi_to_f32(w[k]);
                            }
                    }
                }
                *reinterpret_cast<float2 *>(scr + (m & 1u) * (16 * 64) + part * 64 + 2 * c2) = a2;
            }
            if (trole >= 0) TF_TRACE(trole, n, 12);
            fin_b = e_b;
            fin_tt = e_tt;
            fin_v = e_v;
            if (++e_slot == (uint32_t)ES) e_slot = 0;
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
        asm volatile("bar.sync 1, 512;" ::: "memory");
        finalize((uint32_t)nsteps + 1u);
        if (leader) {
            tma_store_wait_read0();
            tma_store_wait_all0();
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
    p.wr_bytes = (uint32_t)p.nkx * 8192u;
    p.hbox_span = (((uint32_t)(kWin + 2 * p.dmax) * 128u) + 1023u) & ~1023u;
    const uint32_t wspan = p.w1_bytes + ((p.w2_bytes + 1023u) & ~1023u) + p.wr_bytes + 8192u /*ones + bias tiles*/;
    const int es_hi = 2, es_lo = 2;
    for (int min_slots = 4; min_slots >= 2; --min_slots) {
        for (int es = es_hi; es >= es_lo; --es) {
            const uint32_t fixed = wspan + p.hbox_span + (uint32_t)es * kBoxBytes + 8192u /*pooling scratch*/ +
                                   1024u /*barriers + biases*/ + 1024u /*alignment slack*/;
            if (fixed + (uint32_t)min_slots * kBoxBytes > limit) continue;
            int st = (int)((limit - fixed) / kBoxBytes);
            if (st > kTfMaxSlots) st = kTfMaxSlots;
            p.slots = st;
            p.eslots = es;
            p.w1_off = (uint32_t)st * kBoxBytes;
            p.w2_off = p.w1_off + p.w1_bytes;
            p.wr_off = p.w2_off + ((p.w2_bytes + 1023u) & ~1023u);
            p.bias_off = p.wr_off + p.wr_bytes;
            p.hbox_off = p.w1_off + wspan;
            p.out_off = p.hbox_off + p.hbox_span;
            p.scr_off = p.out_off + (uint32_t)es * kBoxBytes;
            p.bar_off = p.scr_off + 8192u;
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

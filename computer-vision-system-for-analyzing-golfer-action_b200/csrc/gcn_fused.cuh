// Fused spatial graph convolution for sm_100a (blocks with Cin >= 64):
//
//   Xg = U_prev * gT * gV                         (deferred attention gates, applied on load)
//   XA[w, p*Cin + c] = sum_v A_p[w,v] * Xg[v, c]  (adjacency contraction, per frame)
//   Y  = relu(XA . Wg + bg)                       (1x1 channel mix, K = 3*Cin)
//
// in ONE kernel: neither XA (3x the input) nor an un-gated copy of the input ever reaches HBM.
// Stage replaced: /root/reference/README.md:27-28 (Spatial Module - Graph Convolution).
//
// A tile is 7 whole frames (119 rows, padded to the 128-row UMMA M).  Both contractions run on
// tcgen05 tensor cores:
//   MMA1 (TS form)  D1[128 x 64]  = Abig_p[128 x 128] . Xg[128 rows x 64 ch]
//        Abig_p = I_7 (x) A_p, the block-diagonal adjacency, lives in TENSOR MEMORY as the
//        A operand (bf16, 64 columns per partition); Xg is the TMA-loaded box itself, used
//        in place as an MN-major B operand (channels contiguous).
//   convert         D1 (fp32, TMEM) -> bf16 -> 128B-swizzled K-major smem chunk
//   MMA2 (SS form)  acc[128 x C] += XAchunk[128 x 64] . WgT[C x 64]^T     (TMA weight ring)
// TMEM map (512 columns): accumulator(s) from column 0 (two of them when 2C <= 256),
// [256,320) D1, [320,512) Abig_0..2.
//
// Warp roles (1024 threads, persistent, 1 CTA/SM):
//   w0 input-box TMA producer   w1 MMA1 issuer   w2 TMEM alloc + MMA2 issuer   w3 weight-ring TMA producer
//   w4-11  convert warps : build Abig (once, w4-7), D1 -> bf16 XA chunks; warps w and w+4 share TMEM
//                          lanes and take 32 of D1's 64 columns each
//   w12-15 gate warps    : multiply each landed box by gT*gV in place, TMA-store it as Xg
// (role split measured: convert 8 / gate 4 / epilogue 16 = 1.94 ms per step over the 5 launches;
//  8/8/8 = 1.98; 4/8/16 = 2.00)
//   w16-31 epilogue warps: 16 columns of every 64-column box per warp.  A thread first DRAINS its share of
//                          the whole accumulator into registers (+bias, ReLU, packed bf16: 8 registers per
//                          box) and hands the accumulator back, then stages and TMA-stores box by box.  At
//                          C = 256 there is room for ONE accumulator in tensor memory, so the drain time is
//                          what the next tile's MMA2 waits for: ~400 cycles this way, ~6000 when each box was
//                          staged before the next was read.
// WHAT BOUNDS IT: the shared-memory data pipe (128 B per clock and SM), not the tensor pipe and not HBM.  ncu on the
// C = 256 launch of the round-2 mid state: l1tex__data_pipe_lsu_wavefronts_mem_shared 68 % of peak + l1tex__data_pipe_tc_
// wavefronts_mem_shared (UMMA operand reads) 36 % = 104 %.  Per 119-row tile at Cin = C = 256, in 128-byte wavefronts:
//   UMMA reads   MMA2 A (XA chunk) 12 x 128 + B (weights) 12 x 256 = 4.6 k,  MMA1 B (the X box, once per partition) 1.5 k
//   TMA writes   weights 3.1 k, X boxes + gate slices 0.75 k;   TMA store reads  Xg 0.5 k, Y 0.5 k
//   LDS / STS    convert 1.5 k, gate 1.6 k (5.1 k before the remap below), epilogue 0.5 k
// = ~14.5 k against a measured tile period of ~14 k cycles (16.5 k before).  With M = 128 and both operands in shared
// memory an N = 256 MMA alone reads 96 B per clock; halving the weight traffic needs cta_group::2, which the TS-form
// MMA1 (a different B operand per CTA) does not fit.  The smaller widths run at the same wall: C = 64: ~2.9 k wavefronts
// per tile, period 3.3 k cycles.
// The input tile moves through a RING of 64-channel boxes (X box + its gT / gV slices, all three
// brought by TMA); a box is released after its three MMA1s (chunk order: box-major, partition
// minor), so the next tile's boxes load and get gated while this tile is still in the MMAs.
#pragma once
#include "umma.cuh"

namespace gs {
namespace gcn {

using namespace tc;

constexpr int kThreadsGcn = 1024;
constexpr int kRoleThreads = 256;   // convert: 8 warps
constexpr int kGateThreads = 128;   // gate: 4 warps
constexpr int kEpiThreads = 512;    // epilogue: 16 warps
constexpr int kFramesPerTile = 7;
constexpr int kRowsPerTile = kFramesPerTile * 17;   // 119
constexpr int kColD1 = 256;
constexpr int kColAbig = 320;
constexpr int kMaxXSlots = 8, kMaxWStages = 4;
constexpr uint32_t kGtOff = 16384;               // slot: [X box 16 KB | gT 8x64 fp32 | gV 17x64 fp32]
constexpr uint32_t kGvOff = 16384 + 2048;
constexpr uint32_t kSlotBytes = 23552;           // 23 KB, keeps every X box 1024 B aligned
constexpr uint32_t kSlotTx = 16384 + 2048 + 17 * 256;

struct Params {
    int Cin, C, T, B;
    int mtiles, ntiles, rows_per_clip;
    int xslots, wstages, eslots, nacc;
    int store_xg;         // 1: the gated input is stored (the next kernel's residual projection reads it); 0: identity
                          //    blocks re-gate the previous block's output themselves (tcn_fused.cuh)
    const float *gT;      // [B,T,Cin]   (maps carry the data; non-null = gating on)
    const float *gV;      // [B,17,Cin]
    const float *A;       // [3,17,17] fp32
    float biasv[256];     // [C] by value: the epilogue warps read it through the constant bank with warp-uniform indices
                          // (16 shared-memory loads per thread and tile less in a kernel bound by the shared-memory pipe)
    __nv_bfloat16 *dbg_xa;       // optional [ntiles*128, 3*Cin] dump of the converted XA chunks
    unsigned long long *trace;   // optional clock64 trace of CTA 0 (tools/trace_gcn.py)
};

// trace layout: [role 0..4][tile 0..kTraceTiles)[event 0..kTraceEv)
constexpr int kTraceTiles = 6, kTraceEv = 64;
#define GCN_TRACE(role, tcount, ev)                                                                     \
    do {                                                                                                \
        if (prm.trace && blockIdx.x == 0 && (tcount) < kTraceTiles && (ev) < kTraceEv)                  \
            prm.trace[((role)*kTraceTiles + (tcount)) * kTraceEv + (ev)] = (unsigned long long)clock64(); \
    } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// MN-major (channels contiguous) SW128 operand: rows of the K dimension are 128 B apart, 8-row groups
// 1024 B apart (SBO); LBO = distance between 64-element MN blocks (unused here: N = 64 = one block).
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(16384 >> 4) << 16;    // LBO
    d |= (uint64_t)(1024 >> 4) << 32;     // SBO
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;               // SWIZZLE_128B
    return d;
}
// A: bf16 from TMEM (K-major), B: bf16 MN-major from smem, D fp32, M=128, N=64
__device__ __forceinline__ uint32_t make_idesc_agg() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

struct Smem {
    uint32_t x_off, w_off, xa_off, epi_off, bar_off, total, w_stage_bytes;
};
__host__ __device__ inline Smem smem_layout(int C, int xslots, int wstages, int eslots) {
    Smem s;
    s.x_off = 0;
    s.w_off = (uint32_t)xslots * kSlotBytes;
    s.w_stage_bytes = (uint32_t)C * 128u;
    s.xa_off = s.w_off + s.w_stage_bytes * wstages;
    s.epi_off = s.xa_off + 2u * 16384u;
    s.bar_off = s.epi_off + (uint32_t)eslots * 16384u;
    s.total = s.bar_off + 512 + 1024 /*bias[C <= 256]*/ + 1024;
    return s;
}

// packed fp32 multiply (sm_100 FMUL2): two IEEE-rounded products per instruction
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// NBOX = C / 64 as a template parameter: the epilogue's per-box loops are straight-line code and the packed results stay in
// registers (with C as a run-time value the compiler indexed them in local memory: stores and loads through the same pipe
// the kernel is bound by).
template <int NBOX>
__global__ void __launch_bounds__(kThreadsGcn, 1)
gcn_fused_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapXg,
                 const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapY,
                 const __grid_constant__ CUtensorMap mapGT, const __grid_constant__ CUtensorMap mapGV,
                 const __grid_constant__ Params prm) {
    extern __shared__ unsigned char smem_raw[];
    // 1024 B alignment by OFFSET from the __shared__ array (not by integer round-trip of the pointer), so
    // the compiler keeps the shared address space: LDS/STS instead of generic LD/ST, no false aliasing
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int XS = prm.xslots, WS = prm.wstages, ES = prm.eslots, NACC = prm.nacc;
    const Smem lay = smem_layout(prm.C, XS, WS, ES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar_off);
    uint64_t *x_full = bars;                          // [kMaxXSlots]
    uint64_t *x_ready = x_full + kMaxXSlots;          // [kMaxXSlots]
    uint64_t *x_empty = x_ready + kMaxXSlots;         // [kMaxXSlots]
    uint64_t *w_full = x_empty + kMaxXSlots;          // [kMaxWStages]
    uint64_t *w_empty = w_full + kMaxWStages;         // [kMaxWStages]
    uint64_t *d1_full = w_empty + kMaxWStages, *d1_empty = d1_full + 1;
    uint64_t *xa_full = d1_full + 2;                  // [2]
    uint64_t *xa_empty = d1_full + 4;                 // [2]
    uint64_t *acc_full = d1_full + 6;                 // [2]
    uint64_t *acc_empty = d1_full + 8;                // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d1_full + 10);

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int Cin = prm.Cin;
    constexpr int C = NBOX * 64;
    constexpr bool kSplit = (NBOX == 4);   // one accumulator, released in two halves (plan_smem: nacc = 1 exactly when C = 256)
    const int nbc = Cin / 64;       // 64-channel boxes of the input tile
    const int nq = 3 * nbc;         // XA chunks (= K chunks of the channel mix) per tile

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapXg);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapGT);
        tma_prefetch_desc(&mapGV);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMaxXSlots; ++s) {
            mbar_init(&x_full[s], 1);
            mbar_init(&x_ready[s], 1);
            mbar_init(&x_empty[s], 2);      // MMA1 commit + Xg store drained
        }
        for (int s = 0; s < kMaxWStages; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, kRoleThreads);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&xa_full[s], kRoleThreads);
            mbar_init(&xa_empty[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kEpiThreads);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- one-time: block-diagonal adjacency Abig_p = I_7 (x) A_p into tensor memory --------
    if (warp >= 4 && warp < 8) {
        const int r = (warp - 4) * 32 + lane;          // output row (w of frame f)
        const int f = r / 17, w = r % 17;
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t regs[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int k0 = half * 64 + 2 * j;   // columns (input rows) k0, k0+1
                    float a0 = 0.f, a1 = 0.f;
                    if (r < kRowsPerTile) {
                        if (k0 / 17 == f && k0 < kRowsPerTile) a0 = prm.A[(p * 17 + w) * 17 + (k0 % 17)];
                        if ((k0 + 1) / 17 == f && k0 + 1 < kRowsPerTile) a1 = prm.A[(p * 17 + w) * 17 + ((k0 + 1) % 17)];
                    }
                    regs[j] = pack_bf16(a0, a1);
                }
                tmem_st32(tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + (uint32_t)(kColAbig + p * 64 + half * 32),
                          regs);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===== input-box producer: X box + its gT / gV slices per ring slot =====
        // (whole warp in the loop, one elected lane issues: see tc::elect_one)
        int slot = 0, tcount = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tcount) {
            const int tl = tile;
            const int b = tl / prm.mtiles;
            const int mt = tl % prm.mtiles;
            const int row0 = mt * kRowsPerTile;
            for (int cb = 0; cb < nbc; ++cb) {
                unsigned char *sl = smem + lay.x_off + (size_t)slot * kSlotBytes;
                mbar_wait(&x_empty[slot], ph ^ 1);
                if (elect_one()) {
                    GCN_TRACE(0, tcount, cb);
                    mbar_expect_tx(&x_full[slot], kSlotTx);
                    tma_load_3d(sl, &mapX, &x_full[slot], cb * 64, row0, b);
                    tma_load_2d(sl + kGtOff, &mapGT, &x_full[slot], cb * 64, b * prm.T + mt * kFramesPerTile);
                    tma_load_2d(sl + kGvOff, &mapGV, &x_full[slot], cb * 64, b * 17);
                }
                __syncwarp();
                if (++slot == XS) { slot = 0; ph ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ===== weight-ring producer: chunk q = (box cb, partition p) needs WgT[:, p*Cin + cb*64 ..) =====
        int stage = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x) {
            for (int q = 0; q < nq; ++q) {
                mbar_wait(&w_empty[stage], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&w_full[stage], lay.w_stage_bytes);
                    tma_load_2d(smem + lay.w_off + (size_t)stage * lay.w_stage_bytes, &mapW, &w_full[stage],
                                (q % 3) * Cin + (q / 3) * 64, 0);
                }
                __syncwarp();
                if (++stage == WS) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA1 issuer (whole warp in the loop, one elected lane issues): D1 = Abig_p . Xg box =====
        // Two issuer warps: one thread issuing the 8 MMA1s + 4 MMA2s + 4 commits of a chunk needed ~1150 cycles per
        // chunk whatever the MMA shapes (~75 cycles per instruction: tools/trace_gcn.py, the C = 64 launches ran at
        // exactly that cadence), more than the tensor work of a chunk at every width.  The two streams share no
        // barrier: MMA1 waits for x_ready / d1_empty and commits d1_full / x_empty, MMA2 waits for acc_empty /
        // xa_full / w_full and commits xa_empty / w_empty / acc_full; the tensor pipe may interleave them freely
        // (D1 and the accumulator are disjoint TMEM columns, the hand-over goes through the convert warps).
        const uint32_t idesc1 = make_idesc_agg();
        const uint32_t t_d1 = tmem_base + kColD1, t_ab = tmem_base + kColAbig;
        uint32_t d1_cnt = 0;        // running count of D1 uses
        int xslot = 0, tcount = 0;
        uint32_t xph = 0;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tcount) {
            for (int q = 0; q < nq; ++q) {
                const int p = q % 3;
                if (p == 0) {
                    mbar_wait(&x_ready[xslot], xph);      // box gated and visible to the async proxy
                    tc_fence_after();
                }
                mbar_wait(d1_empty, (d1_cnt & 1) ^ 1);
                tc_fence_after();
                const uint32_t xb = smem_u32(smem + lay.x_off + (size_t)xslot * kSlotBytes);
                const uint64_t dx = make_mnmajor_desc(xb);
                const uint32_t ta = t_ab + (uint32_t)(p * 64);
                if (elect_one()) {
                    GCN_TRACE(1, tcount, q);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // 16 input rows per step: +8 TMEM columns of Abig, +2048 B (two 8-row groups) of X
                        umma_bf16_ts(t_d1, ta + (uint32_t)(k * 8), dx + (uint64_t)(k * 128), idesc1, (uint32_t)(k > 0));
                    }
                    umma_commit(d1_full);
                    if (p == 2) umma_commit(&x_empty[xslot]);   // all three MMA1s of this box issued
                }
                __syncwarp();
                ++d1_cnt;
                if (p == 2) {
                    if (++xslot == XS) { xslot = 0; xph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ===== MMA2 issuer: acc (+)= XA chunk . WgT chunk^T =====
        const uint32_t idesc2 = make_idesc_bf16((uint32_t)C);
        const uint32_t idesc_half = make_idesc_bf16(128u);
        uint32_t xa_cnt = 0;        // running count of XA chunks consumed
        uint32_t acc_cnt = 0;       // running count of tiles (accumulator uses)
        int wstage = 0, tcount = 0;
        uint32_t wph = 0;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t as = acc_cnt % (uint32_t)NACC, aph = (acc_cnt / (uint32_t)NACC) & 1;
            const uint32_t td = tmem_base + as * (uint32_t)C;
            for (int j = 0; j < nq; ++j) {
                const uint32_t slot = xa_cnt & 1, ph = (xa_cnt >> 1) & 1;
                if (j == 0) mbar_wait(&acc_empty[as], aph ^ 1);   // epilogue has drained this accumulator (its low half: kSplit)
                mbar_wait(&xa_full[slot], ph);
                mbar_wait(&w_full[wstage], wph);
                tc_fence_after();
                const uint64_t da = make_kmajor_desc(smem_u32(smem + lay.xa_off + slot * 16384u), 128);
                const uint64_t db = make_kmajor_desc(smem_u32(smem + lay.w_off + (size_t)wstage * lay.w_stage_bytes), 128);
                if (kSplit && j == 0) {
                    // one accumulator (C = 256): the tile's first chunk is issued as two N = 128 halves, the low one as soon
                    // as the epilogue has read columns [0,128) of the previous tile: the tensor pipe restarts half a drain
                    // (~700 cycles of ~1500: the TMEM read-out of 128 KB) earlier
                    if (elect_one()) {
                        GCN_TRACE(1, tcount, 16 + j);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_half, (uint32_t)(k > 0));
                    }
                    __syncwarp();
                    mbar_wait(&acc_empty[1], aph ^ 1);            // high half drained
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(td + 128u, da + (uint64_t)(2 * k), db + (uint64_t)(1024 + 2 * k), idesc_half, (uint32_t)(k > 0));
                        umma_commit(&xa_empty[slot]);
                        umma_commit(&w_empty[wstage]);
                    }
                } else if (elect_one()) {
                    GCN_TRACE(1, tcount, 16 + j);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (uint32_t)((j > 0) | (k > 0)));
                    umma_commit(&xa_empty[slot]);
                    umma_commit(&w_empty[wstage]);
                    if (j == nq - 1) umma_commit(&acc_full[as]);
                }
                __syncwarp();
                ++xa_cnt;
                if (++wstage == WS) { wstage = 0; wph ^= 1; }
            }
            ++acc_cnt;
        }
    } else if (warp >= 4 && warp < 12) {
        // ===== convert warps: D1 (fp32, TMEM) -> bf16 XA chunk in swizzled smem; 32 columns per warp =====
        const int ew = (warp - 4) & 3, half = (warp - 4) >> 2;
        const int r = ew * 32 + lane;
        const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
        uint32_t d1_cnt = 0, xa_cnt = 0;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tcount) {
            for (int q = 0; q < nq; ++q) {
                uint32_t v[32];
                mbar_wait(d1_full, d1_cnt & 1);
                if (threadIdx.x == 128) GCN_TRACE(2, tcount, q);
                tc_fence_after();
                tmem_ld32(tmem_base + lane_base + kColD1 + (uint32_t)(half * 32), v);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(d1_empty);              // D1 may be overwritten by the next MMA1
                ++d1_cnt;
                const uint32_t slot = xa_cnt & 1, ph = (xa_cnt >> 1) & 1;
                mbar_wait(&xa_empty[slot], ph ^ 1);  // MMA2 that last read this slot has retired
                if (threadIdx.x == 128) GCN_TRACE(2, tcount, 16 + q);
                unsigned char *box = smem + lay.xa_off + slot * 16384u;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int cc = half * 4 + c4;
                    uint4 pk;
                    pk.x = pack_bf16(__uint_as_float(v[c4 * 8 + 0]), __uint_as_float(v[c4 * 8 + 1]));
                    pk.y = pack_bf16(__uint_as_float(v[c4 * 8 + 2]), __uint_as_float(v[c4 * 8 + 3]));
                    pk.z = pack_bf16(__uint_as_float(v[c4 * 8 + 4]), __uint_as_float(v[c4 * 8 + 5]));
                    pk.w = pack_bf16(__uint_as_float(v[c4 * 8 + 6]), __uint_as_float(v[c4 * 8 + 7]));
                    *reinterpret_cast<uint4 *>(box + (size_t)r * 128 + ((cc ^ (r & 7)) << 4)) = pk;
                    if (prm.dbg_xa)
                        *reinterpret_cast<uint4 *>(prm.dbg_xa + ((size_t)tile * 128 + r) * (size_t)(3 * Cin) +
                                                   (q % 3) * Cin + (q / 3) * 64 + cc * 8) = pk;
                }
                fence_proxy_async_smem();
                mbar_arrive(&xa_full[slot]);
                if (threadIdx.x == 128) GCN_TRACE(2, tcount, 32 + q);
                ++xa_cnt;
            }
        }
    } else if (warp >= 12 && warp < 16) {
        // ===== gate warps: Xg = X * gT * gV in place (gates read from the slot), TMA-store Xg =====
        const int gt_id = threadIdx.x - 384;          // 0..127
        const bool leader = (gt_id == 0);
        int slot = 0, prev_slot = -1, tcount = 0;
        uint32_t ph = 0;
        // (clip, tile in clip) advance by gridDim.x tiles without a division per tile
        const int step_b = (int)gridDim.x / prm.mtiles, step_m = (int)gridDim.x % prm.mtiles;
        int b = (int)blockIdx.x / prm.mtiles, mt = (int)blockIdx.x % prm.mtiles;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tcount) {
            const int row0 = mt * kRowsPerTile;
            for (int cb = 0; cb < nbc; ++cb) {
                unsigned char *sl = smem + lay.x_off + (size_t)slot * kSlotBytes;
                mbar_wait(&x_full[slot], ph);
                if (leader) GCN_TRACE(3, tcount, cb);
                if (prm.gT) {
                    // The kernel is bound by the SHARED-MEMORY data pipe (header comment), so the mapping minimises
                    // wavefronts: thread = (4-channel group g, joints j, j + 8) for all 7 frames, plus one row of joint
                    // 16.  A half-warp reads one whole 128-byte row (1 wavefront), a quarter-warp reads 128 contiguous
                    // bytes of a gate row (no bank conflict), gV stays in registers over the frames and gT over the two
                    // joints: ~400 wavefronts per box.  The earlier mapping (16-byte x vectors, four 16-byte gate
                    // loads per vector at a 32-byte lane stride = 2-way conflicts) needed ~1300, and ncu counted 44 M
                    // shared-load bank conflicts per C = 256 launch.
                    const float4 *sgt = reinterpret_cast<const float4 *>(sl + kGtOff);   // [8 frames][16 groups]
                    const float4 *sgv = reinterpret_cast<const float4 *>(sl + kGvOff);   // [17 joints][16 groups]
                    const int g = gt_id & 15, j = gt_id >> 4;
                    // whole frames only: rows of frames >= fv are TMA zero fill and must stay zero (0 x NaN of a stale
                    // gate would reach the other frames' rows through the zeros of Abig)
                    int fv = (prm.rows_per_clip - row0) / 17;
                    fv = fv < kFramesPerTile ? fv : kFramesPerTile;
                    unsigned char *colp = sl + (g & 1) * 8;
                    auto gate_row = [&](unsigned char *p, uint2 x, const float4 &t, const float4 &v) {
                        const __nv_bfloat162 *xp = reinterpret_cast<const __nv_bfloat162 *>(&x);
                        // (x * gT) * gV, same rounding order as the scalar form, two products per FMUL2
                        const float2 a = fmul2(fmul2(__bfloat1622float2(xp[0]), make_float2(t.x, t.y)), make_float2(v.x, v.y));
                        const float2 c2 = fmul2(fmul2(__bfloat1622float2(xp[1]), make_float2(t.z, t.w)), make_float2(v.z, v.w));
                        *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(c2.x, c2.y));
                    };
                    // A rolled loop over the frames (one frame = the rows of joints j and j + 8), software-pipelined by
                    // hand: the next frame's three loads are issued before this frame's arithmetic and stores (the
                    // compiler cannot move a load above a store that may alias).  A box still takes ~2000 cycles in
                    // these warps, ~50 cycles per LDS / STS whatever the form of the loop (row by row, all loads first,
                    // rolled: tools/trace_gcn.py): the instructions queue for a data pipe that is full.
                    const float4 gv0 = sgv[j * 16 + g], gv1 = sgv[(j + 8) * 16 + g];
                    if (j < fv) {       // joint 16: thread (g, j < 7) takes the row of frame j
                        const int r = j * 17 + 16;
                        unsigned char *p = colp + (size_t)r * 128 + (((g >> 1) ^ (r & 7)) << 4);
                        gate_row(p, *reinterpret_cast<const uint2 *>(p), sgt[j * 16 + g], sgv[16 * 16 + g]);
                    }
                    const uint32_t gsw = (uint32_t)(g >> 1);
                    unsigned char *p = colp + (size_t)j * 128 + ((gsw ^ (uint32_t)(j & 7)) << 4);   // row j of frame 0
                    float4 t = sgt[g];
                    uint2 x0 = *reinterpret_cast<const uint2 *>(p), x1 = *reinterpret_cast<const uint2 *>(p + 1024);
#pragma unroll 1
                    for (int f = 0; f < fv; ++f) {
                        // prefetch frame f + 1 (the last pass re-reads frame 6's gT row and an in-box row: unused)
                        const int fn = f + 1 < kFramesPerTile ? f + 1 : f;
                        unsigned char *rown = colp + (size_t)(fn * 17 + j) * 128;
                        unsigned char *pn = rown + ((gsw ^ ((uint32_t)(fn * 17 + j) & 7)) << 4);
                        const float4 tn = sgt[fn * 16 + g];
                        const uint2 x0n = *reinterpret_cast<const uint2 *>(pn), x1n = *reinterpret_cast<const uint2 *>(pn + 1024);
                        gate_row(p, x0, t, gv0);
                        gate_row(p + 1024, x1, t, gv1);
                        p = pn; t = tn; x0 = x0n; x1 = x1n;
                    }
                    if (leader) GCN_TRACE(3, tcount, 32 + cb);
                    fence_proxy_async_smem();
                    if (leader) GCN_TRACE(3, tcount, 36 + cb);
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (leader) {
                    mbar_arrive(&x_ready[slot]);
                    GCN_TRACE(3, tcount, 16 + cb);
                    if (prm.store_xg) {
                        tma_store_3d(&mapXg, sl, cb * 64, row0, b);
                        tma_store_commit();
                        GCN_TRACE(3, tcount, 48 + cb);
                        // the PREVIOUS box's store has been in flight for a whole box period: drain it now
                        // and release its slot (deferred wait keeps the gate pipeline moving)
                        if (prev_slot >= 0) {
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            mbar_arrive(&x_empty[prev_slot]);
                        }
                        GCN_TRACE(3, tcount, 52 + cb);
                        prev_slot = slot;
                    } else {
                        mbar_arrive(&x_empty[slot]);     // nothing reads the box but the MMA1s
                    }
                }
                if (++slot == XS) { slot = 0; ph ^= 1; }
            }
            b += step_b;
            mt += step_m;
            if (mt >= prm.mtiles) { mt -= prm.mtiles; ++b; }
        }
        if (leader) {
            tma_store_wait_read0();
            if (prev_slot >= 0) mbar_arrive(&x_empty[prev_slot]);
            tma_store_wait_all0();
        }
    } else if (warp >= 16) {
        // ===== epilogue warps: drain the accumulator to packed registers, release it, then stage + store =====
        const int ew = (warp - 16) & 3, cq = (warp - 16) >> 2;      // TMEM lane quarter, 16-column quarter of a box
        const int r = ew * 32 + lane;
        const bool leader = (threadIdx.x == 512);
        const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
        uint32_t acc_cnt = 0, ecnt = 0;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < prm.ntiles; tile += gridDim.x, ++tcount) {
            const int tl = tile;
            const int b = tl / prm.mtiles;
            const int mt = tl % prm.mtiles;
            const uint32_t as = acc_cnt % (uint32_t)NACC, aph = (acc_cnt / (uint32_t)NACC) & 1;
            mbar_wait(&acc_full[as], aph);
            if (leader) GCN_TRACE(4, tcount, 0);
            tc_fence_after();
            uint32_t pk[NBOX][8];                    // NBOX boxes x 16 bf16
            // cq comes from the warp index, so the bias index is warp-uniform: constant-bank operands, no LDS
#pragma unroll
            for (int qb = 0; qb < NBOX; ++qb) {
                uint32_t v[16];
                tmem_ld16(tmem_base + lane_base + as * (uint32_t)C + (uint32_t)(qb * 64 + cq * 16), v);
                tmem_ld_wait();
                const float *bq = prm.biasv + qb * 64 + cq * 16;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    pk[qb][e] = pack_bf16(fmaxf(__uint_as_float(v[2 * e]) + bq[2 * e], 0.f),
                                          fmaxf(__uint_as_float(v[2 * e + 1]) + bq[2 * e + 1], 0.f));
                if (kSplit && qb == 1) {             // columns [0,128) read: the low half of the next tile's first chunk may start
                    tc_fence_before();
                    mbar_arrive(&acc_empty[0]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[kSplit ? 1 : as]);   // accumulator fully read: the next tile's MMA2 may start
            if (leader) GCN_TRACE(4, tcount, 2);
#pragma unroll
            for (int qb = 0; qb < NBOX; ++qb) {
                {
                    const uint32_t es = ecnt % (uint32_t)ES;
                    // staging slot reuse: the store issued ES boxes ago must have finished reading it
                    if (leader && ecnt >= (uint32_t)ES) {
                        if (ES == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    unsigned char *rowp = smem + lay.epi_off + es * 16384u + (size_t)r * 128;
                    *reinterpret_cast<uint4 *>(rowp + (((cq * 2) ^ (r & 7)) << 4)) =
                        make_uint4(pk[qb][0], pk[qb][1], pk[qb][2], pk[qb][3]);
                    *reinterpret_cast<uint4 *>(rowp + (((cq * 2 + 1) ^ (r & 7)) << 4)) =
                        make_uint4(pk[qb][4], pk[qb][5], pk[qb][6], pk[qb][7]);
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (leader) {
                        // Y is stored JOINT-MAJOR [B,V,T,C] through a (C,V,T,B) map with a (64,17,7,1) box: the staged
                        // rows (frame, joint) land transposed for free (tcn_fused.cuh reads frame windows per joint)
                        tma_store_4d(&mapY, smem + lay.epi_off + es * 16384u, qb * 64, 0, mt * kFramesPerTile, b);
                        tma_store_commit();
                    }
                    ++ecnt;
                }
            }
            if (leader) GCN_TRACE(4, tcount, 1);
            ++acc_cnt;
        }
        if (leader) tma_store_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

struct LaunchGcn {
    CUtensorMap mapX, mapXg, mapW, mapY, mapGT, mapGV;
    Params prm;
    double flops = 0, bytes = 0;
};

// Largest rings that fit 227 KB: epilogue staging, then the weight ring (latency of the L2
// stream), then as many input boxes as remain (at least 3).
inline bool plan_smem(Params &p) {
    p.nacc = (2 * p.C <= 256) ? 2 : 1;
    const int es_c[2] = {2, 1};
    const int ws_c[3] = {4, 3, 2};
    // two staging slots first (measured: the store of box n draining behind box n+1 is worth more than a
    // deeper weight ring), then the deepest weight ring, then as many input boxes as remain
    for (int es : es_c) {
        for (int ws : ws_c) {
            for (int xs = kMaxXSlots; xs >= 3; --xs) {
                if (smem_layout(p.C, xs, ws, es).total <= 227u * 1024u) {
                    p.xslots = xs;
                    p.wstages = ws;
                    p.eslots = es;
                    return true;
                }
            }
        }
    }
    return false;
}

inline int launch(Ctx *ctx, int kid, LaunchGcn &L, cudaStream_t st) {
    if (!plan_smem(L.prm)) {
        set_error("gcn_fused: shared memory plan does not fit (Cin=%d C=%d)", L.prm.Cin, L.prm.C);
        return GS_ERR_UNSUPPORTED;
    }
    const Smem lay = smem_layout(L.prm.C, L.prm.xslots, L.prm.wstages, L.prm.eslots);
    int grid = L.prm.ntiles < ctx->sm_count ? L.prm.ntiles : ctx->sm_count;
    if (grid < 1) return GS_OK;
    auto kern = L.prm.C == 64 ? gcn_fused_kernel<1> : (L.prm.C == 128 ? gcn_fused_kernel<2> : gcn_fused_kernel<4>);
    if (L.prm.C != 64 && L.prm.C != 128 && L.prm.C != 256) {
        set_error("gcn_fused: C must be 64, 128 or 256 (C=%d)", L.prm.C);
        return GS_ERR_UNSUPPORTED;
    }
    int rc = ensure_dyn_smem(ctx, (const void *)kern, lay.total);
    if (rc != GS_OK) return rc;
    {
        LaunchScope ls(ctx, kid, st, L.flops, L.bytes);
        kern<<<grid, kThreadsGcn, lay.total, st>>>(L.mapX, L.mapXg, L.mapW, L.mapY, L.mapGT, L.mapGV, L.prm);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace gcn
}  // namespace gs

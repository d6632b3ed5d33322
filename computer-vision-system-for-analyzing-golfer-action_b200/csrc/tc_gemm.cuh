// tcgen05 / TMEM / TMA "tap GEMM" for sm_100a — the dense contractions of the bf16 path.
//
//   Out[b, r, n] = act( sum_chunks  A_chunk[b, r + shift, k..k+KC) . B_chunk[n, k..k+KC)  + bias[n]
//                       (+ residual[b, r, n]) )            bf16 in, fp32 accumulate in TMEM, bf16 out
//
// A "chunk" is one TMA box of the A operand (128 rows x KC bf16, rows optionally shifted
// in time: OOB rows of the [C, T*V, B] tensor map are zero-filled by TMA, which is exactly
// the zero padding of the dilated temporal convolution) times one TMA box of weights,
// accumulated into a column range [n_off, n_off+n_size) of the CTA's TMEM accumulator.
// One chunk program therefore expresses
//   * the graph-conv channel mix        (K = 3*Cin in 64-wide chunks, N = C),
//   * the branch 1x1 reduce             (K = C, N = C),
//   * the multi-branch dilated conv     (per branch r, 3 shifted taps, K = N = C/R, n_off = r*C/R)
//     fused with the residual projection (extra chunks, K = Cin, N = C) and the final add+ReLU.
// Stages replaced: /root/reference/README.md:27-30.
//
// Warp roles (640 threads, persistent over (clip, row-tile) tiles, 1 CTA per SM):
//   warp 0  TMA producer (A boxes, streamed weights)     warp 1  tcgen05.mma issuer
//   warp 2  TMEM allocator, then residual-box TMA producer of group 1      warp 3  ... of group 0
//   warps 4-11 / 12-19  two epilogue groups of 8 warps; group g drains accumulator buffer g, so two
//           tiles are in their epilogues at once.  Warps w and w+4 of a group share TMEM lanes (tile
//           rows) and split each 64-column box in halves:  tcgen05.ld -> +bias (+residual read from
//           the TMA-loaded box in the staging slot) -> ReLU -> bf16 written back IN PLACE to the
//           swizzled slot -> TMA store; optionally the clip-pooling partial sums (PT / PVpart of
//           segment_common.cuh) are taken from the staged tile before it leaves.
// Pipelines: smem ring (full/empty mbarriers), double-buffered TMEM accumulator (tfull/tempty),
// per-group staging slots (res_full/res_empty).  Measured (profiles/): with 4 epilogue warps the
// epilogue (1700 cycles per 64-column box) bounded every N=256 launch, not HBM or the tensor pipe.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace gs {
namespace tc {

constexpr int kTileM = 128;
constexpr int kMaxChunks = 32;
constexpr int kThreads = 640;
constexpr int kGroupThreads = 256;   // one epilogue group: 8 warps

struct Chunk {
    int a_map;    // 0 / 1: which A tensor map
    int b_map;    // 0 / 1: which B tensor map
    int a_k;      // channel coordinate in the A map
    int a_shift;  // row shift (frames * V) applied to the tile's first row
    int b_k;      // k coordinate in the B map
    int b_row;    // row (n) coordinate in the B map
    int n_off;    // accumulator column offset
    int n_size;   // N of this MMA (multiple of 16, 16..256)
    int accum;    // 0: first chunk touching these columns (overwrite), 1: accumulate
};

struct Program {
    int nchunks;
    int kc;             // K elements per chunk: 16, 32 or 64  (swizzle = 2*kc bytes)
    int N;              // accumulator columns used (multiple of 64, <= 256)
    int mtiles;         // row tiles per clip
    int ntiles;         // total tiles = B * mtiles
    int rev;            // 1: walk the tiles from the last clip down (the producer's most recent output is in L2)
    int rows_per_clip;  // T * V
    int relu;
    int has_residual;   // add residual[b, r, n] (bf16, TMA-loaded through mapRes) before the activation
    int tile_rows;      // output rows per tile: 128, or 119 (7 whole frames) when stats are taken
    int stats;          // 1: write PT[b,t,n] (sum over joints) and PVpart[b,tile,v,n] (sum over the tile's frames)
    int T;              // frames per clip (stats)
    int out_joint_major; // 1: mapOut is the 4-D (C, V, T, B) view of a joint-major [B,V,T,C] tensor and a tile is
                        //    7 whole frames (tile_rows = 119): the store transposes (frame, joint) rows for free
    unsigned long long *trace;   // optional clock64 trace of CTA 0: [5 roles][kTraceTiles][kTraceEv] (tools/trace_tc.py)
    int a_bytes;        // bytes of one A box
    int b_bytes[2];     // bytes of one B box per B map
    // shared-memory plan (host: plan_smem)
    int resident;       // 1: every chunk's weight box is loaded ONCE per CTA and stays in smem;
                        // 0: weight boxes stream through the ring next to the A boxes
    int stages;         // ring depth
    int eslots;         // epilogue staging slots (64-column boxes)
    uint32_t stage_bytes, a_span, w_off, out_off, bar_off, smem_total;
    uint32_t b_off[kMaxChunks];   // resident mode: byte offset of chunk c's weight box from w_off
    Chunk ch[kMaxChunks];
};

// ------------------------------------------------------------------ PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait for the phase with the given parity.  The retry loop lives INSIDE one asm statement (as in
// CUTLASS's ClusterBarrier::wait): a C++ loop on the per-thread try_wait predicate makes the compiler
// treat everything after it as potentially divergent, which pushes the MMA / TMA operands into vector
// registers and costs an R2UR round trip per tcgen05.mma (~200-290 cycles each, measured).
// mbarrier.try_wait suspends in hardware; a pipeline bug shows up as a hang caught by the caller's timeout.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Warp index as a value the compiler knows is warp-uniform (CUTLASS canonical_warp_idx_sync): role
// branches on it are uniform branches, so operands of the single-thread instructions stay in uniform
// registers.
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// One lane of a fully active warp.  The single-thread instructions (tcgen05.mma / commit, TMA) are
// issued under this predicate from WARP-UNIFORM control flow: measured with experiments/mma_probe.cu,
// the same tcgen05.mma costs ~175 cycles per issue from a `lane == 0` divergent branch (operands
// shuttled into uniform registers) and runs at the tensor-pipe rate (128 cycles at N=256) this way.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "elect.sync _|P1, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 x bf16 -> fp32, cta_group::1
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): rows of
// `row_bytes` (= swizzle span: 32 / 64 / 128 B), 8-row groups SBO = 8*row_bytes apart.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address  [0,14)
    d |= (uint64_t)1 << 16;                               // LBO (unused for swizzled K-major) [16,30)
    d |= (uint64_t)((8u * row_bytes) >> 4) << 32;         // SBO [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version = 1 (sm_100)
    d |= layout << 61;                                    // swizzle mode [61,64)
    return d;
}
// kind::f16 instruction descriptor: A=B=bf16 (1), D=fp32 (1), both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

constexpr int kMaxStages = 8;
constexpr int kTraceTiles = 8, kTraceEv = 32;
#define TC_TRACE(role, tcount, ev)                                                                        \
    do {                                                                                                  \
        if (prog.trace && blockIdx.x == 0 && (tcount) < kTraceTiles && (ev) < kTraceEv)                    \
            prog.trace[((role)*kTraceTiles + (tcount)) * kTraceEv + (ev)] = (unsigned long long)clock64(); \
    } while (0)

__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapRes,
               const __grid_constant__ Program prog, const float *__restrict__ bias,
               float *__restrict__ PT, float *__restrict__ PVpart) {
    extern __shared__ unsigned char smem_raw[];
    // 1024 B alignment by OFFSET from the __shared__ array (not by integer round-trip of the pointer), so
    // the compiler keeps the shared address space: LDS/STS instead of generic LD/ST, no false aliasing
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int STAGES = prog.stages, ES = prog.eslots;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + prog.bar_off);
    uint64_t *empty = full + kMaxStages;
    uint64_t *tfull = empty + kMaxStages;
    uint64_t *tempty = tfull + 2;
    uint64_t *wres = tempty + 2;
    uint64_t *res_full = wres + 1;          // [2 groups][2 slots]
    uint64_t *res_empty = res_full + 4;     // [2 groups][2 slots]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_empty + 4);
    float *sbias = reinterpret_cast<float *>(smem + prog.bar_off + 512);

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const int N = prog.N;
    const uint32_t tmem_cols = (2 * N <= 128) ? 128u : (2 * N <= 256 ? 256u : 512u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0);
        tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB0);
        tma_prefetch_desc(&mapB1);
        tma_prefetch_desc(&mapOut);
        tma_prefetch_desc(&mapRes);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], kGroupThreads);     // every thread of the draining epilogue group arrives
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 1);
        }
        mbar_init(wres, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    for (int k = threadIdx.x; k < N; k += kThreads) sbias[k] = bias[k];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (whole warp in the loop, one elected lane issues) =====
        if (prog.resident) {
            // weights: every chunk's box once per CTA (stays resident for all tiles)
            if (elect_one()) {
                uint32_t total = 0;
                for (int c = 0; c < prog.nchunks; ++c) total += (uint32_t)prog.b_bytes[prog.ch[c].b_map];
                mbar_expect_tx(wres, total);
                for (int c = 0; c < prog.nchunks; ++c) {
                    const Chunk &ch = prog.ch[c];
                    tma_load_2d(smem + prog.w_off + prog.b_off[c], ch.b_map ? &mapB1 : &mapB0, wres, ch.b_k, ch.b_row);
                }
            }
            __syncwarp();
        }
        int stage = 0, tcount = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < prog.ntiles; tile += gridDim.x, ++tcount) {
            const int tl = prog.rev ? prog.ntiles - 1 - tile : tile;
            const int b = tl / prog.mtiles;
            const int row0 = (tl % prog.mtiles) * prog.tile_rows;
            for (int c = 0; c < prog.nchunks; ++c) {
                const Chunk &ch = prog.ch[c];
                mbar_wait(&empty[stage], phase ^ 1);
                unsigned char *sa = smem + (size_t)stage * prog.stage_bytes;
                if (elect_one()) {
                    TC_TRACE(0, tcount, c);
                    if (prog.resident) {
                        mbar_expect_tx(&full[stage], (uint32_t)prog.a_bytes);
                    } else {
                        mbar_expect_tx(&full[stage], (uint32_t)prog.a_bytes + (uint32_t)prog.b_bytes[ch.b_map]);
                        tma_load_2d(sa + prog.a_span, ch.b_map ? &mapB1 : &mapB0, &full[stage], ch.b_k, ch.b_row);
                    }
                    tma_load_3d(sa, ch.a_map ? &mapA1 : &mapA0, &full[stage], ch.a_k, row0 + ch.a_shift, b);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp in the loop, one elected lane issues) =====
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t row_bytes = (uint32_t)prog.kc * 2;
        const int ksteps = prog.kc / 16;
        if (prog.resident) mbar_wait(wres, 0);
        int tcount = 0;
        for (int tile = blockIdx.x; tile < prog.ntiles; tile += gridDim.x, ++tcount) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(acc * N);
            for (int c = 0; c < prog.nchunks; ++c) {
                const Chunk &ch = prog.ch[c];
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * prog.stage_bytes);
                const uint32_t sb = prog.resident ? smem_u32(smem + prog.w_off + prog.b_off[c]) : sa + prog.a_span;
                const uint64_t da = make_kmajor_desc(sa, row_bytes);
                const uint64_t db = make_kmajor_desc(sb, row_bytes);
                const uint32_t idesc = make_idesc_bf16((uint32_t)ch.n_size);
                const uint32_t td = tacc + (uint32_t)ch.n_off;
                if (elect_one()) {
                    TC_TRACE(1, tcount, c);
                    for (int k = 0; k < ksteps; ++k) {
                        // +32 bytes of K per step: descriptor start address is in 16 B units
                        umma_bf16(td, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (uint32_t)(ch.accum | (k > 0)));
                    }
                    umma_commit(&empty[stage]);          // frees this smem stage once the MMAs retire
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) {
                TC_TRACE(1, tcount, 31);
                umma_commit(&tfull[acc]);   // accumulator complete -> epilogue
            }
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp == 2 || warp == 3) {
        // ===== residual producers (warp 3 -> group 0, warp 2 -> group 1): one 64-column box of the
        // residual per epilogue box, straight into the group's staging slot (the epilogue adds it and
        // overwrites it in place).  One warp per group so neither group ever waits for the other's slots.
        if (prog.has_residual) {
            const int g = 3 - warp;
            uint32_t ecnt = 0;
            const uint32_t box_bytes = (uint32_t)prog.tile_rows * 128u;
            int it = g;
            for (int tile = blockIdx.x + g * gridDim.x; tile < prog.ntiles; tile += 2 * gridDim.x, it += 2) {
                const int tl = prog.rev ? prog.ntiles - 1 - tile : tile;
                const int b = tl / prog.mtiles;
                const int row0 = (tl % prog.mtiles) * prog.tile_rows;
                for (int q = 0; q < N / 64; ++q) {
                    const uint32_t sl = ecnt % (uint32_t)ES, ph = (ecnt / (uint32_t)ES) & 1;
                    uint64_t *rf = &res_full[g * 2 + (int)sl];
                    mbar_wait(&res_empty[g * 2 + (int)sl], ph ^ 1);
                    if (elect_one()) {
                        TC_TRACE(2, it, q);
                        mbar_expect_tx(rf, box_bytes);
                        tma_load_3d(smem + prog.out_off + (size_t)(g * ES + (int)sl) * (kTileM * 128), &mapRes, rf, q * 64,
                                    row0, b);
                    }
                    __syncwarp();
                    ++ecnt;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: two groups of 8 warps; warp w owns TMEM lanes [32(w%4), +32) = tile rows and
        // columns [32h, 32h+32) of each 64-column box, h = (w-4)/4 % 2 =====
        const int g = (warp - 4) >> 3;                 // group = accumulator buffer
        const int ew = (warp - 4) & 3;
        const int half = ((warp - 4) >> 2) & 1;
        const int gt = threadIdx.x - 128 - g * kGroupThreads;    // thread inside the group, 0..255
        const int r = ew * 32 + lane;                  // row inside the tile
        const bool leader = (gt == 0);
        const int bar_id = 1 + g;
        unsigned char *sout = smem + prog.out_off + (size_t)(g * ES) * (kTileM * 128);
        uint32_t ecnt = 0;
        int pending = -1;                               // slot whose TMA store has not been drained yet (leader)
        int it = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < prog.ntiles; tile += 2 * gridDim.x, it += 2) {
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1;
            const int tl = prog.rev ? prog.ntiles - 1 - tile : tile;
            const int b = tl / prog.mtiles;
            const int mt = tl % prog.mtiles;
            const int row0 = mt * prog.tile_rows;
            mbar_wait(&tfull[g], acc_phase);
            if (leader) TC_TRACE(3 + g, it >> 1, 0);
            tc_fence_after();
            for (int q = 0; q < N / 64; ++q) {
                const uint32_t es = ecnt % (uint32_t)ES, eph = (ecnt / (uint32_t)ES) & 1;
                unsigned char *box = sout + (size_t)es * (kTileM * 128);
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(g * N + q * 64 + half * 32);
                tmem_ld32(taddr, v);
                if (prog.has_residual) {
                    mbar_wait(&res_full[g * 2 + (int)es], eph);   // slot is ours and holds the residual box
                } else if (ES == 1) {
                    // single slot, no producer: the leader drained the previous store after issuing it
                    asm volatile("bar.sync %0, 256;" ::"r"(bar_id + 2) : "memory");
                }
                if (leader) TC_TRACE(3 + g, it >> 1, 1 + 4 * q);
                tmem_ld_wait();
                if (leader) TC_TRACE(3 + g, it >> 1, 2 + 4 * q);
                if (q == N / 64 - 1) {
                    // every TMEM read of this accumulator is done: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(&tempty[g]);
                }
                unsigned char *rowp = box + (size_t)r * 128;
                const float *bq = sbias + q * 64 + half * 32;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(bq + cc * 8);
                    const float4 b1 = *reinterpret_cast<const float4 *>(bq + cc * 8 + 4);
                    float f[8];
                    f[0] = __uint_as_float(v[cc * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[cc * 8 + 1]) + b0.y;
                    f[2] = __uint_as_float(v[cc * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[cc * 8 + 3]) + b0.w;
                    f[4] = __uint_as_float(v[cc * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[cc * 8 + 5]) + b1.y;
                    f[6] = __uint_as_float(v[cc * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[cc * 8 + 7]) + b1.w;
                    // 128B-swizzled box row: 16 B chunk index XOR (row & 7) — matches the TMA maps
                    uint4 *sp = reinterpret_cast<uint4 *>(rowp + (((half * 4 + cc) ^ (r & 7)) << 4));
                    if (prog.has_residual) {
                        const uint4 rr = *sp;
                        const __nv_bfloat162 *rp = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 t = __bfloat1622float2(rp[e]);
                            f[2 * e] += t.x;
                            f[2 * e + 1] += t.y;
                        }
                    }
                    if (prog.relu) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                    }
                    uint4 packed;
                    __nv_bfloat162 *pp = reinterpret_cast<__nv_bfloat162 *>(&packed);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pp[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                    *sp = packed;
                }
                fence_proxy_async_smem();
                // two slots, no producer: the store issued a whole box ago (the only one pending) must have
                // read its slot before the NEXT box overwrites it; that box starts after this barrier
                if (leader && !prog.has_residual && ES == 2) tma_store_wait_read0();
                asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
                if (leader) {
                    TC_TRACE(3 + g, it >> 1, 3 + 4 * q);
                    if (prog.out_joint_major) tma_store_4d(&mapOut, box, q * 64, 0, mt * 7, b);
                    else tma_store_3d(&mapOut, box, q * 64, row0, b);
                    tma_store_commit();
                    if (!prog.has_residual) {
                        if (ES == 1) tma_store_wait_read0();   // published to the group by the next box's barrier
                    } else {
                        // the previous box's store has had a whole box period: drain it and hand its slot back
                        // (every thread finished its pooling reads of that box before this box's barrier)
                        if (pending >= 0) {
                            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                            mbar_arrive(&res_empty[g * 2 + pending]);
                        }
                        pending = (int)es;
                    }
                }
                if (leader) TC_TRACE(3 + g, it >> 1, 17 + 2 * q);
                if (prog.stats) {
                    // clip-pooling partial sums from the staged bf16 tile (7 frames x 17 joints x 64 channels):
                    // thread = (channel pair c2, part); frames / joints are split over the 8 parts
                    const int c2 = gt & 31, part = gt >> 5;
                    const int nfv = min(7, prog.T - mt * 7);                 // valid frames of this tile
                    const uint32_t coff = (uint32_t)(c2 & 3) * 4u;
                    const int cchunk = c2 >> 2;
                    const unsigned char *colp = box + coff;
                    auto ldv = [&](int row) {
                        return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(
                            colp + (size_t)row * 128 + ((cchunk ^ (row & 7)) << 4)));
                    };
                    // frame sums: part p (< 7) owns frame p
                    if (part < nfv) {
                        float2 t[17];
#pragma unroll
                        for (int vv = 0; vv < 17; ++vv) t[vv] = ldv(part * 17 + vv);
                        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                        for (int vv = 0; vv < 17; ++vv) {
                            acc.x += t[vv].x;
                            acc.y += t[vv].y;
                        }
                        *reinterpret_cast<float2 *>(PT + ((size_t)b * prog.T + (size_t)(mt * 7 + part)) * N + q * 64 + 2 * c2) = acc;
                    }
                    if (leader) TC_TRACE(3 + g, it >> 1, 18 + 2 * q);
                    // joint sums over the tile's frames: part p owns joints p, p+8, p+16
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int vv = part + 8 * k;
                        if (vv < 17) {
                            float2 t[7];
#pragma unroll
                            for (int f = 0; f < 7; ++f) t[f] = ldv(f * 17 + vv);
                            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                            for (int f = 0; f < 7; ++f)
                                if (f < nfv) {
                                    acc.x += t[f].x;
                                    acc.y += t[f].y;
                                }
                            *reinterpret_cast<float2 *>(PVpart + (((size_t)b * prog.mtiles + mt) * 17 + vv) * N + q * 64 + 2 * c2) = acc;
                        }
                    }
                }
                if (leader) TC_TRACE(3 + g, it >> 1, 4 + 4 * q);
                ++ecnt;
            }
        }
        if (leader) {
            tma_store_wait_read0();
            if (pending >= 0) mbar_arrive(&res_empty[g * 2 + pending]);
            tma_store_wait_all0();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ------------------------------------------------------------------ host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
    return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 [batch, rows, width] activation (row-major, width contiguous) -> 3-D map, box (box_w, box_rows, 1)
inline int make_act_map(CUtensorMap *m, const void *base, int width, int rows, int batch, int box_w,
                        int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GS_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)width * 2 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_w * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(act %dx%dx%d box %dx%d) failed: %d", width, rows, batch, box_w, box_rows,
                  (int)r);
        return GS_ERR_CUDA;
    }
    return GS_OK;
}

// bf16 [nrows, kwidth] weight matrix (K contiguous) -> 2-D map, box (box_k, box_rows)
inline int make_weight_map(CUtensorMap *m, const void *base, int kwidth, int nrows, int box_k, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kwidth, (cuuint64_t)nrows};
    cuuint64_t strides[1] = {(cuuint64_t)kwidth * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(box_k * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weight %dx%d box %dx%d) failed: %d", kwidth, nrows, box_k, box_rows,
                  (int)r);
        return GS_ERR_CUDA;
    }
    return GS_OK;
}

// fp32 [nrows, width] matrix -> 2-D map, box (box_w, box_rows), no swizzle (gate slices)
inline int make_f32_map(CUtensorMap *m, const void *base, int width, long long nrows, int box_w, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)nrows};
    cuuint64_t strides[1] = {(cuuint64_t)width * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(f32 %dx%lld box %dx%d) failed: %d", width, nrows, box_w, box_rows, (int)r);
        return GS_ERR_CUDA;
    }
    return GS_OK;
}

struct Launch {
    CUtensorMap mapA0, mapA1, mapB0, mapB1, mapOut, mapRes;   // mapRes: residual boxes (has_residual)
    Program prog;
    const float *bias = nullptr;
    float *PT = nullptr, *PVpart = nullptr;                    // prog.stats outputs
    double flops = 0, bytes = 0;   // algorithmic, for the profiler
};

// Fill the shared-memory plan of `p` (nchunks, kc, N, a_bytes, b_bytes, ch[], stats must be set).
// Weights stay resident when they fit next to >= 3 A stages; otherwise they stream with A.
// Staging: 2 epilogue groups x eslots 16 KB boxes (the pooling sums need eslots = 2).
inline bool plan_smem(Program &p, bool allow_resident = true) {
    const uint32_t limit = 227u * 1024u;
    const uint32_t a_span = ((uint32_t)p.a_bytes + 1023u) & ~1023u;
    const uint32_t bars = 512 + 1024, slack = 1024;   // barriers + bias[N <= 256]
    uint32_t wtotal = 0;
    for (int c = 0; c < p.nchunks; ++c) {
        p.b_off[c] = wtotal;
        wtotal += ((uint32_t)p.b_bytes[p.ch[c].b_map] + 1023u) & ~1023u;
    }
    uint32_t bmax = 0;
    for (int c = 0; c < p.nchunks; ++c) {
        const uint32_t b = ((uint32_t)p.b_bytes[p.ch[c].b_map] + 1023u) & ~1023u;
        bmax = b > bmax ? b : bmax;
    }
    auto fill = [&](int resident, int st, int es) {
        p.resident = resident;
        p.stages = st;
        p.eslots = es;
        p.a_span = a_span;
        p.stage_bytes = resident ? a_span : a_span + bmax;
        p.w_off = resident ? a_span * st : 0;
        p.out_off = resident ? p.w_off + wtotal : p.stage_bytes * st;
        p.bar_off = p.out_off + 2u * (uint32_t)es * 16384u;
        p.smem_total = p.bar_off + bars + slack;
    };
    const int es_lo = (p.stats || p.has_residual) ? 2 : 1;
    auto try_resident = [&](int es, int min_st) {
        for (int st = kMaxStages; st >= min_st; --st)
            if (wtotal + a_span * st + 2u * es * 16384u + bars + slack <= limit) {
                fill(1, st, es);
                return true;
            }
        return false;
    };
    auto try_stream = [&](int es, int min_st) {
        for (int st = kMaxStages; st >= min_st; --st)
            if ((a_span + bmax) * st + 2u * es * 16384u + bars + slack <= limit) {
                fill(0, st, es);
                return true;
            }
        return false;
    };
    // preference: two staging slots per group (the store of box n drains behind box n+1) with resident
    // weights and a deep A ring; then the same with streamed weights; only then a single slot
    if (allow_resident && try_resident(2, 4)) return true;
    if (try_stream(2, 3)) return true;
    if (allow_resident && try_resident(2, 3)) return true;
    if (es_lo == 1) {
        if (allow_resident && try_resident(1, 3)) return true;
        if (try_stream(1, 2)) return true;
    }
    if (try_stream(2, 2)) return true;
    return false;
}

inline int launch(Ctx *ctx, int kid, Launch &L, cudaStream_t st) {
    static const bool no_resident = getenv("GOLFER_TC_STREAM_WEIGHTS") != nullptr;
    if (!plan_smem(L.prog, !no_resident)) {
        set_error("tc_gemm: shared memory plan does not fit (kc=%d N=%d chunks=%d)", L.prog.kc, L.prog.N,
                  L.prog.nchunks);
        return GS_ERR_UNSUPPORTED;
    }
    int grid = L.prog.ntiles < ctx->sm_count ? L.prog.ntiles : ctx->sm_count;
    if (grid < 1) return GS_OK;
    GS_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.prog.smem_total));
    {
        LaunchScope ls(ctx, kid, st, L.flops, L.bytes);
        tc_gemm_kernel<<<grid, kThreads, L.prog.smem_total, st>>>(L.mapA0, L.mapA1, L.mapB0, L.mapB1, L.mapOut, L.mapRes,
                                                                  L.prog, L.bias, L.PT, L.PVpart);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace tc
}  // namespace gs

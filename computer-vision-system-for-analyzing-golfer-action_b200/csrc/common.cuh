// Shared declarations for libgolfer_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/golfer_b200.h"

namespace gs {

void set_error(const char *fmt, ...);

#define GS_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            gs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return GS_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define GS_KERNEL_CHECK()                                                                  \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            gs::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GS_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- per-block folded parameters on the device (fp32 master copies) -------------
struct BlockParams {
    int cin, c, cr, cs, cj;
    bool has_res;
    const float *A;    // [P,V,V]
    const float *Wg;   // [P*cin, c]
    const float *bg;   // [c]
    const float *W1;   // [c, c]
    const float *b1;   // [c]
    const float *W2;   // [R,3,cr,cr]
    const float *b2;   // [c]
    const float *Wr;   // [cin, c] or null
    const float *br;
    const float *seW1, *seb1, *seW2, *seb2;       // [c,cs] [cs] [cs,c] [c]
    const float *jW, *jb, *jWt, *jbt, *jWv, *jbv; // [c,cj] [cj] [cj,c] [c] [cj,c] [c]
};

struct Bf16Path;   // segment_bf16.cu
struct EmbedPath;  // align_embed.cu

// ---- per-kernel profiling (bench.py roofline): CUDA events around each launch ----
enum KernelId {
    K_AGG = 0, K_GEMM_GCN, K_GEMM_TCN1, K_GEMM_RES, K_TCONV, K_STATS, K_SE, K_STJ, K_HEAD, K_FEAT,
    K_DTW, K_DTW_GENERIC, K_PAIRCOST, K_COMPARE,
    K_B_FRONT, K_B_AGG, K_B_GEMM_GCN, K_B_GEMM_TCN1, K_B_TCONV, K_B_MISC, K_DTW_BACKTRACK, K_POSE,
    K_EMBED, K_EMBED_COST,
    K_COUNT
};
const char *kernel_name(int id);

struct ProfSlot {
    cudaEvent_t e0, e1;
    int id, blk;
};

struct Profiler {
    bool on = false;
    std::vector<ProfSlot> pool;
    size_t used = 0;
    double ms[K_COUNT] = {0};
    double flops[K_COUNT] = {0};
    double bytes[K_COUNT] = {0};
    int64_t launches[K_COUNT] = {0};
    // the same, split by network block (index GS_MAX_BLOCKS = launches outside any block)
    double ms_blk[K_COUNT][GS_MAX_BLOCKS + 1] = {{0}};
    int64_t launches_blk[K_COUNT][GS_MAX_BLOCKS + 1] = {{0}};
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    bool has_net = false;
    gs_config cfg{};
    int max_B = 0, max_T = 0;
    int64_t launches = 0;
    size_t ws_bytes = 0;
    int cur_block = GS_MAX_BLOCKS;   // network block the forward pass is in (profiler attribution)

    // weights
    float *d_blob = nullptr;      // whole folded blob (fp32)
    std::vector<float> h_blob;    // host copy (bf16 path repacks weights from it)
    size_t blob_floats = 0;
    const float *in_scale = nullptr, *in_shift = nullptr, *headW = nullptr, *headb = nullptr;
    float *headWT = nullptr;      // head weights transposed to [K][C] (head_kernel stages rows)
    std::vector<BlockParams> blocks;

    // segmentation workspace (element type depends on precision)
    void *bufX = nullptr;   // gated block input            [rows, Cmax]
    void *bufXA = nullptr;  // adjacency-aggregated input   [rows, 3*Cmax]
    void *bufY = nullptr;   // GCN output                   [rows, Cmax]
    void *bufH = nullptr;   // branch 1x1 output            [rows, Cmax]
    void *bufU[2] = {nullptr, nullptr};  // pre-attention block output, ping-pong
    void *bufR = nullptr;   // residual projection          [rows, Cmax]
    float *PT = nullptr;    // sum over joints  [B,T,Cmax]
    float *PV = nullptr;    // sum over frames  [B,V,Cmax]
    float *PVpart = nullptr;// per-chunk partials of PV
    float *seS = nullptr;   // SE gate [B,Cmax]
    float *gT = nullptr;    // s * a_t [B,T,Cmax]
    float *gV = nullptr;    // a_v     [B,V,Cmax]
    float *d_skel = nullptr;        // staging for the host entry points
    float *d_logits = nullptr;
    uint8_t *d_labels = nullptr;

    // alignment scratch (grow-only)
    void *align_ws = nullptr;
    size_t align_ws_bytes = 0;
    float *d_al_a = nullptr, *d_al_b = nullptr, *d_al_cost = nullptr;
    int32_t *d_al_path = nullptr, *d_al_plen = nullptr;
    size_t al_host_cap[5] = {0, 0, 0, 0, 0};

    cudaStream_t own_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_copy[2] = {nullptr, nullptr};
    // gs_segment_host, bf16 path: the input arrives in chunks on the copy stream; only block 0's input kernel is
    // launched per chunk (behind ev_front[k]), every later kernel runs once on the whole batch
    static constexpr int kFrontChunks = 4;
    cudaEvent_t ev_front[kFrontChunks] = {nullptr, nullptr, nullptr, nullptr};
    int front_nchunks = 0, front_b0[kFrontChunks] = {0, 0, 0, 0}, front_nb[kFrontChunks] = {0, 0, 0, 0};
    // gs_segment_host_submit / _wait: up to two batches in flight.  The input copy of batch n+1 runs under the kernels
    // of batch n (it waits for ev_pipe_front: batch n's input kernels have consumed d_skel), the result copy of batch n
    // runs on a third stream under the kernels of batch n+1 (outputs alternate between two device buffers).
    cudaStream_t pipe_d2h = nullptr;
    cudaEvent_t ev_pipe_front = nullptr, ev_pipe_head[2] = {nullptr, nullptr}, ev_pipe_done[2] = {nullptr, nullptr};
    bool pipe_front_valid = false, pipe_done_valid[2] = {false, false};
    float *d_logits2 = nullptr;
    uint8_t *d_labels2 = nullptr;
    long long pipe_next = 0;      // next ticket
    bool pipe_chain = false;      // the previous entry-point call on this context was a submit (no cross-stream wait needed)
    // gs_align_host_submit / _wait: the same for the alignment (two batches in flight, two sets of staging buffers;
    // the call is bound by the input copy, so what the pipeline hides is the last chunk's sweep and result copy)
    float *al2_a = nullptr, *al2_b = nullptr, *al2_cost = nullptr;
    int32_t *al2_path = nullptr, *al2_plen = nullptr;
    size_t al2_cap[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev_al_chunk[2] = {nullptr, nullptr}, ev_al_done[2] = {nullptr, nullptr};
    bool al_done_valid[2] = {false, false};
    long long al_pipe_next = 0;
    bool al_pipe_chain = false;
    bool ev_valid = false;
    // Every entry point shares this context's workspace, whatever stream it runs on: each call records
    // `ev_last` behind its work and the next call's stream(s) wait for it first, so sequential calls from
    // one thread on different streams (a device call on the caller's stream, then a host call on the
    // context's own streams) never overlap on the workspace.
    cudaEvent_t ev_last = nullptr;
    bool ev_last_valid = false;
    cudaStream_t last_stream = nullptr;

    Bf16Path *bf16 = nullptr;
    EmbedPath *embed = nullptr;          // align_embed.cu
    Profiler prof;
};

// Brackets ONE kernel launch: counts it, and when profiling is on records a CUDA event
// pair on the launch stream plus the launch's algorithmic flops / bytes.
struct LaunchScope {
    Ctx *c;
    int slot = -1;
    cudaStream_t st;
    LaunchScope(Ctx *ctx, int id, cudaStream_t s, double flops = 0, double bytes = 0) : c(ctx), st(s) {
        c->launches += 1;
        Profiler &p = c->prof;
        if (!p.on) return;
        p.launches[id] += 1;
        p.launches_blk[id][c->cur_block] += 1;
        p.flops[id] += flops;
        p.bytes[id] += bytes;
        if (p.used == p.pool.size()) {
            if (p.pool.size() >= 32768) return;
            ProfSlot ns{};
            if (cudaEventCreate(&ns.e0) != cudaSuccess || cudaEventCreate(&ns.e1) != cudaSuccess) return;
            p.pool.push_back(ns);
        }
        slot = (int)p.used++;
        p.pool[slot].id = id;
        p.pool[slot].blk = c->cur_block;
        cudaEventRecord(p.pool[slot].e0, st);
    }
    ~LaunchScope() {
        if (slot >= 0) cudaEventRecord(c->prof.pool[slot].e1, st);
    }
};

// api.cu: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the occupancy query are made once per
// (device, kernel) and remembered, not repeated on every call of an entry point
int ensure_dyn_smem(Ctx *ctx, const void *kernel, size_t bytes);
int cached_occupancy(Ctx *ctx, const void *kernel, int nthreads, size_t smem_bytes, int *blocks_per_sm);

// align.cu
int align_launch(Ctx *ctx, const float *a, const float *b, int N, int Ta, int Tb, int V, int Cc,
                 float *cost, int32_t *path, int32_t *plen, cudaStream_t st, const uint8_t *la = nullptr,
                 const uint8_t *lb = nullptr, float penalty = 0.f);
int backtrack_launch(Ctx *ctx, const uint32_t *dirs, int N, int ra, int rb, bool swap, int32_t *path, int32_t *plen,
                     cudaStream_t st);
int dtw_costmat_launch(Ctx *ctx, const float *cm, int N, int ra, int rb, bool swap, float *cost, int32_t *path,
                       int32_t *plen, cudaStream_t st);
int pair_cost_launch(Ctx *ctx, const float *a, const float *b, int N, int Ta, int Tb, int V, int Cc,
                     float *out, cudaStream_t st);
int compare_launch(Ctx *ctx, const float *a, const float *b, const int32_t *path, const int32_t *plen,
                   int N, int Ta, int Tb, int V, int Cc, float *out, cudaStream_t st);

// align_embed.cu: learned alignment embedding (encoder -> tensor-core cost matrix -> DP)
int align_embed_set_encoder(Ctx *ctx, const float *blob, size_t nfloats);
int align_embed_launch(Ctx *ctx, const float *a, const float *b, int N, int Ta, int Tb, int V, int Cc, float *cost,
                       int32_t *path, int32_t *plen, float *cost_matrix_out, cudaStream_t st);
void align_embed_destroy(Ctx *ctx);

// pose.cu
int normalize_pose_launch(Ctx *ctx, const float *kp, float *out, int B, int T, int V, float min_score,
                          cudaStream_t st);

// segment_fp32.cu
int segment_fp32_forward(Ctx *ctx, const float *skel, float *logits, uint8_t *labels, int B, int T,
                         int upto_block, float *feat_out, cudaStream_t st);
// segment_bf16.cu
int bf16_path_create(Ctx *ctx);
void bf16_path_destroy(Ctx *ctx);
int bf16_debug_read(Ctx *ctx, const char *name, void *host, size_t nbytes);
int segment_bf16_forward(Ctx *ctx, const float *skel, float *logits, uint8_t *labels, int B, int T,
                         int upto_block, float *feat_out, cudaStream_t st);

}  // namespace gs

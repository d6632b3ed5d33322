/* golfer_b200.h — C ABI of libgolfer_b200.so (sm_100a only).
 *
 * Drop-in boundary for the skeleton-sequence hot path named by BASELINE.json's
 * north_star:   segment(skel[B,T,V,C]) -> logits[B,T,K]
 *               align(a, b)            -> (cost, path)
 *
 * Reference interface each entry point replaces: the reference ships no code,
 * hence no plugin / operator / FFI interface exists to cite (SURVEY.md 8b:
 * "not in reference").  The stages these functions implement are named at
 *   /root/reference/README.md:17-18  action-segmentation model      -> gs_segment*
 *   /root/reference/README.md:27-34  GCN / multi-branch TCN / channel / ST-joint attention
 *   /root/reference/README.md:21-22  temporal-alignment model        -> gs_align*
 *   /root/reference/README.md:44-49  alignment section
 *   /root/reference/README.md:50-52  "Compare 2 skeleton"            -> gs_compare
 * The signatures follow the ABI proposed in SURVEY.md 8b.
 *
 * Conventions
 *   - plain pointers and sizes; no C++ / torch types cross this boundary;
 *   - every function returns 0 on success or a negative gs_status; it never throws;
 *     gs_last_error() returns a thread-local message for the last failure;
 *   - "dev" pointers are device memory on the context's device, "host" pointers are
 *     host memory (pinned memory makes the host entry points asynchronous copies);
 *   - the caller owns all input/output buffers; the library allocates its workspace
 *     in gs_create (sized for max_B x max_T) and grows the alignment scratch only
 *     when a larger (N,Ta,Tb) than ever seen arrives; steady state is allocation-free;
 *   - device entry points are asynchronous on `cuda_stream` (a cudaStream_t passed as
 *     void*) and never synchronise the host;
 *   - one context per device; a context is not re-entrant; contexts are independent;
 *   - there is NO CPU fallback: without a usable sm_100 device gs_create fails with
 *     GS_ERR_NO_DEVICE.
 */
#ifndef GOLFER_B200_H
#define GOLFER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 1
#define GS_MAX_BLOCKS 8
#define GS_MAX_BRANCHES 8

typedef enum gs_status {
    GS_OK = 0,
    GS_ERR_INVALID = -1,    /* bad argument / shape / config */
    GS_ERR_NO_DEVICE = -2,  /* no CUDA device, or not compute capability 10.x */
    GS_ERR_CUDA = -3,       /* a CUDA call failed (message has the CUDA error string) */
    GS_ERR_NOMEM = -4,
    GS_ERR_UNSUPPORTED = -5 /* valid request this build has no kernel for */
} gs_status;

typedef enum gs_precision {
    GS_PREC_FP32 = 0, /* fp32 storage + fp32 CUDA-core math: the 1e-5 parity path */
    GS_PREC_BF16 = 1  /* bf16 storage + tcgen05 (fp32 accumulate): the throughput path */
} gs_precision;

/* Mirrors golfer_b200.config.GolfSegConfig (frozen "v0" assumption set). */
typedef struct gs_config {
    int32_t num_joints;     /* V = 17 */
    int32_t in_channels;    /* 3: x, y, confidence */
    int32_t num_partitions; /* P = 3 */
    int32_t num_blocks;
    int32_t widths[GS_MAX_BLOCKS];
    int32_t num_branches;   /* R */
    int32_t kernel_size;    /* 3 */
    int32_t dilations[GS_MAX_BRANCHES];
    int32_t se_reduction;
    int32_t stj_reduction;
    int32_t num_classes;    /* K */
    int32_t precision;      /* gs_precision */
} gs_config;

typedef struct gs_ctx gs_ctx;

/* ABI version of the loaded library (compare with GS_ABI_VERSION). */
int gs_abi_version(void);
const char *gs_last_error(void);

/* Create a context on `device`.  `cfg` + `weights_blob` (host memory; layout in
 * params.py:pack_blob — header {magic,n_floats,n_blocks,0} then BN-folded fp32
 * parameters) enable gs_segment*; pass cfg = NULL, weights_blob = NULL for an
 * alignment-only context.  Workspace is sized for max_B clips of max_T frames. */
int gs_create(gs_ctx **out, int device, const gs_config *cfg, const void *weights_blob,
              size_t weights_nbytes, int max_B, int max_T);
int gs_destroy(gs_ctx *ctx);

/* skel_dev [B,T,V,Cin] fp32 -> logits_dev [B,T,K] fp32 and, if non-NULL,
 * labels_dev [B,T] u8 = first arg-max over K.  B <= max_B, T <= max_T. */
int gs_segment(gs_ctx *ctx, const float *skel_dev, float *logits_dev, uint8_t *labels_dev,
               int B, int T, void *cuda_stream);

/* Same through HOST buffers: host->device copy of skel, the kernels, device->host
 * copy of logits (and labels), chunked over clips so copies overlap compute;
 * returns after the results are in host memory.  Either output may be NULL. */
int gs_segment_host(gs_ctx *ctx, const float *skel_host, float *logits_host,
                    uint8_t *labels_host, int B, int T);

/* Pipelined form of gs_segment_host for a stream of batches (GS_PREC_BF16 contexts): _submit enqueues the copies and
 * kernels of one batch and returns at once with a ticket; _wait returns when that batch's results are in its host
 * buffers.  Up to two batches are in flight: the input copy of batch n+1 overlaps the kernels of batch n, the result
 * copy of batch n overlaps the kernels of batch n+1 (a third submit first waits for the oldest batch).  The host
 * buffers of a batch (pinned memory, or the copies do not overlap) must stay untouched until its _wait returns.
 * Any other entry point may follow a submit; it runs after every batch submitted so far. */
int gs_segment_host_submit(gs_ctx *ctx, const float *skel_host, float *logits_host,
                           uint8_t *labels_host, int B, int T, int *ticket);
int gs_segment_host_wait(gs_ctx *ctx, int ticket);

/* Debug / parity hook: run the network through block `block` (0-based) and write that
 * block's output AFTER both attention gates as fp32 [B,T,V,C_block] to out_dev. */
int gs_segment_features(gs_ctx *ctx, const float *skel_dev, int block, float *out_dev,
                        int B, int T, void *cuda_stream);

/* a_dev [N,Ta,V,Cc], b_dev [N,Tb,V,Cc] fp32 (Cc >= 2; channels 0,1 = x,y) ->
 * cost_dev [N] fp32 = D[Ta-1][Tb-1]; path_dev [N, Ta+Tb-1, 2] int32 = (i,j) cells from
 * (0,0) to (Ta-1,Tb-1), rows beyond path_len filled with -1; path_len_dev [N] int32.
 * Arithmetic contract (bit-exact vs oracle/align.py): per-joint sqrt(dx*dx+dy*dy) with
 * individually rounded IEEE fp32 ops (no FMA), joints summed in index order, divided
 * by V; tie-break diagonal > up > left.  path_dev / path_len_dev may both be NULL
 * (cost only). */
int gs_align(gs_ctx *ctx, const float *a_dev, const float *b_dev, int N, int Ta, int Tb,
             int V, int Cc, float *cost_dev, int32_t *path_dev, int32_t *path_len_dev,
             void *cuda_stream);

/* Phase-conditioned alignment: gs_align with the per-frame phase labels that
 * gs_segment produced for both clips (labels_a_dev [N,Ta], labels_b_dev [N,Tb] u8, device memory,
 * so segment -> align needs no host trip).  A cell whose two frames carry different labels costs
 * `penalty` more:  c'[i,j] = c[i,j] + (labels_a[i] != labels_b[j] ? penalty : 0), one more
 * individually rounded fp32 add after the division by V; the DTW recurrence, tie-break and outputs
 * are those of gs_align.  penalty = +inf forbids phase-crossing cells (a path through them costs
 * inf).  Bit-exact vs oracle/align.py:align_phase_ref. */
int gs_align_phase(gs_ctx *ctx, const float *a_dev, const float *b_dev, const uint8_t *labels_a_dev,
                   const uint8_t *labels_b_dev, float penalty, int N, int Ta, int Tb, int V, int Cc,
                   float *cost_dev, int32_t *path_dev, int32_t *path_len_dev, void *cuda_stream);

/* Learned alignment embedding (/root/reference/README.md:44-47: the alignment model is
 * trained).  Contract: oracle/embed.py, AlignEmbedConfig v0 - a per-frame MLP 34 -> 128 (ReLU) -> 128 over the
 * (x, y) of V = 17 joints, frame cost c[i,j] = sqrt(max(|fa_i|^2 + |fb_j|^2 - 2 fa_i.fb_j, 0)) (the embeddings
 * rounded to bf16, the Gram term on tcgen05 tensor cores with fp32 accumulation), then the DP, tie-break and
 * backtrack of gs_align.
 * gs_set_align_encoder: host fp32 blob {W1 [34,128], b1 [128], W2 [128,128], b2 [128]} (oracle/embed.py:
 *   pack_embed_blob), copied to the device; may be called again to replace the weights.
 * gs_align_embed: as gs_align; cost_matrix_dev (may be NULL) additionally receives the cost matrices, [N,Ta,Tb]
 *   fp32 when Ta >= Tb and [N,Tb,Ta] (transposed) when Ta < Tb.  Parity: cost matrix within 1e-2 relative of the
 *   fp32 oracle; total and path bit-exact against the oracle's DP on that cost matrix. */
int gs_set_align_encoder(gs_ctx *ctx, const float *blob_host, size_t nbytes);
int gs_align_embed(gs_ctx *ctx, const float *a_dev, const float *b_dev, int N, int Ta, int Tb, int V, int Cc,
                   float *cost_dev, int32_t *path_dev, int32_t *path_len_dev, float *cost_matrix_dev,
                   void *cuda_stream);

/* Same through HOST buffers (copies inside; returns when results are on the host). */
int gs_align_host(gs_ctx *ctx, const float *a_host, const float *b_host, int N, int Ta,
                  int Tb, int V, int Cc, float *cost_host, int32_t *path_host,
                  int32_t *path_len_host);

/* Pipelined form of gs_align_host for a stream of batches (as gs_segment_host_submit / _wait): _submit enqueues one
 * batch and returns a ticket, _wait returns when its results are in its host buffers; two batches in flight on two
 * sets of staging buffers.  The call is bound by the input copy; the pipeline hides the last chunk's sweep and the
 * result copy behind the next batch's input copy. */
int gs_align_host_submit(gs_ctx *ctx, const float *a_host, const float *b_host, int N, int Ta,
                         int Tb, int V, int Cc, float *cost_host, int32_t *path_host,
                         int32_t *path_len_host, int *ticket);
int gs_align_host_wait(gs_ctx *ctx, int ticket);

/* Materialise the pairwise cost matrices only: cost_matrix_dev [N,Ta,Tb] fp32
 * (debug / parity hook for oracle/align.py:pair_cost). */
int gs_pair_cost(gs_ctx *ctx, const float *a_dev, const float *b_dev, int N, int Ta, int Tb,
                 int V, int Cc, float *cost_matrix_dev, void *cuda_stream);

/* "Compare 2 skeleton" (README.md:50-52): per aligned step l < path_len[n], per joint v,
 * out_dev[n,l,v] = sqrt(dx*dx+dy*dy) between a[n,path[l,0],v] and b[n,path[l,1],v];
 * rows l >= path_len[n] are written as 0.  out_dev [N, Ta+Tb-1, V] fp32. */
int gs_compare(gs_ctx *ctx, const float *a_dev, const float *b_dev, const int32_t *path_dev,
               const int32_t *path_len_dev, int N, int Ta, int Tb, int V, int Cc,
               float *out_dev, void *cuda_stream);

/* Input adapter from pose estimation (README.md:15): kp_dev [B,T,V,3] fp32 =
 * (x, y, score) keypoints in image coordinates, COCO-17 joint order (V >= 13) ->
 * skel_dev [B,T,V,3] fp32 = hip-centred, torso-scaled (x, y) plus the score, the layout
 * gs_segment consumes.  Frames whose hips score under min_score take the centre of the latest
 * earlier valid frame (leading frames: the first valid one); the scale is the clip's mean
 * shoulder-centre to hip-centre distance over frames with all four joints valid (1 if none);
 * joints scoring under min_score are written as (0,0,0).  Bit-exact vs oracle/pose.py
 * (individually rounded IEEE fp32 ops, sequential sums).  kp_dev == skel_dev is NOT allowed. */
int gs_normalize_pose(gs_ctx *ctx, const float *kp_dev, float *skel_dev, int B, int T, int V,
                      float min_score, void *cuda_stream);

/* Debug hook: copy a named internal device buffer ("X", "XA", "Y", "H", "gcn_trace") to host
 * memory after synchronising the device.  Not part of the product surface. */
int gs_debug_read(gs_ctx *ctx, const char *name, void *host_out, size_t nbytes);

/* Per-kernel profiling for the roofline report: while enabled, every kernel launch is
 * bracketed by a CUDA event pair on its stream.  gs_profile_read synchronises the device,
 * folds the pending event pairs, and returns for kernel index `kernel`
 * (0 <= kernel < gs_profile_kernels()) its name, summed device time, launch count and
 * the summed ALGORITHMIC flops / bytes of those launches (un-padded). */
int gs_profile_enable(gs_ctx *ctx, int on);
int gs_profile_reset(gs_ctx *ctx);
int gs_profile_kernels(void);
int gs_profile_read(gs_ctx *ctx, int kernel, const char **name, double *total_ms,
                    int64_t *launches, double *alg_flops, double *alg_bytes);

/* The same split by network block: `block` in [0, num_blocks) or GS_MAX_BLOCKS for launches
 * outside any block (head, alignment). */
int gs_profile_read_block(gs_ctx *ctx, int kernel, int block, double *total_ms, int64_t *launches);

/* Number of kernels this context has launched since creation (bench.py gpu_launches). */
int64_t gs_launch_count(const gs_ctx *ctx);
/* Bytes of device workspace currently held by the context. */
size_t gs_workspace_bytes(const gs_ctx *ctx);
/* CUDA-event time (ms) of the kernels of the most recent gs_segment / gs_align call,
 * measured on its stream; blocks until that call has finished.  < 0 on error. */
float gs_last_kernel_ms(gs_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GOLFER_B200_H */

// C ABI of libgolfer_b200.so (include/golfer_b200.h): context lifetime, weight blob
// parsing, workspace, and the device / host entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace gs {

static thread_local char g_err[512] = "";

namespace {
std::mutex g_attr_mu;
std::map<std::pair<int, const void *>, size_t> g_dyn_smem;                        // (device, kernel) -> opted-in bytes
std::map<std::tuple<int, const void *, int, size_t>, int> g_occupancy;            // (device, kernel, threads, smem)
}  // namespace

int ensure_dyn_smem(Ctx *ctx, const void *kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return GS_OK;
    std::lock_guard<std::mutex> lk(g_attr_mu);
    size_t &have = g_dyn_smem[{ctx->device, kernel}];
    if (bytes <= have) return GS_OK;
    GS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
    return GS_OK;
}

int cached_occupancy(Ctx *ctx, const void *kernel, int nthreads, size_t smem_bytes, int *blocks_per_sm) {
    std::lock_guard<std::mutex> lk(g_attr_mu);
    const auto key = std::make_tuple(ctx->device, kernel, nthreads, smem_bytes);
    auto it = g_occupancy.find(key);
    if (it == g_occupancy.end()) {
        int n = 0;
        GS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, nthreads, smem_bytes));
        it = g_occupancy.emplace(key, n < 1 ? 1 : n).first;
    }
    *blocks_per_sm = it->second;
    return GS_OK;
}

const char *kernel_name(int id) {
    static const char *names[K_COUNT] = {
        "aggregate_fp32", "gemm_gcn_fp32", "gemm_tcn1x1_fp32", "gemm_res_fp32", "tconv_fp32", "stats", "se_gate",
        "stj_gate", "head", "features", "dtw_wavefront", "dtw_generic", "pair_cost", "compare",
        "bf16_front", "bf16_aggregate", "bf16_gemm_gcn", "bf16_gemm_tcn1x1", "bf16_tconv", "bf16_misc", "dtw_backtrack", "normalize_pose",
        "embed_encoder", "embed_cost_gemm"};
    return (id >= 0 && id < K_COUNT) ? names[id] : "?";
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

constexpr uint32_t kBlobMagic = 0x30575347u;  // "GSW0"

template <typename T>
int dmalloc(Ctx *ctx, T **p, size_t count) {
    const size_t bytes = count * sizeof(T);
    cudaError_t e = cudaMalloc((void **)p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
        return GS_ERR_NOMEM;
    }
    ctx->ws_bytes += bytes;
    return GS_OK;
}

int validate_cfg(const gs_config *c) {
    if (c->num_joints != 17 || c->num_partitions != 3 || c->kernel_size != 3) {
        set_error("this build has kernels for V=17, P=3, 3-tap branches only");
        return GS_ERR_UNSUPPORTED;
    }
    if (c->num_blocks < 1 || c->num_blocks > GS_MAX_BLOCKS || c->num_branches < 1 ||
        c->num_branches > GS_MAX_BRANCHES || c->in_channels < 1 || c->num_classes < 1 ||
        c->num_classes > 32 || c->se_reduction < 1 || c->stj_reduction < 1) {
        set_error("gs_config out of range");
        return GS_ERR_INVALID;
    }
    for (int i = 0; i < c->num_blocks; ++i) {
        const int w = c->widths[i];
        if (w < 8 || w > 1024 || w % 8 || w % c->num_branches || w % c->se_reduction || w % c->stj_reduction) {
            set_error("width[%d]=%d invalid (8..1024, multiple of 8, divisible by branches and reductions)", i, w);
            return GS_ERR_INVALID;
        }
        if ((w / c->stj_reduction) % 2) {
            set_error("width[%d]/stj_reduction must be even", i);
            return GS_ERR_INVALID;
        }
    }
    for (int r = 0; r < c->num_branches; ++r)
        if (c->dilations[r] < 1) {
            set_error("dilation[%d] must be >= 1", r);
            return GS_ERR_INVALID;
        }
    if (c->precision != GS_PREC_FP32 && c->precision != GS_PREC_BF16) {
        set_error("unknown precision %d", c->precision);
        return GS_ERR_INVALID;
    }
    return GS_OK;
}

// Walk the folded blob in params.py:fold_params order.  Every tensor is re-based to a
// 16-byte aligned offset in the device copy (kernels read weights with vector loads);
// ctx->h_blob mirrors the device layout.
int parse_blob(Ctx *ctx, const float *host, size_t nfloats) {
    const gs_config &c = ctx->cfg;
    const int V = c.num_joints, P = c.num_partitions, R = c.num_branches;
    size_t off = 0, doff = 0;
    std::vector<float> &staged = ctx->h_blob;
    auto take = [&](size_t n) {
        const float *p = ctx->d_blob + doff;
        if (off + n <= nfloats && doff + n <= staged.size())
            memcpy(staged.data() + doff, host + off, n * sizeof(float));
        off += n;
        doff = (doff + n + 3) & ~(size_t)3;
        return p;
    };
    ctx->in_scale = take((size_t)V * c.in_channels);
    ctx->in_shift = take((size_t)V * c.in_channels);
    int cin = c.in_channels;
    ctx->blocks.clear();
    for (int i = 0; i < c.num_blocks; ++i) {
        BlockParams b{};
        b.cin = cin;
        b.c = c.widths[i];
        b.cr = b.c / R;
        b.cs = b.c / c.se_reduction;
        b.cj = b.c / c.stj_reduction;
        b.has_res = (cin != b.c);
        b.A = take((size_t)P * V * V);
        b.Wg = take((size_t)P * cin * b.c);
        b.bg = take(b.c);
        b.W1 = take((size_t)b.c * b.c);
        b.b1 = take(b.c);
        b.W2 = take((size_t)R * 3 * b.cr * b.cr);
        b.b2 = take(b.c);
        if (b.has_res) {
            b.Wr = take((size_t)cin * b.c);
            b.br = take(b.c);
        }
        b.seW1 = take((size_t)b.c * b.cs);
        b.seb1 = take(b.cs);
        b.seW2 = take((size_t)b.cs * b.c);
        b.seb2 = take(b.c);
        b.jW = take((size_t)b.c * b.cj);
        b.jb = take(b.cj);
        b.jWt = take((size_t)b.cj * b.c);
        b.jbt = take(b.c);
        b.jWv = take((size_t)b.cj * b.c);
        b.jbv = take(b.c);
        ctx->blocks.push_back(b);
        cin = b.c;
    }
    ctx->headW = take((size_t)cin * c.num_classes);
    ctx->headb = take(c.num_classes);
    if (off != nfloats || doff > staged.size()) {
        set_error("weight blob holds %zu floats, config needs %zu", nfloats, off);
        return GS_ERR_INVALID;
    }
    if (cudaMemcpy(ctx->d_blob, staged.data(), doff * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("weight upload failed");
        return GS_ERR_CUDA;
    }
    return GS_OK;
}

int alloc_workspace(Ctx *ctx) {
    const gs_config &c = ctx->cfg;
    int cmax = c.in_channels;
    for (int i = 0; i < c.num_blocks; ++i) cmax = cmax > c.widths[i] ? cmax : c.widths[i];
    const size_t frames = (size_t)ctx->max_B * ctx->max_T;
    const size_t rows = frames * c.num_joints;
    const size_t esz = c.precision == GS_PREC_BF16 ? 2 : 4;
    int rc;
    auto bytes = [&](void **p, size_t n) { return dmalloc(ctx, (unsigned char **)p, n); };
    if ((rc = bytes(&ctx->bufX, rows * cmax * esz))) return rc;
    // aggregated input: a real tensor on the fp32 path only; the bf16 path never stores it (gcn_fused.cuh) except as a
    // debug dump of whole 128-row tiles when GOLFER_DEBUG_XA=1 (tests/test_gpu_kernels.py)
    size_t xa_bytes = rows * cmax * 3 * esz;
    if (c.precision == GS_PREC_BF16) {
        const char *e = getenv("GOLFER_DEBUG_XA");
        xa_bytes = (e && e[0] == '1') ? (size_t)ctx->max_B * ((ctx->max_T + 6) / 7) * 128 * 3 * cmax * 2 : 0;
    }
    if ((rc = bytes(&ctx->bufXA, xa_bytes))) return rc;
    if ((rc = bytes(&ctx->bufY, rows * cmax * esz))) return rc;
    if ((rc = bytes(&ctx->bufH, rows * cmax * esz))) return rc;
    if ((rc = bytes(&ctx->bufR, rows * cmax * esz))) return rc;
    if ((rc = bytes(&ctx->bufU[0], rows * cmax * esz))) return rc;
    if ((rc = bytes(&ctx->bufU[1], rows * cmax * esz))) return rc;
    // pooling partials: 30-frame chunks (stats_kernel; the bf16 path pools per >= 96-frame tile of tcn_fused.cuh)
    const int nchunk = (ctx->max_T + 29) / 30;
    if ((rc = dmalloc(ctx, &ctx->PT, frames * cmax))) return rc;
    if ((rc = dmalloc(ctx, &ctx->PV, (size_t)ctx->max_B * c.num_joints * cmax))) return rc;
    if ((rc = dmalloc(ctx, &ctx->PVpart, (size_t)ctx->max_B * nchunk * c.num_joints * cmax))) return rc;
    if ((rc = dmalloc(ctx, &ctx->seS, (size_t)ctx->max_B * cmax))) return rc;
    if ((rc = dmalloc(ctx, &ctx->gT, frames * cmax))) return rc;
    if ((rc = dmalloc(ctx, &ctx->gV, (size_t)ctx->max_B * c.num_joints * cmax))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_skel, rows * c.in_channels))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_logits, frames * c.num_classes))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_labels, frames))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_logits2, frames * c.num_classes))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_labels2, frames))) return rc;
    return GS_OK;
}

void free_ctx(Ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->bf16) bf16_path_destroy(ctx);
    align_embed_destroy(ctx);
    void *ptrs[] = {ctx->headWT, ctx->d_blob, ctx->bufX, ctx->bufXA, ctx->bufY, ctx->bufH, ctx->bufR, ctx->bufU[0],
                    ctx->bufU[1], ctx->PT, ctx->PV, ctx->PVpart, ctx->seS, ctx->gT, ctx->gV, ctx->d_skel,
                    ctx->d_logits, ctx->d_labels, ctx->d_logits2, ctx->d_labels2, ctx->align_ws, ctx->d_al_a, ctx->d_al_b, ctx->d_al_cost,
                    ctx->d_al_path, ctx->d_al_plen, ctx->al2_a, ctx->al2_b, ctx->al2_cost, ctx->al2_path, ctx->al2_plen};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 2; ++i) {
        if (ctx->own_stream[i]) cudaStreamDestroy(ctx->own_stream[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    }
    for (cudaEvent_t e : ctx->ev_front)
        if (e) cudaEventDestroy(e);
    if (ctx->pipe_d2h) cudaStreamDestroy(ctx->pipe_d2h);
    for (cudaEvent_t e : {ctx->ev_pipe_front, ctx->ev_pipe_head[0], ctx->ev_pipe_head[1], ctx->ev_pipe_done[0], ctx->ev_pipe_done[1],
                          ctx->ev_al_chunk[0], ctx->ev_al_chunk[1], ctx->ev_al_done[0], ctx->ev_al_done[1]})
        if (e) cudaEventDestroy(e);
    for (ProfSlot &ps : ctx->prof.pool) {
        cudaEventDestroy(ps.e0);
        cudaEventDestroy(ps.e1);
    }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
    delete ctx;
}

int forward(Ctx *ctx, const float *skel, float *logits, uint8_t *labels, int B, int T, int upto,
            float *feat, cudaStream_t st) {
    if (ctx->cfg.precision == GS_PREC_BF16)
        return segment_bf16_forward(ctx, skel, logits, labels, B, T, upto, feat, st);
    return segment_fp32_forward(ctx, skel, logits, labels, B, T, upto, feat, st);
}

int check_segment_args(Ctx *ctx, const void *in, int B, int T) {
    if (!ctx || !ctx->has_net) {
        set_error("context has no network (created without cfg/weights)");
        return GS_ERR_INVALID;
    }
    if (!in || B < 1 || T < 1 || B > ctx->max_B || T > ctx->max_T) {
        set_error("bad segment arguments: B=%d (max %d) T=%d (max %d)", B, ctx->max_B, T, ctx->max_T);
        return GS_ERR_INVALID;
    }
    return GS_OK;
}

// workspace ordering between calls on different streams (Ctx::ev_last)
int order_after_previous(Ctx *ctx, cudaStream_t st) {
    ctx->pipe_chain = false;      // any entry point but the matching submit breaks a chain of submits
    ctx->al_pipe_chain = false;
    if (ctx->ev_last_valid && ctx->last_stream != st) GS_CUDA(cudaStreamWaitEvent(st, ctx->ev_last, 0));
    return GS_OK;
}
int mark_done(Ctx *ctx, cudaStream_t st) {
    GS_CUDA(cudaEventRecord(ctx->ev_last, st));
    ctx->ev_last_valid = true;
    ctx->last_stream = st;
    return GS_OK;
}

template <typename T>
int grow(Ctx *ctx, T **p, size_t *cap, size_t count) {
    if (count <= *cap) return GS_OK;
    if (*p) {
        cudaFree(*p);
        ctx->ws_bytes -= *cap * sizeof(T);
        *p = nullptr;
        *cap = 0;
    }
    int rc = dmalloc(ctx, p, count);
    if (rc == GS_OK) *cap = count;
    return rc;
}

}  // namespace
}  // namespace gs

using namespace gs;

extern "C" {

int gs_abi_version(void) { return GS_ABI_VERSION; }

const char *gs_last_error(void) { return g_err; }

int gs_create(gs_ctx **out, int device, const gs_config *cfg, const void *weights_blob,
              size_t weights_nbytes, int max_B, int max_T) {
    if (!out) {
        set_error("out is NULL");
        return GS_ERR_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        set_error("no CUDA device visible: libgolfer_b200 has no CPU fallback");
        return GS_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range (%d visible)", device, ndev);
        return GS_ERR_INVALID;
    }
    cudaDeviceProp prop;
    GS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                  prop.minor);
        return GS_ERR_NO_DEVICE;
    }
    GS_CUDA(cudaSetDevice(device));
    Ctx *ctx = new Ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    int rc = GS_OK;
    do {
        for (int i = 0; i < 2; ++i) {
            if (cudaStreamCreateWithFlags(&ctx->own_stream[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming) != cudaSuccess) {
                set_error("stream/event creation failed");
                rc = GS_ERR_CUDA;
                break;
            }
        }
        {
            bool ok = cudaStreamCreateWithFlags(&ctx->pipe_d2h, cudaStreamNonBlocking) == cudaSuccess &&
                      cudaEventCreateWithFlags(&ctx->ev_pipe_front, cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; i < 2 && ok; ++i)
                ok = cudaEventCreateWithFlags(&ctx->ev_pipe_head[i], cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&ctx->ev_pipe_done[i], cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&ctx->ev_al_chunk[i], cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&ctx->ev_al_done[i], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) {
                set_error("stream/event creation failed");
                rc = GS_ERR_CUDA;
                break;
            }
        }
        for (int i = 0; i < Ctx::kFrontChunks && !rc; ++i)
            if (cudaEventCreateWithFlags(&ctx->ev_front[i], cudaEventDisableTiming) != cudaSuccess) {
                set_error("stream/event creation failed");
                rc = GS_ERR_CUDA;
            }
        if (rc) break;
        if (cudaEventCreate(&ctx->ev_start) != cudaSuccess || cudaEventCreate(&ctx->ev_stop) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_last, cudaEventDisableTiming) != cudaSuccess) {
            set_error("event creation failed");
            rc = GS_ERR_CUDA;
            break;
        }
        if (!cfg) break;  // alignment-only context
        if ((rc = validate_cfg(cfg))) break;
        if (!weights_blob || weights_nbytes < 16 || max_B < 1 || max_T < 1) {
            set_error("weights blob / max_B / max_T missing");
            rc = GS_ERR_INVALID;
            break;
        }
        ctx->cfg = *cfg;
        ctx->max_B = max_B;
        ctx->max_T = max_T;
        const uint32_t *head = (const uint32_t *)weights_blob;
        const size_t nfloats = head[1];
        if (head[0] != kBlobMagic || (nfloats + 4) * 4 != weights_nbytes || (int)head[2] != cfg->num_blocks) {
            set_error("weight blob header mismatch (magic %08x, n_floats %u, blocks %u, nbytes %zu)", head[0],
                      head[1], head[2], weights_nbytes);
            rc = GS_ERR_INVALID;
            break;
        }
        const size_t padded = nfloats + 4 * (size_t)(32 * cfg->num_blocks + 16);   // room for per-tensor alignment
        if ((rc = dmalloc(ctx, &ctx->d_blob, padded))) break;
        ctx->blob_floats = nfloats;
        const float *body = (const float *)weights_blob + 4;
        ctx->h_blob.assign(padded, 0.f);
        if ((rc = parse_blob(ctx, body, nfloats))) break;
        if ((rc = alloc_workspace(ctx))) break;
        {   // head weights [C,K] -> [K,C]
            const int C = cfg->widths[cfg->num_blocks - 1], K = cfg->num_classes;
            const float *hw = ctx->h_blob.data() + (ctx->headW - ctx->d_blob);
            std::vector<float> t((size_t)K * C);
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < K; ++k) t[(size_t)k * C + c] = hw[(size_t)c * K + k];
            if ((rc = dmalloc(ctx, &ctx->headWT, t.size()))) break;
            if (cudaMemcpy(ctx->headWT, t.data(), t.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("head weight upload failed");
                rc = GS_ERR_CUDA;
                break;
            }
        }
        ctx->has_net = true;
        if (cfg->precision == GS_PREC_BF16 && (rc = bf16_path_create(ctx))) break;
    } while (0);
    if (rc != GS_OK) {
        free_ctx(ctx);
        return rc;
    }
    *out = (gs_ctx *)ctx;
    return GS_OK;
}

int gs_destroy(gs_ctx *h) {
    free_ctx((Ctx *)h);
    return GS_OK;
}

int gs_segment(gs_ctx *h, const float *skel_dev, float *logits_dev, uint8_t *labels_dev, int B, int T,
               void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_segment_args(ctx, skel_dev, B, T);
    if (rc) return rc;
    if (!logits_dev) {
        set_error("logits_dev is NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if ((rc = order_after_previous(ctx, st))) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_start, st));
    rc = forward(ctx, skel_dev, logits_dev, labels_dev, B, T, -1, nullptr, st);
    if (rc) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_stop, st));
    ctx->ev_valid = true;
    return mark_done(ctx, st);
}

int gs_segment_features(gs_ctx *h, const float *skel_dev, int block, float *out_dev, int B, int T,
                        void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_segment_args(ctx, skel_dev, B, T);
    if (rc) return rc;
    if (!out_dev || block < 0 || block >= ctx->cfg.num_blocks) {
        set_error("bad block index %d", block);
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if ((rc = order_after_previous(ctx, st))) return rc;
    if ((rc = forward(ctx, skel_dev, nullptr, nullptr, B, T, block, out_dev, st))) return rc;
    return mark_done(ctx, st);
}

int gs_segment_host(gs_ctx *h, const float *skel_host, float *logits_host, uint8_t *labels_host, int B,
                    int T) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_segment_args(ctx, skel_host, B, T);
    if (rc) return rc;
    GS_CUDA(cudaSetDevice(ctx->device));
    const gs_config &c = ctx->cfg;
    const size_t in_per = (size_t)T * c.num_joints * c.in_channels;
    const size_t out_per = (size_t)T * c.num_classes;
    cudaStream_t sc = ctx->own_stream[0], sx = ctx->own_stream[1];
    if (ctx->cfg.precision == GS_PREC_BF16) {
        // bf16 path: the 15.7 MB input (B = 256) arrives in up to 4 chunks on the copy stream and ONLY block 0's input
        // kernel runs per chunk (frames are independent there); every later kernel runs once on the whole batch.  Splitting
        // the whole forward pass instead (the fp32 path below) costs the persistent kernels their efficiency: 2 chunks
        // of 64 + 192 clips measured 57.3 k clips/s end to end against 63.4 k device-resident.
        if ((rc = order_after_previous(ctx, sc))) return rc;
        if ((rc = order_after_previous(ctx, sx))) return rc;
        GS_CUDA(cudaEventRecord(ctx->ev_start, sc));
        const int nch = B >= 4 * Ctx::kFrontChunks ? Ctx::kFrontChunks : 1;
        const int per = (B + nch - 1) / nch;
        ctx->front_nchunks = 0;
        for (int k = 0, b0 = 0; b0 < B; ++k, b0 += per) {
            const int nb = (B - b0) < per ? (B - b0) : per;
            GS_CUDA(cudaMemcpyAsync(ctx->d_skel + b0 * in_per, skel_host + b0 * in_per, nb * in_per * 4,
                                    cudaMemcpyHostToDevice, sx));
            GS_CUDA(cudaEventRecord(ctx->ev_front[k], sx));
            ctx->front_b0[k] = b0;
            ctx->front_nb[k] = nb;
            ctx->front_nchunks = k + 1;
        }
        rc = forward(ctx, ctx->d_skel, ctx->d_logits, labels_host ? ctx->d_labels : nullptr, B, T, -1, nullptr, sc);
        ctx->front_nchunks = 0;
        if (rc) return rc;
        if (logits_host)
            GS_CUDA(cudaMemcpyAsync(logits_host, ctx->d_logits, B * out_per * 4, cudaMemcpyDeviceToHost, sc));
        if (labels_host)
            GS_CUDA(cudaMemcpyAsync(labels_host, ctx->d_labels, (size_t)B * T, cudaMemcpyDeviceToHost, sc));
        GS_CUDA(cudaEventRecord(ctx->ev_stop, sc));
        ctx->ev_valid = true;
        if ((rc = mark_done(ctx, sc))) return rc;
        GS_CUDA(cudaStreamSynchronize(sc));
        return GS_OK;
    }
    // fp32 path: clips are independent: split the batch in two chunks; the H2D of the second (copy stream) overlaps
    // the kernels of the first (compute stream); D2H follows each chunk.  A quarter-size first chunk gets the
    // kernels started early.  Measured at B = 256 (e2e clips/s): 1 chunk 42.8 k, 2 equal 42.1 k, 2 with a
    // quarter first 43.2 k, 3 chunks 35-41 k, 4 equal 37.4 k: the persistent kernels lose efficiency on small
    // batches faster than the 15.7 MB input copy (~0.35 ms) is worth hiding.
    const int nchunks = B >= 16 ? 2 : 1;
    const int first = B >= 16 ? B / 4 : B;
    const int rest = nchunks > 1 ? (B - first + nchunks - 2) / (nchunks - 1) : B;
    if ((rc = order_after_previous(ctx, sc))) return rc;
    if ((rc = order_after_previous(ctx, sx))) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_start, sc));
    for (int k = 0, b0 = 0; b0 < B; ++k) {
        const int per = k == 0 ? first : rest;
        const int nb = (B - b0) < per ? (B - b0) : per;
        const int e = k & 1;
        GS_CUDA(cudaMemcpyAsync(ctx->d_skel + b0 * in_per, skel_host + b0 * in_per, nb * in_per * 4,
                                cudaMemcpyHostToDevice, sx));
        GS_CUDA(cudaEventRecord(ctx->ev_copy[e], sx));
        GS_CUDA(cudaStreamWaitEvent(sc, ctx->ev_copy[e], 0));
        // workspace is shared between chunks: kernels of chunk k run after chunk k-1 on `sc`
        rc = forward(ctx, ctx->d_skel + b0 * in_per, ctx->d_logits + b0 * out_per,
                     labels_host ? ctx->d_labels + (size_t)b0 * T : nullptr, nb, T, -1, nullptr, sc);
        if (rc) return rc;
        if (logits_host)
            GS_CUDA(cudaMemcpyAsync(logits_host + b0 * out_per, ctx->d_logits + b0 * out_per, nb * out_per * 4,
                                    cudaMemcpyDeviceToHost, sc));
        if (labels_host)
            GS_CUDA(cudaMemcpyAsync(labels_host + (size_t)b0 * T, ctx->d_labels + (size_t)b0 * T, (size_t)nb * T,
                                    cudaMemcpyDeviceToHost, sc));
        b0 += nb;
    }
    GS_CUDA(cudaEventRecord(ctx->ev_stop, sc));
    ctx->ev_valid = true;
    if ((rc = mark_done(ctx, sc))) return rc;
    GS_CUDA(cudaStreamSynchronize(sc));
    return GS_OK;
}

int gs_segment_host_submit(gs_ctx *h, const float *skel_host, float *logits_host, uint8_t *labels_host, int B, int T,
                           int *ticket) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_segment_args(ctx, skel_host, B, T);
    if (rc) return rc;
    if (!ticket || (!logits_host && !labels_host)) {
        set_error("gs_segment_host_submit: ticket and at least one output buffer must be set");
        return GS_ERR_INVALID;
    }
    if (ctx->cfg.precision != GS_PREC_BF16) {
        set_error("gs_segment_host_submit: bf16 contexts only (the fp32 path has the synchronous entry point)");
        return GS_ERR_UNSUPPORTED;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    const gs_config &c = ctx->cfg;
    const size_t in_per = (size_t)T * c.num_joints * c.in_channels;
    const size_t out_per = (size_t)T * c.num_classes;
    cudaStream_t sc = ctx->own_stream[0], sx = ctx->own_stream[1], sd = ctx->pipe_d2h;
    const int slot = (int)(ctx->pipe_next & 1);
    // at most two batches in flight: the batch that used this slot (ticket - 2) must be complete, its device output
    // buffer and the caller's host buffers of that batch are then free
    if (ctx->pipe_done_valid[slot]) GS_CUDA(cudaEventSynchronize(ctx->ev_pipe_done[slot]));
    const bool chained = ctx->pipe_chain;
    if (!chained) {
        // first submit after another entry point: order all three streams behind that call
        if (ctx->ev_last_valid) {
            GS_CUDA(cudaStreamWaitEvent(sc, ctx->ev_last, 0));
            GS_CUDA(cudaStreamWaitEvent(sx, ctx->ev_last, 0));
            GS_CUDA(cudaStreamWaitEvent(sd, ctx->ev_last, 0));
        }
    } else if (ctx->pipe_front_valid) {
        // the previous batch's input kernels must have read d_skel before it is overwritten
        GS_CUDA(cudaStreamWaitEvent(sx, ctx->ev_pipe_front, 0));
    }
    const int nch = B >= 4 * Ctx::kFrontChunks ? Ctx::kFrontChunks : 1;
    const int per = (B + nch - 1) / nch;
    ctx->front_nchunks = 0;
    for (int k = 0, b0 = 0; b0 < B; ++k, b0 += per) {
        const int nb = (B - b0) < per ? (B - b0) : per;
        GS_CUDA(cudaMemcpyAsync(ctx->d_skel + b0 * in_per, skel_host + b0 * in_per, nb * in_per * 4,
                                cudaMemcpyHostToDevice, sx));
        GS_CUDA(cudaEventRecord(ctx->ev_front[k], sx));
        ctx->front_b0[k] = b0;
        ctx->front_nb[k] = nb;
        ctx->front_nchunks = k + 1;
    }
    float *dl = slot ? ctx->d_logits2 : ctx->d_logits;
    uint8_t *db = slot ? ctx->d_labels2 : ctx->d_labels;
    rc = forward(ctx, ctx->d_skel, dl, labels_host ? db : nullptr, B, T, -1, nullptr, sc);
    ctx->front_nchunks = 0;
    if (rc) return rc;
    ctx->pipe_front_valid = true;             // recorded by the forward pass behind its input kernels
    GS_CUDA(cudaEventRecord(ctx->ev_pipe_head[slot], sc));
    GS_CUDA(cudaStreamWaitEvent(sd, ctx->ev_pipe_head[slot], 0));
    if (logits_host) GS_CUDA(cudaMemcpyAsync(logits_host, dl, B * out_per * 4, cudaMemcpyDeviceToHost, sd));
    if (labels_host) GS_CUDA(cudaMemcpyAsync(labels_host, db, (size_t)B * T, cudaMemcpyDeviceToHost, sd));
    GS_CUDA(cudaEventRecord(ctx->ev_pipe_done[slot], sd));
    ctx->pipe_done_valid[slot] = true;
    // everything of this batch is behind ev_last (the result copy is the last thing it does)
    if ((rc = mark_done(ctx, sd))) return rc;
    ctx->pipe_chain = true;
    ctx->al_pipe_chain = false;
    *ticket = (int)(ctx->pipe_next & 0x7fffffff);
    ctx->pipe_next += 1;
    return GS_OK;
}

int gs_segment_host_wait(gs_ctx *h, int ticket) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx) {
        set_error("gs_segment_host_wait: null context");
        return GS_ERR_INVALID;
    }
    const long long next = ctx->pipe_next;
    const long long t = (next & ~0x7fffffffLL) | (long long)(unsigned)ticket;
    if (ticket < 0 || t >= next) {
        set_error("gs_segment_host_wait: ticket %d was never issued", ticket);
        return GS_ERR_INVALID;
    }
    if (t < next - 2) return GS_OK;           // older batches were completed when their slot was reused
    GS_CUDA(cudaSetDevice(ctx->device));
    GS_CUDA(cudaEventSynchronize(ctx->ev_pipe_done[(int)(t & 1)]));
    return GS_OK;
}

static int check_align_args(Ctx *ctx, const void *a, const void *b, int N, int Ta, int Tb, int V, int Cc) {
    if (!ctx || !a || !b || N < 1 || Ta < 1 || Tb < 1 || V < 1 || Cc < 2) {
        set_error("bad align arguments: N=%d Ta=%d Tb=%d V=%d Cc=%d", N, Ta, Tb, V, Cc);
        return GS_ERR_INVALID;
    }
    if ((long long)Ta + Tb - 1 > 65535) {
        set_error("Ta+Tb-1 must be <= 65535");
        return GS_ERR_UNSUPPORTED;
    }
    return GS_OK;
}

int gs_align(gs_ctx *h, const float *a_dev, const float *b_dev, int N, int Ta, int Tb, int V, int Cc,
             float *cost_dev, int32_t *path_dev, int32_t *path_len_dev, void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!cost_dev || ((path_dev == nullptr) != (path_len_dev == nullptr))) {
        set_error("cost_dev must be set; path_dev and path_len_dev must both be set or both NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if ((rc = order_after_previous(ctx, st))) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_start, st));
    rc = align_launch(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc, cost_dev, path_dev, path_len_dev, st);
    if (rc) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_stop, st));
    ctx->ev_valid = true;
    return mark_done(ctx, st);
}

int gs_align_phase(gs_ctx *h, const float *a_dev, const float *b_dev, const uint8_t *labels_a_dev,
                   const uint8_t *labels_b_dev, float penalty, int N, int Ta, int Tb, int V, int Cc,
                   float *cost_dev, int32_t *path_dev, int32_t *path_len_dev, void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!cost_dev || ((path_dev == nullptr) != (path_len_dev == nullptr))) {
        set_error("cost_dev must be set; path_dev and path_len_dev must both be set or both NULL");
        return GS_ERR_INVALID;
    }
    if (!labels_a_dev || !labels_b_dev || penalty != penalty) {
        set_error("align_phase: labels_a / labels_b must be set and penalty must not be NaN");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if ((rc = order_after_previous(ctx, st))) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_start, st));
    rc = align_launch(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc, cost_dev, path_dev, path_len_dev, st, labels_a_dev,
                      labels_b_dev, penalty);
    if (rc) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_stop, st));
    ctx->ev_valid = true;
    return mark_done(ctx, st);
}

int gs_align_host(gs_ctx *h, const float *a_host, const float *b_host, int N, int Ta, int Tb, int V, int Cc,
                  float *cost_host, int32_t *path_host, int32_t *path_len_host) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_host, b_host, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!cost_host || ((path_host == nullptr) != (path_len_host == nullptr))) {
        set_error("cost_host must be set; path_host and path_len_host must both be set or both NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    const size_t na = (size_t)N * Ta * V * Cc, nb = (size_t)N * Tb * V * Cc;
    const size_t maxL = (size_t)Ta + Tb - 1;
    if ((rc = grow(ctx, &ctx->d_al_a, &ctx->al_host_cap[0], na))) return rc;
    if ((rc = grow(ctx, &ctx->d_al_b, &ctx->al_host_cap[1], nb))) return rc;
    if ((rc = grow(ctx, &ctx->d_al_cost, &ctx->al_host_cap[2], (size_t)N))) return rc;
    if (path_host) {
        if ((rc = grow(ctx, &ctx->d_al_path, &ctx->al_host_cap[3], (size_t)N * maxL * 2))) return rc;
        if ((rc = grow(ctx, &ctx->d_al_plen, &ctx->al_host_cap[4], (size_t)N))) return rc;
    }
    // Pairs are independent: chunk so the H2D of chunk k+1 overlaps the DTW of chunk k.  The call is bound by the
    // host-to-device copy (334 MB for 4096 pairs of 300 frames: ~6 ms at 55 GB/s against 4 ms of sweep), so what is
    // left to hide is the tail behind the last copy: the last chunk's sweep and its results.
    // Chunks are whole rounds of the persistent sweep (a multiple of the SM count: one CTA per SM, pairs one after the
    // other), about 16 of them: 4096 pairs -> 14 chunks of 296.  Measured (pairs/s end to end): 4 chunks 566 k,
    // 8 chunks 600 k.
    int per = N;
    if (N >= 64) {
        const int m = (N + 16 * ctx->sm_count - 1) / (16 * ctx->sm_count);
        per = m * ctx->sm_count;
        if (N < 4 * per) per = (N + 3) / 4;
    }
    cudaStream_t sc = ctx->own_stream[0], sx = ctx->own_stream[1];
    if ((rc = order_after_previous(ctx, sc))) return rc;
    if ((rc = order_after_previous(ctx, sx))) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_start, sc));
    for (int k = 0, n0 = 0; n0 < N; ++k, n0 += per) {
        const int cnt = (N - n0) < per ? (N - n0) : per;
        const int e = k & 1;
        const size_t oa = (size_t)n0 * Ta * V * Cc, ob = (size_t)n0 * Tb * V * Cc;
        GS_CUDA(cudaMemcpyAsync(ctx->d_al_a + oa, a_host + oa, (size_t)cnt * Ta * V * Cc * 4,
                                cudaMemcpyHostToDevice, sx));
        GS_CUDA(cudaMemcpyAsync(ctx->d_al_b + ob, b_host + ob, (size_t)cnt * Tb * V * Cc * 4,
                                cudaMemcpyHostToDevice, sx));
        GS_CUDA(cudaEventRecord(ctx->ev_copy[e], sx));
        GS_CUDA(cudaStreamWaitEvent(sc, ctx->ev_copy[e], 0));
        rc = align_launch(ctx, ctx->d_al_a + oa, ctx->d_al_b + ob, cnt, Ta, Tb, V, Cc, ctx->d_al_cost + n0,
                          path_host ? ctx->d_al_path + (size_t)n0 * maxL * 2 : nullptr,
                          path_host ? ctx->d_al_plen + n0 : nullptr, sc);
        if (rc) return rc;
        GS_CUDA(cudaMemcpyAsync(cost_host + n0, ctx->d_al_cost + n0, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sc));
        if (path_host) {
            GS_CUDA(cudaMemcpyAsync(path_host + (size_t)n0 * maxL * 2, ctx->d_al_path + (size_t)n0 * maxL * 2,
                                    (size_t)cnt * maxL * 8, cudaMemcpyDeviceToHost, sc));
            GS_CUDA(cudaMemcpyAsync(path_len_host + n0, ctx->d_al_plen + n0, (size_t)cnt * 4,
                                    cudaMemcpyDeviceToHost, sc));
        }
    }
    GS_CUDA(cudaEventRecord(ctx->ev_stop, sc));
    ctx->ev_valid = true;
    if ((rc = mark_done(ctx, sc))) return rc;
    GS_CUDA(cudaStreamSynchronize(sc));
    return GS_OK;
}

int gs_align_host_submit(gs_ctx *h, const float *a_host, const float *b_host, int N, int Ta, int Tb, int V, int Cc,
                         float *cost_host, int32_t *path_host, int32_t *path_len_host, int *ticket) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_host, b_host, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!ticket || !cost_host || ((path_host == nullptr) != (path_len_host == nullptr))) {
        set_error("ticket and cost_host must be set; path_host and path_len_host must both be set or both NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    const int slot = (int)(ctx->al_pipe_next & 1);
    // at most two batches in flight: the batch that used this set of staging buffers (ticket - 2) must be complete
    if (ctx->al_done_valid[slot]) GS_CUDA(cudaEventSynchronize(ctx->ev_al_done[slot]));
    const size_t na = (size_t)N * Ta * V * Cc, nb = (size_t)N * Tb * V * Cc;
    const size_t maxL = (size_t)Ta + Tb - 1;
    float **pa = slot ? &ctx->al2_a : &ctx->d_al_a, **pb = slot ? &ctx->al2_b : &ctx->d_al_b;
    float **pc = slot ? &ctx->al2_cost : &ctx->d_al_cost;
    int32_t **pp = slot ? &ctx->al2_path : &ctx->d_al_path, **pl = slot ? &ctx->al2_plen : &ctx->d_al_plen;
    size_t *cap = slot ? ctx->al2_cap : ctx->al_host_cap;
    if ((rc = grow(ctx, pa, &cap[0], na))) return rc;
    if ((rc = grow(ctx, pb, &cap[1], nb))) return rc;
    if ((rc = grow(ctx, pc, &cap[2], (size_t)N))) return rc;
    if (path_host) {
        if ((rc = grow(ctx, pp, &cap[3], (size_t)N * maxL * 2))) return rc;
        if ((rc = grow(ctx, pl, &cap[4], (size_t)N))) return rc;
    }
    int per = N;          // chunks of whole sweep rounds, as in gs_align_host
    if (N >= 64) {
        const int m = (N + 16 * ctx->sm_count - 1) / (16 * ctx->sm_count);
        per = m * ctx->sm_count;
        if (N < 4 * per) per = (N + 3) / 4;
    }
    cudaStream_t sc = ctx->own_stream[0], sx = ctx->own_stream[1], sd = ctx->pipe_d2h;
    if (!ctx->al_pipe_chain && ctx->ev_last_valid) {
        // first submit after another entry point: order all three streams behind that call
        GS_CUDA(cudaStreamWaitEvent(sc, ctx->ev_last, 0));
        GS_CUDA(cudaStreamWaitEvent(sx, ctx->ev_last, 0));
        GS_CUDA(cudaStreamWaitEvent(sd, ctx->ev_last, 0));
    }
    for (int k = 0, n0 = 0; n0 < N; ++k, n0 += per) {
        const int cnt = (N - n0) < per ? (N - n0) : per;
        const int e = k & 1;
        const size_t oa = (size_t)n0 * Ta * V * Cc, ob = (size_t)n0 * Tb * V * Cc;
        GS_CUDA(cudaMemcpyAsync(*pa + oa, a_host + oa, (size_t)cnt * Ta * V * Cc * 4, cudaMemcpyHostToDevice, sx));
        GS_CUDA(cudaMemcpyAsync(*pb + ob, b_host + ob, (size_t)cnt * Tb * V * Cc * 4, cudaMemcpyHostToDevice, sx));
        GS_CUDA(cudaEventRecord(ctx->ev_copy[e], sx));
        GS_CUDA(cudaStreamWaitEvent(sc, ctx->ev_copy[e], 0));
        rc = align_launch(ctx, *pa + oa, *pb + ob, cnt, Ta, Tb, V, Cc, *pc + n0,
                          path_host ? *pp + (size_t)n0 * maxL * 2 : nullptr, path_host ? *pl + n0 : nullptr, sc);
        if (rc) return rc;
        // results leave on the third stream: the next chunk's sweep does not wait for them
        GS_CUDA(cudaEventRecord(ctx->ev_al_chunk[e], sc));
        GS_CUDA(cudaStreamWaitEvent(sd, ctx->ev_al_chunk[e], 0));
        GS_CUDA(cudaMemcpyAsync(cost_host + n0, *pc + n0, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sd));
        if (path_host) {
            GS_CUDA(cudaMemcpyAsync(path_host + (size_t)n0 * maxL * 2, *pp + (size_t)n0 * maxL * 2, (size_t)cnt * maxL * 8,
                                    cudaMemcpyDeviceToHost, sd));
            GS_CUDA(cudaMemcpyAsync(path_len_host + n0, *pl + n0, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sd));
        }
    }
    GS_CUDA(cudaEventRecord(ctx->ev_al_done[slot], sd));
    ctx->al_done_valid[slot] = true;
    if ((rc = mark_done(ctx, sd))) return rc;      // the last result copy is the last thing this batch does
    ctx->al_pipe_chain = true;
    ctx->pipe_chain = false;
    *ticket = (int)(ctx->al_pipe_next & 0x7fffffff);
    ctx->al_pipe_next += 1;
    return GS_OK;
}

int gs_align_host_wait(gs_ctx *h, int ticket) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx) {
        set_error("gs_align_host_wait: null context");
        return GS_ERR_INVALID;
    }
    const long long next = ctx->al_pipe_next;
    const long long t = (next & ~0x7fffffffLL) | (long long)(unsigned)ticket;
    if (ticket < 0 || t >= next) {
        set_error("gs_align_host_wait: ticket %d was never issued", ticket);
        return GS_ERR_INVALID;
    }
    if (t < next - 2) return GS_OK;           // older batches were completed when their buffers were reused
    GS_CUDA(cudaSetDevice(ctx->device));
    GS_CUDA(cudaEventSynchronize(ctx->ev_al_done[(int)(t & 1)]));
    return GS_OK;
}

int gs_set_align_encoder(gs_ctx *h, const float *blob_host, size_t nbytes) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx || !blob_host || nbytes % 4) {
        set_error("gs_set_align_encoder: ctx / blob missing or size not a multiple of 4");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    return align_embed_set_encoder(ctx, blob_host, nbytes / 4);
}

int gs_align_embed(gs_ctx *h, const float *a_dev, const float *b_dev, int N, int Ta, int Tb, int V, int Cc,
                   float *cost_dev, int32_t *path_dev, int32_t *path_len_dev, float *cost_matrix_dev, void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!cost_dev || ((path_dev == nullptr) != (path_len_dev == nullptr))) {
        set_error("cost_dev must be set; path_dev and path_len_dev must both be set or both NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if ((rc = order_after_previous(ctx, st))) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_start, st));
    rc = align_embed_launch(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc, cost_dev, path_dev, path_len_dev, cost_matrix_dev, st);
    if (rc) return rc;
    GS_CUDA(cudaEventRecord(ctx->ev_stop, st));
    ctx->ev_valid = true;
    return mark_done(ctx, st);
}

int gs_pair_cost(gs_ctx *h, const float *a_dev, const float *b_dev, int N, int Ta, int Tb, int V, int Cc,
                 float *cost_matrix_dev, void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!cost_matrix_dev) {
        set_error("cost_matrix_dev is NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    return pair_cost_launch(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc, cost_matrix_dev, (cudaStream_t)cuda_stream);
}

int gs_compare(gs_ctx *h, const float *a_dev, const float *b_dev, const int32_t *path_dev,
               const int32_t *path_len_dev, int N, int Ta, int Tb, int V, int Cc, float *out_dev,
               void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    int rc = check_align_args(ctx, a_dev, b_dev, N, Ta, Tb, V, Cc);
    if (rc) return rc;
    if (!path_dev || !path_len_dev || !out_dev) {
        set_error("path / path_len / out is NULL");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    return compare_launch(ctx, a_dev, b_dev, path_dev, path_len_dev, N, Ta, Tb, V, Cc, out_dev,
                          (cudaStream_t)cuda_stream);
}

int gs_normalize_pose(gs_ctx *h, const float *kp_dev, float *skel_dev, int B, int T, int V, float min_score,
                      void *cuda_stream) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx) {
        set_error("ctx is NULL");
        return GS_ERR_INVALID;
    }
    if (!kp_dev || !skel_dev || B < 0 || T < 1 || V < 13 || !(min_score == min_score)) {
        set_error("normalize_pose: need kp/skel pointers, B >= 0, T >= 1, V >= 13 (COCO hips are joints 11, 12), "
                  "a non-NaN min_score; got B=%d T=%d V=%d", B, T, V);
        return GS_ERR_INVALID;
    }
    if (B == 0) return GS_OK;
    GS_CUDA(cudaSetDevice(ctx->device));
    return normalize_pose_launch(ctx, kp_dev, skel_dev, B, T, V, min_score, (cudaStream_t)cuda_stream);
}

int gs_debug_read(gs_ctx *h, const char *name, void *host_out, size_t nbytes) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx || !name || !host_out) {
        set_error("bad gs_debug_read arguments");
        return GS_ERR_INVALID;
    }
    GS_CUDA(cudaSetDevice(ctx->device));
    return bf16_debug_read(ctx, name, host_out, nbytes);
}

int gs_profile_enable(gs_ctx *h, int on) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx) return GS_ERR_INVALID;
    ctx->prof.on = on != 0;
    return GS_OK;
}

int gs_profile_reset(gs_ctx *h) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx) return GS_ERR_INVALID;
    Profiler &p = ctx->prof;
    p.used = 0;
    for (int i = 0; i < K_COUNT; ++i) {
        p.ms[i] = p.flops[i] = p.bytes[i] = 0, p.launches[i] = 0;
        for (int b = 0; b <= GS_MAX_BLOCKS; ++b) p.ms_blk[i][b] = 0, p.launches_blk[i][b] = 0;
    }
    return GS_OK;
}

int gs_profile_kernels(void) { return K_COUNT; }

static int profile_fold(Ctx *ctx) {
    Profiler &p = ctx->prof;
    if (p.used) {   // fold the recorded event pairs into the per-kernel totals
        GS_CUDA(cudaSetDevice(ctx->device));
        GS_CUDA(cudaDeviceSynchronize());
        for (size_t i = 0; i < p.used; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.pool[i].e0, p.pool[i].e1) == cudaSuccess) {
                p.ms[p.pool[i].id] += ms;
                p.ms_blk[p.pool[i].id][p.pool[i].blk] += ms;
            }
        }
        p.used = 0;
    }
    return GS_OK;
}

int gs_profile_read_block(gs_ctx *h, int kernel, int block, double *total_ms, int64_t *launches) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx || kernel < 0 || kernel >= K_COUNT || block < 0 || block > GS_MAX_BLOCKS) return GS_ERR_INVALID;
    int rc = profile_fold(ctx);
    if (rc) return rc;
    if (total_ms) *total_ms = ctx->prof.ms_blk[kernel][block];
    if (launches) *launches = ctx->prof.launches_blk[kernel][block];
    return GS_OK;
}

int gs_profile_read(gs_ctx *h, int kernel, const char **name, double *total_ms, int64_t *launches,
                    double *alg_flops, double *alg_bytes) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx || kernel < 0 || kernel >= K_COUNT) return GS_ERR_INVALID;
    Profiler &p = ctx->prof;
    int rc = profile_fold(ctx);
    if (rc) return rc;
    if (name) *name = kernel_name(kernel);
    if (total_ms) *total_ms = p.ms[kernel];
    if (launches) *launches = p.launches[kernel];
    if (alg_flops) *alg_flops = p.flops[kernel];
    if (alg_bytes) *alg_bytes = p.bytes[kernel];
    return GS_OK;
}

int64_t gs_launch_count(const gs_ctx *h) { return h ? ((const Ctx *)h)->launches : -1; }

size_t gs_workspace_bytes(const gs_ctx *h) { return h ? ((const Ctx *)h)->ws_bytes : 0; }

float gs_last_kernel_ms(gs_ctx *h) {
    Ctx *ctx = (Ctx *)h;
    if (!ctx || !ctx->ev_valid) return -1.f;
    if (cudaEventSynchronize(ctx->ev_stop) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop) != cudaSuccess) return -1.f;
    return ms;
}

}  // extern "C"

// Input adapter between pose estimation and the segmentation network.
//
// Stage replaced: /root/reference/README.md:15 (pose-estimation output feeding the skeleton models).
// Arithmetic contract: oracle/pose.py — hip-centred (forward-filled over frames whose hips are below
// the score threshold), scaled by the clip's mean torso length (sequential fp32 sum over valid frames),
// low-score joints masked to (0,0,0); every op an individually rounded IEEE fp32 op.
//
// One CTA per clip.  Phase A (all threads): per-frame hip centre, torso length and validity flags into
// shared memory.  Phase B (one thread): the two order-dependent scans (centre fill, length sum) — T
// steps on shared memory.  Phase C (all threads): one coalesced pass over the clip's keypoints.
// HBM-bound by construction: 12 B in + 12 B out per keypoint, plus 4 joints per frame re-read from L2.
#include "common.cuh"

namespace gs {

namespace {

constexpr int kLShoulder = 5, kRShoulder = 6, kLHip = 11, kRHip = 12;

__global__ void __launch_bounds__(1024)
normalize_pose_kernel(const float *__restrict__ kp, float *__restrict__ out, int T, int V, float thr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *cx = reinterpret_cast<float *>(smem_raw);     // hip centre, then filled centre
    float *cy = cx + T;
    float *len = cy + T;
    unsigned char *flags = reinterpret_cast<unsigned char *>(len + T);   // bit0 hip_ok, bit1 len_ok
    __shared__ float s_scale;
    const size_t clip = (size_t)blockIdx.x * T * V * 3;
    const float *k = kp + clip;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const float *f = k + (size_t)t * V * 3;
        const float lsx = f[kLShoulder * 3], lsy = f[kLShoulder * 3 + 1], lss = f[kLShoulder * 3 + 2];
        const float rsx = f[kRShoulder * 3], rsy = f[kRShoulder * 3 + 1], rss = f[kRShoulder * 3 + 2];
        const float lhx = f[kLHip * 3], lhy = f[kLHip * 3 + 1], lhs = f[kLHip * 3 + 2];
        const float rhx = f[kRHip * 3], rhy = f[kRHip * 3 + 1], rhs = f[kRHip * 3 + 2];
        const float hx = __fmul_rn(0.5f, __fadd_rn(lhx, rhx)), hy = __fmul_rn(0.5f, __fadd_rn(lhy, rhy));
        const float sx = __fmul_rn(0.5f, __fadd_rn(lsx, rsx)), sy = __fmul_rn(0.5f, __fadd_rn(lsy, rsy));
        const float dx = __fsub_rn(sx, hx), dy = __fsub_rn(sy, hy);
        cx[t] = hx;
        cy[t] = hy;
        len[t] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const bool hip_ok = lhs >= thr && rhs >= thr;
        const bool len_ok = hip_ok && lss >= thr && rss >= thr;
        flags[t] = (unsigned char)((hip_ok ? 1 : 0) | (len_ok ? 2 : 0));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float lx = 0.f, ly = 0.f;
        for (int t = 0; t < T; ++t)
            if (flags[t] & 1) { lx = cx[t]; ly = cy[t]; break; }     // leading frames take the first valid centre
        float acc = 0.f;
        int n = 0;
        for (int t = 0; t < T; ++t) {
            const unsigned char fl = flags[t];
            if (fl & 1) { lx = cx[t]; ly = cy[t]; }
            cx[t] = lx;
            cy[t] = ly;
            if (fl & 2) { acc = __fadd_rn(acc, len[t]); ++n; }
        }
        float scale = 1.f;
        if (n > 0) {
            const float mean = __fdiv_rn(acc, (float)n);
            if (mean > 0.f) scale = mean;
        }
        s_scale = scale;
    }
    __syncthreads();
    const float scale = s_scale;
    float *o = out + clip;
    for (int e = threadIdx.x; e < T * V; e += blockDim.x) {
        const int t = e / V;
        const float x = k[(size_t)e * 3], y = k[(size_t)e * 3 + 1], s = k[(size_t)e * 3 + 2];
        const bool keep = s >= thr;
        o[(size_t)e * 3] = keep ? __fdiv_rn(__fsub_rn(x, cx[t]), scale) : 0.f;
        o[(size_t)e * 3 + 1] = keep ? __fdiv_rn(__fsub_rn(y, cy[t]), scale) : 0.f;
        o[(size_t)e * 3 + 2] = keep ? s : 0.f;
    }
}

}  // namespace

int normalize_pose_launch(Ctx *ctx, const float *kp, float *out, int B, int T, int V, float min_score,
                          cudaStream_t st) {
    const size_t smem = (size_t)T * 13 + 16;
    if (smem > 227 * 1024) {
        set_error("normalize_pose: T = %d needs %zu bytes of shared memory (limit 227 KB)", T, smem);
        return GS_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        GS_CUDA(cudaFuncSetAttribute(normalize_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        LaunchScope ls(ctx, K_POSE, st, 8.0 * B * T * V, 24.0 * B * T * V);
        normalize_pose_kernel<<<B, T * V >= 4096 ? 1024 : 256, smem, st>>>(kp, out, T, V, min_score);
    }
    GS_KERNEL_CHECK();
    return GS_OK;
}

}  // namespace gs

// bf16 throughput path (tcgen05 / TMA) — placeholder until the kernels land.
#include "segment_common.cuh"

namespace gs {

int bf16_path_create(Ctx *ctx) {
    (void)ctx;
    return GS_OK;
}

void bf16_path_destroy(Ctx *ctx) { (void)ctx; }

int segment_bf16_forward(Ctx *, const float *, float *, uint8_t *, int, int, int, float *, cudaStream_t) {
    set_error("bf16 path not built yet");
    return GS_ERR_UNSUPPORTED;
}

}  // namespace gs

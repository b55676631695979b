// warp-cooperative meanclip instantiations, 8 lanes per pixel
#include "stack_meanclip_coop.cuh"

namespace apgpu_stack {

int stack_dispatch_meanclip_coop_p8(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    COOP_CASE(40, 256, 8) COOP_CASE(50, 320, 8) COOP_CASE(64, 400, 8)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

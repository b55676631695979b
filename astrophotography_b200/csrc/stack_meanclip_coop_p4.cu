// warp-cooperative meanclip instantiations, 4 lanes per pixel
#include "stack_meanclip_coop.cuh"

namespace apgpu_stack {

int stack_dispatch_meanclip_coop_p4(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    COOP_CASE(40, 128, 4) COOP_CASE(50, 160, 4) COOP_CASE(64, 200, 4) COOP_CASE(80, 256, 4)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

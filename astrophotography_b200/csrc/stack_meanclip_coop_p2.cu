// warp-cooperative meanclip instantiations, 2 lanes per pixel
#include "stack_meanclip_coop.cuh"

namespace apgpu_stack {

int stack_dispatch_meanclip_coop_p2(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    COOP_CASE(64, 100, 2) COOP_CASE(80, 128, 2)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

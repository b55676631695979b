// lane-split meanclip instantiations, 8 lanes per pixel
#include "stack_meanclip_split.cuh"

namespace apgpu_stack {

int stack_dispatch_meanclip_split_p8(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    SPLIT_CASE(80, 512, 8) SPLIT_CASE(100, 640, 8) SPLIT_CASE(128, 800, 8)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

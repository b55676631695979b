// lane-split meanclip instantiations, 4 lanes per pixel
#include "stack_meanclip_split.cuh"

namespace apgpu_stack {

int stack_dispatch_meanclip_split_p4(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    if (a.N > 160 && a.N <= 200) return launch_meanclip_split<50, 160, 4>(frames, a, st, done_pix);
    SPLIT_CASE(64, 200, 4) SPLIT_CASE(80, 256, 4) SPLIT_CASE(100, 320, 4) SPLIT_CASE(128, 400, 4)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

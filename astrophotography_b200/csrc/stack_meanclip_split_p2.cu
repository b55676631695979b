// lane-split meanclip instantiations, 2 lanes per pixel
#include "stack_meanclip_split.cuh"

namespace apgpu_stack {

int stack_dispatch_meanclip_split_p2(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    if (a.N > 80 && a.N <= 100) return launch_meanclip_split<50, 80, 2>(frames, a, st, done_pix);
    SPLIT_CASE(64, 100, 2) SPLIT_CASE(80, 128, 2) SPLIT_CASE(100, 160, 2)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

// lane-cooperative MODE_MEDUNC instantiations, 4 lanes per pixel
#include "stack_median_coop.cuh"
#define MEDCOOP_MODE MODE_MEDUNC

namespace apgpu_stack {

int stack_dispatch_median_coop_medunc_p4(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    MEDCOOP_CASE(56, 200, 4) MEDCOOP_CASE(64, 224, 4)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

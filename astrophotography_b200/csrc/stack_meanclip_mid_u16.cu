// meanclip<NB, NLO> instantiations, part "mid", uint16_t frames (split so that nvcc compiles the buckets in parallel)
#include "stack_meanclip.cuh"
#define MC_T uint16_t

namespace apgpu_stack {

int stack_dispatch_meanclip_mid(int nb, const uint16_t* const* frames, const StackArgs& a, cudaStream_t st, int flags) {
    MC_CASE(80, 64) MC_CASE(100, 80) MC_CASE(128, 100)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

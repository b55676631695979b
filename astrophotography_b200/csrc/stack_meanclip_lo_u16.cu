// meanclip<NB, NLO> instantiations, part "lo", uint16_t frames (split so that nvcc compiles the buckets in parallel)
#include "stack_meanclip.cuh"
#define MC_T uint16_t

namespace apgpu_stack {

int stack_dispatch_meanclip_lo(int nb, const uint16_t* const* frames, const StackArgs& a, cudaStream_t st, int flags) {
    MC_CASE(8, 2) MC_CASE(16, 8) MC_CASE(24, 16) MC_CASE(32, 24) MC_CASE(48, 32) MC_CASE(64, 48)
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

// sorted<NB, NLO, MODE_MEDMAD1> instantiations, uint16_t frames, bucket part 0 (see stack_common.cuh)
#include "stack_sorted.cuh"

namespace apgpu_stack {
template int dispatch_sorted_part<MODE_MEDMAD1, uint16_t, 0>(int, const uint16_t* const*, const StackArgs&, cudaStream_t);
}  // namespace apgpu_stack

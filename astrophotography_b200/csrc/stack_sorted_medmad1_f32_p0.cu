// sorted<NB, NLO, MODE_MEDMAD1> instantiations, float frames, bucket part 0 (see stack_common.cuh)
#include "stack_sorted.cuh"

namespace apgpu_stack {
template int dispatch_sorted_part<MODE_MEDMAD1, float, 0>(int, const float* const*, const StackArgs&, cudaStream_t);
}  // namespace apgpu_stack

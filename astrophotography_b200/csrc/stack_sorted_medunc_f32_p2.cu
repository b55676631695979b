// sorted<NB, NLO, MODE_MEDUNC> instantiations, float frames, bucket part 2 (see stack_common.cuh)
#include "stack_sorted.cuh"

namespace apgpu_stack {
template int dispatch_sorted_part<MODE_MEDUNC, float, 2>(int, const float* const*, const StackArgs&, cudaStream_t);
}  // namespace apgpu_stack

// sorted<NB, NLO, MODE_MED> instantiations, float frames, bucket part 1 (see stack_common.cuh)
#include "stack_sorted.cuh"

namespace apgpu_stack {
template int dispatch_sorted_part<MODE_MED, float, 1>(int, const float* const*, const StackArgs&, cudaStream_t);
}  // namespace apgpu_stack

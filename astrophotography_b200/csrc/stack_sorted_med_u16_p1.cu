// sorted<NB, NLO, MODE_MED> instantiations, uint16_t frames, bucket part 1 (see stack_common.cuh)
#include "stack_sorted.cuh"

namespace apgpu_stack {
template int dispatch_sorted_part<MODE_MED, uint16_t, 1>(int, const uint16_t* const*, const StackArgs&, cudaStream_t);
}  // namespace apgpu_stack

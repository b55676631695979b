// sorted<NB, NLO, MODE_MED> instantiations
#include "stack_sorted.cuh"

namespace apgpu_stack {

int stack_dispatch_sorted_med(int nb, const float* const* frames, const StackArgs& a, cudaStream_t st) {
    return dispatch_sorted<MODE_MED>(nb, frames, a, st);
}

}  // namespace apgpu_stack

// sorted<NB, NLO, MODE_MEDMAD1> instantiations
#include "stack_sorted.cuh"

namespace apgpu_stack {

int stack_dispatch_sorted_medmad1(int nb, const float* const* frames, const StackArgs& a, cudaStream_t st) {
    return dispatch_sorted<MODE_MEDMAD1>(nb, frames, a, st);
}

}  // namespace apgpu_stack

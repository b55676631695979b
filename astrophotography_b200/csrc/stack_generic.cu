// Frame-stack reducer: the generic float64 kernel (every parameter combination).  See stack_common.cuh.
#include "stack_common.cuh"

namespace apgpu_stack {

template <int CAP, typename T>
__global__ void __launch_bounds__(TPB)
stack_generic_kernel(const __grid_constant__ FramePtrs<CAP, T> fp, const __grid_constant__ StackArgs a) {
    int64_t p = a.pix0 + (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (p >= a.pix0 + a.npix) return;
    generic_pixel<CAP, FramePtrs<CAP, T>>(fp, a, p);
}

template <int CAP, typename T>
int launch_generic(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<CAP, T> fp;
    for (int i = 0; i < CAP; ++i) fp.p[i] = i < a.N ? frames[i] : nullptr;
    int64_t blocks = (a.npix + TPB - 1) / TPB;
    stack_generic_kernel<CAP, T><<<(unsigned)blocks, TPB, 0, st>>>(fp, a);
    APGPU_LAUNCH_CHECK("stack_generic_kernel");
    return APGPU_OK;
}

template <typename T>
int launch_generic_any(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    if (a.N <= 32) return launch_generic<32, T>(frames, a, st);
    if (a.N <= 128) return launch_generic<128, T>(frames, a, st);
    return launch_generic<1024, T>(frames, a, st);
}

int stack_launch_generic(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    return launch_generic_any<float>(frames, a, st);
}
int stack_launch_generic(const uint16_t* const* frames, const StackArgs& a, cudaStream_t st) {
    return launch_generic_any<uint16_t>(frames, a, st);
}

}  // namespace apgpu_stack

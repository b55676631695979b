// Image arithmetic (sm_100a): ADD / SUB / MUL / DIV of a float32 image with a scalar or a second image.
//
// Reference: ApImArith.process_files, AstroPhotography/core/ApImArith.py:321-333 -- np.add / np.subtract /
// np.multiply / np.divide (data1, data2, out=result) with result of data1's dtype.  For a float32 image numpy
// computes: with a Python-float scalar in float32 (the scalar is a weak type: it is rounded to float32 first), with
// a float32 image in float32, with a float64 image in float64 followed by a cast of the result to float32.
// Purely HBM-bound: 128-bit streaming accesses, every operation an explicitly rounded intrinsic.
#include "apgpu_common.cuh"

namespace {

constexpr int IA_THREADS = 256;

template <typename T> __device__ __forceinline__ T ia_op(int op, T a, T b);
template <> __device__ __forceinline__ float ia_op<float>(int op, float a, float b) {
    switch (op) {
        case APGPU_OP_ADD: return __fadd_rn(a, b);
        case APGPU_OP_SUB: return __fsub_rn(a, b);
        case APGPU_OP_MUL: return __fmul_rn(a, b);
        default: return __fdiv_rn(a, b);
    }
}
template <> __device__ __forceinline__ double ia_op<double>(int op, double a, double b) {
    switch (op) {
        case APGPU_OP_ADD: return __dadd_rn(a, b);
        case APGPU_OP_SUB: return __dsub_rn(a, b);
        case APGPU_OP_MUL: return __dmul_rn(a, b);
        default: return __ddiv_rn(a, b);
    }
}

// b_kind: 0 scalar, 1 float32 image, 2 float64 image
__global__ void __launch_bounds__(IA_THREADS)
imarith_kernel(const float* __restrict__ a, const void* __restrict__ b, int b_kind, float scalar, int op,
               float* __restrict__ out, int64_t npix, bool vec_ok) {
    const int64_t nvec = vec_ok ? npix / 4 : 0;
    const int64_t stride = (int64_t)gridDim.x * IA_THREADS;
    const int64_t t = (int64_t)blockIdx.x * IA_THREADS + threadIdx.x;
    for (int64_t i = t; i < nvec; i += stride) {
        const float4 x = ld_stream(reinterpret_cast<const float4*>(a) + i);
        float4 o;
        if (b_kind == 0) {
            o.x = ia_op(op, x.x, scalar); o.y = ia_op(op, x.y, scalar); o.z = ia_op(op, x.z, scalar); o.w = ia_op(op, x.w, scalar);
        } else if (b_kind == 1) {
            const float4 y = ld_stream(reinterpret_cast<const float4*>(b) + i);
            o.x = ia_op(op, x.x, y.x); o.y = ia_op(op, x.y, y.y); o.z = ia_op(op, x.z, y.z); o.w = ia_op(op, x.w, y.w);
        } else {
            const double2 y0 = __ldcs(reinterpret_cast<const double2*>(b) + 2 * i);
            const double2 y1 = __ldcs(reinterpret_cast<const double2*>(b) + 2 * i + 1);
            o.x = (float)ia_op(op, (double)x.x, y0.x); o.y = (float)ia_op(op, (double)x.y, y0.y);
            o.z = (float)ia_op(op, (double)x.z, y1.x); o.w = (float)ia_op(op, (double)x.w, y1.y);
        }
        st_stream(reinterpret_cast<float4*>(out) + i, o);
    }
    for (int64_t i = nvec * 4 + t; i < npix; i += stride) {
        const float x = a[i];
        if (b_kind == 0) out[i] = ia_op(op, x, scalar);
        else if (b_kind == 1) out[i] = ia_op(op, x, reinterpret_cast<const float*>(b)[i]);
        else out[i] = (float)ia_op(op, (double)x, reinterpret_cast<const double*>(b)[i]);
    }
}

}  // namespace

extern "C" int apgpu_imarith_f32(const float* a, const void* b, int b_kind, double scalar, int op,
                                 float* out, int64_t npix, apgpu_stream_t stream) {
    APGPU_REQUIRE(a && out, "imarith: null image pointer");
    APGPU_REQUIRE(b_kind >= 0 && b_kind <= 2 && (b_kind == 0 || b), "imarith: bad second operand (kind %d)", b_kind);
    APGPU_REQUIRE(op >= APGPU_OP_ADD && op <= APGPU_OP_DIV, "imarith: bad operation %d", op);
    APGPU_REQUIRE(npix >= 0, "imarith: bad npix");
    if (npix == 0) return APGPU_OK;
    const bool vec_ok = apgpu_aligned(a, 16) && apgpu_aligned(out, 16) && (b_kind == 0 || apgpu_aligned(b, 16));
    int64_t blocks = (npix / 4 + IA_THREADS - 1) / IA_THREADS;
    const int64_t cap = (int64_t)APGPU_NUM_SMS * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    imarith_kernel<<<(unsigned)blocks, IA_THREADS, 0, (cudaStream_t)stream>>>(a, b, b_kind, (float)scalar, op, out, npix, vec_ok);
    APGPU_LAUNCH_CHECK("imarith_kernel");
    return APGPU_OK;
}

// Bad-pixel neighbourhood-median repair (sm_100a).
//
// Reference: ApFixBadPixels.fix_bad_pixels, AstroPhotography/core/ApFixBadPixels.py
// :334-443 -- a Python loop over the bad pixels, each replaced by np.median of
// the good pixels of the ORIGINAL data inside the (2*dp+1)^2 window clipped to
// the image (:383-394), provided at least min_valid=4 good ones exist (:397).
//
// Here: a dense copy kernel at copy bandwidth (4 B read + 4 B write per pixel, 128-bit
// streaming accesses) followed by a persistent mask-scan kernel (1 B per pixel, eight
// 16-byte loads in flight per thread) in which only the warps that meet bad pixels work:
// they repair them one after the other COOPERATIVELY -- the 32 lanes fetch the <= 24 donors
// of the 5x5 window in parallel, sort them with a bitonic network of warp shuffles (+inf
// for non-donors) and read back the middle order statistics (badpix_common.cuh).  Donors
// come from the input image only, so every pixel is independent: no halo exchange.
//
// np.median semantics kept bit-exact: odd count -> middle value; even count ->
// float32(a + b) / 2 (numpy takes the mean of the two middle float32 values in
// float32); any NaN donor -> NaN.
#include "badpix_common.cuh"

namespace {
using namespace apgpu_badpix;

// dense pass: the produced rows of the band at copy bandwidth (128-bit streaming accesses, four independent
// vectors in flight per thread)
constexpr int CP_UNROLL = 4;

__global__ void __launch_bounds__(BP_THREADS)
badpix_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t nvec) {
    const int64_t base = (int64_t)blockIdx.x * (BP_THREADS * CP_UNROLL) + threadIdx.x;
    float4 v[CP_UNROLL];
#pragma unroll
    for (int u = 0; u < CP_UNROLL; ++u) {
        const int64_t i = base + (int64_t)u * BP_THREADS;
        if (i < nvec) v[u] = ld_stream(reinterpret_cast<const float4*>(src) + i);
    }
#pragma unroll
    for (int u = 0; u < CP_UNROLL; ++u) {
        const int64_t i = base + (int64_t)u * BP_THREADS;
        if (i < nvec) st_stream(reinterpret_cast<float4*>(dst) + i, v[u]);
    }
}

__global__ void __launch_bounds__(BP_THREADS)
badpix_copy_scalar_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t i0, int64_t n) {
    const int64_t i = i0 + (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

template <typename MaskT>
int launch_bp(const float* data, const MaskT* mask, int64_t H, int64_t W, int64_t band_row0,
              int64_t row0, int64_t nrows, int dp, int min_valid, float* out, int64_t* counts,
              cudaStream_t st) {
    // 1. every produced pixel, unrepaired; 2. the mask scan overwrites the repaired ones (donors come from `data`)
    const float* first = data + (row0 - band_row0) * W;
    const int64_t n = nrows * W;
    const int64_t nvec = (apgpu_aligned(first, 16) && apgpu_aligned(out, 16)) ? n / 4 : 0;
    if (nvec > 0) {
        const int64_t per_block = BP_THREADS * CP_UNROLL;
        badpix_copy_kernel<<<(unsigned)((nvec + per_block - 1) / per_block), BP_THREADS, 0, st>>>(first, out, nvec);
        APGPU_LAUNCH_CHECK("badpix_copy_kernel");
    }
    if (nvec * 4 < n) {
        const int64_t rem = n - nvec * 4;
        badpix_copy_scalar_kernel<<<(unsigned)((rem + BP_THREADS - 1) / BP_THREADS), BP_THREADS, 0, st>>>(first, out, nvec * 4, n);
        APGPU_LAUNCH_CHECK("badpix_copy_scalar_kernel");
    }
    return launch_repair_scan<MaskT>(PlainImage{data}, mask, H, W, band_row0, row0, nrows, dp, min_valid, out, false,
                                     counts, st);
}

}  // namespace

extern "C" int apgpu_fix_badpix_f32(const float* data, const void* mask, int mask_dtype,
                                    int64_t H_image, int64_t W, int64_t band_row0, int64_t band_rows,
                                    int64_t row0, int64_t nrows, int deltapix, int min_valid,
                                    float* out, int64_t* counts, apgpu_stream_t stream) {
    APGPU_REQUIRE(data && mask && out && counts, "fix_badpix: null pointer");
    APGPU_REQUIRE(H_image > 0 && W > 0, "fix_badpix: bad image shape %lld x %lld", (long long)H_image, (long long)W);
    APGPU_REQUIRE(deltapix >= 1 && deltapix <= 7, "fix_badpix: deltapix %d outside supported range 1..7", deltapix);
    APGPU_REQUIRE(min_valid >= 1, "fix_badpix: min_valid must be >= 1");
    APGPU_REQUIRE(band_row0 >= 0 && band_rows > 0 && band_row0 + band_rows <= H_image, "fix_badpix: bad band");
    APGPU_REQUIRE(nrows >= 0 && row0 >= band_row0 && row0 + nrows <= band_row0 + band_rows, "fix_badpix: rows outside band");
    // the band must hold every halo row that exists in the image
    int64_t need_lo = row0 - deltapix < 0 ? 0 : row0 - deltapix;
    int64_t need_hi = row0 + nrows + deltapix > H_image ? H_image : row0 + nrows + deltapix;
    APGPU_REQUIRE(nrows == 0 || (band_row0 <= need_lo && band_row0 + band_rows >= need_hi),
                  "fix_badpix: band [%lld,%lld) lacks the %d-row halo of rows [%lld,%lld)",
                  (long long)band_row0, (long long)(band_row0 + band_rows), deltapix,
                  (long long)row0, (long long)(row0 + nrows));
    if (nrows == 0) return APGPU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (mask_dtype) {
        case APGPU_MASK_U8:  return launch_bp<uint8_t>(data, (const uint8_t*)mask, H_image, W, band_row0, row0, nrows, deltapix, min_valid, out, counts, st);
        case APGPU_MASK_I16: return launch_bp<int16_t>(data, (const int16_t*)mask, H_image, W, band_row0, row0, nrows, deltapix, min_valid, out, counts, st);
        case APGPU_MASK_I32: return launch_bp<int32_t>(data, (const int32_t*)mask, H_image, W, band_row0, row0, nrows, deltapix, min_valid, out, counts, st);
        case APGPU_MASK_F32: return launch_bp<float>(data, (const float*)mask, H_image, W, band_row0, row0, nrows, deltapix, min_valid, out, counts, st);
        case APGPU_MASK_F64: return launch_bp<double>(data, (const double*)mask, H_image, W, band_row0, row0, nrows, deltapix, min_valid, out, counts, st);
        default: break;
    }
    apgpu_set_error("fix_badpix: unknown mask_dtype %d", mask_dtype);
    return APGPU_ERR_ARG;
}

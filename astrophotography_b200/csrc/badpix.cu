// Bad-pixel neighbourhood-median repair (sm_100a).
//
// Reference: ApFixBadPixels.fix_bad_pixels, AstroPhotography/core/ApFixBadPixels.py
// :334-443 -- a Python loop over the bad pixels, each replaced by np.median of
// the good pixels of the ORIGINAL data inside the (2*dp+1)^2 window clipped to
// the image (:383-394), provided at least min_valid=4 good ones exist (:397).
//
// Here: one dense pass, HBM-bound (4 B read + mask + 4 B write per pixel).  Each
// thread copies four adjacent pixels with 128-bit accesses.  When a warp meets
// bad pixels it repairs them one after the other COOPERATIVELY: the 32 lanes
// fetch the <= 24 donors of the 5x5 window in parallel (neighbour rows are L1/L2
// hits: the warp has just streamed them), sort them with a bitonic network of
// warp shuffles (+inf for non-donors) and read back the middle order
// statistics.  Donors come from the input image only, so every pixel is
// independent: no halo exchange, no second pass.
//
// np.median semantics kept bit-exact: odd count -> middle value; even count ->
// float32(a + b) / 2 (numpy takes the mean of the two middle float32 values in
// float32); any NaN donor -> NaN.
#include "badpix_common.cuh"

namespace {
using namespace apgpu_badpix;

// grid.x covers the 4-pixel groups of one row (whole warps), grid.y the rows of the band.
template <int DP, typename MaskT>
__global__ void __launch_bounds__(BP_THREADS)
fix_badpix_kernel(const float* __restrict__ data, const MaskT* __restrict__ mask,
                  int64_t H, int64_t W, int64_t band_row0, int64_t row0, int64_t nrows,
                  int dp, int min_valid, float* __restrict__ out,
                  unsigned long long* __restrict__ counts, bool vec_ok) {
    const int64_t groups_per_row = (W + 3) / 4;
    const int64_t g = (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // whole warps stay in the loop together (a warp whose first group is past the row leaves as one)
    const bool warp_active = (g - lane) < groups_per_row;
    unsigned nbad = 0, nfix = 0;
    for (int64_t rel = blockIdx.y; warp_active && rel < nrows; rel += gridDim.y) {
        const int64_t r = row0 + rel;
        const int64_t c0 = g * 4;
        const int64_t in_base = (r - band_row0) * W + c0;
        const int64_t out_base = rel * W + c0;
        float px[4] = {0.f, 0.f, 0.f, 0.f};
        bool bad[4] = {false, false, false, false};
        const int nvalid = g < groups_per_row ? (int)((W - c0) < 4 ? (W - c0) : 4) : 0;
        if (vec_ok && nvalid == 4) {
            const float4 d = ld_stream(reinterpret_cast<const float4*>(data + in_base));
            px[0] = d.x; px[1] = d.y; px[2] = d.z; px[3] = d.w;
            if (sizeof(MaskT) == 1) {
                const uchar4 m4 = __ldcs(reinterpret_cast<const uchar4*>(mask + in_base));
                bad[0] = m4.x != 0; bad[1] = m4.y != 0; bad[2] = m4.z != 0; bad[3] = m4.w != 0;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) bad[k] = mask_bad(mask, in_base + k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < nvalid) { px[k] = data[in_base + k]; bad[k] = mask_bad(mask, in_base + k); }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (DP == 1 || DP == 2) {
                unsigned todo = __ballot_sync(0xffffffffu, bad[k]);
                nbad += bad[k] ? 1u : 0u;
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int64_t cc = __shfl_sync(0xffffffffu, c0, src) + k;
                    float res = 0.f;
                    const bool ok = repair_warp<(DP == 1 || DP == 2) ? DP : 1, MaskT>(
                        PlainImage{data}, mask, H, W, band_row0, r, cc, min_valid, res);
                    if (ok && lane == src) { px[k] = res; ++nfix; }
                }
            } else if (bad[k]) {
                ++nbad;
                float res;
                if (repair_any<MaskT>(PlainImage{data}, mask, H, W, band_row0, r, c0 + k, dp, min_valid, res)) { px[k] = res; ++nfix; }
            }
        }
        if (vec_ok && nvalid == 4) {
            st_stream(reinterpret_cast<float4*>(out + out_base), make_float4(px[0], px[1], px[2], px[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < nvalid) out[out_base + k] = px[k];
        }
    }
    // warp-aggregated counters
    for (int off = 16; off > 0; off >>= 1) {
        nbad += __shfl_down_sync(0xffffffffu, nbad, off);
        nfix += __shfl_down_sync(0xffffffffu, nfix, off);
    }
    if (lane == 0 && nbad) {
        atomicAdd(&counts[0], (unsigned long long)nbad);
        if (nfix) atomicAdd(&counts[1], (unsigned long long)nfix);
    }
}

template <typename MaskT>
int launch_bp(const float* data, const MaskT* mask, int64_t H, int64_t W, int64_t band_row0,
              int64_t row0, int64_t nrows, int dp, int min_valid, float* out, int64_t* counts,
              cudaStream_t st) {
    int64_t groups = (W + 3) / 4;
    dim3 grid((unsigned)((groups + BP_THREADS - 1) / BP_THREADS),
              (unsigned)(nrows < 65535 ? nrows : 65535));
    bool vec_ok = (W % 4 == 0) && apgpu_aligned(data, 16) && apgpu_aligned(out, 16) && apgpu_aligned(mask, 4);
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
    if (dp == 1)
        fix_badpix_kernel<1, MaskT><<<grid, BP_THREADS, 0, st>>>(data, mask, H, W, band_row0, row0, nrows, dp, min_valid, out, c, vec_ok);
    else if (dp == 2)
        fix_badpix_kernel<2, MaskT><<<grid, BP_THREADS, 0, st>>>(data, mask, H, W, band_row0, row0, nrows, dp, min_valid, out, c, vec_ok);
    else
        fix_badpix_kernel<0, MaskT><<<grid, BP_THREADS, 0, st>>>(data, mask, H, W, band_row0, row0, nrows, dp, min_valid, out, c, vec_ok);
    APGPU_LAUNCH_CHECK("fix_badpix_kernel");
    return APGPU_OK;
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

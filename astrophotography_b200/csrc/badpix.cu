// Bad-pixel neighbourhood-median repair (sm_100a).
//
// Reference: ApFixBadPixels.fix_bad_pixels, AstroPhotography/core/ApFixBadPixels.py
// :334-443 -- a Python loop over the bad pixels, each replaced by np.median of
// the good pixels of the ORIGINAL data inside the (2*dp+1)^2 window clipped to
// the image (:383-394), provided at least min_valid=4 good ones exist (:397).
//
// Here: one dense pass, HBM-bound (4 B read + mask + 4 B write per pixel).  Each
// thread copies four adjacent pixels with 128-bit accesses; a lane that meets a
// bad pixel gathers the <= 24 donors of the 5x5 window (neighbour rows are L1/L2
// hits: the warp has just streamed them) into registers, sorts them with a
// Batcher network (+inf padding for non-donors) and picks the middle order
// statistics.  Donors come from the input image only, so every pixel is
// independent: no halo exchange, no second pass.
//
// np.median semantics kept bit-exact: odd count -> middle value; even count ->
// float32(a + b) / 2 (numpy takes the mean of the two middle float32 values in
// float32); any NaN donor -> NaN.
#include <math.h>

#include "apgpu_common.cuh"
#include "sort_networks.inc"

namespace {

template <typename MaskT>
__device__ __forceinline__ bool mask_bad(const MaskT* m, int64_t i) { return m[i] != (MaskT)0; }

#define CE_V(i, j) { float lo_ = fminf(v[i], v[j]); float hi_ = fmaxf(v[i], v[j]); v[i] = lo_; v[j] = hi_; }

template <int NV> __device__ __forceinline__ void sort_donors(float (&v)[NV]);
template <> __device__ __forceinline__ void sort_donors<8>(float (&v)[8]) { APGPU_SORTNET_8(CE_V) }
template <> __device__ __forceinline__ void sort_donors<24>(float (&v)[24]) { APGPU_SORTNET_24(CE_V) }

template <int NV>
__device__ __forceinline__ float pick(const float (&v)[NV], int k) {
    float r = v[0];
#pragma unroll
    for (int i = 1; i < NV; ++i) r = (i == k) ? v[i] : r;
    return r;
}

// Median of the good neighbours of bad pixel (r, c); returns false when fewer
// than min_valid good neighbours exist.  DP in {1, 2}: register network.
template <int DP, typename MaskT>
__device__ __noinline__ bool repair_small(const float* __restrict__ data, const MaskT* __restrict__ mask,
                                          int64_t H, int64_t W, int64_t band_row0,
                                          int64_t r, int64_t c, int min_valid, float& result) {
    constexpr int NV = (2 * DP + 1) * (2 * DP + 1) - 1;
    float v[NV];
    int ngood = 0;
    bool anynan = false;
    int k = 0;
#pragma unroll
    for (int dr = -DP; dr <= DP; ++dr) {
#pragma unroll
        for (int dc = -DP; dc <= DP; ++dc) {
            if (dr == 0 && dc == 0) continue;       // the centre is bad: never a donor
            int64_t rr = r + dr, cc = c + dc;
            float val = INFINITY;
            if (rr >= 0 && rr < H && cc >= 0 && cc < W) {
                int64_t idx = (rr - band_row0) * W + cc;
                if (!mask_bad(mask, idx)) {
                    float x = data[idx];
                    ++ngood;
                    if (x != x) anynan = true; else val = x;
                }
            }
            v[k++] = val;
        }
    }
    if (ngood < min_valid) return false;
    if (anynan) { result = NAN; return true; }
    sort_donors<NV>(v);
    float a = pick<NV>(v, (ngood - 1) >> 1);
    float b = pick<NV>(v, ngood >> 1);
    result = (ngood & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
    return true;
}

// Any deltapix up to 7: donors in local memory, insertion sort.
template <typename MaskT>
__device__ __noinline__ bool repair_any(const float* __restrict__ data, const MaskT* __restrict__ mask,
                                        int64_t H, int64_t W, int64_t band_row0,
                                        int64_t r, int64_t c, int dp, int min_valid, float& result) {
    float v[224];
    int ngood = 0;          // good neighbours (donors), NaN ones included
    int nv = 0;             // non-NaN donors held sorted in v[0..nv)
    bool anynan = false;
    for (int dr = -dp; dr <= dp; ++dr) {
        for (int dc = -dp; dc <= dp; ++dc) {
            if (dr == 0 && dc == 0) continue;
            int64_t rr = r + dr, cc = c + dc;
            if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
            int64_t idx = (rr - band_row0) * W + cc;
            if (mask_bad(mask, idx)) continue;
            float x = data[idx];
            ++ngood;
            if (x != x) { anynan = true; continue; }
            int m = nv++;
            while (m > 0 && v[m - 1] > x) { v[m] = v[m - 1]; --m; }
            v[m] = x;
        }
    }
    if (ngood < min_valid) return false;
    if (anynan) { result = NAN; return true; }
    float a = v[(ngood - 1) >> 1], b = v[ngood >> 1];
    result = (ngood & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
    return true;
}

constexpr int BP_THREADS = 256;

// grid.x covers the 4-pixel groups of one row, grid.y the rows of the band.
template <int DP, typename MaskT>
__global__ void __launch_bounds__(BP_THREADS)
fix_badpix_kernel(const float* __restrict__ data, const MaskT* __restrict__ mask,
                  int64_t H, int64_t W, int64_t band_row0, int64_t row0, int64_t nrows,
                  int dp, int min_valid, float* __restrict__ out,
                  unsigned long long* __restrict__ counts, bool vec_ok) {
    int64_t groups_per_row = (W + 3) / 4;
    int64_t g = (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    unsigned nbad = 0, nfix = 0;
    for (int64_t rel = blockIdx.y; rel < nrows; rel += gridDim.y) {
        if (g >= groups_per_row) break;
        int64_t r = row0 + rel;
        int64_t c0 = g * 4;
        int64_t in_base = (r - band_row0) * W + c0;
        int64_t out_base = rel * W + c0;
        float px[4];
        bool bad[4];
        int nvalid = (int)((W - c0) < 4 ? (W - c0) : 4);
        if (vec_ok && nvalid == 4) {
            float4 d = ld_stream(reinterpret_cast<const float4*>(data + in_base));
            px[0] = d.x; px[1] = d.y; px[2] = d.z; px[3] = d.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) px[k] = (k < nvalid) ? data[in_base + k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) bad[k] = (k < nvalid) && mask_bad(mask, in_base + k);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (bad[k]) {
                ++nbad;
                float res;
                bool ok;
                if (DP == 1 || DP == 2)
                    ok = repair_small<(DP == 1 || DP == 2) ? DP : 1, MaskT>(
                        data, mask, H, W, band_row0, r, c0 + k, min_valid, res);
                else
                    ok = repair_any<MaskT>(data, mask, H, W, band_row0, r, c0 + k, dp, min_valid, res);
                if (ok) { px[k] = res; ++nfix; }
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
    if ((threadIdx.x & 31) == 0 && nbad) {
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
    bool vec_ok = (W % 4 == 0) && apgpu_aligned(data, 16) && apgpu_aligned(out, 16);
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

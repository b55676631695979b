// Bad-pixel repair: device routines shared by the stand-alone kernel (badpix.cu) and the fused
// calibrate + repair kernel (calibrate_repair.cu).  `Src` supplies the ORIGINAL image the donors come from:
// a plain float32 image, or the calibrated value recomputed from raw / bias / dark / flat on the fly.
#pragma once
#include <math.h>

#include "apgpu_common.cuh"

namespace apgpu_badpix {

struct PlainImage {
    const float* data;
    __device__ __forceinline__ float at(int64_t i) const { return data[i]; }
};

template <typename MaskT>
__device__ __forceinline__ bool mask_bad(const MaskT* m, int64_t i) { return m[i] != (MaskT)0; }

// Any deltapix up to 7: donors in local memory, insertion sort.
template <typename MaskT, typename Src>
__device__ __noinline__ bool repair_any(const Src& data, const MaskT* __restrict__ mask,
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
            float x = data.at(idx);
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

// Warp-cooperative median of the good neighbours of bad pixel (r, c), DP in {1, 2}:
// lane l fetches neighbour l of the (2*DP+1)^2 window, the 32 lanes sort their
// values with a bitonic network of shuffles (+inf for non-donors), and the middle
// order statistics are read back by shuffle.  ~80 warp instructions per bad pixel
// instead of a ~700-instruction single-lane detour that idles the other 31 lanes.
template <int DP, typename MaskT, typename Src>
__device__ __forceinline__ bool repair_warp(const Src& data, const MaskT* __restrict__ mask,
                                            int64_t H, int64_t W, int64_t band_row0,
                                            int64_t r, int64_t c, int min_valid, float& result) {
    constexpr int WD = 2 * DP + 1;
    const int lane = threadIdx.x & 31;
    const int dr = lane / WD - DP, dc = lane % WD - DP;
    float v = INFINITY;
    bool good = false, isnan_ = false;
    if (lane < WD * WD && !(dr == 0 && dc == 0)) {
        const int64_t rr = r + dr, cc = c + dc;
        if (rr >= 0 && rr < H && cc >= 0 && cc < W) {
            const int64_t idx = (rr - band_row0) * W + cc;
            if (!mask_bad(mask, idx)) {
                const float x = data.at(idx);
                good = true;
                if (x != x) isnan_ = true; else v = x;
            }
        }
    }
    const int ngood = __popc(__ballot_sync(0xffffffffu, good));
    const bool anynan = __ballot_sync(0xffffffffu, isnan_) != 0u;
    if (ngood < min_valid) return false;
    if (anynan) { result = NAN; return true; }
    // bitonic sort, ascending across lanes
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = ((lane & k) == 0);            // ascending block
            const bool lower = ((lane & j) == 0);
            v = (up == lower) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    const float a = __shfl_sync(0xffffffffu, v, (ngood - 1) >> 1);
    const float b = __shfl_sync(0xffffffffu, v, ngood >> 1);
    result = (ngood & 1) ? a : __fmul_rn(__fadd_rn(a, b), 0.5f);
    return true;
}

}  // namespace apgpu_badpix

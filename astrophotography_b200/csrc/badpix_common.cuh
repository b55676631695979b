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
    __device__ __forceinline__ float4 vec4(int64_t i) const { return ld_stream(reinterpret_cast<const float4*>(data + i)); }
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
            // (mask and donor value are fetched together: one memory round trip per repair instead of two)
            const bool mb = mask_bad(mask, idx);
            const float x = data.at(idx);
            if (!mb) {
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


// ---------------------------------------------------------------------------
// Repair pass: scan the mask, repair where it is set, write ONLY the repaired pixels.
// ---------------------------------------------------------------------------
// The image itself is produced by a separate pure streaming kernel (a copy, or the calibration arithmetic) that
// runs at copy bandwidth and never sees the mask.  Measured on the way here (61 Mpix, 82 k bad pixels, round 2):
//   * repairs inside the streaming loop (round 1): 199 us -- 205 us even with an EMPTY mask, against 82 us for a
//     plain copy: the per-warp control flow and the short per-CTA work, not the repairs, held it at 32 % of the
//     DRAM throughput (the repairs themselves cost ~25 us);
//   * streaming CTAs and repair CTAs interleaved in one grid: 235 us (the latency-bound repair CTAs get a quarter
//     of the resident slots);
//   * a scan with one small CTA per row segment: 120 us for 61 MB of mask (19 k CTAs that each wait for one load).
//   * one global atomicAdd pair per warp for the two counters: ~1.2 ns each on ONE address -- 30 k of them were
//     half of the scan's time; the counters are summed per CTA in shared memory first.
//   * the repairs of one warp are serial (~3 us each: two dependent gathers from HBM + the shuffle network), so
//     what counts is how many warps are resident: 32 registers per thread (64 warps per SM) instead of 64.
// So the scan is a PERSISTENT grid of 8 CTAs per SM, two 16-byte mask loads in flight per thread (1024 pixels per
// warp and step), and only warps that find bad pixels do anything else: the 32 lanes repair them one after the
// other cooperatively (repair_warp) and write the repaired ones.
//   mask / src rows:  image row r lives at index (r - band_row0) * W   out rows: image row r at (r - row0) * W
__device__ __forceinline__ float store_order(float v, bool big_endian) {
    return big_endian ? __uint_as_float(__byte_perm(__float_as_uint(v), 0u, 0x0123)) : v;
}

constexpr int SCAN_VEC = 2;                       // 16-byte mask vectors in flight per thread (1024 pixels per warp and step)

// uint8 masks whose first produced pixel is 16-byte aligned: linear walk over the produced pixels
template <int DP, typename Src>
__global__ void __launch_bounds__(BP_THREADS, 8)
repair_scan_linear_kernel(const Src src, const uint8_t* __restrict__ mask, int64_t H, int64_t W, int64_t band_row0,
                          int64_t row0, int64_t nrows, int dp, int min_valid, float* __restrict__ out,
                          bool out_big_endian, unsigned long long* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t in_off = (row0 - band_row0) * W;
    const int64_t npx = nrows * W;
    const int64_t nvec = (npx + 15) / 16;                      // the last vector may be partial
    const int64_t warps = ((int64_t)gridDim.x * BP_THREADS) >> 5;
    const int64_t wid = ((int64_t)blockIdx.x * BP_THREADS + threadIdx.x) >> 5;
    unsigned nbad = 0, nfix = 0;
    for (int64_t v0 = wid * (32 * SCAN_VEC); v0 < nvec; v0 += warps * (32 * SCAN_VEC)) {
        uint4 m[SCAN_VEC];
#pragma unroll
        for (int q = 0; q < SCAN_VEC; ++q) {
            const int64_t v = v0 + q * 32 + lane;
            m[q] = make_uint4(0, 0, 0, 0);
            if (v * 16 + 16 <= npx) {
                m[q] = __ldcs(reinterpret_cast<const uint4*>(mask + in_off) + v);
            } else if (v < nvec) {                              // the image's last few pixels, byte by byte
                uint32_t wds[4] = {0, 0, 0, 0};
                for (int k = 0; v * 16 + k < npx; ++k) wds[k >> 2] |= (uint32_t)mask[in_off + v * 16 + k] << (8 * (k & 3));
                m[q] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
            }
        }
#pragma unroll
        for (int q = 0; q < SCAN_VEC; ++q) {
            const bool any = (m[q].x | m[q].y | m[q].z | m[q].w) != 0u;
            unsigned todo = __ballot_sync(0xffffffffu, any);
            if (!todo) continue;
            const uint32_t wds[4] = {m[q].x, m[q].y, m[q].z, m[q].w};
            uint32_t badbits = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    badbits |= ((wds[a] >> (8 * b)) & 0xffu) ? (1u << (4 * a + b)) : 0u;
            nbad += __popc(badbits);
            while (todo) {
                const int s = __ffs(todo) - 1;
                todo &= todo - 1;
                uint32_t bits = __shfl_sync(0xffffffffu, badbits, s);
                const int64_t j0 = (v0 + q * 32 + s) * 16;     // produced-pixel index of lane s's first pixel
                while (bits) {
                    const int k = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int64_t j = j0 + k;
                    const int64_t rel = j / W;
                    const int64_t c = j - rel * W;
                    float res = 0.f;
                    bool ok;
                    if (DP == 1 || DP == 2) {
                        ok = repair_warp<(DP == 1 || DP == 2) ? DP : 1, uint8_t>(src, mask, H, W, band_row0, row0 + rel, c, min_valid, res);
                    } else {
                        ok = false;
                        if (lane == s) ok = repair_any<uint8_t>(src, mask, H, W, band_row0, row0 + rel, c, dp, min_valid, res);
                    }
                    if (ok && lane == s) { out[j] = store_order(res, out_big_endian); ++nfix; }
                }
            }
        }
    }
    // per-CTA sums first: one pair of global atomics per CTA (same-address atomics serialise at ~1.2 ns each)
    __shared__ unsigned cta_bad, cta_fix;
    if (threadIdx.x == 0) { cta_bad = 0; cta_fix = 0; }
    __syncthreads();
    for (int off = 16; off > 0; off >>= 1) {
        nbad += __shfl_down_sync(0xffffffffu, nbad, off);
        nfix += __shfl_down_sync(0xffffffffu, nfix, off);
    }
    if (lane == 0 && nbad) {
        atomicAdd(&cta_bad, nbad);
        if (nfix) atomicAdd(&cta_fix, nfix);
    }
    __syncthreads();
    if (threadIdx.x == 0 && cta_bad) {
        atomicAdd(&counts[0], (unsigned long long)cta_bad);
        if (cta_fix) atomicAdd(&counts[1], (unsigned long long)cta_fix);
    }
}

// any mask type / any row length: one CTA per row segment, four mask elements per thread
template <int DP, typename MaskT, typename Src>
__global__ void __launch_bounds__(BP_THREADS)
repair_scan_rows_kernel(const Src src, const MaskT* __restrict__ mask, int64_t H, int64_t W, int64_t band_row0,
                        int64_t row0, int64_t nrows, int dp, int min_valid, float* __restrict__ out, bool out_big_endian,
                        unsigned long long* __restrict__ counts) {
    constexpr int PXT = 4;
    const int64_t groups_per_row = (W + PXT - 1) / PXT;
    const int64_t g = (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool warp_active = (g - lane) < groups_per_row;          // whole warps stay in the loop together
    unsigned nbad = 0, nfix = 0;
    for (int64_t rel = blockIdx.y; warp_active && rel < nrows; rel += gridDim.y) {
        const int64_t r = row0 + rel;
        const int64_t c0 = g * PXT;
        const int64_t in_base = (r - band_row0) * W + c0;
        const int64_t out_base = rel * W + c0;
        const int nvalid = g < groups_per_row ? (int)((W - c0) < PXT ? (W - c0) : PXT) : 0;
        uint32_t badbits = 0;
#pragma unroll
        for (int k = 0; k < PXT; ++k)
            if (k < nvalid && mask_bad(mask, in_base + k)) badbits |= 1u << k;
        if (!__any_sync(0xffffffffu, badbits != 0)) continue;
        nbad += __popc(badbits);
        for (int k = 0; k < PXT; ++k) {
            const bool bad = (badbits >> k) & 1u;
            if (DP == 1 || DP == 2) {
                unsigned todo = __ballot_sync(0xffffffffu, bad);
                while (todo) {
                    const int s = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int64_t cc = __shfl_sync(0xffffffffu, c0, s) + k;
                    float res = 0.f;
                    const bool ok = repair_warp<(DP == 1 || DP == 2) ? DP : 1, MaskT>(src, mask, H, W, band_row0, r, cc, min_valid, res);
                    if (ok && lane == s) { out[out_base + k] = store_order(res, out_big_endian); ++nfix; }
                }
            } else if (bad) {
                float res;
                if (repair_any<MaskT>(src, mask, H, W, band_row0, r, c0 + k, dp, min_valid, res)) {
                    out[out_base + k] = store_order(res, out_big_endian);
                    ++nfix;
                }
            }
        }
    }
    // per-CTA sums first: one pair of global atomics per CTA (same-address atomics serialise at ~1.2 ns each)
    __shared__ unsigned cta_bad, cta_fix;
    if (threadIdx.x == 0) { cta_bad = 0; cta_fix = 0; }
    __syncthreads();
    for (int off = 16; off > 0; off >>= 1) {
        nbad += __shfl_down_sync(0xffffffffu, nbad, off);
        nfix += __shfl_down_sync(0xffffffffu, nfix, off);
    }
    if (lane == 0 && nbad) {
        atomicAdd(&cta_bad, nbad);
        if (nfix) atomicAdd(&cta_fix, nfix);
    }
    __syncthreads();
    if (threadIdx.x == 0 && cta_bad) {
        atomicAdd(&counts[0], (unsigned long long)cta_bad);
        if (cta_fix) atomicAdd(&counts[1], (unsigned long long)cta_fix);
    }
}

template <typename MaskT, typename Src>
int launch_repair_scan(const Src& src, const MaskT* mask, int64_t H, int64_t W, int64_t band_row0, int64_t row0,
                       int64_t nrows, int dp, int min_valid, float* out, bool out_big_endian, int64_t* counts,
                       cudaStream_t st) {
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
    if constexpr (sizeof(MaskT) == 1) {
        const int64_t in_off = (row0 - band_row0) * W;
        if (apgpu_aligned(reinterpret_cast<const uint8_t*>(mask) + in_off, 16)) {
            const uint8_t* m8 = reinterpret_cast<const uint8_t*>(mask);
            const int64_t nvec = (nrows * W + 15) / 16;
            int64_t blocks = (nvec + (int64_t)BP_THREADS * SCAN_VEC - 1) / ((int64_t)BP_THREADS * SCAN_VEC);
            if (blocks > (int64_t)APGPU_NUM_SMS * 8) blocks = (int64_t)APGPU_NUM_SMS * 8;
            if (blocks < 1) blocks = 1;
#define APGPU_SCANL(DP_) repair_scan_linear_kernel<DP_, Src><<<(unsigned)blocks, BP_THREADS, 0, st>>>( \
        src, m8, H, W, band_row0, row0, nrows, dp, min_valid, out, out_big_endian, c)
            if (dp == 1) APGPU_SCANL(1); else if (dp == 2) APGPU_SCANL(2); else APGPU_SCANL(0);
#undef APGPU_SCANL
            APGPU_LAUNCH_CHECK("repair_scan_linear_kernel");
            return APGPU_OK;
        }
    }
    const int64_t groups = (W + 3) / 4;
    dim3 grid((unsigned)((groups + BP_THREADS - 1) / BP_THREADS), (unsigned)(nrows < 65535 ? nrows : 65535));
#define APGPU_SCANR(DP_) repair_scan_rows_kernel<DP_, MaskT, Src><<<grid, BP_THREADS, 0, st>>>( \
        src, mask, H, W, band_row0, row0, nrows, dp, min_valid, out, out_big_endian, c)
    if (dp == 1) APGPU_SCANR(1); else if (dp == 2) APGPU_SCANR(2); else APGPU_SCANR(0);
#undef APGPU_SCANR
    APGPU_LAUNCH_CHECK("repair_scan_rows_kernel");
    return APGPU_OK;
}

}  // namespace apgpu_badpix

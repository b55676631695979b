// Fused science-frame calibration + bad-pixel repair (sm_100a): one launch per frame for the batch driver.
//
// Reference: ApCalibrate.calibrate, AstroPhotography/core/ApCalibrate.py:439-479 -- the numpy calibration
// arithmetic (:439-474) immediately followed by ApFixBadPixels.fix_bad_pixels on its result (:478-479).  Run as
// two kernels that intermediate image costs 8 more bytes per pixel of HBM traffic (4 written, 4 read again); here
// every pixel is calibrated in registers and written once, and a bad pixel's donors -- the CALIBRATED values of
// its good neighbours -- are recomputed from raw / bias / dark / flat on the fly (<= 24 neighbours of ~1 % of the
// pixels: the rows are L1/L2 hits, the warp has just streamed them).  The arithmetic is the same explicitly
// rounded float32 sequence as calibrate.cu and the same warp-cooperative median as badpix.cu (badpix_common.cuh),
// so the output equals apgpu_calibrate_* followed by apgpu_fix_badpix_f32 bit for bit.
//
// The raw frame may be float32, host-order uint16, or the data unit of a BITPIX=16 / BZERO=32768 FITS file as it
// is on disk (big-endian int16 + 32768, decoded in the load); the output may be written as big-endian float32,
// i.e. as the data unit of the BITPIX=-32 FITS file the caller is about to write.
#include "badpix_common.cuh"

namespace {
using namespace apgpu_badpix;

struct CalParams {
    const void* raw;
    const float* bias;
    const float* dark;
    const float* nflat;      // may be null
    float ped, r;
    int raw_kind;            // APGPU_RAW_F32 / APGPU_RAW_U16 / APGPU_RAW_U16_FITS
    int has_ped, biased;
};

__device__ __forceinline__ float decode_raw(const CalParams& p, int64_t i) {
    float v;
    if (p.raw_kind == APGPU_RAW_F32) {
        v = reinterpret_cast<const float*>(p.raw)[i];
    } else {
        uint32_t u = reinterpret_cast<const uint16_t*>(p.raw)[i];
        if (p.raw_kind == APGPU_RAW_U16_FITS) u = (__byte_perm(u, 0u, 0x4401) ^ 0x8000u) & 0xffffu;
        v = (float)u;                                      // ApCalibrate.py:304-307 astype(float32)
    }
    return p.has_ped ? __fadd_rn(v, p.ped) : v;           // :322 ext_data += pedestal
}

__device__ __forceinline__ float cal1(const CalParams& p, float raw, float b, float d, float nf) {
    const float t = __fsub_rn(raw, b);                     // :439 raw - bias
    const float ds = p.biased ? __fsub_rn(d, b) : d;       // :442 dark - bias
    float o = __fsub_rn(t, __fmul_rn(p.r, ds));            // :450-451
    if (p.nflat) {
        const float q = __fdiv_rn(o, nf);                  // :462-464 np.where(nf != 0, o/nf, o)
        o = (nf != 0.0f) ? q : o;
    }
    return o;
}

// the calibrated image as a pixel source for the repair routines
struct CalibratedImage {
    CalParams p;
    __device__ __forceinline__ float at(int64_t i) const {
        return cal1(p, decode_raw(p, i), p.bias[i], p.dark[i], p.nflat ? p.nflat[i] : 1.f);
    }
};

__device__ __forceinline__ float to_big_endian(float v) {
    return __uint_as_float(__byte_perm(__float_as_uint(v), 0u, 0x0123));
}

template <int DP>
__global__ void __launch_bounds__(BP_THREADS)
calibrate_repair_kernel(const __grid_constant__ CalParams p, const uint8_t* __restrict__ mask,
                        int64_t H, int64_t W, int dp, int min_valid, float* __restrict__ out, int out_big_endian,
                        unsigned long long* __restrict__ counts, bool vec_ok) {
    const CalibratedImage img{p};
    const int64_t groups_per_row = (W + 3) / 4;
    const int64_t g = (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool warp_active = (g - lane) < groups_per_row;          // whole warps stay in the loop together
    unsigned nbad = 0, nfix = 0;
    for (int64_t r = blockIdx.y; warp_active && r < H; r += gridDim.y) {
        const int64_t c0 = g * 4;
        const int64_t base = r * W + c0;
        float px[4] = {0.f, 0.f, 0.f, 0.f};
        bool bad[4] = {false, false, false, false};
        const int nvalid = g < groups_per_row ? (int)((W - c0) < 4 ? (W - c0) : 4) : 0;
        if (vec_ok && nvalid == 4) {
            float rv[4];
            if (p.raw_kind == APGPU_RAW_F32) {
                const float4 v = ld_stream(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.raw) + base));
                rv[0] = v.x; rv[1] = v.y; rv[2] = v.z; rv[3] = v.w;
            } else {
                const ushort4 u = __ldcs(reinterpret_cast<const ushort4*>(reinterpret_cast<const uint16_t*>(p.raw) + base));
                uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (p.raw_kind == APGPU_RAW_U16_FITS) w4[k] = (__byte_perm(w4[k], 0u, 0x4401) ^ 0x8000u) & 0xffffu;
                    rv[k] = (float)w4[k];
                }
            }
            if (p.has_ped) {
#pragma unroll
                for (int k = 0; k < 4; ++k) rv[k] = __fadd_rn(rv[k], p.ped);
            }
            const float4 b = ld_stream(reinterpret_cast<const float4*>(p.bias + base));
            const float4 d = ld_stream(reinterpret_cast<const float4*>(p.dark + base));
            float4 f = make_float4(1.f, 1.f, 1.f, 1.f);
            if (p.nflat) f = ld_stream(reinterpret_cast<const float4*>(p.nflat + base));
            px[0] = cal1(p, rv[0], b.x, d.x, f.x); px[1] = cal1(p, rv[1], b.y, d.y, f.y);
            px[2] = cal1(p, rv[2], b.z, d.z, f.z); px[3] = cal1(p, rv[3], b.w, d.w, f.w);
            if (mask) {
                const uchar4 m4 = __ldcs(reinterpret_cast<const uchar4*>(mask + base));
                bad[0] = m4.x != 0; bad[1] = m4.y != 0; bad[2] = m4.z != 0; bad[3] = m4.w != 0;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < nvalid) { px[k] = img.at(base + k); bad[k] = mask && mask[base + k] != 0; }
            }
        }
        if (mask) {
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
                        const bool ok = repair_warp<(DP == 1 || DP == 2) ? DP : 1, uint8_t>(img, mask, H, W, 0, r, cc, min_valid, res);
                        if (ok && lane == src) { px[k] = res; ++nfix; }
                    }
                } else if (bad[k]) {
                    ++nbad;
                    float res;
                    if (repair_any<uint8_t>(img, mask, H, W, 0, r, c0 + k, dp, min_valid, res)) { px[k] = res; ++nfix; }
                }
            }
        }
        if (out_big_endian) {
#pragma unroll
            for (int k = 0; k < 4; ++k) px[k] = to_big_endian(px[k]);
        }
        if (vec_ok && nvalid == 4) {
            st_stream(reinterpret_cast<float4*>(out + base), make_float4(px[0], px[1], px[2], px[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < nvalid) out[base + k] = px[k];
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        nbad += __shfl_down_sync(0xffffffffu, nbad, off);
        nfix += __shfl_down_sync(0xffffffffu, nfix, off);
    }
    if (lane == 0 && nbad) {
        atomicAdd(&counts[0], (unsigned long long)nbad);
        if (nfix) atomicAdd(&counts[1], (unsigned long long)nfix);
    }
}

}  // namespace

extern "C" int apgpu_calibrate_repair(const void* raw, int raw_kind, float pedestal, int has_pedestal,
                                      const float* bias, const float* dark, const float* normflat,
                                      float exp_ratio, int dark_still_biased,
                                      const uint8_t* mask, int64_t H, int64_t W, int deltapix, int min_valid,
                                      float* out, int out_big_endian, int64_t* counts, apgpu_stream_t stream) {
    APGPU_REQUIRE(raw && bias && dark && out, "calibrate_repair: null image pointer");
    APGPU_REQUIRE(raw_kind == APGPU_RAW_F32 || raw_kind == APGPU_RAW_U16 || raw_kind == APGPU_RAW_U16_FITS,
                  "calibrate_repair: bad raw_kind %d", raw_kind);
    APGPU_REQUIRE(H > 0 && W > 0 && H * W < ((int64_t)1 << 40), "calibrate_repair: bad image shape %lld x %lld",
                  (long long)H, (long long)W);
    APGPU_REQUIRE(!mask || counts, "calibrate_repair: a mask needs the counts output");
    APGPU_REQUIRE(!mask || (deltapix >= 1 && deltapix <= 7), "calibrate_repair: deltapix %d outside supported range 1..7", deltapix);
    APGPU_REQUIRE(!mask || min_valid >= 1, "calibrate_repair: min_valid must be >= 1");
    CalParams p;
    p.raw = raw; p.bias = bias; p.dark = dark; p.nflat = normflat;
    p.ped = pedestal; p.r = exp_ratio; p.raw_kind = raw_kind;
    p.has_ped = has_pedestal != 0; p.biased = dark_still_biased != 0;
    const int64_t groups = (W + 3) / 4;
    dim3 grid((unsigned)((groups + BP_THREADS - 1) / BP_THREADS), (unsigned)(H < 65535 ? H : 65535));
    const size_t raw_align = raw_kind == APGPU_RAW_F32 ? 16 : 8;
    const bool vec_ok = (W % 4 == 0) && apgpu_aligned(raw, raw_align) && apgpu_aligned(bias, 16) && apgpu_aligned(dark, 16) &&
                        apgpu_aligned(out, 16) && (!normflat || apgpu_aligned(normflat, 16)) && (!mask || apgpu_aligned(mask, 4));
    unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
    cudaStream_t st = (cudaStream_t)stream;
    if (!mask || deltapix == 1)
        calibrate_repair_kernel<1><<<grid, BP_THREADS, 0, st>>>(p, mask, H, W, deltapix, min_valid, out, out_big_endian, c, vec_ok);
    else if (deltapix == 2)
        calibrate_repair_kernel<2><<<grid, BP_THREADS, 0, st>>>(p, mask, H, W, deltapix, min_valid, out, out_big_endian, c, vec_ok);
    else
        calibrate_repair_kernel<0><<<grid, BP_THREADS, 0, st>>>(p, mask, H, W, deltapix, min_valid, out, out_big_endian, c, vec_ok);
    APGPU_LAUNCH_CHECK("calibrate_repair_kernel");
    return APGPU_OK;
}

// Fused science-frame calibration + bad-pixel repair (sm_100a): one call per frame for the batch driver.
//
// Reference: ApCalibrate.calibrate, AstroPhotography/core/ApCalibrate.py:439-479 -- the numpy calibration
// arithmetic (:439-474) immediately followed by ApFixBadPixels.fix_bad_pixels on its result (:478-479).  Run as
// two kernels that intermediate image costs 8 more bytes per pixel of HBM traffic (4 written, 4 read again); here
// a streaming kernel calibrates every pixel at copy bandwidth and a persistent mask scan then overwrites the
// repaired ones (badpix_common.cuh); a bad pixel's donors -- the CALIBRATED values of its good neighbours -- are recomputed from raw / bias / dark /
// flat on the fly (<= 24 neighbours of ~1 % of the pixels), so the intermediate image is never read back.  The
// arithmetic is the same explicitly rounded float32 sequence as calibrate.cu and the same warp-cooperative median as badpix.cu (badpix_common.cuh),
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
    // four consecutive pixels starting at a multiple of 4 (128-bit streaming loads)
    __device__ __forceinline__ float4 vec4(int64_t i) const {
        float rv[4];
        if (p.raw_kind == APGPU_RAW_F32) {
            const float4 v = ld_stream(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.raw) + i));
            rv[0] = v.x; rv[1] = v.y; rv[2] = v.z; rv[3] = v.w;
        } else {
            const ushort4 w = __ldcs(reinterpret_cast<const ushort4*>(reinterpret_cast<const uint16_t*>(p.raw) + i));
            uint32_t w4[4] = {w.x, w.y, w.z, w.w};
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
        const float4 b = ld_stream(reinterpret_cast<const float4*>(p.bias + i));
        const float4 d = ld_stream(reinterpret_cast<const float4*>(p.dark + i));
        const float4 f = p.nflat ? ld_stream(reinterpret_cast<const float4*>(p.nflat + i)) : make_float4(1.f, 1.f, 1.f, 1.f);
        return make_float4(cal1(p, rv[0], b.x, d.x, f.x), cal1(p, rv[1], b.y, d.y, f.y),
                           cal1(p, rv[2], b.z, d.z, f.z), cal1(p, rv[3], b.w, d.w, f.w));
    }
};

__device__ __forceinline__ float to_big_endian(float v) {
    return __uint_as_float(__byte_perm(__float_as_uint(v), 0u, 0x0123));
}

// dense pass: the calibration arithmetic at copy bandwidth -- 128-bit streaming accesses, two independent
// vectors in flight per thread (like calibrate_vec4_kernel), raw decode and output byte order chosen at run time
constexpr int CR_UNROLL = 2;

__global__ void __launch_bounds__(BP_THREADS)
calibrate_stream_kernel(const __grid_constant__ CalParams p, float* __restrict__ out, int out_big_endian, int64_t nvec) {
    const int64_t base = (int64_t)blockIdx.x * (BP_THREADS * CR_UNROLL) + threadIdx.x;
    float rv[CR_UNROLL][4];
    float4 b[CR_UNROLL], d[CR_UNROLL], f[CR_UNROLL];
#pragma unroll
    for (int u = 0; u < CR_UNROLL; ++u) {
        const int64_t i = base + (int64_t)u * BP_THREADS;
        if (i < nvec) {
            if (p.raw_kind == APGPU_RAW_F32) {
                const float4 v = ld_stream(reinterpret_cast<const float4*>(p.raw) + i);
                rv[u][0] = v.x; rv[u][1] = v.y; rv[u][2] = v.z; rv[u][3] = v.w;
            } else {
                const ushort4 w = __ldcs(reinterpret_cast<const ushort4*>(p.raw) + i);
                uint32_t w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (p.raw_kind == APGPU_RAW_U16_FITS) w4[k] = (__byte_perm(w4[k], 0u, 0x4401) ^ 0x8000u) & 0xffffu;
                    rv[u][k] = (float)w4[k];
                }
            }
            b[u] = ld_stream(reinterpret_cast<const float4*>(p.bias) + i);
            d[u] = ld_stream(reinterpret_cast<const float4*>(p.dark) + i);
            f[u] = p.nflat ? ld_stream(reinterpret_cast<const float4*>(p.nflat) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
    }
#pragma unroll
    for (int u = 0; u < CR_UNROLL; ++u) {
        const int64_t i = base + (int64_t)u * BP_THREADS;
        if (i < nvec) {
            if (p.has_ped) {
#pragma unroll
                for (int k = 0; k < 4; ++k) rv[u][k] = __fadd_rn(rv[u][k], p.ped);
            }
            float4 o;
            o.x = cal1(p, rv[u][0], b[u].x, d[u].x, f[u].x); o.y = cal1(p, rv[u][1], b[u].y, d[u].y, f[u].y);
            o.z = cal1(p, rv[u][2], b[u].z, d[u].z, f[u].z); o.w = cal1(p, rv[u][3], b[u].w, d[u].w, f[u].w);
            if (out_big_endian) { o.x = to_big_endian(o.x); o.y = to_big_endian(o.y); o.z = to_big_endian(o.z); o.w = to_big_endian(o.w); }
            st_stream(reinterpret_cast<float4*>(out) + i, o);
        }
    }
}

__global__ void __launch_bounds__(BP_THREADS)
calibrate_stream_scalar_kernel(const __grid_constant__ CalParams p, float* __restrict__ out, int out_big_endian,
                               int64_t i0, int64_t n) {
    const CalibratedImage img{p};
    const int64_t i = i0 + (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    if (i < n) {
        const float v = img.at(i);
        out[i] = out_big_endian ? to_big_endian(v) : v;
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
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t npix = H * W;
    const size_t raw_align = raw_kind == APGPU_RAW_F32 ? 16 : 8;
    const bool vec_ok = apgpu_aligned(raw, raw_align) && apgpu_aligned(bias, 16) && apgpu_aligned(dark, 16) &&
                        apgpu_aligned(out, 16) && (!normflat || apgpu_aligned(normflat, 16));
    // 1. every pixel calibrated, unrepaired, at streaming speed
    const int64_t nvec = vec_ok ? npix / 4 : 0;
    if (nvec > 0) {
        const int64_t per_block = BP_THREADS * CR_UNROLL;
        calibrate_stream_kernel<<<(unsigned)((nvec + per_block - 1) / per_block), BP_THREADS, 0, st>>>(p, out, out_big_endian, nvec);
        APGPU_LAUNCH_CHECK("calibrate_stream_kernel");
    }
    if (nvec * 4 < npix) {
        const int64_t rem = npix - nvec * 4;
        calibrate_stream_scalar_kernel<<<(unsigned)((rem + BP_THREADS - 1) / BP_THREADS), BP_THREADS, 0, st>>>(
            p, out, out_big_endian, nvec * 4, npix);
        APGPU_LAUNCH_CHECK("calibrate_stream_scalar_kernel");
    }
    // 2. the mask scan overwrites the repaired pixels; their donors are recomputed from the inputs
    if (mask)
        return launch_repair_scan<uint8_t>(CalibratedImage{p}, mask, H, W, 0, 0, H, deltapix, min_valid, out,
                                           out_big_endian != 0, counts, st);
    return APGPU_OK;
}

// Fused science-frame calibration and flat normalisation (sm_100a).
//
// Reference arithmetic: AstroPhotography/core/ApCalibrate.py:439-474 (calibrate),
// :178-190 (_generate_flat), :303-326 (_read_fits conversion + PEDESTAL).
//
// calibrate: purely HBM-bound -- one read each of raw, bias, dark, normflat and
// one write (20 B/pixel, 18 B with uint16 raw); 128-bit streaming loads/stores,
// two independent vectors in flight per thread, no shared memory (no reuse).
// Every float32 operation is an explicitly rounded intrinsic so that nvcc can
// never contract a multiply-subtract into an FMA: the reference (numpy) rounds
// each of its five array operations separately.
#include "apgpu_common.cuh"

namespace {

template <bool FLAT, bool BIASED>
__device__ __forceinline__ float cal1(float raw, float b, float d, float nf, float r) {
    float t = __fsub_rn(raw, b);                       // :439 raw - bias
    float ds = BIASED ? __fsub_rn(d, b) : d;           // :442 dark - bias
    float o = __fsub_rn(t, __fmul_rn(r, ds));          // :450-451
    if (FLAT) {
        float q = __fdiv_rn(o, nf);                    // :462-464 np.where(nf != 0, o/nf, o)
        o = (nf != 0.0f) ? q : o;
    }
    return o;
}

template <bool PED>
__device__ __forceinline__ float4 raw4(const float4* p, int64_t i, float) { return ld_stream(p + i); }
template <bool PED>
__device__ __forceinline__ float4 raw4(const ushort4* p, int64_t i, float ped) {
    ushort4 u = __ldcs(p + i);
    float4 v = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
    if (PED) {                                          // :322 ext_data += pedestal (float32 add)
        v.x = __fadd_rn(v.x, ped); v.y = __fadd_rn(v.y, ped);
        v.z = __fadd_rn(v.z, ped); v.w = __fadd_rn(v.w, ped);
    }
    return v;
}
template <bool PED>
__device__ __forceinline__ float raw1(const float* p, int64_t i, float) { return ld_stream(p + i); }
template <bool PED>
__device__ __forceinline__ float raw1(const uint16_t* p, int64_t i, float ped) {
    float v = (float)__ldcs(p + i);
    return PED ? __fadd_rn(v, ped) : v;
}

template <typename T> struct Vec4Of;
template <> struct Vec4Of<float> { using type = float4; };
template <> struct Vec4Of<uint16_t> { using type = ushort4; };

constexpr int CAL_THREADS = 256;
constexpr int CAL_UNROLL = 2;      // independent 128-bit vectors per thread per array

template <typename RawT, bool FLAT, bool BIASED, bool PED>
__global__ void __launch_bounds__(CAL_THREADS)
calibrate_vec4_kernel(const RawT* __restrict__ raw, float ped,
                      const float* __restrict__ bias, const float* __restrict__ dark,
                      const float* __restrict__ nflat, float r,
                      float* __restrict__ out, int64_t nvec) {
    using RV = typename Vec4Of<RawT>::type;
    const RV* raw4p = reinterpret_cast<const RV*>(raw);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
    const float4* d4 = reinterpret_cast<const float4*>(dark);
    const float4* f4 = reinterpret_cast<const float4*>(nflat);
    float4* o4 = reinterpret_cast<float4*>(out);
    int64_t base = (int64_t)blockIdx.x * (CAL_THREADS * CAL_UNROLL) + threadIdx.x;
    float4 vr[CAL_UNROLL], vb[CAL_UNROLL], vd[CAL_UNROLL], vf[CAL_UNROLL];
#pragma unroll
    for (int u = 0; u < CAL_UNROLL; ++u) {
        int64_t i = base + (int64_t)u * CAL_THREADS;
        if (i < nvec) {
            vr[u] = raw4<PED>(raw4p, i, ped);
            vb[u] = ld_stream(b4 + i);
            vd[u] = ld_stream(d4 + i);
            if (FLAT) vf[u] = ld_stream(f4 + i);
        }
    }
#pragma unroll
    for (int u = 0; u < CAL_UNROLL; ++u) {
        int64_t i = base + (int64_t)u * CAL_THREADS;
        if (i < nvec) {
            float4 o;
            o.x = cal1<FLAT, BIASED>(vr[u].x, vb[u].x, vd[u].x, FLAT ? vf[u].x : 1.f, r);
            o.y = cal1<FLAT, BIASED>(vr[u].y, vb[u].y, vd[u].y, FLAT ? vf[u].y : 1.f, r);
            o.z = cal1<FLAT, BIASED>(vr[u].z, vb[u].z, vd[u].z, FLAT ? vf[u].z : 1.f, r);
            o.w = cal1<FLAT, BIASED>(vr[u].w, vb[u].w, vd[u].w, FLAT ? vf[u].w : 1.f, r);
            st_stream(o4 + i, o);
        }
    }
}

// Scalar kernel: unaligned pointers and the < 4 pixel tail.
template <typename RawT, bool FLAT, bool BIASED, bool PED>
__global__ void __launch_bounds__(CAL_THREADS)
calibrate_scalar_kernel(const RawT* __restrict__ raw, float ped,
                        const float* __restrict__ bias, const float* __restrict__ dark,
                        const float* __restrict__ nflat, float r,
                        float* __restrict__ out, int64_t i0, int64_t n) {
    int64_t i = i0 + (int64_t)blockIdx.x * CAL_THREADS + threadIdx.x;
    if (i < n) {
        float v = raw1<PED>(raw, i, ped);
        out[i] = cal1<FLAT, BIASED>(v, bias[i], dark[i], FLAT ? nflat[i] : 1.f, r);
    }
}

template <typename RawT, bool FLAT, bool BIASED, bool PED>
int launch_cal(const RawT* raw, float ped, const float* bias, const float* dark,
               const float* nflat, float r, float* out, int64_t npix, cudaStream_t st) {
    bool vec_ok = apgpu_aligned(raw, sizeof(RawT) * 4) && apgpu_aligned(bias, 16) &&
                  apgpu_aligned(dark, 16) && apgpu_aligned(out, 16) &&
                  (!FLAT || apgpu_aligned(nflat, 16));
    int64_t nvec = vec_ok ? npix / 4 : 0;
    if (nvec > 0) {
        int64_t per_block = CAL_THREADS * CAL_UNROLL;
        int64_t blocks = (nvec + per_block - 1) / per_block;
        calibrate_vec4_kernel<RawT, FLAT, BIASED, PED><<<(unsigned)blocks, CAL_THREADS, 0, st>>>(
            raw, ped, bias, dark, nflat, r, out, nvec);
        APGPU_LAUNCH_CHECK("calibrate_vec4_kernel");
    }
    int64_t done = nvec * 4;
    if (done < npix) {
        int64_t rem = npix - done;
        int64_t blocks = (rem + CAL_THREADS - 1) / CAL_THREADS;
        calibrate_scalar_kernel<RawT, FLAT, BIASED, PED><<<(unsigned)blocks, CAL_THREADS, 0, st>>>(
            raw, ped, bias, dark, nflat, r, out, done, npix);
        APGPU_LAUNCH_CHECK("calibrate_scalar_kernel");
    }
    return APGPU_OK;
}

template <typename RawT>
int dispatch_cal(const RawT* raw, float ped, bool has_ped, const float* bias, const float* dark,
                 const float* nflat, float r, int biased, float* out, int64_t npix,
                 cudaStream_t st) {
    APGPU_REQUIRE(raw && bias && dark && out, "calibrate: null image pointer");
    APGPU_REQUIRE(npix >= 0 && npix < ((int64_t)1 << 40), "calibrate: bad npix %lld", (long long)npix);
    if (npix == 0) return APGPU_OK;
    bool flat = nflat != nullptr;
#define CAL_CASE(F, B, P) \
    if (flat == F && (biased != 0) == B && has_ped == P) \
        return launch_cal<RawT, F, B, P>(raw, ped, bias, dark, nflat, r, out, npix, st);
    CAL_CASE(false, false, false) CAL_CASE(false, false, true)
    CAL_CASE(false, true, false)  CAL_CASE(false, true, true)
    CAL_CASE(true, false, false)  CAL_CASE(true, false, true)
    CAL_CASE(true, true, false)   CAL_CASE(true, true, true)
#undef CAL_CASE
    return APGPU_ERR_ARG;
}

// ------------------------------------------------------------------------
// Flat normalisation: np.nanmean(flat) reproduced bit-for-bit.
//
// numpy sums a C-contiguous float32 image with FLOAT_pairwise_sum
// (numpy/_core/src/umath/loops_utils.h.src): a binary recursion that splits
// n into n2 = n/2 - (n/2)%8 and n-n2 until n <= 128, sums each leaf with 8
// interleaved accumulators, and adds the halves left+right on the way up.
// The tree is a pure function of npix, so it can be evaluated in parallel:
// one thread per node at the deepest depth D at which every node still
// exists (all depth-(D-1) nodes have n > 128); that thread evaluates its
// subtree serially; then the 2^D results are folded pairwise (p[i] =
// p[2i] + p[2i+1]) level by level -- the same additions in the same order.
// np.nanmean replaces NaN by 0 before summing and divides the float32 sum by
// the count of non-NaN in float64, casting the quotient back to float32
// (numpy/lib/_nanfunctions_impl.py: _replace_nan, _divide_by_count).
// ------------------------------------------------------------------------
__device__ __forceinline__ float nan0(float x) { return (x != x) ? 0.0f : x; }

__device__ float pw_subtree(const float* __restrict__ a, int64_t lo, int64_t n, unsigned long long& cnt) {
    if (n < 8) {
        float res = 0.0f;
        for (int64_t i = 0; i < n; ++i) {
            float x = a[lo + i];
            cnt += (x == x);
            res = __fadd_rn(res, nan0(x));
        }
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float x = a[lo + j];
            cnt += (x == x);
            r[j] = nan0(x);
        }
        int64_t m = n - (n % 8);
        for (int64_t i = 8; i < m; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float x = a[lo + i + j];
                cnt += (x == x);
                r[j] = __fadd_rn(r[j], nan0(x));
            }
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (int64_t i = m; i < n; ++i) {
            float x = a[lo + i];
            cnt += (x == x);
            res = __fadd_rn(res, nan0(x));
        }
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    float l = pw_subtree(a, lo, n2, cnt);
    float rr = pw_subtree(a, lo + n2, n - n2, cnt);
    return __fadd_rn(l, rr);
}

constexpr int FN_THREADS = 256;

// Stage A: node sums at depth D, folded inside the block down to one value
// per block (blockDim = min(2^D, 256), a power of two).
__global__ void flat_norm_nodes_kernel(const float* __restrict__ flat, int64_t npix, int depth,
                                       float* __restrict__ partial,
                                       unsigned long long* __restrict__ count) {
    __shared__ float sh[FN_THREADS];
    __shared__ unsigned long long shc;
    if (threadIdx.x == 0) shc = 0;
    __syncthreads();
    int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t lo = 0, n = npix;
    for (int lvl = depth - 1; lvl >= 0; --lvl) {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        if ((node >> lvl) & 1) { lo += n2; n -= n2; } else { n = n2; }
    }
    unsigned long long c = 0;
    float v = pw_subtree(flat, lo, n, c);
    atomicAdd(&shc, c);
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int width = blockDim.x >> 1; width >= 1; width >>= 1) {
        float s = 0.f;
        if ((int)threadIdx.x < width) s = __fadd_rn(sh[2 * threadIdx.x], sh[2 * threadIdx.x + 1]);
        __syncthreads();
        if ((int)threadIdx.x < width) sh[threadIdx.x] = s;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sh[0];
        atomicAdd(count, shc);
    }
}

// Stage B: one block folds the per-block values (count is a power of two)
// and writes norm = float32(float64(0.0f + sum) / count).
__global__ void flat_norm_finish_kernel(float* __restrict__ p, float* __restrict__ q, int64_t count_nodes,
                                        const unsigned long long* __restrict__ count,
                                        float* __restrict__ norm_out) {
    float* src = p;
    float* dst = q;
    for (int64_t width = count_nodes >> 1; width >= 1; width >>= 1) {
        for (int64_t i = threadIdx.x; i < width; i += blockDim.x)
            dst[i] = __fadd_rn(src[2 * i], src[2 * i + 1]);
        __syncthreads();
        float* t = src; src = dst; dst = t;
    }
    if (threadIdx.x == 0) {
        float total = __fadd_rn(0.0f, src[0]);     // add.reduce starts from the identity
        double c = (double)(*count);
        norm_out[0] = (float)__ddiv_rn((double)total, c);
    }
}

__global__ void flat_divide_kernel(const float* __restrict__ flat, const float* __restrict__ norm,
                                   float* __restrict__ out, int64_t npix) {
    float nv = norm[0];
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < npix) {
        float4 v = ld_stream(reinterpret_cast<const float4*>(flat + i));
        v.x = __fdiv_rn(v.x, nv); v.y = __fdiv_rn(v.y, nv);
        v.z = __fdiv_rn(v.z, nv); v.w = __fdiv_rn(v.w, nv);
        *reinterpret_cast<float4*>(out + i) = v;
    } else {
        for (; i < npix; ++i) out[i] = __fdiv_rn(flat[i], nv);
    }
}

__global__ void flat_divide_scalar_kernel(const float* __restrict__ flat, const float* __restrict__ norm,
                                          float* __restrict__ out, int64_t npix) {
    float nv = norm[0];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) out[i] = __fdiv_rn(flat[i], nv);
}

// Deepest depth at which every node of numpy's pairwise tree still exists.
int flat_norm_depth(int64_t npix) {
    // Track the distinct node sizes per level (few: sizes differ by < 8*level).
    int64_t sizes[256];
    int ns = 1;
    sizes[0] = npix;
    int depth = 0;
    while (depth < 40) {
        bool all_split = true;
        for (int i = 0; i < ns; ++i) if (sizes[i] <= 128) all_split = false;
        if (!all_split) break;
        int64_t next[512];
        int nn = 0;
        for (int i = 0; i < ns; ++i) {
            int64_t n2 = sizes[i] / 2; n2 -= n2 % 8;
            int64_t ch[2] = {n2, sizes[i] - n2};
            for (int k = 0; k < 2; ++k) {
                bool seen = false;
                for (int j = 0; j < nn; ++j) if (next[j] == ch[k]) seen = true;
                if (!seen && nn < 512) next[nn++] = ch[k];
            }
        }
        if (nn > 256) break;       // cannot happen for image sizes; stay safe
        for (int i = 0; i < nn; ++i) sizes[i] = next[i];
        ns = nn;
        ++depth;
    }
    return depth;
}

}  // namespace

extern "C" int apgpu_calibrate_f32(const float* raw, const float* bias, const float* dark,
                                   const float* normflat, float exp_ratio, int dark_still_biased,
                                   float* out, int64_t npix, apgpu_stream_t stream) {
    return dispatch_cal<float>(raw, 0.f, false, bias, dark, normflat, exp_ratio,
                               dark_still_biased, out, npix, (cudaStream_t)stream);
}

extern "C" int apgpu_calibrate_u16(const uint16_t* raw, float pedestal, int has_pedestal,
                                   const float* bias, const float* dark, const float* normflat,
                                   float exp_ratio, int dark_still_biased,
                                   float* out, int64_t npix, apgpu_stream_t stream) {
    return dispatch_cal<uint16_t>(raw, pedestal, has_pedestal != 0, bias, dark, normflat,
                                  exp_ratio, dark_still_biased, out, npix, (cudaStream_t)stream);
}

extern "C" size_t apgpu_flat_norm_workspace_bytes(int64_t npix) {
    if (npix <= 0) return 64;
    int depth = flat_norm_depth(npix);
    int64_t nodes = (int64_t)1 << depth;
    int64_t blocks = nodes >= FN_THREADS ? nodes / FN_THREADS : 1;
    return 64 + (size_t)blocks * sizeof(float) * 2;
}

extern "C" int apgpu_flat_norm_f32(const float* flat, int64_t npix, void* workspace,
                                   size_t workspace_bytes, float* norm_out, apgpu_stream_t stream) {
    APGPU_REQUIRE(flat && workspace && norm_out, "flat_norm: null pointer");
    APGPU_REQUIRE(npix > 0, "flat_norm: npix must be positive (got %lld)", (long long)npix);
    APGPU_REQUIRE(workspace_bytes >= apgpu_flat_norm_workspace_bytes(npix),
                  "flat_norm: workspace too small (%zu < %zu)", workspace_bytes,
                  apgpu_flat_norm_workspace_bytes(npix));
    APGPU_REQUIRE(apgpu_aligned(workspace, 8), "flat_norm: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int depth = flat_norm_depth(npix);
    int64_t nodes = (int64_t)1 << depth;
    int threads = nodes >= FN_THREADS ? FN_THREADS : (int)nodes;
    int64_t blocks = nodes / threads;
    unsigned long long* count = reinterpret_cast<unsigned long long*>(workspace);
    float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 64);
    float* q = p + blocks;
    APGPU_CUDA(cudaMemsetAsync(count, 0, 8, st));
    flat_norm_nodes_kernel<<<(unsigned)blocks, threads, 0, st>>>(flat, npix, depth, p, count);
    APGPU_LAUNCH_CHECK("flat_norm_nodes_kernel");
    flat_norm_finish_kernel<<<1, 1024, 0, st>>>(p, q, blocks, count, norm_out);
    APGPU_LAUNCH_CHECK("flat_norm_finish_kernel");
    return APGPU_OK;
}

extern "C" int apgpu_flat_divide_f32(const float* flat, const float* norm, float* normflat,
                                     int64_t npix, apgpu_stream_t stream) {
    APGPU_REQUIRE(flat && norm && normflat, "flat_divide: null pointer");
    APGPU_REQUIRE(npix >= 0, "flat_divide: bad npix");
    if (npix == 0) return APGPU_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (apgpu_aligned(flat, 16) && apgpu_aligned(normflat, 16)) {
        int64_t nthreads = (npix + 3) / 4;
        int64_t blocks = (nthreads + 255) / 256;
        flat_divide_kernel<<<(unsigned)blocks, 256, 0, st>>>(flat, norm, normflat, npix);
    } else {
        int64_t blocks = (npix + 255) / 256;
        flat_divide_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(flat, norm, normflat, npix);
    }
    APGPU_LAUNCH_CHECK("flat_divide_kernel");
    return APGPU_OK;
}

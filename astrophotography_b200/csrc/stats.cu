// Whole-image sigma-clipped statistics and threshold mask (sm_100a).
//
// Replaces astropy.stats.sigma_clipped_stats(data, sigma=...) as called by the
// reference's mask producer, AstroPhotography/core/ApFindBadPixels.py:191, and the
// threshold mask that follows it (:194-209).  (the step
// that produces the mask the repair kernel consumes.)
//
// sigma_clipped_stats defaults: cenfunc = median, stdfunc = population std,
// maxiters = 5; every iteration keeps x in [c - sigma*s, c + sigma*s], so the
// surviving set is always "finite x inside an interval [L, U]" and never has to
// be materialised.  One iteration on the device:
//   moments   count and float64 sum of the survivors      -> n, mean
//   moments2  float64 sum of (x - mean)^2                  -> std
//   select    exact median by 3-pass radix select (11+11+10 bits) on the
//             order-preserving uint32 key of the float; for even n both middle
//             order statistics are selected
//   update    L = max(L, c - sigma*s), U = min(U, c + sigma*s)
// All passes are stream-ordered with the running state in device memory: no
// host synchronisation until the three results are read back.
// Parity: the median is exact; mean/std are float64 tree reductions whose
// summation order differs from numpy's pairwise order (relative 1e-15), so a
// clip decision can differ only for a sample within 1e-15 of a bound.
#include <float.h>
#include <math.h>

#include "apgpu_common.cuh"

namespace {

struct StatsState {
    double L, U;                  // current survivor interval
    double sum, sumsq, mean, std, median;
    unsigned long long n;         // survivors
    unsigned long long rank;      // rank still to resolve inside the current prefix
    unsigned int prefix, prefix_mask;
    double sel[2];                // the one or two middle order statistics
    unsigned int hist[2048];
};

__device__ __forceinline__ unsigned int float_key(float x) {
    unsigned int b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
    unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
__device__ __forceinline__ bool survivor(float x, double L, double U) {
    return fabsf(x) <= FLT_MAX && (double)x >= L && (double)x <= U;
}

constexpr int ST_THREADS = 256;
constexpr int ST_BLOCKS = APGPU_NUM_SMS * 8;

__global__ void stats_init_kernel(StatsState* st) {
    st->L = -(double)INFINITY;
    st->U = (double)INFINITY;
    st->n = 0; st->sum = 0.0; st->sumsq = 0.0;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < ST_THREADS / 32) r = sh[threadIdx.x];
    if (threadIdx.x < 32) for (int off = 16; off > 0; off >>= 1) r += __shfl_down_sync(0xffffffffu, r, off);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(ST_THREADS)
stats_moments_kernel(const float* __restrict__ x, int64_t n, StatsState* st, int second) {
    __shared__ double sh[ST_THREADS / 32];
    const double L = st->L, U = st->U, mean = st->mean;
    double acc = 0.0;
    unsigned long long cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * ST_THREADS) {
        float v = x[i];
        if (survivor(v, L, U)) {
            if (second) { double d = (double)v - mean; acc += d * d; }
            else { acc += (double)v; ++cnt; }
        }
    }
    double tot = block_sum(acc, sh);
    double totc = block_sum((double)cnt, sh);
    if (threadIdx.x == 0) {
        if (second) atomicAdd(&st->sumsq, tot);
        else { atomicAdd(&st->sum, tot); atomicAdd(&st->n, (unsigned long long)totc); }
    }
}

// after the first moments pass: mean; after the second: std.  Also arms the select.
__global__ void stats_finish_moments_kernel(StatsState* st, int second, int which_rank) {
    if (threadIdx.x == 0) {
        if (!second) {
            st->mean = st->n ? st->sum / (double)st->n : (double)NAN;
            st->sumsq = 0.0;
        } else {
            st->std = st->n ? sqrt(st->sumsq / (double)st->n) : (double)NAN;
        }
        if (which_rank >= 0) {
            unsigned long long n = st->n;
            st->rank = which_rank == 0 ? (n ? (n - 1) / 2 : 0) : n / 2;
            st->prefix = 0; st->prefix_mask = 0;
        }
    }
    if (which_rank >= 0)
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) st->hist[i] = 0;
}

__global__ void __launch_bounds__(ST_THREADS)
stats_hist_kernel(const float* __restrict__ x, int64_t n, StatsState* st, int shift, int bits) {
    __shared__ unsigned int sh[2048];
    for (int i = threadIdx.x; i < 2048; i += ST_THREADS) sh[i] = 0;
    __syncthreads();
    const double L = st->L, U = st->U;
    const unsigned int prefix = st->prefix, pmask = st->prefix_mask;
    const unsigned int dmask = (1u << bits) - 1u;
    for (int64_t i = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * ST_THREADS) {
        float v = x[i];
        if (survivor(v, L, U)) {
            unsigned int k = float_key(v);
            if ((k & pmask) == prefix) atomicAdd(&sh[(k >> shift) & dmask], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += ST_THREADS)
        if (sh[i]) atomicAdd(&st->hist[i], sh[i]);
}

// single thread: find the digit holding the wanted rank, extend the prefix
__global__ void stats_pick_kernel(StatsState* st, int shift, int bits, int which_rank, int last) {
    if (threadIdx.x == 0) {
        unsigned long long r = st->rank;
        unsigned int nb = 1u << bits, d = 0;
        for (; d < nb; ++d) {
            unsigned int h = st->hist[d];
            if (r < h) break;
            r -= h;
        }
        if (d >= nb) d = nb - 1;
        st->rank = r;
        st->prefix |= d << shift;
        st->prefix_mask |= ((1u << bits) - 1u) << shift;
        if (last) st->sel[which_rank] = st->n ? (double)key_float(st->prefix) : (double)NAN;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) st->hist[i] = 0;
}

__global__ void stats_update_kernel(StatsState* st, double sigma, int do_clip, double* out3) {
    unsigned long long n = st->n;
    double med = (n & 1) ? st->sel[0] : (st->sel[0] + st->sel[1]) / 2.0;
    if (n == 0) med = (double)NAN;
    st->median = med;
    if (do_clip && n) {
        double lo = med - st->std * sigma;
        double hi = med + st->std * sigma;
        if (lo > st->L) st->L = lo;
        if (hi < st->U) st->U = hi;
    }
    if (out3) { out3[0] = st->mean; out3[1] = med; out3[2] = st->std; out3[3] = (double)n; }
    st->n = 0; st->sum = 0.0; st->sumsq = 0.0;
}

__global__ void __launch_bounds__(256)
threshold_mask_kernel(const float* __restrict__ x, int64_t n, double lo, double hi,
                      uint8_t* __restrict__ mask, unsigned long long* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    unsigned int bad = 0;
    if (i < n) {
        double v = (double)x[i];
        bad = (v < lo) || (v > hi);          // NaN compares false on both sides, like numpy
        mask[i] = (uint8_t)bad;
    }
    unsigned int tot = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(count, (unsigned long long)tot);
}

int one_round(const float* x, int64_t n, StatsState* st, double sigma, int do_clip, double* out3, cudaStream_t s) {
    stats_moments_kernel<<<ST_BLOCKS, ST_THREADS, 0, s>>>(x, n, st, 0);
    APGPU_LAUNCH_CHECK("stats_moments_kernel");
    stats_finish_moments_kernel<<<1, 256, 0, s>>>(st, 0, -1);
    APGPU_LAUNCH_CHECK("stats_finish_moments_kernel");
    stats_moments_kernel<<<ST_BLOCKS, ST_THREADS, 0, s>>>(x, n, st, 1);
    APGPU_LAUNCH_CHECK("stats_moments_kernel");
    for (int which = 0; which < 2; ++which) {
        stats_finish_moments_kernel<<<1, 256, 0, s>>>(st, 1, which);
        APGPU_LAUNCH_CHECK("stats_finish_moments_kernel");
        const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
        for (int pass = 0; pass < 3; ++pass) {
            stats_hist_kernel<<<ST_BLOCKS, ST_THREADS, 0, s>>>(x, n, st, shifts[pass], bits[pass]);
            APGPU_LAUNCH_CHECK("stats_hist_kernel");
            stats_pick_kernel<<<1, 256, 0, s>>>(st, shifts[pass], bits[pass], which, pass == 2);
            APGPU_LAUNCH_CHECK("stats_pick_kernel");
        }
    }
    stats_update_kernel<<<1, 1, 0, s>>>(st, sigma, do_clip, out3);
    APGPU_LAUNCH_CHECK("stats_update_kernel");
    return APGPU_OK;
}

}  // namespace

extern "C" size_t apgpu_image_stats_workspace_bytes(int64_t npix) {
    (void)npix;
    return sizeof(StatsState) + 64;
}

// out4 (device, 4 doubles): mean, median, std, count of the survivors.
extern "C" int apgpu_sigma_clipped_stats_f32(const float* data, int64_t npix, double sigma, int maxiters,
                                             void* workspace, size_t workspace_bytes, double* out4,
                                             apgpu_stream_t stream) {
    APGPU_REQUIRE(data && workspace && out4, "sigma_clipped_stats: null pointer");
    APGPU_REQUIRE(npix > 0, "sigma_clipped_stats: npix must be positive");
    APGPU_REQUIRE(maxiters >= 0 && maxiters <= 64, "sigma_clipped_stats: maxiters %d outside 0..64", maxiters);
    APGPU_REQUIRE(workspace_bytes >= apgpu_image_stats_workspace_bytes(npix), "sigma_clipped_stats: workspace too small");
    APGPU_REQUIRE(apgpu_aligned(workspace, 8), "sigma_clipped_stats: workspace must be 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    StatsState* st = reinterpret_cast<StatsState*>(workspace);
    stats_init_kernel<<<1, 1, 0, s>>>(st);
    APGPU_LAUNCH_CHECK("stats_init_kernel");
    // Each clip round narrows [L, U]; a round in which nothing is rejected leaves the
    // state unchanged, so running all `maxiters` rounds equals astropy's early exit.
    for (int it = 0; it < maxiters; ++it) {
        int rc = one_round(data, npix, st, sigma, 1, nullptr, s);
        if (rc) return rc;
    }
    return one_round(data, npix, st, sigma, 0, out4, s);
}

extern "C" int apgpu_threshold_mask_f32(const float* data, int64_t npix, double lo, double hi,
                                        uint8_t* mask, int64_t* nbad, apgpu_stream_t stream) {
    APGPU_REQUIRE(data && mask && nbad, "threshold_mask: null pointer");
    APGPU_REQUIRE(npix >= 0, "threshold_mask: bad npix");
    if (npix == 0) return APGPU_OK;
    int64_t blocks = (npix + 255) / 256;
    threshold_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        data, npix, lo, hi, mask, reinterpret_cast<unsigned long long*>(nbad));
    APGPU_LAUNCH_CHECK("threshold_mask_kernel");
    return APGPU_OK;
}

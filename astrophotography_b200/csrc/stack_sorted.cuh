// Frame-stack reducer: sorted<NB, NLO, MODE> register-resident Batcher-network kernels.  See stack_common.cuh.
#pragma once
#include <stdlib.h>
#include <string.h>
#include "stack_common.cuh"
#include "sort_networks.inc"

namespace apgpu_stack {

// ---------------------------------------------------------------------------
// sorted<NB, NLO, MODE>: Batcher network in registers, N in (NLO, NB]
// ---------------------------------------------------------------------------
constexpr int STPB = 256;         // threads per CTA of the sorted kernels (lock-stepped, see sort_regs)

#define CE_X(i, j) { float lo_ = fminf(x[i], x[j]); float hi_ = fmaxf(x[i], x[j]); x[i] = lo_; x[j] = hi_; }
// Every second comparator takes its maximum off the ALU pipe: FMNMX (min / max) issues once per two
// cycles per scheduler on the ALU pipe and the network is nothing but FMNMX, while the FMA pipe idles.
// lo = fminf(a, b) is one of the two inputs bit for bit, so hi = bits(a) + bits(b) - bits(lo) in wrapping
// integer arithmetic IS the other input, exactly (two IMAD on the FMA pipe; the multipliers +1 / -1 come
// from the kernel arguments so that ptxas cannot fold them back into ALU-pipe IADD3).  Pixels holding a
// NaN are marked for the generic routine anyway.
__device__ __forceinline__ int imad_fma_pipe(int a, int m, int c) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
    return d;
}
#define CE_Y(i, j) { const float lo_ = fminf(x[i], x[j]);                                              \
                     const int t_ = imad_fma_pipe(__float_as_int(x[i]), one_, __float_as_int(x[j]));   \
                     x[j] = __int_as_float(imad_fma_pipe(__float_as_int(lo_), mone_, t_)); x[i] = lo_; }

// Every 128 comparators the network has a CTA barrier: the 8 warps of a CTA walk the ~35 KB of
// straight-line code together, so one instruction-cache fill serves all of them (ncu before:
// `no_instruction` was the top stall of the median/MAD kernel).  The network is branch-free
// and data-independent, so the barrier costs no load imbalance.
#define SY_X() __syncthreads();
template <int NB, bool MIX> __device__ __forceinline__ void sort_regs(float (&x)[NB], int one_, int mone_);
#define APGPU_DEF_SORT(n) \
    template <> __device__ __forceinline__ void sort_regs<n, true>(float (&x)[n], int one_, int mone_) { APGPU_SORTNET_##n(CE_X, CE_Y, SY_X) } \
    template <> __device__ __forceinline__ void sort_regs<n, false>(float (&x)[n], int, int) { APGPU_SORTNET_##n(CE_X, CE_X, SY_X) }
APGPU_DEF_SORT(4) APGPU_DEF_SORT(8) APGPU_DEF_SORT(12) APGPU_DEF_SORT(16) APGPU_DEF_SORT(20)
APGPU_DEF_SORT(24) APGPU_DEF_SORT(32) APGPU_DEF_SORT(40) APGPU_DEF_SORT(48) APGPU_DEF_SORT(56)
APGPU_DEF_SORT(64) APGPU_DEF_SORT(72) APGPU_DEF_SORT(80) APGPU_DEF_SORT(90) APGPU_DEF_SORT(100)
APGPU_DEF_SORT(112) APGPU_DEF_SORT(128) APGPU_DEF_SORT(160) APGPU_DEF_SORT(200)

// one staged sample as its float32 value (uint16 stages hold raw 16-bit samples)
__device__ __forceinline__ float staged_value(const float* s, const StackArgs&) { return *s; }
__device__ __forceinline__ float staged_value(const uint16_t* s, const StackArgs& a) {
    return u16_biased((uint32_t)*s, a) - U16_BIAS;
}

// MAD = median of |x - med| over the sorted column s[base .. base+N) parked in shared memory ([row][thread],
// -inf / +inf guard rows and padding, so that every out-of-range row has an infinite deviation).  Left of the
// median the deviations grow towards row `base`, right of it towards row `base+N`: two sorted lists,
//     L[j] = med - s[l0 - j]   (j = 0 .. nL-1),   R[j] = s[l0 + 1 + j] - med   (j = 0 .. nR-1),
// whose (k+1)-th smallest element is found by bisecting on how many come from L (O(log N) shared-memory
// reads, float64, exact).
__device__ __forceinline__ double mad_parked(const float* s, const int base, const int N, const double med) {
    const int l0 = base + ((N - 1) >> 1);
    const int nL = l0 - base + 1, nR = N - nL;
    auto devL = [&](int j) { return fabs(__dsub_rn((double)s[(l0 - j) * STPB], med)); };
    auto devR = [&](int j) { return fabs(__dsub_rn((double)s[(l0 + 1 + j) * STPB], med)); };
    const int k1 = (N - 1) >> 1;                       // 0-based rank of the lower middle deviation
    int lo_i = k1 + 1 - nR > 0 ? k1 + 1 - nR : 0;
    int hi_i = k1 + 1 < nL ? k1 + 1 : nL;
    while (lo_i < hi_i) {
        const int mid = (lo_i + hi_i) >> 1;
        if (devL(mid) < devR(k1 - mid)) lo_i = mid + 1; else hi_i = mid;
    }
    // lo_i samples of the k1+1 smallest deviations come from L, k1+1-lo_i from R
    const double la = lo_i > 0 ? devL(lo_i - 1) : -1.0;
    const double ra = (k1 - lo_i) >= 0 ? devR(k1 - lo_i) : -1.0;
    const double d1 = la > ra ? la : ra;
    const double lb = devL(lo_i), rb = devR(k1 + 1 - lo_i);      // the next deviation up (inf past the ends)
    const double d2 = lb < rb ? lb : rb;
    return (N & 1) ? d1 : __dmul_rn(__dadd_rn(d1, d2), 0.5);
}

// TMA = true (equally spaced frames): the CTA's 256-pixel x N-frame tile arrives by ONE tensor-map bulk copy
// into shared memory (the same region later holds the parked sorted columns) and the threads read their
// column with LDS at immediate offsets: no LDG, no per-sample 64-bit address arithmetic on the ALU pipe
// that the comparators saturate.
template <int NB, int NLO, int MODE, bool MIX, bool TMA, typename T>
__global__ void __launch_bounds__(STPB, (NB <= 32 ? 4 : (NB <= 100 ? 2 : 1)))
stack_sorted_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FramePtrs<NB, T> fp,
                    const __grid_constant__ StackArgs a) {
    // [NB + 2][STPB] sorted columns + guard rows (median/MAD modes); the TMA stage is the same region
    extern __shared__ __align__(128) float col[];
    constexpr size_t PARK = (size_t)(NB + 2) * STPB;
    const T* const stage = reinterpret_cast<const T*>(col);
    uint64_t* const bar = reinterpret_cast<uint64_t*>(col + PARK);
    const int N = a.N;
    // Pad to NB with -inf / +inf split so that the real samples sit centred in
    // the sorted array: the median is then at the compile-time index NB/2-1
    // (and NB/2 for even N) whatever N is.
    const int npad = NB - N;
    const int nneg = npad >> 1;          // -inf pads; the other npad-nneg are +inf
    const int64_t pend = a.pix0 + a.npix;
    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            // pixels past the end of the band are out of bounds of the tensor map: zero-filled, and counted
            mbar_expect_tx(bar, (uint32_t)N * STPB * sizeof(T));
            tma_load_2d(col, &tmap, (int32_t)(a.pix0 + (int64_t)blockIdx.x * STPB), 0, bar, l2_evict_first_policy());
        }
        __syncthreads();
    }
    {
        // no early exit: the sort contains CTA barriers.  Threads past the end redo the last pixel
        // and skip the write.
        int64_t p = a.pix0 + (int64_t)blockIdx.x * STPB + threadIdx.x;
        const bool valid = p < pend;
        if (!valid) p = pend - 1;
        const uint32_t p32 = (uint32_t)p;    // host guarantees H*W < 2^32: one IMAD.WIDE per address
        float x[NB];
        float z = 0.f;
        double sum_all = 0.0;
        // Padding slots (i >= N, only possible for i >= NLO) are loaded like real ones -- the host points
        // them at frame 0 -- and replaced afterwards by uniform selects: no predicated loads.
        if constexpr (TMA) {
            while (!mbar_try_wait(bar, 0)) {}
#pragma unroll
            for (int i = 0; i < NB; ++i) x[i] = (i < NLO || i < N) ? staged_value(stage + i * STPB + threadIdx.x, a) : 0.f;
            // uint16 rows are half as wide as the float32 columns parked below: every thread must have read
            // its samples before any thread parks
            if constexpr (sizeof(T) != sizeof(float) && MODE != MODE_MED) __syncthreads();
        } else {
#pragma unroll
            for (int i = 0; i < NB; ++i) x[i] = load_sample(fp.p[i] + p32, a);
        }
        // Non-finite detection and, for the median/MAD mode, the sum of all samples (the mean when nothing
        // is clipped, ~99 % of the pixels).  float64 output: frame-order float64 sum, bit-identical to
        // np.nanmean.  float32 output: float32 sum of the samples shifted by the first frame (two packed
        // accumulators on the FMA pipe instead of 100 F2F + 100 DADD; the result is within ~1e-8 relative of
        // the float64 sum, far inside the 1e-6 contract).  Either sum is NaN / inf iff a sample is.
        const float pivot = x[0];
        float2 facc = make_float2(0.f, 0.f);
        auto real_or = [&](int i, float pad) { return (i < NLO || i < N) ? x[i] : pad; };   // i < NLO: compile-time true
        if (MODE == MODE_MEDMAD1) {
            if (a.out_f64) {
#pragma unroll
                for (int i = 0; i < NB; ++i) sum_all = __dadd_rn(sum_all, (double)real_or(i, -0.f));   // x + (-0) == x
                z = (float)(sum_all - sum_all);                                  // NaN iff a sample is NaN / inf
            } else {
                const float2 negpiv = make_float2(-pivot, -pivot);
#pragma unroll
                for (int i = 0; i + 1 < NB; i += 2)
                    facc = __fadd2_rn(facc, __fadd2_rn(make_float2(real_or(i, pivot), real_or(i + 1, pivot)), negpiv));
                z = (facc.x + facc.y) * 0.f;                                     // NaN iff a sample is NaN / inf
            }
        } else {
#pragma unroll
            for (int i = 0; i < NB; ++i) z = fmaf(real_or(i, 0.f), 0.f, z);
        }
#pragma unroll
        for (int i = NLO; i < NB; ++i) x[i] = (i < N) ? x[i] : ((i - N < nneg) ? -INFINITY : INFINITY);
        const bool nonfinite = (z != z);          // handled after the (barrier-carrying) sort

        sort_regs<NB, MIX>(x, a.one, a.minus_one);
        [&]() {
            if (!valid) return;
            if (nonfinite) { mark_pixel(a, p); return; }

            constexpr int C = NB / 2;
            const double med = (N & 1) ? (double)x[C - 1]
                                       : __dmul_rn(__dadd_rn((double)x[C - 1], (double)x[C]), 0.5);
            if (MODE == MODE_MED) {
                write_pixel(a, p, med, 0, (double)NAN, 0);
                return;
            }

            // Park the sorted column in shared memory ([row][thread]: conflict-free for
            // any per-thread row index) for the data-dependent selection below.  Row 0
            // and row NB+1 are -inf / +inf guards, so together with the +-inf padding
            // every row outside the real samples has an infinite deviation from the
            // median and the merge below needs no bounds checks.
            float* s = col + threadIdx.x + STPB;          // s[i * STPB] = sorted sample i, i in [-1, NB]
            s[-STPB] = -INFINITY;
            s[NB * STPB] = INFINITY;
#pragma unroll
            for (int i = 0; i < NB; ++i) s[i * STPB] = x[i];
            const int base = nneg;               // real samples occupy rows [base, base + N)
            const double mad = mad_parked(s, base, N, med);
            const double sd = __dmul_rn(MAD_TO_STD, mad);
            if (MODE == MODE_MEDUNC) {
                // Combiner.median_combine: uncertainty = mad_std(kept) / sqrt(n_kept), nothing clipped
                write_pixel(a, p, med, 0, __ddiv_rn(sd, __dsqrt_rn((double)N)), 0);
                return;
            }
            const double lo = __dsub_rn(med, __dmul_rn(sd, a.klo));
            const double hi = __dadd_rn(med, __dmul_rn(sd, a.khi));
            int sa = base, sb = base + N;
            while (sa < sb && (double)s[sa * STPB] < lo) ++sa;
            while (sa < sb && (double)s[(sb - 1) * STPB] > hi) --sb;
            const int nk = sb - sa;
#ifdef APGPU_DEBUG_MEDMAD
            if (p == a.pix0) printf("dbg N=%d NB=%d base=%d med=%.6f mad=%.6f lo=%.6f hi=%.6f sa=%d sb=%d klo=%f\n",
                N, NB, base, med, mad, lo, hi, sa, sb, a.klo);
#endif
            double mean;
            if (nk == N) {
                mean = a.out_f64 ? __ddiv_rn(sum_all, (double)N)
                                 : __dadd_rn((double)pivot, __ddiv_rn((double)(facc.x + facc.y), (double)N));
            } else {
                double acc = 0.0;
                for (int i = sa; i < sb; ++i) acc = __dadd_rn(acc, (double)s[i * STPB]);
                mean = __ddiv_rn(acc, (double)nk);      // nk >= 1: the median itself always survives
            }
            double unc = (double)NAN;
            if (a.uncert) {
                double acc = 0.0;
                for (int i = sa; i < sb; ++i) {
                    double d = __dsub_rn((double)s[i * STPB], mean);
                    acc = __dadd_rn(acc, __dmul_rn(d, d));
                }
                unc = __ddiv_rn(__dsqrt_rn(__ddiv_rn(acc, (double)nk)), __dsqrt_rn((double)nk));
            }
            write_pixel(a, p, mean, N - nk, unc, 0);

        }();
    }
}

// Plain median on equally spaced frames: the shared-memory stage is free again as soon as the threads
// hold their samples, so each CTA walks a run of consecutive 256-pixel tiles and the tensor-map copy of the
// NEXT tile is issued before the sort of the current one starts -- its HBM latency hides behind the
// ~1700 comparators (the one-tile-per-CTA kernel spent 18 % of its stall samples waiting for its tile).
template <int NB, int NLO, bool MIX, typename T>
__global__ void __launch_bounds__(STPB, (NB <= 32 ? 4 : (NB <= 100 ? 2 : 1)))
stack_median_tmap_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FramePtrs<NB, T> fp,
                         const __grid_constant__ StackArgs a) {
    extern __shared__ __align__(128) float col[];   // [NB][STPB] stage
    const T* const stage = reinterpret_cast<const T*>(col);
    uint64_t* bar = reinterpret_cast<uint64_t*>(col + (size_t)NB * STPB);
    const int N = a.N;
    const int npad = NB - N;
    const int nneg = npad >> 1;          // -inf pads; the other npad-nneg are +inf
    const int pix0 = (int)a.pix0, pend = pix0 + (int)a.npix;     // the host takes this path below 2^31 pixels
    const int ntiles = (int)((a.npix + STPB - 1) / STPB);
    int tile = (int)blockIdx.x * a.tiles_per_warp;
    const int tile_end = min(tile + a.tiles_per_warp, ntiles);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    auto issue = [&](int t) {
        // pixels past the end of the band are out of bounds of the tensor map: zero-filled, and counted
        mbar_expect_tx(bar, (uint32_t)N * STPB * sizeof(T));
        tma_load_2d(col, &tmap, pix0 + t * STPB, 0, bar, l2_evict_first_policy());
    };
    if (threadIdx.x == 0 && tile < tile_end) issue(tile);
    uint32_t parity = 0;
    for (; tile < tile_end; ++tile) {
        while (!mbar_try_wait(bar, parity)) {}
        parity ^= 1u;
        float x[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) x[i] = (i < NLO || i < N) ? staged_value(stage + i * STPB + threadIdx.x, a) : 0.f;
        float z = 0.f;
#pragma unroll
        for (int i = 0; i < NB; ++i) z = fmaf(x[i], 0.f, z);                  // NaN iff a sample is NaN / inf
        __syncthreads();                 // every thread has consumed its column: the stage can be refilled
        if (threadIdx.x == 0 && tile + 1 < tile_end) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(tile + 1);
        }
#pragma unroll
        for (int i = NLO; i < NB; ++i) x[i] = (i < N) ? x[i] : ((i - N < nneg) ? -INFINITY : INFINITY);
        sort_regs<NB, MIX>(x, a.one, a.minus_one);
        const int p = pix0 + tile * STPB + threadIdx.x;
        if (p < pend) {
            if (z != z) {
                mark_pixel(a, (int64_t)p);
            } else {
                constexpr int C = NB / 2;
                const double med = (N & 1) ? (double)x[C - 1]
                                           : __dmul_rn(__dadd_rn((double)x[C - 1], (double)x[C]), 0.5);
                write_pixel(a, (int64_t)p, med, 0, (double)NAN, 0);
            }
        }
    }
}

template <int NB, int NLO, int MODE, bool MIX, typename T>
int launch_sorted_mix(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<NB, T> fp;
    for (int i = 0; i < NB; ++i) fp.p[i] = i < a.N ? frames[i] : frames[0];   // padding: loaded, then replaced
    int64_t blocks = (a.npix + STPB - 1) / STPB;
    const size_t park = (size_t)(NB + 2) * STPB * sizeof(float);
    CUtensorMap tmap;
    const bool tma = stack_is_cube(frames, a.N, a.pix0 + a.npix) && (a.pix0 * (int64_t)sizeof(T)) % 16 == 0 &&
                     encode_stack_tensor_map(&tmap, frames[0], (uint64_t)(a.pix0 + a.npix), a.N,
                                             (uint64_t)((const char*)frames[1] - (const char*)frames[0]), STPB);
    if constexpr (MODE == MODE_MED) {
      if (tma) {
        StackArgs at = a;
        at.tiles_per_warp = stack_median_tiles_per_cta();
        const int64_t grid = (blocks + at.tiles_per_warp - 1) / at.tiles_per_warp;
        const size_t smem = (size_t)NB * STPB * sizeof(float) + sizeof(uint64_t);
        APGPU_CUDA(cudaFuncSetAttribute(stack_median_tmap_kernel<NB, NLO, MIX, T>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stack_median_tmap_kernel<NB, NLO, MIX, T><<<(unsigned)grid, STPB, smem, st>>>(tmap, fp, at);
        stack_note_staging(1);
        APGPU_LAUNCH_CHECK("stack_median_tmap_kernel");
        return APGPU_OK;
      }
    }
    if constexpr (MODE != MODE_MED) {
      if (tma) {
        const size_t smem = park + sizeof(uint64_t);
        APGPU_CUDA(cudaFuncSetAttribute(stack_sorted_kernel<NB, NLO, MODE, MIX, true, T>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stack_sorted_kernel<NB, NLO, MODE, MIX, true, T><<<(unsigned)blocks, STPB, smem, st>>>(tmap, fp, a);
        stack_note_staging(1);
        APGPU_LAUNCH_CHECK("stack_sorted_kernel");
        return APGPU_OK;
      }
    }
    {
        // any frame pointers, gathered loads
        memset(&tmap, 0, sizeof(tmap));
        const size_t smem = (MODE != MODE_MED) ? park : 0;
        if (smem > 48 * 1024)
            APGPU_CUDA(cudaFuncSetAttribute(stack_sorted_kernel<NB, NLO, MODE, MIX, false, T>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stack_sorted_kernel<NB, NLO, MODE, MIX, false, T><<<(unsigned)blocks, STPB, smem, st>>>(tmap, fp, a);
        stack_note_staging(0);
    }
    APGPU_LAUNCH_CHECK("stack_sorted_kernel");
    return APGPU_OK;
}

template <int NB, int NLO, int MODE, typename T>
int launch_sorted(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    // the mixed-pipe comparators win everywhere (N=64 median 66 -> 81 %) except in the 255-register
    // median/MAD kernel of the (160, 200] bucket (measured: tools/time_sorted.py)
    constexpr bool MIX = !(NB > 160 && MODE != MODE_MED);
    return launch_sorted_mix<NB, NLO, MODE, MIX, T>(frames, a, st);
}

#define SO_CASE(NB_, NLO_) \
    if constexpr (sorted_part_of(NB_) == PART) { if (nb == NB_) return launch_sorted<NB_, NLO_, MODE, T>(frames, a, st); }

template <int MODE, typename T, int PART>
int dispatch_sorted_part(int nb, const T* const* frames, const StackArgs& a, cudaStream_t st) {
    if constexpr (sorted_fine_buckets<MODE, T>()) {
        SO_CASE(4, 0) SO_CASE(8, 4) SO_CASE(12, 8) SO_CASE(16, 12) SO_CASE(20, 16) SO_CASE(24, 20)
        SO_CASE(32, 24) SO_CASE(40, 32) SO_CASE(48, 40) SO_CASE(56, 48) SO_CASE(64, 56) SO_CASE(72, 64)
        SO_CASE(80, 72) SO_CASE(90, 80) SO_CASE(100, 90) SO_CASE(112, 100) SO_CASE(128, 112)
        SO_CASE(160, 128) SO_CASE(200, 160)
    } else {
        SO_CASE(4, 0) SO_CASE(8, 4) SO_CASE(16, 8) SO_CASE(24, 16) SO_CASE(32, 24) SO_CASE(48, 32)
        SO_CASE(64, 48) SO_CASE(80, 64) SO_CASE(100, 80) SO_CASE(128, 100) SO_CASE(160, 128) SO_CASE(200, 160)
    }
    return APGPU_ERR_UNSUPPORTED;
}

}  // namespace apgpu_stack

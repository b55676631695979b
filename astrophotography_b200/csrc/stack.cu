// Frame-stack reducer (sm_100a): per-pixel median / sigma-clipped mean over N frames.
//
// Replaces ccdproc.combine(...) as called by the reference at
// AstroPhotography/scripts/ap_combine_darks.py:411-420 (settings :394-399) and
// generalises it to astropy.stats.sigma_clip's iterative kappa-sigma clip.
//
// Layout: N separate H x W float32 frames (frame-major; never transposed in
// HBM).  One thread owns one pixel; a warp reads 128 contiguous bytes of each
// frame, so every HBM access is a full coalesced line and each input byte is
// read exactly once (4*N + 5 B per pixel algorithmic traffic).
//
// Three kernel families behind one entry point:
//   generic<CAP>      every parameter combination, N <= 1024.  float64
//                     arithmetic in the oracle's operation order (explicitly
//                     rounded intrinsics, no FMA contraction): bit-identical to
//                     the numpy restatement.  Values live in local memory.
//   meanclip<NB,NLO>  iterative kappa-sigma clip about the MEAN with the
//                     population STD, then the mean of the survivors.
//                     Register-resident (N <= 200), float32 arithmetic on
//                     pivot-shifted values with a rigorous error bound: a pixel
//                     whose decision could differ from the float64 oracle
//                     (a sample within the bound of a clip threshold) is redone
//                     by the generic routine, so rejection maps are identical.
//   sorted<NB,MODE>   register-resident Batcher merge-exchange network
//                     (N <= 128): plain median (MODE_MED), or the reference's
//                     ApMasterCal setting -- one median/MAD clip pass then the
//                     mean (MODE_MEDMAD1) -- with the sorted column parked in
//                     shared memory for the data-dependent MAD selection.
//   A pixel holding NaN/inf samples leaves the fast kernels for the generic
//   routine, which owns the reference's non-finite semantics.
#include <float.h>
#include <math.h>

#include "apgpu_common.cuh"
#include "sort_networks.inc"

namespace {

constexpr double MAD_TO_STD = 1.482602218505602;   // astropy.stats.mad_std scale
constexpr int TPB = 128;                           // threads (= pixels) per block
constexpr int SMEM_MAX_BYTES = 227 * 1024;         // opt-in dynamic shared memory per CTA / per SM budget
constexpr int MEANCLIP_MAX_TAIL = 40;              // widest meanclip bucket (160, 200]

struct StackArgs {
    int N, method, maxiters, cen, dev;
    double klo, khi;
    int64_t pix0, npix;          // flat pixel range [pix0, pix0 + npix)
    void* out; int out_f64;
    void* nrej; int nrej_u16;
    void* uncert;
    uint8_t* allmasked;
    // meanclip<NB, NLO>: 1.0f for a real frame, 0.0f for padding, for frames NLO .. NB-1.  The
    // kernels load the padding slots unconditionally (the host points them at frame 0, the
    // staged kernels zero the rows) and multiply the pivot-shifted value by this mask: no
    // per-sample predicates in the load phase.
    float tailmask[MEANCLIP_MAX_TAIL];
};

template <int CAP> struct FramePtrs { const float* p[CAP]; };

__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= FLT_MAX; }

__device__ __forceinline__ void write_pixel(const StackArgs& a, int64_t p, double data, int nrej,
                                            double unc, int allm) {
    if (a.out_f64) reinterpret_cast<double*>(a.out)[p] = data;
    else st_stream(reinterpret_cast<float*>(a.out) + p, (float)data);
    if (a.nrej) {
        if (a.nrej_u16) reinterpret_cast<uint16_t*>(a.nrej)[p] = (uint16_t)nrej;
        else reinterpret_cast<uint8_t*>(a.nrej)[p] = (uint8_t)nrej;
    }
    if (a.uncert) {
        if (a.out_f64) reinterpret_cast<double*>(a.uncert)[p] = unc;
        else reinterpret_cast<float*>(a.uncert)[p] = (float)unc;
    }
    if (a.allmasked) a.allmasked[p] = (uint8_t)allm;
}

// ---------------------------------------------------------------------------
// generic routine: float64, oracle operation order
// ---------------------------------------------------------------------------
__device__ void shell_sort(float* s, int n) {
    const int gaps[8] = {701, 301, 132, 57, 23, 10, 4, 1};
    for (int g = 0; g < 8; ++g) {
        int gap = gaps[g];
        if (gap >= n && gap != 1) continue;
        for (int i = gap; i < n; ++i) {
            float t = s[i];
            int j = i;
            while (j >= gap && s[j - gap] > t) { s[j] = s[j - gap]; j -= gap; }
            s[j] = t;
        }
    }
}

// median of the sorted range s[sa, sb): nanmedian's (lo + hi) / 2 in float64
__device__ __forceinline__ double median_sorted(const float* s, int sa, int sb) {
    int m = sb - sa;
    if (m <= 0) return (double)NAN;
    double lo = (double)s[sa + ((m - 1) >> 1)];
    if (m & 1) return lo;
    double hi = (double)s[sa + (m >> 1)];
    return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

// median of |x - med| over the sorted range: the deviations left of the median
// grow towards sa and those right of it grow towards sb, so the k-th smallest
// comes out of a two-pointer merge -- no second sort.
__device__ __forceinline__ double mad_sorted(const float* s, int sa, int sb, double med) {
    int m = sb - sa;
    if (m <= 0) return (double)NAN;
    int l = sa + ((m - 1) >> 1), r = l + 1;
    int k1 = (m - 1) >> 1, k2 = m >> 1;
    double d1 = 0.0, d2 = 0.0;
    for (int t = 0; t <= k2; ++t) {
        double dl = (l >= sa) ? fabs(__dsub_rn((double)s[l], med)) : (double)INFINITY;
        double dr = (r < sb) ? fabs(__dsub_rn((double)s[r], med)) : (double)INFINITY;
        double d;
        if (dl <= dr) { d = dl; --l; } else { d = dr; ++r; }
        if (t == k1) d1 = d;
        if (t == k2) d2 = d;
    }
    return (m & 1) ? d1 : __dmul_rn(__dadd_rn(d1, d2), 0.5);
}

template <int CAP>
__device__ __noinline__ void generic_pixel(const FramePtrs<CAP>& fp, const StackArgs& a, int64_t p) {
    float v[CAP];      // frame order; NaN marks a sample that is not (or no longer) used
    float s[CAP];      // the used samples, ascending
    const int N = a.N;
    const bool clip = a.maxiters != 0;
    const bool need_sorted = (a.method != APGPU_METHOD_AVERAGE) ||
                             (clip && (a.cen == APGPU_CEN_MEDIAN || a.dev == APGPU_DEV_MAD_STD));
    int nk = 0;
    for (int i = 0; i < N; ++i) {
        float x = ld_stream(fp.p[i] + p);
        // sigma_clip rejects non-finite samples up front; without clipping the
        // nan-functions only skip NaN.
        bool ok = clip ? finite_f(x) : (x == x);
        v[i] = ok ? x : NAN;
        if (ok) { if (need_sorted) s[nk] = x; ++nk; }
    }
    if (need_sorted) shell_sort(s, nk);
    int sa = 0, sb = nk;

    auto mean_kept = [&](int cnt) -> double {          // np.nanmean: sequential sum / count
        double acc = 0.0;
        for (int i = 0; i < N; ++i) if (v[i] == v[i]) acc = __dadd_rn(acc, (double)v[i]);
        return __ddiv_rn(acc, (double)cnt);
    };
    auto std_kept = [&](int cnt, double avg) -> double {   // np.nanstd, ddof=0
        double acc = 0.0;
        for (int i = 0; i < N; ++i)
            if (v[i] == v[i]) { double d = __dsub_rn((double)v[i], avg); acc = __dadd_rn(acc, __dmul_rn(d, d)); }
        return __dsqrt_rn(__ddiv_rn(acc, (double)cnt));
    };

    if (clip) {
        int it = 0;
        while (a.maxiters < 0 || it < a.maxiters) {
            ++it;
            if (nk == 0) break;
            double avg = 0.0, med = 0.0;
            if (a.cen == APGPU_CEN_MEAN || a.dev == APGPU_DEV_STD) avg = mean_kept(nk);
            if (a.cen == APGPU_CEN_MEDIAN || a.dev == APGPU_DEV_MAD_STD) med = median_sorted(s, sa, sb);
            double c = (a.cen == APGPU_CEN_MEAN) ? avg : med;
            double sd = (a.dev == APGPU_DEV_STD) ? std_kept(nk, avg)
                                                 : __dmul_rn(MAD_TO_STD, mad_sorted(s, sa, sb, med));
            double lo = __dsub_rn(c, __dmul_rn(sd, a.klo));
            double hi = __dadd_rn(c, __dmul_rn(sd, a.khi));
            int changed = 0;
            for (int i = 0; i < N; ++i) {
                float x = v[i];
                if (x == x && ((double)x < lo || (double)x > hi)) { v[i] = NAN; ++changed; }
            }
            if (need_sorted) {
                while (sa < sb && (double)s[sa] < lo) ++sa;
                while (sa < sb && (double)s[sb - 1] > hi) --sb;
            }
            nk -= changed;
            if (changed == 0) break;
        }
    }

    double data, unc = (double)NAN;
    if (nk == 0) {
        data = (double)NAN;
    } else if (a.method == APGPU_METHOD_AVERAGE) {
        data = mean_kept(nk);
    } else if (a.method == APGPU_METHOD_MEDIAN) {
        data = median_sorted(s, sa, sb);
    } else if (a.method == APGPU_METHOD_MIN) {
        data = (double)s[sa];
    } else {
        data = (double)s[sb - 1];
    }
    if (a.uncert && nk > 0) {
        double dev;
        if (a.method == APGPU_METHOD_MEDIAN) {
            dev = __dmul_rn(MAD_TO_STD, mad_sorted(s, sa, sb, median_sorted(s, sa, sb)));
        } else {
            dev = std_kept(nk, mean_kept(nk));
        }
        unc = __ddiv_rn(dev, __dsqrt_rn((double)nk));
    }
    write_pixel(a, p, data, N - nk, unc, nk == 0);
}

template <int CAP>
__global__ void __launch_bounds__(TPB)
stack_generic_kernel(const __grid_constant__ FramePtrs<CAP> fp, const __grid_constant__ StackArgs a) {
    int64_t p = a.pix0 + (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (p >= a.pix0 + a.npix) return;
    generic_pixel<CAP>(fp, a, p);
}

// ---------------------------------------------------------------------------
// meanclip<NB, NLO>: kappa-sigma clip about the mean, N in (NLO, NB]
// ---------------------------------------------------------------------------
// i < NLO is known at compile time to be a real frame; only the NB-NLO tail
// elements carry a (warp-uniform) runtime predicate.
#define APGPU_ACTIVE(i) ((i) < NLO || (i) < N)

template <int K>
__device__ __forceinline__ float tree_sum(const float (&v)[K]) {
    float t[K];
#pragma unroll
    for (int k = 0; k < K; ++k) t[k] = v[k];
#pragma unroll
    for (int w = K / 2; w >= 1; w /= 2) {
#pragma unroll
        for (int k = 0; k < w; ++k) t[k] = t[2 * k] + t[2 * k + 1];
    }
    return t[0];
}

__device__ __forceinline__ float med3(float a, float b, float c) {
    return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}

// Sweep design (see DESIGN.md "meanclip"): the samples stay in registers as
// y = x - pivot, two per 64-bit register pair so that the Blackwell packed
// FADD2 / FFMA2 instructions update two samples per issue slot.  A rejected
// sample is overwritten with 0 (it then adds nothing to the running sums).  A
// sweep walks the samples in groups of 8: the common path per group is 4 FADD2
// (t = y - c), 4 FADD2 + 4 FFMA2 (sums), the |t| maximum, one compare -- no
// per-sample predicates or selects.  Only a group whose largest |y - c| reaches
// the inner clip bound is revisited, sample by sample, in a second "rare" pass.
constexpr int meanclip_min_blocks(int NB) {
    return NB <= 32 ? 6 : (NB <= 48 ? 5 : (NB <= 100 ? 4 : (NB <= 128 ? 3 : 2)));
}

// The per-pixel work, given the N raw samples of pixel p in y[] (padding = 0).
template <int NB, int NLO, bool SYM>
__device__ __forceinline__ void meanclip_pixel(float2 (&y)[NB / 2], const FramePtrs<NB>& fp,
                                               const StackArgs& a, const int64_t p) {
    static_assert(NB % 2 == 0, "meanclip buckets must be even");
    const int N = a.N;
    constexpr int NP = NB / 2;                         // register pairs
    constexpr int GP = 4;                              // pairs (8 samples) per group
    constexpr int NG = (NP + GP - 1) / GP;
    static_assert(NG <= 64, "flag word too small");

    // Pivot: median of the first three frames (robust to one outlier).  All
    // float32 arithmetic below is on y = x - pivot: sums stay small and the
    // variance is free of catastrophic cancellation.
    const float pivot = med3(y[0].x, y[0].y, y[1].x);
    const float2 negpiv = make_float2(-pivot, -pivot);
    float S1 = 0.f, S2 = 0.f;
#pragma unroll
    for (int gidx = 0; gidx < NG; ++gidx) {
        float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < GP; ++k) {
            const int j = gidx * GP + k;
            if (j < NP) {
                float2 d = __fadd2_rn(y[j], negpiv);
                if (2 * j >= NLO)                          // bucket tail: padding beyond N becomes y = 0
                    d = __fmul2_rn(d, make_float2(a.tailmask[2 * j - NLO], a.tailmask[2 * j + 1 - NLO]));
                y[j] = d;
                s1 = __fadd2_rn(s1, d);
                s2 = __ffma2_rn(d, d, s2);
            }
        }
        S1 += s1.x + s1.y;
        S2 += s2.x + s2.y;
    }
    // NaN input poisons S1/S2, inf input (or overflow) makes S2 infinite: the
    // generic routine owns those semantics.
    if (!(S2 <= FLT_MAX) || !(fabsf(S1) <= FLT_MAX)) { generic_pixel<NB>(fp, a, p); return; }

    int nk = N;
    const float klo = (float)a.klo, khi = (float)a.khi;
    const float kmax = fmaxf(klo, khi);
    bool uncertain = false;
    int it = 0;
    while (a.maxiters != 0 && (a.maxiters < 0 || it < a.maxiters)) {
        ++it;
        if (S2 == 0.f) break;            // all survivors equal the pivot: bounds [c,c], nothing to reject
        const float fn = (float)nk;
        const float c = S1 / fn;
        const float ex2 = S2 / fn;
        const float var = ex2 - c * c;
        const float sd = sqrtf(fmaxf(var, 0.f));
        // Bound on |threshold_f32 - threshold_exact| + |y_f32 - y_exact|:
        // group-wise summation (GP + 1 + NG terms deep), unit roundoff doubled for safety.
        const float u2 = 1.1920929e-7f;                       // 2^-23
        const float m = (float)(GP + NG + 9);
        const float g = m * u2 * (sqrtf(ex2) + 1.5f * kmax * ex2 / sd) + 6.f * u2 * (fabsf(c) + kmax * sd);
        if (!(g < 0.25f * kmax * sd)) { uncertain = true; break; }   // degenerate (var ~ 0): let float64 decide
        // inner (certainly kept inside) and outer (certainly rejected outside) bounds on t = y - c
        const float lo_in = -klo * sd + g, lo_out = -klo * sd - g;
        const float hi_in = khi * sd - g, hi_out = khi * sd + g;
        const float t_in = fminf(-lo_in, hi_in);              // symmetric inner bound on |t|
        const float ylo_out = c + lo_out, ylo_in = c + lo_in, yhi_in = c + hi_in, yhi_out = c + hi_out;
        // A rejected sample is overwritten with y = 0 (the pivot), so the pivot itself must sit
        // strictly inside the inner bounds: then zeros are never rejected (again) and add nothing.
        if (!(ylo_in < 0.f && yhi_in > 0.f)) { uncertain = true; break; }
        const float2 negc = make_float2(-c, -c);
        const int nk_before = nk;
        uint64_t flags = 0;              // bit g: group g holds a sample outside the inner bounds
        // test pass: tight straight-line code, no per-sample predicates, no sums (the sums of
        // the survivors only change when something is rejected)
#pragma unroll
        for (int gidx = 0; gidx < NG; ++gidx) {
            float tmax = 0.f, tmin = 0.f;
#pragma unroll
            for (int k = 0; k < GP; ++k) {
                const int j = gidx * GP + k;
                if (j < NP) {
                    const float2 t = __fadd2_rn(y[j], negc);
                    if (SYM) {
                        tmax = fmaxf(tmax, fmaxf(fabsf(t.x), fabsf(t.y)));
                    } else {
                        tmax = fmaxf(tmax, fmaxf(t.x, t.y));
                        tmin = fminf(tmin, fminf(t.x, t.y));
                    }
                }
            }
            const bool flagged = SYM ? (tmax >= t_in) : (tmax >= hi_in || tmin <= lo_in);
            if (flagged) flags |= (uint64_t)1 << gidx;
        }
        if (flags == 0) break;           // every survivor is certainly inside the bounds: converged
        // update pass: flagged groups sample by sample, the others with packed sums
        float n1 = 0.f, n2 = 0.f;
#pragma unroll
        for (int gidx = 0; gidx < NG; ++gidx) {
            if ((flags >> gidx) & 1) {
                float g1 = 0.f, g2 = 0.f, vmax = 0.f, vmin = 0.f;
#pragma unroll
                for (int k = 0; k < 2 * GP; ++k) {
                    const int i = gidx * 2 * GP + k;
                    if (i < NB) {
                        // compare y against bounds shifted by c (not t = y - c: keeps the
                        // compiler from holding every t of the test pass live in registers)
                        float v = (i & 1) ? y[i >> 1].y : y[i >> 1].x;
                        const bool keep = (v >= ylo_out) && (v <= yhi_out);
                        nk -= keep ? 0 : 1;                      // certainly rejected
                        v = keep ? v : 0.f;
                        if (i & 1) y[i >> 1].y = v; else y[i >> 1].x = v;
                        vmax = fmaxf(vmax, v);
                        vmin = fminf(vmin, v);
                        g1 += v;
                        g2 = fmaf(v, v, g2);
                    }
                }
                // a survivor inside the guard band: float64 must decide
                if (!(vmin > ylo_in && vmax < yhi_in)) uncertain = true;
                n1 += g1;
                n2 += g2;
            } else {
                float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < GP; ++k) {
                    const int j = gidx * GP + k;
                    if (j < NP) {
                        s1 = __fadd2_rn(s1, y[j]);
                        s2 = __ffma2_rn(y[j], y[j], s2);
                    }
                }
                n1 += s1.x + s1.y;
                n2 += s2.x + s2.y;
            }
        }
        if (uncertain) break;
        S1 = n1;
        S2 = n2;
        if (nk == nk_before || nk == 0) break;
    }
    if (uncertain || nk == 0) { generic_pixel<NB>(fp, a, p); return; }

    // mean of the survivors = pivot + sum(y)/nk (rejected samples are zeros).
    // float32 output: the float32 sums of the small shifted values are accurate
    // to ~1e-8 of max(|mean|, sigma).  float64 output: the shifted values are
    // summed in float64 (exact), leaving only the rounding of y = x - pivot
    // itself (none when x and pivot are within a factor 2, Sterbenz).
    double sum1 = (double)S1, sum2 = (double)S2;
    if (a.out_f64) {
        sum1 = 0.0; sum2 = 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const double d0 = (double)y[j].x, d1 = (double)y[j].y;
            sum1 = __dadd_rn(__dadd_rn(sum1, d0), d1);
            sum2 = __dadd_rn(__dadd_rn(sum2, __dmul_rn(d0, d0)), __dmul_rn(d1, d1));
        }
    }
    const double cy = __ddiv_rn(sum1, (double)nk);
    const double mean = __dadd_rn((double)pivot, cy);
    double unc_out = (double)NAN;
    if (a.uncert) {
        double var = __dsub_rn(__ddiv_rn(sum2, (double)nk), __dmul_rn(cy, cy));
        unc_out = __ddiv_rn(__dsqrt_rn(var > 0.0 ? var : 0.0), __dsqrt_rn((double)nk));
    }
    write_pixel(a, p, mean, N - nk, unc_out, 0);
}

// Direct kernel: one block per 128-pixel tile, samples loaded straight from
// global memory.
template <int NB, int NLO, bool SYM>
__global__ void __launch_bounds__(TPB, meanclip_min_blocks(NB))
stack_meanclip_kernel(const __grid_constant__ FramePtrs<NB> fp, const __grid_constant__ StackArgs a) {
    const int64_t p = a.pix0 + (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (p >= a.pix0 + a.npix) return;
    const uint32_t p32 = (uint32_t)p;                  // host guarantees H*W < 2^32
    float2 y[NB / 2];
#pragma unroll
    for (int j = 0; j < NB / 2; ++j) {                 // padding slots point at frame 0 (masked later)
        y[j].x = ld_stream(fp.p[2 * j] + p32);
        y[j].y = ld_stream(fp.p[2 * j + 1] + p32);
    }
    meanclip_pixel<NB, NLO, SYM>(y, fp, a, p);
}

// ---------------------------------------------------------------------------
// TMA-staged persistent kernel (opt-in: APGPU_STACK_USE_TMA)
// ---------------------------------------------------------------------------
// meanclip_min_blocks(NB)/2 persistent 256-thread CTAs per SM walk the 256-pixel
// tiles of the band.  For each tile one warp issues N bulk asynchronous copies
// (cp.async.bulk global -> shared, 1 KB contiguous of each frame, completion
// counted on an mbarrier); the threads pull their column out of shared memory
// into registers (conflict-free LDS), release the stage with one __syncthreads,
// and the copies for the CTA's NEXT tile are issued before the arithmetic on the
// current one starts.  Measured on B200 (profiles/): correct, but 20 % SLOWER
// than the direct kernel at N=100 -- the per-tile CTA barrier makes every warp
// wait for the CTA's slowest pixel (clip iteration counts differ per pixel) --
// so the dispatcher only uses it on request.
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) {
    return (uint32_t)__cvta_generic_to_shared(ptr);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "APGPU_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra APGPU_DONE;\n"
        "bra APGPU_WAIT;\n"
        "APGPU_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Threads (= pixels) per CTA of the TMA-staged kernel: 256, so that every bulk copy
// moves 1 KB (512-byte copies are TMA-issue bound).
constexpr int TTPB = 256;

template <int NB>
__device__ __forceinline__ void issue_tile_copies(const FramePtrs<NB>& fp, int N, int64_t pix, float* stage,
                                                  uint64_t* bar) {
    // called by warp 0
    const int lane = threadIdx.x & 31;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)N * TTPB * sizeof(float));
    __syncwarp();
    for (int i = lane; i < N; i += 32)
        bulk_copy_g2s(stage + i * TTPB, fp.p[i] + pix, TTPB * sizeof(float), bar);
}

template <int NB, int NLO, bool SYM>
__global__ void __launch_bounds__(TTPB, meanclip_min_blocks(NB) / 2)
stack_meanclip_tma_kernel(const __grid_constant__ FramePtrs<NB> fp, const __grid_constant__ StackArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);                       // [NB][TTPB]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NB * TTPB * sizeof(float));
    const int N = a.N;
    const int64_t ntiles = a.npix / TTPB;                                    // full tiles only (host launches the tail)
    if (threadIdx.x == 0) mbar_init(bar, 1);
    for (int i = N * TTPB + threadIdx.x; i < NB * TTPB; i += TTPB) stage[i] = 0.f;   // padding rows: never copied into
    __syncthreads();
    int64_t tile = blockIdx.x;
    uint32_t parity = 0;
    if (tile < ntiles && threadIdx.x < 32) issue_tile_copies<NB>(fp, N, a.pix0 + tile * TTPB, stage, bar);
    for (; tile < ntiles; tile += gridDim.x) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        float2 y[NB / 2];
#pragma unroll
        for (int j = 0; j < NB / 2; ++j) {
            y[j].x = stage[(2 * j) * TTPB + threadIdx.x];
            y[j].y = stage[(2 * j + 1) * TTPB + threadIdx.x];
        }
        __syncthreads();                                                     // every thread has drained the stage
        const int64_t next = tile + gridDim.x;
        if (next < ntiles && threadIdx.x < 32) issue_tile_copies<NB>(fp, N, a.pix0 + next * TTPB, stage, bar);
        meanclip_pixel<NB, NLO, SYM>(y, fp, a, a.pix0 + tile * TTPB + threadIdx.x);
    }
}

// ---------------------------------------------------------------------------
// cp.async-staged persistent kernel: warp-granular software pipeline
// ---------------------------------------------------------------------------
// Every warp is its own pipeline: it owns a private [N][32-pixel] shared-memory
// stage, fills it with 16-byte asynchronous copies (cp.async / LDGSTS: one warp
// instruction moves the 128-byte rows of four frames), drains it into registers,
// immediately re-arms it with the copies for its NEXT tile and only then does
// the arithmetic.  No CTA barrier anywhere, so a warp never waits for another
// warp's slow pixel (the flaw of the CTA-wide TMA variant above), the HBM latency
// of tile t+1 hides behind the compute of tile t without extra registers, and
// the instruction stream has 1/4 of the load instructions and no per-sample
// 64-bit address arithmetic.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int WT = 32;      // pixels per warp tile

__device__ __forceinline__ void issue_warp_tile(const float* const* ptab, int N, int64_t pix, float* stage, int lane) {
    const int sub = lane >> 3;            // which of the 4 frames of this instruction
    const int off = (lane & 7) * 4;       // 4 pixels (16 bytes) per lane
    for (int i0 = 0; i0 < N; i0 += 4) {
        const int i = i0 + sub;
        if (i < N) cp_async16(stage + i * WT + off, ptab[i] + pix + off);
    }
    cp_async_commit();
}

template <int NB, int NLO, bool SYM>
__global__ void __launch_bounds__(TPB, meanclip_min_blocks(NB))
stack_meanclip_cpasync_kernel(const __grid_constant__ FramePtrs<NB> fp, const __grid_constant__ StackArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float** ptab = reinterpret_cast<const float**>(smem_raw);           // [NB] frame pointers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* stage = reinterpret_cast<float*>(smem_raw + (size_t)NB * sizeof(float*)) + (size_t)warp * NB * WT;
    const int N = a.N;
    for (int i = threadIdx.x; i < N; i += TPB) ptab[i] = fp.p[i];
    for (int i = N * WT + lane; i < NB * WT; i += 32) stage[i] = 0.f;         // padding rows: never copied into
    __syncthreads();
    const int64_t ntiles = a.npix / WT;                                       // full warp tiles (host launches the tail)
    const int64_t nwarps = (int64_t)gridDim.x * (TPB / 32);
    int64_t tile = (int64_t)blockIdx.x * (TPB / 32) + warp;
    if (tile < ntiles) issue_warp_tile(ptab, N, a.pix0 + tile * WT, stage, lane);
    for (; tile < ntiles; tile += nwarps) {
        cp_async_wait_all();
        __syncwarp();                                                         // every lane's copies are visible
        float2 y[NB / 2];
#pragma unroll
        for (int j = 0; j < NB / 2; ++j) {
            y[j].x = stage[(2 * j) * WT + lane];
            y[j].y = stage[(2 * j + 1) * WT + lane];
        }
        __syncwarp();                                                         // stage drained by every lane
        const int64_t next = tile + nwarps;
        if (next < ntiles) issue_warp_tile(ptab, N, a.pix0 + next * WT, stage, lane);
        meanclip_pixel<NB, NLO, SYM>(y, fp, a, a.pix0 + tile * WT + lane);
    }
}

// ---------------------------------------------------------------------------
// meanclip_smem<SYM, CAP>: the same algorithm with the pixel's samples parked in
// shared memory instead of registers, any N that fits (N <= ~450).
// ---------------------------------------------------------------------------
// Each thread owns one pixel and one shared-memory column of float4 groups
// ([group][thread] layout: 128-bit accesses, conflict-free).  Because shared
// memory can be indexed dynamically, every pass is a real loop: the code is a
// few hundred instructions whatever N is (the register kernels unroll N-fold
// and become instruction-fetch bound beyond ~64 frames), registers stay low,
// and N is a run-time value.  Sweeps follow the meanclip design above: groups
// of 8 samples, branch only when a group's largest |y - c| reaches the inner
// bound.
constexpr int SM_U = 4;      // float4 groups (16 frames) loaded per unrolled step of the load loop

template <bool SYM, int CAP>
__global__ void __launch_bounds__(TPB)
stack_meanclip_smem_kernel(const __grid_constant__ FramePtrs<CAP> fp, const __grid_constant__ StackArgs a) {
    extern __shared__ float4 tile4[];                  // [n4e][TPB]
    const int64_t p = a.pix0 + (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (p >= a.pix0 + a.npix) return;
    const uint32_t p32 = (uint32_t)p;
    const int N = a.N;
    const int n4 = (N + 3) >> 2;
    const int n4e = (n4 + 1) & ~1;                     // even number of groups: sweeps take two per step
    float4* col = tile4 + threadIdx.x;                 // col[g * TPB] = samples 4g .. 4g+3 of this pixel

    const float pivot = med3(ld_stream(fp.p[0] + p32), ld_stream(fp.p[1] + p32), ld_stream(fp.p[2] + p32));
    float S1 = 0.f, S2 = 0.f;
    for (int g0 = 0; g0 < n4e; g0 += SM_U) {
        float4 v[SM_U];
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
            const int i = 4 * (g0 + u);
            // samples beyond N are padded with the pivot: y = 0, which adds nothing anywhere
            v[u].x = (i + 0 < N) ? ld_stream(fp.p[i + 0] + p32) : pivot;
            v[u].y = (i + 1 < N) ? ld_stream(fp.p[i + 1] + p32) : pivot;
            v[u].z = (i + 2 < N) ? ld_stream(fp.p[i + 2] + p32) : pivot;
            v[u].w = (i + 3 < N) ? ld_stream(fp.p[i + 3] + p32) : pivot;
        }
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
            if (g0 + u < n4e) {
                float4 y;
                y.x = v[u].x - pivot; y.y = v[u].y - pivot; y.z = v[u].z - pivot; y.w = v[u].w - pivot;
                const float g1 = (y.x + y.y) + (y.z + y.w);
                const float g2 = fmaf(y.w, y.w, fmaf(y.z, y.z, fmaf(y.y, y.y, y.x * y.x)));
                S1 += g1;
                S2 += g2;
                col[(g0 + u) * TPB] = y;
            }
        }
    }
    if (!(S2 <= FLT_MAX) || !(fabsf(S1) <= FLT_MAX)) { generic_pixel<CAP>(fp, a, p); return; }

    int nk = N;
    const float klo = (float)a.klo, khi = (float)a.khi;
    const float kmax = fmaxf(klo, khi);
    bool uncertain = false;
    int it = 0;
    while (a.maxiters != 0 && (a.maxiters < 0 || it < a.maxiters)) {
        ++it;
        if (S2 == 0.f) break;
        const float fn = (float)nk;
        const float c = S1 / fn;
        const float ex2 = S2 / fn;
        const float var = ex2 - c * c;
        const float sd = sqrtf(fmaxf(var, 0.f));
        const float u2 = 1.1920929e-7f;                       // 2^-23
        const float m = (float)(n4 + 12);                     // group-wise summation depth, doubled roundoff
        const float g = m * u2 * (sqrtf(ex2) + 1.5f * kmax * ex2 / sd) + 6.f * u2 * (fabsf(c) + kmax * sd);
        if (!(g < 0.25f * kmax * sd)) { uncertain = true; break; }
        const float lo_in = -klo * sd + g, lo_out = -klo * sd - g;
        const float hi_in = khi * sd - g, hi_out = khi * sd + g;
        const float t_in = fminf(-lo_in, hi_in);
        const float ylo_out = c + lo_out, ylo_in = c + lo_in, yhi_in = c + hi_in, yhi_out = c + hi_out;
        if (!(ylo_in < 0.f && yhi_in > 0.f)) { uncertain = true; break; }   // zeros (rejected/padding) must stay inside
        const int nk_before = nk;
        float n1 = 0.f, n2 = 0.f;
        for (int gq = 0; gq < n4e; gq += 2) {
            float4 q0 = col[gq * TPB], q1 = col[(gq + 1) * TPB];
            float yv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            float g1 = 0.f, g2 = 0.f, tmax = 0.f, tmin = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t = yv[k] - c;
                if (SYM) {
                    tmax = fmaxf(tmax, fabsf(t));
                } else {
                    tmax = fmaxf(tmax, t);
                    tmin = fminf(tmin, t);
                }
                g1 += yv[k];
                g2 = fmaf(yv[k], yv[k], g2);
            }
            const bool flagged = SYM ? (tmax >= t_in) : (tmax >= hi_in || tmin <= lo_in);
            if (flagged) {
                g1 = 0.f; g2 = 0.f;
                float vmax = 0.f, vmin = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float v = yv[k];
                    const bool keep = (v >= ylo_out) && (v <= yhi_out);
                    nk -= keep ? 0 : 1;
                    v = keep ? v : 0.f;
                    yv[k] = v;
                    vmax = fmaxf(vmax, v);
                    vmin = fminf(vmin, v);
                    g1 += v;
                    g2 = fmaf(v, v, g2);
                }
                if (!(vmin > ylo_in && vmax < yhi_in)) uncertain = true;
                col[gq * TPB] = make_float4(yv[0], yv[1], yv[2], yv[3]);
                col[(gq + 1) * TPB] = make_float4(yv[4], yv[5], yv[6], yv[7]);
            }
            n1 += g1;
            n2 += g2;
        }
        if (uncertain) break;
        S1 = n1;
        S2 = n2;
        if (nk == nk_before || nk == 0) break;
    }
    if (uncertain || nk == 0) { generic_pixel<CAP>(fp, a, p); return; }

    double sum1 = (double)S1, sum2 = (double)S2;
    if (a.out_f64) {
        sum1 = 0.0; sum2 = 0.0;
        for (int gq = 0; gq < n4e; ++gq) {
            const float4 q = col[gq * TPB];
            const float yv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double d = (double)yv[k];
                sum1 = __dadd_rn(sum1, d);
                sum2 = __dadd_rn(sum2, __dmul_rn(d, d));
            }
        }
    }
    const double cy = __ddiv_rn(sum1, (double)nk);
    const double mean = __dadd_rn((double)pivot, cy);
    double unc_out = (double)NAN;
    if (a.uncert) {
        double var = __dsub_rn(__ddiv_rn(sum2, (double)nk), __dmul_rn(cy, cy));
        unc_out = __ddiv_rn(__dsqrt_rn(var > 0.0 ? var : 0.0), __dsqrt_rn((double)nk));
    }
    write_pixel(a, p, mean, N - nk, unc_out, 0);
}

// ---------------------------------------------------------------------------
// sorted<NB, NLO, MODE>: Batcher network in registers, N in (NLO, NB]
// ---------------------------------------------------------------------------
constexpr int STPB = 256;         // threads per CTA of the sorted kernels (lock-stepped, see sort_regs)
constexpr int MODE_MED = 0;       // method=median, no clipping
constexpr int MODE_MEDMAD1 = 1;   // one median/MAD clip pass, then the mean (ApMasterCal)

#define CE_X(i, j) { float lo_ = fminf(x[i], x[j]); float hi_ = fmaxf(x[i], x[j]); x[i] = lo_; x[j] = hi_; }

// Every 128 comparators the network has a CTA barrier: the 8 warps of a CTA walk the ~35 KB of
// straight-line code together, so one instruction-cache fill serves all of them (ncu before:
// `no_instruction` was the top stall of the median/MAD kernel).  The network is branch-free
// and data-independent, so the barrier costs no load imbalance.
#define SY_X() __syncthreads();
template <int NB> __device__ __forceinline__ void sort_regs(float (&x)[NB]);
#define APGPU_DEF_SORT(n) \
    template <> __device__ __forceinline__ void sort_regs<n>(float (&x)[n]) { APGPU_SORTNET_##n(CE_X, SY_X) }
APGPU_DEF_SORT(4) APGPU_DEF_SORT(8) APGPU_DEF_SORT(12) APGPU_DEF_SORT(16) APGPU_DEF_SORT(20)
APGPU_DEF_SORT(24) APGPU_DEF_SORT(32) APGPU_DEF_SORT(40) APGPU_DEF_SORT(48) APGPU_DEF_SORT(56)
APGPU_DEF_SORT(64) APGPU_DEF_SORT(72) APGPU_DEF_SORT(80) APGPU_DEF_SORT(90) APGPU_DEF_SORT(100)
APGPU_DEF_SORT(112) APGPU_DEF_SORT(128)

template <int NB, int NLO, int MODE>
__global__ void __launch_bounds__(STPB, (NB <= 32 ? 4 : (NB <= 100 ? 2 : 1)))
stack_sorted_kernel(const __grid_constant__ FramePtrs<NB> fp, const __grid_constant__ StackArgs a) {
    extern __shared__ float col[];       // MODE_MEDMAD1: [NB + 2][TPB] sorted columns + guard rows
    // no early exit: the sort contains CTA barriers.  Threads past the end redo the last pixel
    // and skip the write.
    const int64_t pend = a.pix0 + a.npix;
    int64_t p = a.pix0 + (int64_t)blockIdx.x * STPB + threadIdx.x;
    const bool valid = p < pend;
    if (!valid) p = pend - 1;
    const int N = a.N;
    // Pad to NB with -inf / +inf split so that the real samples sit centred in
    // the sorted array: the median is then at the compile-time index NB/2-1
    // (and NB/2 for even N) whatever N is.
    const int npad = NB - N;
    const int nneg = npad >> 1;          // -inf pads; the other npad-nneg are +inf
    const uint32_t p32 = (uint32_t)p;    // host guarantees H*W < 2^32: one IMAD.WIDE per address
    float x[NB];
    float z = 0.f;
    double sum_all = 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (APGPU_ACTIVE(i)) {
            x[i] = ld_stream(fp.p[i] + p32);
        } else {
            x[i] = (i - N < nneg) ? -INFINITY : INFINITY;
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        if (APGPU_ACTIVE(i)) {
            z = fmaf(x[i], 0.f, z);
            if (MODE == MODE_MEDMAD1) sum_all = __dadd_rn(sum_all, (double)x[i]);   // frame order, as nanmean
        }
    }
    const bool nonfinite = (z != z);          // handled after the (barrier-carrying) sort

    sort_regs<NB>(x);
    if (!valid) return;
    if (nonfinite) { generic_pixel<NB>(fp, a, p); return; }

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
    // MAD = median of |x - med|.  Left of the median the deviations grow towards row
    // `base`, right of it towards row `base+N`: two sorted lists,
    //     L[j] = med - s[l0 - j]   (j = 0 .. nL-1),   R[j] = s[l0 + 1 + j] - med   (j = 0 .. nR-1),
    // whose (k+1)-th smallest element is found by bisecting on how many come from L
    // (O(log N) shared-memory reads, float64, exact).  Guard rows / +-inf padding give
    // every out-of-range index an infinite deviation.
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
    const double mad = (N & 1) ? d1 : __dmul_rn(__dadd_rn(d1, d2), 0.5);
    const double sd = __dmul_rn(MAD_TO_STD, mad);
    const double lo = __dsub_rn(med, __dmul_rn(sd, a.klo));
    const double hi = __dadd_rn(med, __dmul_rn(sd, a.khi));
    int sa = base, sb = base + N;
    while (sa < sb && (double)s[sa * STPB] < lo) ++sa;
    while (sa < sb && (double)s[(sb - 1) * STPB] > hi) --sb;
    const int nk = sb - sa;
#ifdef APGPU_DEBUG_MEDMAD
    if (p == a.pix0) printf("dbg N=%d NB=%d base=%d med=%.6f d1=%.6f d2=%.6f mad=%.6f lo=%.6f hi=%.6f sa=%d sb=%d klo=%f\n",
        N, NB, base, med, d1, d2, mad, lo, hi, sa, sb, a.klo);
#endif
    double mean;
    if (nk == N) {
        mean = __ddiv_rn(sum_all, (double)N);
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
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
enum Family { FAM_GENERIC = 0, FAM_MEANCLIP = 1, FAM_SORT_MED = 2, FAM_SORT_MEDMAD1 = 3, FAM_MEANCLIP_SMEM = 4 };

constexpr int MEANCLIP_SMEM_MAX_N = 4 * ((SMEM_MAX_BYTES / (TPB * 16)) & ~1);   // 452
constexpr int MEANCLIP_REG_DEFAULT_MAX_N = 200;  // measured: the register kernel wins wherever it exists (bench.py variants)

struct Bucket { int nb, nlo; };
// (NLO, NB] buckets.  meanclip goes to 200 frames in registers; sorted to 128.
const Bucket MEANCLIP_BUCKETS[] = {{8, 2}, {16, 8}, {24, 16}, {32, 24}, {48, 32}, {64, 48}, {80, 64},
                                   {100, 80}, {128, 100}, {160, 128}, {200, 160}};
const Bucket SORT_BUCKETS[] = {{4, 0}, {8, 4}, {12, 8}, {16, 12}, {20, 16}, {24, 20}, {32, 24}, {40, 32},
                               {48, 40}, {56, 48}, {64, 56}, {72, 64}, {80, 72}, {90, 80}, {100, 90},
                               {112, 100}, {128, 112}};

template <size_t K>
const Bucket* find_bucket(const Bucket (&b)[K], int N) {
    for (size_t i = 0; i < K; ++i) if (N > b[i].nlo && N <= b[i].nb) return &b[i];
    return nullptr;
}

Family choose_family(int N, int method, double klo, double khi, int maxiters, int cen, int dev,
                     bool want_uncert, int flags, const Bucket** bucket) {
    *bucket = nullptr;
    if (flags & APGPU_STACK_FORCE_GENERIC) return FAM_GENERIC;
    if (method == APGPU_METHOD_AVERAGE && (maxiters == 0 || (cen == APGPU_CEN_MEAN && dev == APGPU_DEV_STD)) &&
        N >= 3 && klo > 0.0 && khi > 0.0 && klo < 1e6 && khi < 1e6) {
        const Bucket* rb = find_bucket(MEANCLIP_BUCKETS, N);
        const bool smem_ok = N <= MEANCLIP_SMEM_MAX_N;
        bool use_reg = rb && (N <= MEANCLIP_REG_DEFAULT_MAX_N || !smem_ok);
        if ((flags & APGPU_STACK_PREFER_REGISTERS) && rb) use_reg = true;
        if ((flags & APGPU_STACK_PREFER_SHARED) && smem_ok) use_reg = false;
        if (use_reg) { *bucket = rb; return FAM_MEANCLIP; }
        if (smem_ok) return FAM_MEANCLIP_SMEM;
    }
    if (method == APGPU_METHOD_MEDIAN && maxiters == 0 && !want_uncert) {
        if ((*bucket = find_bucket(SORT_BUCKETS, N))) return FAM_SORT_MED;
    }
    if (method == APGPU_METHOD_AVERAGE && maxiters == 1 && cen == APGPU_CEN_MEDIAN &&
        dev == APGPU_DEV_MAD_STD && N >= 2) {
        if ((*bucket = find_bucket(SORT_BUCKETS, N))) return FAM_SORT_MEDMAD1;
    }
    return FAM_GENERIC;
}

template <int CAP>
int launch_generic(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<CAP> fp;
    for (int i = 0; i < CAP; ++i) fp.p[i] = i < a.N ? frames[i] : nullptr;
    int64_t blocks = (a.npix + TPB - 1) / TPB;
    stack_generic_kernel<CAP><<<(unsigned)blocks, TPB, 0, st>>>(fp, a);
    APGPU_LAUNCH_CHECK("stack_generic_kernel");
    return APGPU_OK;
}

template <int NB, int NLO, bool SYM>
int launch_meanclip_sym(const FramePtrs<NB>& fp, const StackArgs& a, int staging, cudaStream_t st) {
    // staging: 0 direct global loads, 1 CTA-wide TMA bulk copies, 2 warp-granular cp.async pipeline
    StackArgs rest = a;
    if (staging == 2) {
        const int64_t ntiles = a.npix / WT;
        if (ntiles > 0) {
            const size_t smem = (size_t)NB * sizeof(float*) + (size_t)(TPB / 32) * NB * WT * sizeof(float);
            APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_cpasync_kernel<NB, NLO, SYM>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int64_t grid = (int64_t)APGPU_NUM_SMS * meanclip_min_blocks(NB);
            const int64_t need = (ntiles + TPB / 32 - 1) / (TPB / 32);
            if (grid > need) grid = need;
            stack_meanclip_cpasync_kernel<NB, NLO, SYM><<<(unsigned)grid, TPB, smem, st>>>(fp, a);
            APGPU_LAUNCH_CHECK("stack_meanclip_cpasync_kernel");
        }
        rest.pix0 = a.pix0 + ntiles * WT;            // the < 32-pixel tail goes through the direct kernel
        rest.npix = a.npix - ntiles * WT;
    }
    if constexpr (meanclip_min_blocks(NB) % 2 == 0) {
        if (staging == 1) {
            const int64_t ntiles = a.npix / TTPB;
            if (ntiles > 0) {
                const size_t smem = (size_t)NB * TTPB * sizeof(float) + 16;
                APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_tma_kernel<NB, NLO, SYM>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int64_t grid = (int64_t)APGPU_NUM_SMS * (meanclip_min_blocks(NB) / 2);
                if (grid > ntiles) grid = ntiles;
                stack_meanclip_tma_kernel<NB, NLO, SYM><<<(unsigned)grid, TTPB, smem, st>>>(fp, a);
                APGPU_LAUNCH_CHECK("stack_meanclip_tma_kernel");
            }
            rest.pix0 = a.pix0 + ntiles * TTPB;      // the < 256-pixel tail goes through the direct kernel
            rest.npix = a.npix - ntiles * TTPB;
        }
    }
    if (rest.npix > 0) {
        int64_t blocks = (rest.npix + TPB - 1) / TPB;
        stack_meanclip_kernel<NB, NLO, SYM><<<(unsigned)blocks, TPB, 0, st>>>(fp, rest);
        APGPU_LAUNCH_CHECK("stack_meanclip_kernel");
    }
    return APGPU_OK;
}

template <int NB, int NLO>
int launch_meanclip(const float* const* frames, const StackArgs& a_in, cudaStream_t st, int flags) {
    static_assert(NB - NLO <= MEANCLIP_MAX_TAIL, "tail mask too small");
    FramePtrs<NB> fp;
    for (int i = 0; i < NB; ++i) fp.p[i] = i < a_in.N ? frames[i] : frames[0];   // padding: loaded, then masked
    StackArgs a = a_in;
    for (int i = NLO; i < NB; ++i) a.tailmask[i - NLO] = i < a.N ? 1.f : 0.f;
    // asynchronous copies need 16-byte aligned sources: frame base + first pixel of the band
    // default (measured, bench.py variants): the warp-granular cp.async pipeline wins for the
    // shorter stacks (N=30: +8 %), direct loads are level or slightly ahead from N~80 up
    int staging = (flags & APGPU_STACK_USE_TMA) ? 1 : ((flags & APGPU_STACK_DIRECT_LOADS) ? 0 : (NB <= 64 ? 2 : 0));
    if (flags & APGPU_STACK_USE_CPASYNC) staging = 2;
    // the per-warp stages must leave room for meanclip_min_blocks CTAs per SM
    const size_t smem_cta = (size_t)NB * sizeof(float*) + (size_t)(TPB / 32) * NB * WT * sizeof(float);
    if (staging == 2 && smem_cta * meanclip_min_blocks(NB) > (size_t)SMEM_MAX_BYTES) staging = 0;
    for (int i = 0; i < a.N; ++i)
        if (!apgpu_aligned(frames[i] + a.pix0, 16)) staging = 0;
    if ((float)a.klo == (float)a.khi) return launch_meanclip_sym<NB, NLO, true>(fp, a, staging, st);
    return launch_meanclip_sym<NB, NLO, false>(fp, a, staging, st);
}

template <int NB, int NLO, int MODE>
int launch_sorted(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<NB> fp;
    for (int i = 0; i < NB; ++i) fp.p[i] = i < a.N ? frames[i] : nullptr;
    int64_t blocks = (a.npix + STPB - 1) / STPB;
    size_t smem = (MODE == MODE_MEDMAD1) ? (size_t)(NB + 2) * STPB * sizeof(float) : 0;
    if (smem > 48 * 1024)
        APGPU_CUDA(cudaFuncSetAttribute(stack_sorted_kernel<NB, NLO, MODE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stack_sorted_kernel<NB, NLO, MODE><<<(unsigned)blocks, STPB, smem, st>>>(fp, a);
    APGPU_LAUNCH_CHECK("stack_sorted_kernel");
    return APGPU_OK;
}

template <int CAP>
int launch_meanclip_smem(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<CAP> fp;
    for (int i = 0; i < CAP; ++i) fp.p[i] = i < a.N ? frames[i] : nullptr;
    int64_t blocks = (a.npix + TPB - 1) / TPB;
    const int n4e = (((a.N + 3) >> 2) + 1) & ~1;
    size_t smem = (size_t)n4e * TPB * sizeof(float4);
    const bool sym = (float)a.klo == (float)a.khi;
    if (sym) {
        APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_smem_kernel<true, CAP>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stack_meanclip_smem_kernel<true, CAP><<<(unsigned)blocks, TPB, smem, st>>>(fp, a);
    } else {
        APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_smem_kernel<false, CAP>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stack_meanclip_smem_kernel<false, CAP><<<(unsigned)blocks, TPB, smem, st>>>(fp, a);
    }
    APGPU_LAUNCH_CHECK("stack_meanclip_smem_kernel");
    return APGPU_OK;
}

#define MC_CASE(NB_, NLO_) if (b->nb == NB_) return launch_meanclip<NB_, NLO_>(frames, a, st, flags);
#define SO_CASE(NB_, NLO_) if (b->nb == NB_) return launch_sorted<NB_, NLO_, MODE>(frames, a, st);

int dispatch_meanclip(const Bucket* b, const float* const* frames, const StackArgs& a, cudaStream_t st, int flags) {
    MC_CASE(8, 2) MC_CASE(16, 8) MC_CASE(24, 16) MC_CASE(32, 24) MC_CASE(48, 32) MC_CASE(64, 48)
    MC_CASE(80, 64) MC_CASE(100, 80) MC_CASE(128, 100) MC_CASE(160, 128) MC_CASE(200, 160)
    return APGPU_ERR_UNSUPPORTED;
}

template <int MODE>
int dispatch_sorted(const Bucket* b, const float* const* frames, const StackArgs& a, cudaStream_t st) {
    SO_CASE(4, 0) SO_CASE(8, 4) SO_CASE(12, 8) SO_CASE(16, 12) SO_CASE(20, 16) SO_CASE(24, 20)
    SO_CASE(32, 24) SO_CASE(40, 32) SO_CASE(48, 40) SO_CASE(56, 48) SO_CASE(64, 56) SO_CASE(72, 64)
    SO_CASE(80, 72) SO_CASE(90, 80) SO_CASE(100, 90) SO_CASE(112, 100) SO_CASE(128, 112)
    return APGPU_ERR_UNSUPPORTED;
}

thread_local char g_kname[64];

}  // namespace

extern "C" const char* apgpu_stack_kernel_name(int N, int method, double k_lo, double k_hi,
                                               int maxiters, int cen, int dev,
                                               int want_uncert, int out_is_f64, int flags) {
    (void)out_is_f64;
    const Bucket* b = nullptr;
    Family f = choose_family(N, method, k_lo, k_hi, maxiters, cen, dev, want_uncert != 0, flags, &b);
    switch (f) {
        case FAM_MEANCLIP: snprintf(g_kname, sizeof(g_kname), "meanclip<%d>", b->nb); break;
        case FAM_MEANCLIP_SMEM: snprintf(g_kname, sizeof(g_kname), "meanclip_smem"); break;
        case FAM_SORT_MED: snprintf(g_kname, sizeof(g_kname), "sorted_median<%d>", b->nb); break;
        case FAM_SORT_MEDMAD1: snprintf(g_kname, sizeof(g_kname), "sorted_medmad1<%d>", b->nb); break;
        default: snprintf(g_kname, sizeof(g_kname), "generic<%d>", N <= 32 ? 32 : (N <= 128 ? 128 : 1024)); break;
    }
    return g_kname;
}

extern "C" int apgpu_stack_reduce_f32(const float* const* frames, int N, int64_t H, int64_t W,
                                      int64_t row0, int64_t nrows, int method,
                                      double k_lo, double k_hi, int maxiters, int cen, int dev,
                                      void* out_data, int out_is_f64,
                                      void* out_nrej, int nrej_is_u16,
                                      void* out_uncert, uint8_t* out_allmasked,
                                      int flags, apgpu_stream_t stream) {
    APGPU_REQUIRE(frames && out_data, "stack_reduce: null frames/out pointer");
    APGPU_REQUIRE(N >= 1 && N <= APGPU_STACK_MAX_FRAMES, "stack_reduce: N=%d outside 1..%d", N, APGPU_STACK_MAX_FRAMES);
    APGPU_REQUIRE(H > 0 && W > 0 && row0 >= 0 && nrows >= 0 && row0 + nrows <= H,
                  "stack_reduce: bad geometry H=%lld W=%lld row0=%lld nrows=%lld",
                  (long long)H, (long long)W, (long long)row0, (long long)nrows);
    APGPU_REQUIRE(H * W < ((int64_t)1 << 32), "stack_reduce: frames of %lld pixels exceed the 2^32 limit", (long long)(H * W));
    APGPU_REQUIRE(method >= 0 && method <= 3, "stack_reduce: bad method %d", method);
    APGPU_REQUIRE(cen == APGPU_CEN_MEAN || cen == APGPU_CEN_MEDIAN, "stack_reduce: bad cen %d", cen);
    APGPU_REQUIRE(dev == APGPU_DEV_STD || dev == APGPU_DEV_MAD_STD, "stack_reduce: bad dev %d", dev);
    APGPU_REQUIRE(maxiters == 0 || (k_lo == k_lo && k_hi == k_hi), "stack_reduce: NaN clip threshold");
    APGPU_REQUIRE(!out_nrej || nrej_is_u16 || N <= 255, "stack_reduce: uint8 rejection map needs N <= 255 (N=%d)", N);
    for (int i = 0; i < N; ++i) APGPU_REQUIRE(frames[i], "stack_reduce: frame %d is null", i);
    if (nrows == 0) return APGPU_OK;

    StackArgs a;
    a.N = N; a.method = method; a.maxiters = maxiters; a.cen = cen; a.dev = dev;
    a.klo = k_lo; a.khi = k_hi;
    a.pix0 = row0 * W; a.npix = nrows * W;
    a.out = out_data; a.out_f64 = out_is_f64;
    a.nrej = out_nrej; a.nrej_u16 = nrej_is_u16;
    a.uncert = out_uncert; a.allmasked = out_allmasked;
    cudaStream_t st = (cudaStream_t)stream;

    const Bucket* b = nullptr;
    Family f = choose_family(N, method, k_lo, k_hi, maxiters, cen, dev, out_uncert != nullptr, flags, &b);
    switch (f) {
        case FAM_MEANCLIP: return dispatch_meanclip(b, frames, a, st, flags);
        case FAM_MEANCLIP_SMEM:
            return N <= 128 ? launch_meanclip_smem<128>(frames, a, st) : launch_meanclip_smem<512>(frames, a, st);
        case FAM_SORT_MED: return dispatch_sorted<MODE_MED>(b, frames, a, st);
        case FAM_SORT_MEDMAD1: return dispatch_sorted<MODE_MEDMAD1>(b, frames, a, st);
        default: break;
    }
    if (N <= 32) return launch_generic<32>(frames, a, st);
    if (N <= 128) return launch_generic<128>(frames, a, st);
    return launch_generic<1024>(frames, a, st);
}

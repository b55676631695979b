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
// Kernel families behind one entry point (one translation unit per family / instantiation group):
//   generic<CAP>        every parameter combination, N <= 1024.  float64 arithmetic in the
//                       oracle's operation order (explicitly rounded intrinsics, no FMA
//                       contraction): bit-identical to the numpy restatement.  Local memory.
//   meanclip<NB,NLO>    iterative kappa-sigma clip about the MEAN with the population STD,
//                       then the mean of the survivors.  Register-resident, float32 arithmetic
//                       on pivot-shifted values with a rigorous error bound: a pixel whose
//                       decision could differ from the float64 oracle is redone by the generic
//                       routine, so rejection maps are identical.  Fed by a warp-granular
//                       tensor-map TMA pipeline when the frames are equally spaced
//                       (stack_meanclip.cuh), else by direct loads / cp.async.
//   meanclip_coop<NBL,P> the same algorithm for 128 < N <= 512: P lanes share a pixel, P warps
//                       share a 128B-swizzled TMA tile (stack_meanclip_coop.cuh);
//                       meanclip_split (cp.async, 512 < N <= 1024), meanclip_smem (pointer tables).
//   sorted<NB,MODE>     register-resident Batcher merge-exchange network (N <= 200) with
//                       mixed ALU/FMA-pipe comparators: plain median (MODE_MED), or the
//                       reference's ApMasterCal setting -- one median/MAD clip pass then the
//                       mean (MODE_MEDMAD1) -- with the sorted column parked in shared memory
//                       for the data-dependent MAD selection (stack_sorted.cuh).
//   A pixel holding NaN/inf samples leaves the fast kernels for the generic
//   routine, which owns the reference's non-finite semantics.
#pragma once
#include <float.h>
#include <math.h>

#include <cuda.h>

#include "apgpu_common.cuh"

namespace apgpu_stack {

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
    int tiles_per_warp;          // tensor-map staged kernels: warp tiles per warp (per group) per CTA
    int box_rows, nchunks;       // warp-cooperative kernel: rows per TMA box, boxes per tile
    int one, minus_one;          // +1 / -1 as run-time values (sorted kernels: IMAD on the FMA pipe)
    // uint16 frames (apgpu_stack_reduce_u16): a raw 16-bit sample r becomes the float
    //     int_as_float(prmt(r, 0x4B000000, u16_sel) ^ u16_xor) = 2^23 + value      (U16_BIAS + value, exact)
    // u16_sel picks the byte order (0x7610 host order, 0x7601 FITS big-endian), u16_xor = 0x8000 folds the
    // BZERO = 32768 offset of FITS unsigned frames (stored value s = v - 32768 as int16: v = s ^ 0x8000).
    uint32_t u16_sel, u16_xor;
    float sample_bias;           // what the staged samples carry on top of their value: U16_BIAS or 0
};

constexpr float U16_BIAS = 8388608.f;               // 2^23

// decode one raw 16-bit sample (zero-extended in r) to U16_BIAS + value: two ALU-pipe instructions
__device__ __forceinline__ float u16_biased(uint32_t r, const StackArgs& a) {
    return __uint_as_float(__byte_perm(r, 0x4B000000u, a.u16_sel) ^ a.u16_xor);
}
// one sample of a frame as its float32 value (direct global load)
__device__ __forceinline__ float load_sample(const float* p, const StackArgs&) { return ld_stream(p); }
__device__ __forceinline__ float load_sample(const uint16_t* p, const StackArgs& a) {
    return u16_biased((uint32_t)__ldcs(p), a) - U16_BIAS;
}
// the same, still carrying sample_bias (kernels that shift by a pivot anyway)
__device__ __forceinline__ float load_sample_biased(const float* p, const StackArgs&) { return ld_stream(p); }
__device__ __forceinline__ float load_sample_biased(const uint16_t* p, const StackArgs& a) {
    return u16_biased((uint32_t)__ldcs(p), a);
}
// a staged (shared-memory) sample, biased
__device__ __forceinline__ float staged_sample_biased(const float* s, const StackArgs&) { return *s; }
__device__ __forceinline__ float staged_sample_biased(const uint16_t* s, const StackArgs& a) {
    return u16_biased((uint32_t)*s, a);
}
template <typename T> __host__ __device__ constexpr float sample_bias_of() { return sizeof(T) == 2 ? U16_BIAS : 0.f; }

template <int CAP, typename T = float> struct FramePtrs {
    typedef T sample_t;
    const T* p[CAP];
    __device__ __forceinline__ const T* frame(int i) const { return p[i]; }
};
// equally spaced frames (a [N][H*W] cube): frame i starts stride bytes after frame i-1
template <typename T = float> struct CubeFramesT {
    typedef T sample_t;
    const char* base;
    int64_t stride;
    __device__ __forceinline__ const T* frame(int i) const {
        return reinterpret_cast<const T*>(base + (int64_t)i * stride);
    }
};
typedef CubeFramesT<float> CubeFrames;

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

// Pixels a fast kernel cannot finish (non-finite samples, a survivor inside the float32 guard band, everything
// rejected) are MARKED instead of being finished in place: a NaN with a payload no result can carry goes to the
// output image, all-ones to the rejection map, and stack_launch_marked() -- one small launch after the fast
// kernels of the call -- scans for the marks and runs generic_pixel on exactly those pixels.  Calling the generic
// routine from inside the fast kernels cost them a stack frame, spills around the call and, in the
// lane-cooperative kernels, stalled the whole tile group for the 30-300 us the routine takes (measured: 7 % of
// the N = 100 kappa-sigma kernel, two thirds of the N = 512 one).
// (the low 4 bits of the payload say why -- read by tools/count_marks.py with APGPU_SKIP_MARKED=1)
constexpr uint32_t STACK_MARK32 = 0x7fc5a5a0u;
constexpr uint64_t STACK_MARK64 = 0x7ff85a5a5a5a5a50ull;
enum { MARK_NONFINITE = 1, MARK_BOUNDS = 2, MARK_BAND = 3, MARK_EMPTY = 4 };
__device__ __forceinline__ void mark_pixel(const StackArgs& a, int64_t p, int why = 0) {
    if (a.out_f64) reinterpret_cast<uint64_t*>(a.out)[p] = STACK_MARK64 | (uint64_t)why;
    else reinterpret_cast<uint32_t*>(a.out)[p] = STACK_MARK32 | (uint32_t)why;
    if (a.nrej) {
        if (a.nrej_u16) reinterpret_cast<uint16_t*>(a.nrej)[p] = 0xffffu;
        else reinterpret_cast<uint8_t*>(a.nrej)[p] = 0xffu;
    }
}
__device__ __forceinline__ bool pixel_is_marked(const StackArgs& a, int64_t p) {
    return a.out_f64 ? (reinterpret_cast<const uint64_t*>(a.out)[p] & ~(uint64_t)15) == STACK_MARK64
                     : (reinterpret_cast<const uint32_t*>(a.out)[p] & ~15u) == STACK_MARK32;
}

// ---------------------------------------------------------------------------
// generic routine: float64, oracle operation order
// ---------------------------------------------------------------------------
static __device__ void shell_sort(float* s, int n) {
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

template <int CAP, typename Frames = FramePtrs<CAP, float>>
__device__ __noinline__ void generic_pixel(const Frames& fp, const StackArgs& a, int64_t p) {
    float v[CAP];      // frame order; NaN marks a sample that is not (or no longer) used
    float s[CAP];      // the used samples, ascending
    const int N = a.N;
    const bool clip = a.maxiters != 0;
    const bool need_sorted = (a.method != APGPU_METHOD_AVERAGE) ||
                             (clip && (a.cen == APGPU_CEN_MEDIAN || a.dev == APGPU_DEV_MAD_STD));
    int nk = 0;
    for (int i = 0; i < N; ++i) {
        float x = load_sample(fp.frame(i) + p, a);
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


// i < NLO is known at compile time to be a real frame; only the NB-NLO tail
// elements carry a (warp-uniform) runtime predicate.
#define APGPU_ACTIVE(i) ((i) < NLO || (i) < N)

__device__ __forceinline__ float med3(float a, float b, float c) {
    return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}

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

// tensor-map TMA helpers (mbarrier phase wait, L2 evict-first policy, 2-D tiled load)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int32_t c0, int32_t c1,
                                            uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

struct Bucket { int nb, nlo; };

// Equally spaced frames (a [N][H*W] cube) can be described by one 2-D TMA tensor map.
bool stack_is_cube_bytes(const void* const* frames, int N, int64_t npix_end, int elem_bytes);
template <typename T> inline bool stack_is_cube(const T* const* frames, int N, int64_t npix_end) {
    return stack_is_cube_bytes(reinterpret_cast<const void* const*>(frames), N, npix_end, (int)sizeof(T));
}
// dim0 = pixel (npix_end of them, contiguous), dim1 = frame (N, stride_bytes apart); box = box_pix x box_rows
// (box_rows = 0: all N frames in one box, N <= 256).  elem_bytes 4 = float32, 2 = uint16.
bool encode_stack_tensor_map_bytes(CUtensorMap* tmap, const void* base, int elem_bytes, uint64_t npix_end, int N,
                                   uint64_t stride_bytes, int box_pix, int box_rows, bool swizzle128);
template <typename T>
inline bool encode_stack_tensor_map(CUtensorMap* tmap, const T* base, uint64_t npix_end, int N,
                                    uint64_t stride_bytes, int box_pix, int box_rows = 0, bool swizzle128 = false) {
    return encode_stack_tensor_map_bytes(tmap, base, (int)sizeof(T), npix_end, N, stride_bytes, box_pix, box_rows, swizzle128);
}

// what the last apgpu_stack_reduce_f32 call of this thread launched (tests / bench bookkeeping)
void stack_note_staging(int staging);
int stack_tmap_tiles_per_warp();
int stack_coop_box_rows_max();
int stack_median_tiles_per_cta();

// cross-translation-unit launchers (one .cu per kernel family so that nvcc compiles them in parallel)
int stack_launch_generic(const float* const* frames, const StackArgs& a, cudaStream_t st);
int stack_launch_generic(const uint16_t* const* frames, const StackArgs& a, cudaStream_t st);
// finish the pixels the fast kernels marked in [a.pix0, a.pix0 + a.npix) (see mark_pixel)
int stack_launch_marked(const float* const* frames, const StackArgs& a, cudaStream_t st);
int stack_launch_marked(const uint16_t* const* frames, const StackArgs& a, cudaStream_t st);
int stack_dispatch_meanclip_lo(int nb, const float* const* frames, const StackArgs& a, cudaStream_t st, int flags);
int stack_dispatch_meanclip_mid(int nb, const float* const* frames, const StackArgs& a, cudaStream_t st, int flags);
int stack_dispatch_meanclip_hi(int nb, const float* const* frames, const StackArgs& a, cudaStream_t st, int flags);
int stack_dispatch_meanclip_lo(int nb, const uint16_t* const* frames, const StackArgs& a, cudaStream_t st, int flags);
int stack_dispatch_meanclip_mid(int nb, const uint16_t* const* frames, const StackArgs& a, cudaStream_t st, int flags);
int stack_dispatch_meanclip_hi(int nb, const uint16_t* const* frames, const StackArgs& a, cudaStream_t st, int flags);
int stack_launch_meanclip_smem(const float* const* frames, const StackArgs& a, cudaStream_t st);
// long stacks (N > 100) on equally spaced frames -- warp-cooperative kernels up to N = 512, lane-split
// cp.async kernels beyond: returns APGPU_ERR_UNSUPPORTED when there
// is no bucket; *done_pix = pixels (from a.pix0) that were reduced, the caller finishes the tail
int stack_dispatch_meanclip_split(const float* const* frames, const StackArgs& a, cudaStream_t st, int flags,
                                  int64_t* done_pix);
int stack_dispatch_meanclip_split_p8(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix);
int stack_dispatch_meanclip_coop_p2(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix);
int stack_dispatch_meanclip_coop_p4(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix);
int stack_dispatch_meanclip_coop_p8(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix);
// lane-cooperative median (200 < N <= 512, equally spaced float32 frames): same contract
#define APGPU_DECL_MEDCOOP(m) \
    int stack_dispatch_median_coop_##m##_p4(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix); \
    int stack_dispatch_median_coop_##m##_p8(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix);
APGPU_DECL_MEDCOOP(med) APGPU_DECL_MEDCOOP(medunc) APGPU_DECL_MEDCOOP(medmad1)
#undef APGPU_DECL_MEDCOOP
// sorted<NB, NLO, MODE> kernels (stack_sorted.cuh).  One translation unit per (mode, sample type, bucket part):
// the unrolled networks are slow to compile (10-25 s per bucket), so the explicit instantiations
// dispatch_sorted_part<MODE, T, PART> are spread over many files that nvcc compiles in parallel.
constexpr int MODE_MED = 0;       // method=median, no clipping
constexpr int MODE_MEDMAD1 = 1;   // one median/MAD clip pass, then the mean (ApMasterCal)
constexpr int MODE_MEDUNC = 2;    // method=median, no clipping, uncertainty = 1.4826 * MAD / sqrt(N)
constexpr int SORT_PARTS = 4;
__host__ __device__ constexpr int sorted_part_of(int nb) { return nb <= 40 ? 0 : (nb <= 80 ? 1 : (nb <= 112 ? 2 : 3)); }
// float32 median and median/MAD clip use the fine bucket list; the uncertainty mode and every uint16 mode the
// coarse one (fewer instantiations; a coarser bucket only costs a little padding work)
template <int MODE, typename T> constexpr bool sorted_fine_buckets() { return sizeof(T) == 4 && MODE != MODE_MEDUNC; }
template <int MODE, typename T, int PART>
int dispatch_sorted_part(int nb, const T* const* frames, const StackArgs& a, cudaStream_t st);


}  // namespace apgpu_stack

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
//                       decision could differ from the float64 oracle is MARKED and redone in
//                       float64 by the scan launch below, so rejection maps are identical.  Fed by a warp-granular
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
//   marked              what a fast kernel cannot finish -- non-finite samples (the generic routine owns the
//                       reference's NaN / inf semantics), float32 guard-band hits, everything rejected -- is
//                       marked in the outputs (stack_common.cuh: mark_pixel) and one scan launch after the fast
//                       kernels of the call (stack_generic.cu: stack_marked_kernel) redoes exactly those pixels:
//                       warp-cooperative float64 for the kappa-sigma family, generic routine otherwise.
#include <stdlib.h>

#include "stack_common.cuh"

namespace apgpu_stack {

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
enum Family { FAM_GENERIC = 0, FAM_MEANCLIP = 1, FAM_SORT_MED = 2, FAM_SORT_MEDMAD1 = 3, FAM_MEANCLIP_SMEM = 4,
              FAM_SORT_MEDUNC = 5 };

constexpr int MEANCLIP_SMEM_MAX_N = 4 * ((SMEM_MAX_BYTES / (TPB * 16)) & ~1);   // 452
// the lane-cooperative kernels take over above this frame count (re-measured after the generic fallback had left the
// kernels: the single-thread meanclip<128> runs N = 104 / 128 at 73 % / 79 % of the roofline, 2 lanes per pixel at
// 64 % / 72 %; from N = 144 on the cooperative kernels lead by 1-5 %)
constexpr int MEANCLIP_COOP_MIN_N = 128;
constexpr int MEANCLIP_REG_DEFAULT_MAX_N = 200;  // measured: the register kernel wins wherever it exists (bench.py variants)

// (NLO, NB] buckets.  meanclip goes to 200 frames in registers; sorted to 128.
const Bucket MEANCLIP_BUCKETS[] = {{8, 2}, {16, 8}, {24, 16}, {32, 24}, {48, 32}, {64, 48}, {80, 64},
                                   {100, 80}, {128, 100}, {160, 128}, {200, 160}};
const Bucket SORT_BUCKETS[] = {{4, 0}, {8, 4}, {12, 8}, {16, 12}, {20, 16}, {24, 20}, {32, 24}, {40, 32},
                               {48, 40}, {56, 48}, {64, 56}, {72, 64}, {80, 72}, {90, 80}, {100, 90},
                               {112, 100}, {128, 112}, {160, 128}, {200, 160}};

const Bucket SORT_BUCKETS_COARSE[] = {{4, 0}, {8, 4}, {16, 8}, {24, 16}, {32, 24}, {48, 32}, {64, 48}, {80, 64},
                                      {100, 80}, {128, 100}, {160, 128}, {200, 160}};

template <size_t K>
const Bucket* find_bucket(const Bucket (&b)[K], int N) {
    for (size_t i = 0; i < K; ++i) if (N > b[i].nlo && N <= b[i].nb) return &b[i];
    return nullptr;
}

// the kappa-sigma-about-the-mean family (meanclip kernels) handles this parameter set
bool meanclip_eligible(int N, int method, double klo, double khi, int maxiters, int cen, int dev, int flags) {
    return !(flags & APGPU_STACK_FORCE_GENERIC) && method == APGPU_METHOD_AVERAGE &&
           (maxiters == 0 || (cen == APGPU_CEN_MEAN && dev == APGPU_DEV_STD)) &&
           N >= 3 && klo > 0.0 && khi > 0.0 && klo < 1e6 && khi < 1e6;
}

// long stacks on the lane-cooperative sorted kernels: -1 = not eligible, else MODE_MED (plain median),
// MODE_MEDUNC (median + uncertainty plane) or MODE_MEDMAD1 (one median/MAD clip pass, then the mean)
int median_coop_mode(int N, int method, int maxiters, int cen, int dev, bool want_uncert, int flags) {
    if ((flags & APGPU_STACK_FORCE_GENERIC) || N <= 200 || N > 512) return -1;
    if (method == APGPU_METHOD_MEDIAN && maxiters == 0) return want_uncert ? MODE_MEDUNC : MODE_MED;
    if (method == APGPU_METHOD_AVERAGE && maxiters == 1 && cen == APGPU_CEN_MEDIAN && dev == APGPU_DEV_MAD_STD)
        return MODE_MEDMAD1;
    return -1;
}

Family choose_family(int N, int method, double klo, double khi, int maxiters, int cen, int dev,
                     bool want_uncert, int flags, const Bucket** bucket) {
    *bucket = nullptr;
    if (flags & APGPU_STACK_FORCE_GENERIC) return FAM_GENERIC;
    if (meanclip_eligible(N, method, klo, khi, maxiters, cen, dev, flags)) {
        const Bucket* rb = find_bucket(MEANCLIP_BUCKETS, N);
        const bool smem_ok = N <= MEANCLIP_SMEM_MAX_N;
        bool use_reg = rb && (N <= MEANCLIP_REG_DEFAULT_MAX_N || !smem_ok);
        if ((flags & APGPU_STACK_PREFER_REGISTERS) && rb) use_reg = true;
        if ((flags & APGPU_STACK_PREFER_SHARED) && smem_ok) use_reg = false;
        if (use_reg) { *bucket = rb; return FAM_MEANCLIP; }
        if (smem_ok) return FAM_MEANCLIP_SMEM;
    }
    if (method == APGPU_METHOD_MEDIAN && maxiters == 0) {
        // with an uncertainty plane (what ApMasterCal always asks for) the MAD of every pixel is needed:
        // full network + parked column, like the median/MAD clip
        if ((*bucket = find_bucket(SORT_BUCKETS, N))) return want_uncert ? FAM_SORT_MEDUNC : FAM_SORT_MED;
    }
    if (method == APGPU_METHOD_AVERAGE && maxiters == 1 && cen == APGPU_CEN_MEDIAN &&
        dev == APGPU_DEV_MAD_STD && N >= 2) {
        if ((*bucket = find_bucket(SORT_BUCKETS, N))) return FAM_SORT_MEDMAD1;
    }
    return FAM_GENERIC;
}

int stack_dispatch_meanclip_split(const float* const* frames, const StackArgs& a, cudaStream_t st, int flags,
                                  int64_t* done_pix) {
    *done_pix = 0;
    if (flags & (APGPU_STACK_DIRECT_LOADS | APGPU_STACK_USE_TMA | APGPU_STACK_USE_CPASYNC | APGPU_STACK_PREFER_SHARED |
                 APGPU_STACK_PREFER_REGISTERS))
        return APGPU_ERR_UNSUPPORTED;
    if (!stack_is_cube(frames, a.N, a.pix0 + a.npix)) return APGPU_ERR_UNSUPPORTED;
    static const bool no_coop = getenv("APGPU_NO_COOP") != nullptr;           // tuning knob
    if (!no_coop && a.N <= 512) {
        // lanes per pixel (measured, tools/time_variant.py): short per-lane arrays (<= 64 samples) keep the
        // unrolled code and the register count small, which matters more than the extra shuffle steps
        static const int force_p = getenv("APGPU_COOP_P") ? atoi(getenv("APGPU_COOP_P")) : 0;   // tuning knob
        // (re-measured once the generic fallback had left the kernels: the longest per-lane arrays that are
        // instantiated win -- N = 160: 2 lanes 75 % against 67 % with 4; N = 320: 4 lanes 73 % against 63 % with 8)
        int P = a.N <= 160 ? 2 : (a.N <= 320 ? 4 : 8);
        if (force_p) P = force_p;
        int rc = APGPU_ERR_UNSUPPORTED;
        if (P == 2) rc = stack_dispatch_meanclip_coop_p2(frames, a, st, done_pix);
        if (P == 4) rc = stack_dispatch_meanclip_coop_p4(frames, a, st, done_pix);
        if (P == 8) rc = stack_dispatch_meanclip_coop_p8(frames, a, st, done_pix);
        if (rc != APGPU_ERR_UNSUPPORTED) return rc;
    }
    if (a.N > 512) return stack_dispatch_meanclip_split_p8(frames, a, st, done_pix);
    return APGPU_ERR_UNSUPPORTED;
}

template <int MODE, typename T>
int stack_dispatch_sorted(int N, const T* const* frames, const StackArgs& a, cudaStream_t st) {
    const Bucket* b = sorted_fine_buckets<MODE, T>() ? find_bucket(SORT_BUCKETS, N) : find_bucket(SORT_BUCKETS_COARSE, N);
    if (!b) return APGPU_ERR_UNSUPPORTED;
    switch (sorted_part_of(b->nb)) {
        case 0: return dispatch_sorted_part<MODE, T, 0>(b->nb, frames, a, st);
        case 1: return dispatch_sorted_part<MODE, T, 1>(b->nb, frames, a, st);
        case 2: return dispatch_sorted_part<MODE, T, 2>(b->nb, frames, a, st);
        default: return dispatch_sorted_part<MODE, T, 3>(b->nb, frames, a, st);
    }
}

template <typename T>
int stack_dispatch_meanclip(int nb, const T* const* frames, const StackArgs& a, cudaStream_t st, int flags) {
    if (nb <= 64) return stack_dispatch_meanclip_lo(nb, frames, a, st, flags);
    if (nb <= 128) return stack_dispatch_meanclip_mid(nb, frames, a, st, flags);
    return stack_dispatch_meanclip_hi(nb, frames, a, st, flags);
}

int stack_median_tiles_per_cta() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("APGPU_MEDIAN_TILES_PER_CTA");   // tuning knob
        int x = e ? atoi(e) : 0;
        v = (x >= 1 && x <= 4096) ? x : 8;
    }
    return v;
}
int stack_coop_box_rows_max() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("APGPU_COOP_BOX_ROWS");      // tuning knob
        int x = e ? atoi(e) : 0;
        v = (x >= 8 && x <= 256) ? x : 256;
    }
    return v;
}
bool stack_is_cube_bytes(const void* const* frames, int N, int64_t npix_end, int elem_bytes) {
    if (N < 2 || npix_end <= 0 || npix_end >= ((int64_t)1 << 31)) return false;
    const int64_t stride = (const char*)frames[1] - (const char*)frames[0];
    if (stride < npix_end * (int64_t)elem_bytes || stride % 16 != 0 || stride >= ((int64_t)1 << 40)) return false;
    if (!apgpu_aligned(frames[0], 16)) return false;
    for (int i = 2; i < N; ++i)
        if ((const char*)frames[i] - (const char*)frames[i - 1] != stride) return false;
    return true;
}

bool encode_stack_tensor_map_bytes(CUtensorMap* tmap, const void* base, int elem_bytes, uint64_t npix_end, int N,
                                   uint64_t stride_bytes, int box_pix, int box_rows, bool swizzle128) {
    // cuTensorMapEncodeTiled lives in the driver (libcuda): fetched through the runtime so that the
    // library keeps linking against cudart only
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !fn) {
            (void)cudaGetLastError();
            return false;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t gdim[2] = {npix_end, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {stride_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_pix, (cuuint32_t)(box_rows > 0 ? box_rows : N)};
    const cuuint32_t estride[2] = {1, 1};
    return encode(tmap, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

thread_local char g_kname[64];
thread_local int g_last_staging = -1;
void stack_note_staging(int staging) { g_last_staging = staging; }
int stack_tmap_tiles_per_warp() {
    // tuning knob (measured on B200, bench.py variants): APGPU_TMAP_TILES_PER_WARP overrides
    static int v = 0;
    if (!v) {
        const char* e = getenv("APGPU_TMAP_TILES_PER_WARP");
        int x = e ? atoi(e) : 0;
        v = (x >= 1 && x <= 4096) ? x : 32;
    }
    return v;
}

}  // namespace apgpu_stack

using namespace apgpu_stack;

extern "C" const char* apgpu_stack_kernel_name(int N, int method, double k_lo, double k_hi,
                                               int maxiters, int cen, int dev,
                                               int want_uncert, int out_is_f64, int flags) {
    (void)out_is_f64;
    const Bucket* b = nullptr;
    Family f = choose_family(N, method, k_lo, k_hi, maxiters, cen, dev, want_uncert != 0, flags, &b);
    // long kappa-sigma stacks: the name of the default path on equally spaced frames (other layouts fall back
    // to the pointer-table kernels named below -- apgpu_stack_last_staging() tells which one ran)
    if (N > MEANCLIP_COOP_MIN_N && meanclip_eligible(N, method, k_lo, k_hi, maxiters, cen, dev, flags) &&
        !(flags & (APGPU_STACK_DIRECT_LOADS | APGPU_STACK_USE_TMA | APGPU_STACK_USE_CPASYNC | APGPU_STACK_PREFER_SHARED |
                   APGPU_STACK_PREFER_REGISTERS))) {
        if (N <= 512) snprintf(g_kname, sizeof(g_kname), "meanclip_coop<%d>", N <= 160 ? 2 : (N <= 320 ? 4 : 8));
        else snprintf(g_kname, sizeof(g_kname), "meanclip_split<8>");
        return g_kname;
    }
    const int cmode = median_coop_mode(N, method, k_lo == k_lo ? maxiters : maxiters, cen, dev, want_uncert != 0, flags);
    if (cmode >= 0) {
        snprintf(g_kname, sizeof(g_kname), "%s_coop<%d>", cmode == MODE_MED ? "median" : (cmode == MODE_MEDUNC ? "median_mad" : "medmad1"),
                 N <= 256 ? 4 : 8);
        return g_kname;
    }
    switch (f) {
        case FAM_MEANCLIP: snprintf(g_kname, sizeof(g_kname), "meanclip<%d>", b->nb); break;
        case FAM_MEANCLIP_SMEM: snprintf(g_kname, sizeof(g_kname), "meanclip_smem"); break;
        case FAM_SORT_MED: snprintf(g_kname, sizeof(g_kname), "sorted_median<%d>", b->nb); break;
        case FAM_SORT_MEDMAD1: snprintf(g_kname, sizeof(g_kname), "sorted_medmad1<%d>", b->nb); break;
        case FAM_SORT_MEDUNC: snprintf(g_kname, sizeof(g_kname), "sorted_median_mad<%d>", b->nb); break;
        default: snprintf(g_kname, sizeof(g_kname), "generic<%d>", N <= 32 ? 32 : (N <= 128 ? 128 : 1024)); break;
    }
    return g_kname;
}

extern "C" int apgpu_stack_last_staging(void) { return g_last_staging; }

namespace apgpu_stack {

template <typename T>
int stack_reduce_fast(const T* const* frames, int N, double k_lo, double k_hi, int flags, StackArgs a, cudaStream_t st,
                      void* out_uncert, bool* used_fast);

template <typename T>
int stack_reduce_impl(const T* const* frames, int u16_format, int N, int64_t H, int64_t W,
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
    APGPU_REQUIRE(u16_format == APGPU_U16_NATIVE || u16_format == APGPU_U16_FITS_BZERO,
                  "stack_reduce: bad uint16 sample format %d", u16_format);
    for (int i = 0; i < N; ++i) APGPU_REQUIRE(frames[i], "stack_reduce: frame %d is null", i);
    for (int i = 0; i < N; ++i)
        APGPU_REQUIRE(apgpu_aligned(frames[i], sizeof(T)), "stack_reduce: frame %d is not aligned to its sample size", i);
    if (nrows == 0) return APGPU_OK;

    StackArgs a;
    a.N = N; a.method = method; a.maxiters = maxiters; a.cen = cen; a.dev = dev;
    a.klo = k_lo; a.khi = k_hi;
    a.pix0 = row0 * W; a.npix = nrows * W;
    a.out = out_data; a.out_f64 = out_is_f64;
    a.nrej = out_nrej; a.nrej_u16 = nrej_is_u16;
    a.uncert = out_uncert; a.allmasked = out_allmasked;
    a.one = 1; a.minus_one = -1; a.tiles_per_warp = 0; a.box_rows = 0; a.nchunks = 0;
    a.u16_sel = u16_format == APGPU_U16_FITS_BZERO ? 0x7601u : 0x7610u;
    a.u16_xor = u16_format == APGPU_U16_FITS_BZERO ? 0x8000u : 0u;
    a.sample_bias = sample_bias_of<T>();
    for (int i = 0; i < MEANCLIP_MAX_TAIL; ++i) a.tailmask[i] = 0.f;
    cudaStream_t st = (cudaStream_t)stream;
    g_last_staging = -1;
    bool used_fast = false;
    const int rc = stack_reduce_fast<T>(frames, N, k_lo, k_hi, flags, a, st, out_uncert, &used_fast);
    if (rc != APGPU_OK || !used_fast) return rc;
    // the pixels the fast kernels left marked (non-finite samples, guard-band hits): generic routine
    static const bool skip_marked = getenv("APGPU_SKIP_MARKED") != nullptr;      // diagnostic knob (tools/count_marks.py)
    if (skip_marked) return APGPU_OK;
    static const bool twice = getenv("APGPU_MARKED_TWICE") != nullptr;          // diagnostic: the second launch is the pure scan
    if (twice) stack_launch_marked(frames, a, st);
    return stack_launch_marked(frames, a, st);
}

template <typename T>
int stack_reduce_fast(const T* const* frames, int N, double k_lo, double k_hi, int flags, StackArgs a, cudaStream_t st,
                      void* out_uncert, bool* used_fast) {
    const int method = a.method, maxiters = a.maxiters, cen = a.cen, dev = a.dev;
    const int64_t pix0_call = a.pix0;
    *used_fast = true;

    // long stacks on equally spaced frames: the lane-split tensor-map kernels take every full warp tile,
    // whatever follows only sees the (< 32-pixel) tail  (float32 frames; uint16 frames use the register
    // kernels up to N = 200)
    if constexpr (sizeof(T) == sizeof(float)) {
        if (N > MEANCLIP_COOP_MIN_N && meanclip_eligible(N, method, k_lo, k_hi, maxiters, cen, dev, flags)) {
            int64_t done = 0;
            const int rc = stack_dispatch_meanclip_split(frames, a, st, flags, &done);
            if (rc != APGPU_OK && rc != APGPU_ERR_UNSUPPORTED) return rc;
            if (rc == APGPU_OK) {
                a.pix0 += done;
                a.npix -= done;
                if (a.npix == 0) return APGPU_OK;
            }
        }
    }
    // long plain medians on equally spaced float32 frames: lane-cooperative selection (stack_median_coop.cuh)
    if constexpr (sizeof(T) == sizeof(float)) {
        const int cmode = median_coop_mode(N, method, maxiters, cen, dev, out_uncert != nullptr, flags);
        if (cmode >= 0 && stack_is_cube(frames, N, a.pix0 + a.npix)) {
            int64_t done = 0;
            int rc;
            if (cmode == MODE_MED)
                rc = N <= 256 ? stack_dispatch_median_coop_med_p4(frames, a, st, &done) : stack_dispatch_median_coop_med_p8(frames, a, st, &done);
            else if (cmode == MODE_MEDUNC)
                rc = N <= 256 ? stack_dispatch_median_coop_medunc_p4(frames, a, st, &done) : stack_dispatch_median_coop_medunc_p8(frames, a, st, &done);
            else
                rc = N <= 256 ? stack_dispatch_median_coop_medmad1_p4(frames, a, st, &done) : stack_dispatch_median_coop_medmad1_p8(frames, a, st, &done);
            if (rc != APGPU_OK && rc != APGPU_ERR_UNSUPPORTED) return rc;
            if (rc == APGPU_OK) {
                a.pix0 += done;
                a.npix -= done;
                if (a.npix == 0) return APGPU_OK;
            }
        }
    }
    const int staging_so_far = g_last_staging;

    const Bucket* b = nullptr;
    Family f = choose_family(N, method, k_lo, k_hi, maxiters, cen, dev, out_uncert != nullptr, flags, &b);
    if (sizeof(T) != sizeof(float) && f == FAM_MEANCLIP_SMEM) f = FAM_GENERIC;
    struct Restore { int v; ~Restore() { if (v >= 0) g_last_staging = v; } } restore{staging_so_far};
    switch (f) {
        case FAM_MEANCLIP: return stack_dispatch_meanclip(b->nb, frames, a, st, flags);
        case FAM_MEANCLIP_SMEM:
            if constexpr (sizeof(T) == sizeof(float)) return stack_launch_meanclip_smem(frames, a, st);
            break;
        case FAM_SORT_MED: return stack_dispatch_sorted<MODE_MED, T>(N, frames, a, st);
        case FAM_SORT_MEDUNC: return stack_dispatch_sorted<MODE_MEDUNC, T>(N, frames, a, st);
        case FAM_SORT_MEDMAD1: return stack_dispatch_sorted<MODE_MEDMAD1, T>(N, frames, a, st);
        default: break;
    }
    // (a generic launch over the remaining range leaves no marks of its own; earlier fast launches may have)
    const int rcg = stack_launch_generic(frames, a, st);
    if (a.pix0 == pix0_call) *used_fast = false;
    return rcg;
}

}  // namespace apgpu_stack

extern "C" int apgpu_stack_reduce_f32(const float* const* frames, int N, int64_t H, int64_t W,
                                      int64_t row0, int64_t nrows, int method,
                                      double k_lo, double k_hi, int maxiters, int cen, int dev,
                                      void* out_data, int out_is_f64,
                                      void* out_nrej, int nrej_is_u16,
                                      void* out_uncert, uint8_t* out_allmasked,
                                      int flags, apgpu_stream_t stream) {
    return stack_reduce_impl<float>(frames, APGPU_U16_NATIVE, N, H, W, row0, nrows, method, k_lo, k_hi, maxiters, cen, dev,
                                    out_data, out_is_f64, out_nrej, nrej_is_u16, out_uncert, out_allmasked, flags, stream);
}

extern "C" int apgpu_stack_reduce_u16(const uint16_t* const* frames, int u16_format, int N, int64_t H, int64_t W,
                                      int64_t row0, int64_t nrows, int method,
                                      double k_lo, double k_hi, int maxiters, int cen, int dev,
                                      void* out_data, int out_is_f64,
                                      void* out_nrej, int nrej_is_u16,
                                      void* out_uncert, uint8_t* out_allmasked,
                                      int flags, apgpu_stream_t stream) {
    return stack_reduce_impl<uint16_t>(frames, u16_format, N, H, W, row0, nrows, method, k_lo, k_hi, maxiters, cen, dev,
                                       out_data, out_is_f64, out_nrej, nrej_is_u16, out_uncert, out_allmasked, flags, stream);
}

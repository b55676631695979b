// Frame-stack reducer: the generic float64 kernel (every parameter combination).  See stack_common.cuh.
#include "stack_common.cuh"

namespace apgpu_stack {

template <int CAP, typename T>
__global__ void __launch_bounds__(TPB)
stack_generic_kernel(const __grid_constant__ FramePtrs<CAP, T> fp, const __grid_constant__ StackArgs a) {
    int64_t p = a.pix0 + (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (p >= a.pix0 + a.npix) return;
    generic_pixel<CAP, FramePtrs<CAP, T>>(fp, a, p);
}

template <int CAP, typename T>
int launch_generic(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<CAP, T> fp;
    for (int i = 0; i < CAP; ++i) fp.p[i] = i < a.N ? frames[i] : nullptr;
    int64_t blocks = (a.npix + TPB - 1) / TPB;
    stack_generic_kernel<CAP, T><<<(unsigned)blocks, TPB, 0, st>>>(fp, a);
    APGPU_LAUNCH_CHECK("stack_generic_kernel");
    return APGPU_OK;
}

template <typename T>
int launch_generic_any(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    if (a.N <= 32) return launch_generic<32, T>(frames, a, st);
    if (a.N <= 128) return launch_generic<128, T>(frames, a, st);
    return launch_generic<1024, T>(frames, a, st);
}

// ---------------------------------------------------------------------------
// the marked pixels of a fast-kernel call (mark_pixel in stack_common.cuh)
// ---------------------------------------------------------------------------
// A persistent grid scans the cheapest plane that carries the marks -- the rejection map when the caller asked for
// one (1-2 bytes per pixel, all-ones; confirmed against the output image, where a legitimate count of 255 / 65535
// has no mark), else the output image itself -- in 16-byte vectors, several in flight per thread; a thread that
// finds a mark runs the generic routine on that pixel.  Marks are rare (~0.1 % of the pixels on clean data), so
// the launch costs the scan: ~1 / (4N) of the stack's bytes.
constexpr int MK_THREADS = 256;
constexpr int MK_UNROLL = 4;
constexpr int MK_LIST = 2048;       // marks a CTA collects before its warps share them out

// which of the 16 / ES elements of the vector could be marks (bit e = element e)?
template <int ES> __device__ __forceinline__ unsigned vector_mark_candidates(const uint4 v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    unsigned m = 0;
    if (ES == 1) {
        // cheap rejection first: no 0xff byte in any word (the common case)
        bool any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const uint32_t n = ~w[k]; any = any || (((n - 0x01010101u) & ~n & 0x80808080u) != 0u); }
        if (any) {
#pragma unroll
            for (int e = 0; e < 16; ++e) m |= (((w[e >> 2] >> (8 * (e & 3))) & 0xffu) == 0xffu) ? (1u << e) : 0u;
        }
    } else if (ES == 2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) m |= (((w[e >> 1] >> (16 * (e & 1))) & 0xffffu) == 0xffffu) ? (1u << e) : 0u;
    } else if (ES == 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) m |= ((w[e] & ~15u) == STACK_MARK32) ? (1u << e) : 0u;
    } else {
#pragma unroll
        for (int e = 0; e < 2; ++e)
            m |= ((w[2 * e] & ~15u) == (uint32_t)STACK_MARK64 && w[2 * e + 1] == (uint32_t)(STACK_MARK64 >> 32)) ? (1u << e) : 0u;
    }
    return m;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;                                           // (commutative: identical in every lane)
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One marked pixel of a kappa-sigma-about-the-mean call (method average, mean centre, std deviation -- or no
// clipping at all), redone by a whole WARP in float64 from the original samples: lane l takes the frames
// i = l, l + 32, ...; the samples are parked in the warp's shared-memory row and every pass is a short ROLLED
// loop -- this code runs cold, one warp at a time, where instruction fetch costs more than the arithmetic (an
// unrolled register-resident version took 40 us per pixel, the generic routine with its serial loads and
// local-memory arrays 100-300 us).  Partial sums are combined by butterfly.  That is another summation order than
// the oracle's frame order; both are within N 2^-53 (|mean| + sd) of the real-number mean / deviation (the oracle's
// sequential sum is the looser of the two), a clip bound within (1 + k) times that, so a sample further than
// g = 16 N 2^-53 (1 + k)(|mean| + sd) from a bound is decided identically -- inside that band the routine gives up
// (returns false) and the caller runs generic_pixel.  The band must be that narrow: a marked pixel is one with a
// sample within the FLOAT32 guard band of a bound, so the chance that it also lies within g is g over that band's
// width -- with g = 2^-34 (...) it was 5e-4, i.e. 3 of the headline stack's 5300 marks and 6 of the N = 512 stack's
// 12600 went to the generic routine, single-lane, and that one call WAS the duration of the launch (0.08 ms at
// N = 100, 0.85 ms at N = 512: ncu source view, profiles/r02_ncu_meanclip_coop512.txt).
template <typename Frames>
__device__ __forceinline__ bool meanstd_pixel_warp(const Frames& fp, const StackArgs& a, const int64_t p, const int lane,
                                                float* __restrict__ x) {
    const int N = a.N;
    const bool clip = a.maxiters != 0;
    int nk = 0;
#pragma unroll 8
    for (int i = lane; i < N; i += 32) {
        const float v = load_sample(fp.frame(i) + p, a);
        // sigma_clip drops non-finite samples up front; without clipping the nan-functions only skip NaN
        const bool ok = clip ? finite_f(v) : (v == v);
        x[i] = ok ? v : NAN;                            // NaN: not (or no longer) used
        nk += ok ? 1 : 0;
    }
    __syncwarp();
    nk = warp_sum(nk);
    auto mean_kept = [&](int cnt) -> double {
        double acc = 0.0;
        for (int i = lane; i < N; i += 32) { const float v = x[i]; if (v == v) acc = __dadd_rn(acc, (double)v); }
        return __ddiv_rn(warp_sum(acc), (double)cnt);
    };
    auto std_kept = [&](int cnt, double avg) -> double {
        double acc = 0.0;
        for (int i = lane; i < N; i += 32) {
            const float v = x[i];
            if (v == v) { const double d = __dsub_rn((double)v, avg); acc = __dadd_rn(acc, __dmul_rn(d, d)); }
        }
        return __dsqrt_rn(__ddiv_rn(warp_sum(acc), (double)cnt));
    };
    if (clip) {
        int it = 0;
        while (a.maxiters < 0 || it < a.maxiters) {
            ++it;
            if (nk == 0) break;
            const double avg = mean_kept(nk);
            const double sd = std_kept(nk, avg);
            const double lo = __dsub_rn(avg, __dmul_rn(sd, a.klo));
            const double hi = __dadd_rn(avg, __dmul_rn(sd, a.khi));
            const double g = (1.0 + fmax(a.klo, a.khi)) * (fabs(avg) + sd) * ((double)N * 1.7763568394002505e-15);   // N 2^-49
            const double lo_out = lo - g, lo_in = lo + g, hi_in = hi - g, hi_out = hi + g;
            bool inband = !(g == g) || !(lo_in <= hi_in);
            int changed = 0;
            for (int i = lane; i < N; i += 32) {        // (every lane only revisits its own elements)
                const float v = x[i];
                const double xv = (double)v;
                if (v == v) {
                    if (xv < lo_out || xv > hi_out) { x[i] = NAN; ++changed; }
                    else if (!(xv >= lo_in && xv <= hi_in)) inband = true;
                }
            }
            if (__any_sync(0xffffffffu, inband)) return false;
            changed = warp_sum(changed);
            nk -= changed;
            if (changed == 0) break;
        }
    }
    double data = (double)NAN, unc = (double)NAN;
    if (nk > 0) {
        data = mean_kept(nk);
        if (a.uncert) unc = __ddiv_rn(std_kept(nk, data), __dsqrt_rn((double)nk));
    }
    if (lane == 0) write_pixel(a, p, data, N - nk, unc, nk == 0);
    return true;
}

// WARP: the call is a kappa-sigma-about-the-mean one -- marked pixels are redone warp-cooperatively
// (meanstd_pixel_warp); otherwise the thread that finds a mark runs the generic routine on it.
template <int CAP, typename T, int ES, bool WARP>
__global__ void __launch_bounds__(MK_THREADS)
stack_marked_kernel(const __grid_constant__ FramePtrs<CAP, T> fp, const __grid_constant__ StackArgs a,
                    const unsigned char* __restrict__ plane, int64_t first_vec_pix, int64_t nvec) {
    constexpr int EPV = 16 / ES;                        // elements (pixels) per vector
    const int64_t pend = a.pix0 + a.npix;
    const int lane = threadIdx.x & 31;
    __shared__ float parked[WARP ? MK_THREADS / 32 : 1][WARP ? CAP : 1];       // (N <= CAP)
    __shared__ uint32_t list[WARP ? MK_LIST : 1];
    __shared__ int nlist;
    // all-ones in the rejection map can only be a mark when the count cannot reach it (N < 255 / 65535): no need
    // to confirm it against the output image then (one dependent load less per mark)
    const bool sure = !(ES == 1 && a.N >= 255);          // (ES 4 / 8: the candidate test IS the mark test)
    auto finish = [&](int64_t p, bool known) {          // WARP: called by all lanes with the same p
        if (p >= a.pix0 && p < pend && (known || pixel_is_marked(a, p))) {
            if constexpr (WARP) {
                if (!meanstd_pixel_warp(fp, a, p, lane, parked[threadIdx.x >> 5]) && lane == 0)
                    generic_pixel<CAP, FramePtrs<CAP, T>>(fp, a, p);
                __syncwarp();
            } else {
                generic_pixel<CAP, FramePtrs<CAP, T>>(fp, a, p);
            }
        }
    };
    const int64_t tid = (int64_t)blockIdx.x * MK_THREADS + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * MK_THREADS;
    // the pixels before the first and after the last whole aligned vector
    const int64_t tail0 = first_vec_pix + nvec * EPV;
    const int64_t unit = WARP ? (tid >> 5) : tid, nunits = WARP ? (nthreads >> 5) : nthreads;
    for (int64_t p = a.pix0 + unit; p < first_vec_pix; p += nunits) finish(p, false);
    for (int64_t p = tail0 + unit; p < pend; p += nunits) finish(p, false);
    const uint4* const vecs = reinterpret_cast<const uint4*>(plane + first_vec_pix * ES);
    // WARP: the marks go to the CTA's shared-memory list while it scans, and its warps share them out once the
    // scan is done -- one mark takes a warp ~10 us of dependent latencies (is-it-marked load, sample loads, a
    // few float64 iterations), so what matters is how many marks the unluckiest warp ends up with: finishing
    // them where they are found left ~5 on one warp where the average is 0.5 (measured 100 us for the 5300 marks
    // of the 100 x 61 Mpixel stack; sharing out after EVERY sweep, with CTA barriers in the loop: 147 us).
    if constexpr (WARP) {
        if (threadIdx.x == 0) nlist = 0;
        __syncthreads();
    }
    for (int64_t b = tid - lane; b < nvec; b += nthreads * MK_UNROLL) {      // (warp-uniform trip count)
        uint4 v[MK_UNROLL];
#pragma unroll
        for (int u = 0; u < MK_UNROLL; ++u) {
            const int64_t i = b + lane + (int64_t)u * nthreads;
            v[u] = i < nvec ? __ldcs(vecs + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < MK_UNROLL; ++u) {
            unsigned cand = vector_mark_candidates<ES>(v[u]);
            const int64_t p0 = first_vec_pix + (b + lane + (int64_t)u * nthreads) * EPV;
            if constexpr (WARP) {
                while (cand) {
                    const int slot = atomicAdd(&nlist, 1);
                    if (slot >= MK_LIST) break;             // list full: the rest is finished in place
                    list[slot] = (uint32_t)(p0 + (__ffs((int)cand) - 1));
                    cand &= cand - 1;
                }
                unsigned hits = __ballot_sync(0xffffffffu, cand != 0u);
                while (hits) {
                    const int src = __ffs((int)hits) - 1;
                    hits &= hits - 1;
                    const int64_t q0 = __shfl_sync(0xffffffffu, p0, src);
                    unsigned c = __shfl_sync(0xffffffffu, cand, src);
                    while (c) { finish(q0 + (__ffs((int)c) - 1), sure); c &= c - 1; }
                }
            } else {
                while (cand) { finish(p0 + (__ffs((int)cand) - 1), sure); cand &= cand - 1; }
            }
        }
    }
    if constexpr (WARP) {
        __syncthreads();
        const int n = nlist < MK_LIST ? nlist : MK_LIST;
        for (int k = threadIdx.x >> 5; k < n; k += MK_THREADS / 32) finish((int64_t)list[k], sure);
    }
}

template <int CAP, typename T>
int launch_marked(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    FramePtrs<CAP, T> fp;
    for (int i = 0; i < CAP; ++i) fp.p[i] = i < a.N ? frames[i] : nullptr;
    const unsigned char* plane = reinterpret_cast<const unsigned char*>(a.nrej ? a.nrej : a.out);
    const int es = a.nrej ? (a.nrej_u16 ? 2 : 1) : (a.out_f64 ? 8 : 4);
    // whole 16-byte vectors of the plane inside the range
    const uintptr_t b0 = (uintptr_t)plane + (uintptr_t)a.pix0 * es, b1 = (uintptr_t)plane + (uintptr_t)(a.pix0 + a.npix) * es;
    const uintptr_t v0 = (b0 + 15) & ~(uintptr_t)15;
    int64_t first_vec_pix = a.pix0 + a.npix, nvec = 0;
    if (v0 + 16 <= b1 && (v0 - (uintptr_t)plane) % es == 0) {
        first_vec_pix = (int64_t)((v0 - (uintptr_t)plane) / es);
        nvec = (int64_t)((b1 - v0) / 16);
    }
    // a plane that cannot be read in vectors is walked pixel by pixel
    const int64_t work = nvec > 0 ? nvec : a.npix;
    int64_t blocks = (work + MK_THREADS - 1) / MK_THREADS;
    if (blocks > (int64_t)APGPU_NUM_SMS * 8) blocks = (int64_t)APGPU_NUM_SMS * 8;
    const bool warp = a.method == APGPU_METHOD_AVERAGE &&
                      (a.maxiters == 0 || (a.cen == APGPU_CEN_MEAN && a.dev == APGPU_DEV_STD));
#define MK_LAUNCH(ES, WARP) stack_marked_kernel<CAP, T, ES, WARP><<<(unsigned)blocks, MK_THREADS, 0, st>>>(fp, a, plane, first_vec_pix, nvec)
    switch (es) {
        case 1: if (warp) MK_LAUNCH(1, true); else MK_LAUNCH(1, false); break;
        case 2: if (warp) MK_LAUNCH(2, true); else MK_LAUNCH(2, false); break;
        case 4: if (warp) MK_LAUNCH(4, true); else MK_LAUNCH(4, false); break;
        default: if (warp) MK_LAUNCH(8, true); else MK_LAUNCH(8, false); break;
    }
#undef MK_LAUNCH
    APGPU_LAUNCH_CHECK("stack_marked_kernel");
    return APGPU_OK;
}

template <typename T>
int launch_marked_any(const T* const* frames, const StackArgs& a, cudaStream_t st) {
    if (a.npix <= 0) return APGPU_OK;
    if (a.N <= 32) return launch_marked<32, T>(frames, a, st);
    if (a.N <= 128) return launch_marked<128, T>(frames, a, st);
    return launch_marked<1024, T>(frames, a, st);
}

int stack_launch_marked(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    return launch_marked_any<float>(frames, a, st);
}
int stack_launch_marked(const uint16_t* const* frames, const StackArgs& a, cudaStream_t st) {
    return launch_marked_any<uint16_t>(frames, a, st);
}

int stack_launch_generic(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    return launch_generic_any<float>(frames, a, st);
}
int stack_launch_generic(const uint16_t* const* frames, const StackArgs& a, cudaStream_t st) {
    return launch_generic_any<uint16_t>(frames, a, st);
}

}  // namespace apgpu_stack

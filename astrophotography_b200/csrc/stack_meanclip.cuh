// Frame-stack reducer: meanclip<NB, NLO> register-resident kappa-sigma kernels.  See stack_common.cuh.
#pragma once
#include "stack_common.cuh"

namespace apgpu_stack {

// ---------------------------------------------------------------------------
// meanclip<NB, NLO>: kappa-sigma clip about the mean, N in (NLO, NB]
// ---------------------------------------------------------------------------
// i < NLO is known at compile time to be a real frame; only the NB-NLO tail
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
constexpr int WT = 32;      // pixels per warp tile of the warp-granular staged kernels

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int J, int NP>
__device__ __forceinline__ float2 pair_or_zero(const float2 (&y)[NP]) {
    if constexpr (J < NP) return y[J];
    else return make_float2(0.f, 0.f);
}
template <int J, int NP>
__device__ __forceinline__ void put_pair(float2 (&y)[NP], const float2 v) {
    if constexpr (J < NP) y[J] = v;
}

struct NoHook { __device__ __forceinline__ void operator()(float) const {} };

// P lanes of a warp share one pixel (lane = r * (32/P) + q holds the samples i = P*j + r of pixel q):
// butterfly all-reductions over those lanes.  Every lane of the group ends up with the identical
// value (the operations are commutative), so the lanes take identical decisions.  P = 1: no-ops.
// Which frame does sample j of lane r (of the P lanes of a pixel) hold?
template <int P> struct InterleavedFrames {            // i = P*j + r
    static __device__ __forceinline__ constexpr int idx(int j, int r) { return P * j + r; }
};
template <int P> struct SwizzledFrames {               // i = 8*(j / RB) + RB*r + j % RB, RB = 8/P (see stack_meanclip_coop.cuh)
    static constexpr int RB = 8 / P;
    static __device__ __forceinline__ constexpr int idx(int j, int r) { return 8 * (j / RB) + RB * r + j % RB; }
};

template <int P> struct LaneGroup {
    static constexpr int PIXW = 32 / P;
    static __device__ __forceinline__ float sum(float v, unsigned m) {
#pragma unroll
        for (int o = PIXW; o < 32; o <<= 1) v += __shfl_xor_sync(m, v, o);
        return v;
    }
    static __device__ __forceinline__ double sum(double v, unsigned m) {
#pragma unroll
        for (int o = PIXW; o < 32; o <<= 1) v = __dadd_rn(v, __shfl_xor_sync(m, v, o));
        return v;
    }
    static __device__ __forceinline__ int sum(int v, unsigned m) {
#pragma unroll
        for (int o = PIXW; o < 32; o <<= 1) v += __shfl_xor_sync(m, v, o);
        return v;
    }
    static __device__ __forceinline__ float max(float v, unsigned m) {
#pragma unroll
        for (int o = PIXW; o < 32; o <<= 1) v = fmaxf(v, __shfl_xor_sync(m, v, o));
        return v;
    }
    static __device__ __forceinline__ float min(float v, unsigned m) {
#pragma unroll
        for (int o = PIXW; o < 32; o <<= 1) v = fminf(v, __shfl_xor_sync(m, v, o));
        return v;
    }
    static __device__ __forceinline__ bool any(bool b, unsigned m) {
        int v = b ? 1 : 0;
#pragma unroll
        for (int o = PIXW; o < 32; o <<= 1) v |= __shfl_xor_sync(m, v, o);
        return v != 0;
    }
};

// after_sums(S2) is called once every sample of the pixel has been consumed (the staged kernels
// re-arm their shared-memory stage there); it runs before any early exit.
//
// NB samples per lane; with P > 1 lanes per pixel the stack holds up to NB * P frames, NLO is the
// bucket's lower bound on the TOTAL frame count, `pivot_in` the pixel's pivot, `gmask` the lanes of
// the pixel and `r` this lane's index among them (only r == 0 writes or marks the pixel).
template <int NB, int NLO, bool SYM, typename AfterSums = NoHook, int P = 1, typename Frames = FramePtrs<NB>,
          typename IMap = InterleavedFrames<P>>
__device__ __forceinline__ void meanclip_pixel(float2 (&y)[NB / 2], const Frames& fp,
                                               const StackArgs& a, const int64_t p,
                                               const AfterSums& after_sums = AfterSums(),
                                               const float pivot_in = 0.f, const unsigned gmask = 0xffffffffu,
                                               const int r = 0) {
    static_assert(NB % 2 == 0, "meanclip buckets must be even");
    using G = LaneGroup<P>;
    const int N = a.N;
    constexpr int NP = NB / 2;                         // register pairs
    constexpr int GP = 4;                              // pairs (8 samples) per group
    constexpr int NG = (NP + GP - 1) / GP;
    static_assert(NG <= 25, "flag word / rare-pass switch too small");

    // Pivot: median of the first three frames (robust to one outlier).  All
    // float32 arithmetic below is on y = x - pivot: sums stay small and the
    // variance is free of catastrophic cancellation.
    const float pivot = (P == 1) ? med3(y[0].x, y[0].y, y[1].x) : pivot_in;
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
                if constexpr (P == 1) {
                    if (2 * j >= NLO)                      // bucket tail: padding beyond N becomes y = 0
                        d = __fmul2_rn(d, make_float2(a.tailmask[2 * j - NLO], a.tailmask[2 * j + 1 - NLO]));
                } else {
                    // only the top of the bucket can lie beyond N (frame index of sample j of lane r: IMap)
                    if (IMap::idx(2 * j, P - 1) > NLO && !(IMap::idx(2 * j, r) < N)) d.x = 0.f;
                    if (IMap::idx(2 * j + 1, P - 1) > NLO && !(IMap::idx(2 * j + 1, r) < N)) d.y = 0.f;
                }
                y[j] = d;
                s1 = __fadd2_rn(s1, d);
                s2 = __ffma2_rn(d, d, s2);
            }
        }
        S1 += s1.x + s1.y;
        S2 += s2.x + s2.y;
    }
    after_sums(S2);
    constexpr unsigned FULL = 0xffffffffu;
    if constexpr (P > 1) {
        S1 = G::sum(S1, FULL);
        S2 = G::sum(S2, FULL);
    }
    // NaN input poisons S1/S2, inf input (or overflow) makes S2 infinite: the
    // generic routine owns those semantics.
    const bool nonfinite = !(S2 <= FLT_MAX) || !(fabsf(S1) <= FLT_MAX);
    if constexpr (P == 1) {
        if (nonfinite) { mark_pixel(a, p, MARK_NONFINITE); return; }
    }

    int nk = N;
    const float klo = (float)a.klo, khi = (float)a.khi;
    const float kmax = fmaxf(klo, khi);
    bool uncertain = nonfinite;
    int why = 0;                                       // (diagnostic: payload of the mark)
    int it = 0;
    // Sums after a rejection are updated by SUBTRACTING the rejected samples' contributions
    // while that is accurate (what is left of sum(y^2) is at least half of the last fully
    // summed value S2_fresh: the rounding errors, bounded relative to the old sums, are then at
    // most doubled relative to the new ones -- m below grows accordingly); a pixel that loses
    // more than that (a cosmic-ray hit) has its sums rebuilt from the registers.
    float S2_fresh = S2;
    int nsub = 0;
    // M_prev: an upper bound on |y - c_prev| over the survivors of the previous iteration.  When
    // M_prev + |c - c_prev| is below the new inner bound, every survivor is certainly inside
    // the new clip limits: converged without another pass over the samples.
    float M_prev = -1.f, c_prev = 0.f;

    // clip bounds of one iteration, all in the pivot-shifted frame
    struct Bounds { float c, t_in, lo_in, hi_in, ylo_out, ylo_in, yhi_in, yhi_out; };
    // returns 0: bounds ready, 1: float64 must decide (degenerate / pivot outside), 2: converged without a pass
    auto make_bounds = [&](Bounds& B) -> int {
        // Approximate reciprocal / square root (MUFU, <= 2 ulp each) instead of the IEEE sequences: this
        // scalar block runs once per clip iteration per pixel and the correctly rounded versions cost
        // ~20 instructions each.  Their few-ulp errors in c and sd are covered by the second term of
        // the guard band g below; anything degenerate (sd = 0, inf, NaN) fails the g test and
        // leaves for the float64 routine.
        const float rn = fast_rcp((float)nk);
        const float c = S1 * rn;
        const float ex2 = S2 * rn;
        const float var = ex2 - c * c;
        const float sd = fast_sqrt(fmaxf(var, 0.f));
        // Bound on |threshold_f32 - threshold_exact| + |y_f32 - y_exact|:
        // group-wise summation (GP + 1 + NG terms deep, + log2 P butterfly steps), unit roundoff
        // doubled for safety; after nsub subtractive updates the summation error is relative to
        // sums up to twice as large, plus one rounding per subtraction.
        const float u2 = 1.1920929e-7f;                       // 2^-23
        constexpr int M0 = GP + NG + 9 + (P > 1 ? 4 : 0);
        const float m = (float)M0 + (nsub ? (float)(M0 + 2 * nsub) : 0.f);
        const float g = 1.0001f * m * u2 * (fast_sqrt(ex2) + 1.5f * kmax * ex2 * fast_rcp(sd)) +
                        12.f * u2 * (fabsf(c) + kmax * sd);
        if (!(g < 0.25f * kmax * sd)) return 1;               // degenerate (var ~ 0): let float64 decide
        // inner (certainly kept inside) and outer (certainly rejected outside) bounds on t = y - c
        B.c = c;
        B.lo_in = -klo * sd + g;
        B.hi_in = khi * sd - g;
        const float lo_out = -klo * sd - g, hi_out = khi * sd + g;
        B.t_in = fminf(-B.lo_in, B.hi_in);                    // symmetric inner bound on |t|
        B.ylo_out = c + lo_out; B.ylo_in = c + B.lo_in; B.yhi_in = c + B.hi_in; B.yhi_out = c + hi_out;
        // A rejected sample is overwritten with y = 0 (the pivot), so the pivot itself must sit
        // strictly inside the inner bounds: then zeros are never rejected (again) and add nothing.
        if (!(B.ylo_in < 0.f && B.yhi_in > 0.f)) return 1;
        if (M_prev >= 0.f && (M_prev + fabsf(c - c_prev)) * 1.000002f < B.t_in) return 2;
        return 0;
    };
    // test pass: tight straight-line code, no per-sample predicates, no sums (the sums of the
    // survivors only change when something is rejected).  flags bit g: group g holds a sample
    // outside the inner bounds; M_nf: max |y - c| over the groups that are not flagged.
    auto test_pass = [&](const Bounds& B, uint64_t& flags, float& M_nf) {
        const float2 negc = make_float2(-B.c, -B.c);
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
            const bool flagged = SYM ? (tmax >= B.t_in) : (tmax >= B.hi_in || tmin <= B.lo_in);
            if (flagged) flags |= (uint64_t)1 << gidx;
            const float tabs = SYM ? tmax : fmaxf(tmax, -tmin);
            M_nf = fmaxf(M_nf, flagged ? 0.f : tabs);
        }
    };
    // Rejection pass over the flagged groups, sample by sample: r1, r2 sums over the samples
    // rejected now, [vmin, vmax] range of the survivors (and zeros) of the flagged groups.
    // Each lane walks its own flagged groups (lowest set bit first); the 8 samples of the group
    // are gathered into scratch registers by a switch on the group index (register arrays need
    // static indices, so lanes with different groups serialise over ~10 instructions of moves),
    // the ~70 instructions of per-sample work then run ONCE per round for all lanes together, and a
    // second switch scatters the survivors back.  (Unrolling the per-sample work into every
    // group instead made the warp execute it once per distinct flagged group: ~7 times per pass.)
    auto rare_pass = [&](const Bounds& B, const uint64_t flags, float& r1, float& r2, int& nrej_it,
                         float& vmax, float& vmin) {
        uint32_t f = (uint32_t)flags;
        while (f) {
            const int g = __ffs((int)f) - 1;
            f &= f - 1;
            float2 t0 = make_float2(0.f, 0.f), t1 = t0, t2 = t0, t3 = t0;
#define MC_GATHER(G) case G: if constexpr (G < NG) { t0 = pair_or_zero<4 * G, NP>(y); t1 = pair_or_zero<4 * G + 1, NP>(y); \
                                                     t2 = pair_or_zero<4 * G + 2, NP>(y); t3 = pair_or_zero<4 * G + 3, NP>(y); } break;
            switch (g) {
                MC_GATHER(0) MC_GATHER(1) MC_GATHER(2) MC_GATHER(3) MC_GATHER(4) MC_GATHER(5) MC_GATHER(6) MC_GATHER(7)
                MC_GATHER(8) MC_GATHER(9) MC_GATHER(10) MC_GATHER(11) MC_GATHER(12) MC_GATHER(13) MC_GATHER(14)
                MC_GATHER(15) MC_GATHER(16) MC_GATHER(17) MC_GATHER(18) MC_GATHER(19) MC_GATHER(20) MC_GATHER(21)
                MC_GATHER(22) MC_GATHER(23) MC_GATHER(24)
                default: break;
            }
#undef MC_GATHER
            float v[8] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, t3.x, t3.y};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // compare y against bounds shifted by c (not t = y - c)
                const bool keep = (v[k] >= B.ylo_out) && (v[k] <= B.yhi_out);
                nrej_it += keep ? 0 : 1;                         // certainly rejected
                const float vr = keep ? 0.f : v[k];
                const float vk = keep ? v[k] : 0.f;
                v[k] = vk;
                vmax = fmaxf(vmax, vk);
                vmin = fminf(vmin, vk);
                r1 += vr;
                r2 = fmaf(vr, vr, r2);
            }
            t0 = make_float2(v[0], v[1]); t1 = make_float2(v[2], v[3]);
            t2 = make_float2(v[4], v[5]); t3 = make_float2(v[6], v[7]);
#define MC_SCATTER(G) case G: if constexpr (G < NG) { put_pair<4 * G, NP>(y, t0); put_pair<4 * G + 1, NP>(y, t1); \
                                                      put_pair<4 * G + 2, NP>(y, t2); put_pair<4 * G + 3, NP>(y, t3); } break;
            switch (g) {
                MC_SCATTER(0) MC_SCATTER(1) MC_SCATTER(2) MC_SCATTER(3) MC_SCATTER(4) MC_SCATTER(5) MC_SCATTER(6) MC_SCATTER(7)
                MC_SCATTER(8) MC_SCATTER(9) MC_SCATTER(10) MC_SCATTER(11) MC_SCATTER(12) MC_SCATTER(13) MC_SCATTER(14)
                MC_SCATTER(15) MC_SCATTER(16) MC_SCATTER(17) MC_SCATTER(18) MC_SCATTER(19) MC_SCATTER(20) MC_SCATTER(21)
                MC_SCATTER(22) MC_SCATTER(23) MC_SCATTER(24)
                default: break;
            }
#undef MC_SCATTER
        }
    };
    // rebuild the sums of the survivors from the registers
    auto resum = [&](float& n1, float& n2) {
#pragma unroll
        for (int gidx = 0; gidx < NG; ++gidx) {
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
    };

    if constexpr (P == 1) {
        while (a.maxiters != 0 && (a.maxiters < 0 || it < a.maxiters)) {
            ++it;
            if (S2 == 0.f) break;            // all survivors equal the pivot: bounds [c,c], nothing to reject
            Bounds B;
            const int st = make_bounds(B);
            if (st == 1) { uncertain = true; why = MARK_BOUNDS; break; }
            if (st == 2) break;              // converged, no pass needed
            uint64_t flags = 0;
            float M_nf = 0.f;
            test_pass(B, flags, M_nf);
            if (flags == 0) break;           // every survivor is certainly inside the bounds: converged
            float r1 = 0.f, r2 = 0.f, vmax = 0.f, vmin = 0.f;
            int nrej_it = 0;
            rare_pass(B, flags, r1, r2, nrej_it, vmax, vmin);
            nk -= nrej_it;
            // a survivor inside the guard band: float64 must decide
            if (!(vmin > B.ylo_in && vmax < B.yhi_in)) { uncertain = true; why = MARK_BAND; break; }
            if (nk == 0) break;
            M_prev = fmaxf(M_nf, fmaxf(vmax - B.c, B.c - vmin)) * 1.000001f;
            c_prev = B.c;
            const float S2n = S2 - r2;
            if (S2n >= 0.5f * S2_fresh) {
                S1 -= r1;
                S2 = S2n;
                ++nsub;
            } else {
                float n1 = 0.f, n2 = 0.f;
                resum(n1, n2);
                S1 = n1;
                S2 = n2;
                S2_fresh = n2;
                nsub = 0;
            }
        }
        if (uncertain || nk == 0) { mark_pixel(a, p, nonfinite ? MARK_NONFINITE : (nk == 0 ? MARK_EMPTY : why)); return; }
    } else {
        // P lanes per pixel: the loop is WARP-uniform (every lane stays in it until the warp's last
        // pixel is done, finished pixels idle through it) so that the butterflies are full-mask
        // shuffles at fixed points of a converged instruction stream.
        bool done = (a.maxiters == 0) || nonfinite;
        for (;;) {
            const bool act = !done && (a.maxiters < 0 || it < a.maxiters);
            if (!__any_sync(FULL, act)) break;
            done = !act;
            bool pass = false;
            Bounds B;
            B.c = 0.f; B.t_in = 0.f; B.lo_in = 0.f; B.hi_in = 0.f; B.ylo_out = 0.f; B.ylo_in = 0.f; B.yhi_in = 0.f; B.yhi_out = 0.f;
            if (act) {
                ++it;
                if (S2 == 0.f) {
                    done = true;
                } else {
                    const int st = make_bounds(B);
                    if (st == 1) { uncertain = true; why = MARK_BOUNDS; done = true; }
                    else if (st == 2) done = true;
                    else pass = true;
                }
            }
            uint64_t flags = 0;
            float M_nf = 0.f;
            if (pass) test_pass(B, flags, M_nf);
            const bool anyf = G::any(flags != 0, FULL);
            M_nf = G::max(M_nf, FULL);
            if (pass && !anyf) { done = true; pass = false; }
            float r1 = 0.f, r2 = 0.f, vmax = 0.f, vmin = 0.f;
            int nrej_it = 0;
            if (pass) rare_pass(B, flags, r1, r2, nrej_it, vmax, vmin);
            r1 = G::sum(r1, FULL); r2 = G::sum(r2, FULL); nrej_it = G::sum(nrej_it, FULL);
            vmax = G::max(vmax, FULL); vmin = G::min(vmin, FULL);
            bool need_resum = false;
            if (pass) {
                nk -= nrej_it;
                if (!(vmin > B.ylo_in && vmax < B.yhi_in)) { uncertain = true; why = MARK_BAND; done = true; }
                else if (nk == 0) done = true;
                else {
                    M_prev = fmaxf(M_nf, fmaxf(vmax - B.c, B.c - vmin)) * 1.000001f;
                    c_prev = B.c;
                    const float S2n = S2 - r2;
                    if (S2n >= 0.5f * S2_fresh) { S1 -= r1; S2 = S2n; ++nsub; }
                    else need_resum = true;
                }
            }
            if (__any_sync(FULL, need_resum)) {
                float n1 = 0.f, n2 = 0.f;
                resum(n1, n2);
                n1 = G::sum(n1, FULL);
                n2 = G::sum(n2, FULL);
                if (need_resum) { S1 = n1; S2 = n2; S2_fresh = n2; nsub = 0; }
            }
        }
    }

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
        if constexpr (P > 1) {
            sum1 = G::sum(sum1, FULL);
            sum2 = G::sum(sum2, FULL);
        }
    }
    if constexpr (P > 1) {
        if (r != 0) return;
        if (uncertain || nk == 0) { mark_pixel(a, p, nonfinite ? MARK_NONFINITE : (nk == 0 ? MARK_EMPTY : why)); return; }
    }
    const double cy = __ddiv_rn(sum1, (double)nk);
    // (uint16 frames are staged as 2^23 + value: pivot - sample_bias is the pivot's value, exactly)
    const double mean = __dadd_rn((double)(pivot - a.sample_bias), cy);
    double unc_out = (double)NAN;
    if (a.uncert) {
        double var = __dsub_rn(__ddiv_rn(sum2, (double)nk), __dmul_rn(cy, cy));
        unc_out = __ddiv_rn(__dsqrt_rn(var > 0.0 ? var : 0.0), __dsqrt_rn((double)nk));
    }
    write_pixel(a, p, mean, N - nk, unc_out, 0);
}

// Direct kernel: one block per 128-pixel tile, samples loaded straight from
// global memory.
template <int NB, int NLO, bool SYM, typename T>
__global__ void __launch_bounds__(TPB, meanclip_min_blocks(NB))
stack_meanclip_kernel(const __grid_constant__ FramePtrs<NB, T> fp, const __grid_constant__ StackArgs a) {
    const int64_t p = a.pix0 + (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (p >= a.pix0 + a.npix) return;
    const uint32_t p32 = (uint32_t)p;                  // host guarantees H*W < 2^32
    float2 y[NB / 2];
#pragma unroll
    for (int j = 0; j < NB / 2; ++j) {                 // padding slots point at frame 0 (masked later)
        y[j].x = load_sample_biased(fp.p[2 * j] + p32, a);
        y[j].y = load_sample_biased(fp.p[2 * j + 1] + p32, a);
    }
    meanclip_pixel<NB, NLO, SYM, NoHook, 1, FramePtrs<NB, T>>(y, fp, a, p);
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
// Tensor-map TMA staged persistent kernel: one bulk tensor copy per warp tile
// ---------------------------------------------------------------------------
// When the N frames are equally spaced in memory (a [N][H*W] cube, which is what the host
// pipeline and torch cubes are) the whole stack is ONE 2-D tensor: dim0 = pixel, dim1 = frame.
// Every warp is its own pipeline: it owns a [N][32-pixel] shared-memory stage and an mbarrier;
// one elected lane issues a single cp.async.bulk.tensor.2d (box = 32 pixels x N frames, 128 B
// per frame row: UTMALDG in SASS) per tile, the lanes read their column out of the stage
// (conflict-free LDS with immediate offsets -- no per-sample address arithmetic, no global-load
// instructions in the instruction stream), and the copy for the warp's NEXT tile is issued
// as soon as every sample has been consumed, so its HBM latency hides behind the clipping
// arithmetic of the current tile.  No CTA barrier anywhere.
template <int NB, int NLO, bool SYM>
__global__ void __launch_bounds__(TPB, meanclip_min_blocks(NB))
stack_meanclip_tmap_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FramePtrs<NB> fp,
                           const __grid_constant__ StackArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * NB * WT;          // [NB][32]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)(TPB / 32) * NB * WT * sizeof(float)) + warp;
    const int N = a.N;
    if (lane == 0) mbar_init(bar, 1);
    for (int i = N * WT + lane; i < NB * WT; i += 32) stage[i] = 0.f;         // padding rows: never copied into
    __syncwarp();
    // one box = the whole tile (cutting it into several concurrent boxes was measured: no gain)
    // (all pixel / tile indices fit 32 bits: the host only takes this path below 2^31 pixels; 64-bit
    // loop state cost spilled registers and a local-memory reload at the end of every tile)
    const int pix0 = (int)a.pix0;
    auto issue = [&](int t) {                       // (policy / byte count rebuilt here: not worth live registers)
        mbar_expect_tx(bar, (uint32_t)a.N * WT * sizeof(float));
        tma_load_2d(stage, &tmap, pix0 + t * WT, 0, bar, l2_evict_first_policy());
    };
    // Each CTA owns a run of 4 * tiles_per_warp consecutive warp tiles (its warps interleave);
    // the hardware CTA scheduler balances the runs over the SMs (a fully persistent grid with a
    // static tile assignment left SMs idle 17 % of the time: SMs do not all see the same
    // memory throughput).
    const int ntiles = (int)(a.npix / WT);                                    // full warp tiles (host launches the tail)
    const int run = (TPB / 32) * a.tiles_per_warp;
    int tile = (int)blockIdx.x * run + warp;
    const int tile_end = min((int)(blockIdx.x + 1) * run, ntiles);
    constexpr int nwarps = TPB / 32;
    uint32_t parity = 0;
    if (tile < tile_end && lane == 0) issue(tile);
    for (; tile < tile_end; tile += nwarps) {
        while (!mbar_try_wait(bar, parity)) {}
        parity ^= 1u;
        float2 y[NB / 2];
#pragma unroll
        for (int j = 0; j < NB / 2; ++j) {
            y[j].x = stage[(2 * j) * WT + lane];
            y[j].y = stage[(2 * j + 1) * WT + lane];
        }
        const int next = tile + nwarps;
        // re-arm the stage once the sums (which depend on every staged sample of every lane of
        // this warp instruction stream) exist: the predicate below carries that dependence
        // (every lane's LDS reads of the stage are ordered before the async-proxy write: converge first)
        auto rearm = [&](float s2) {
            __syncwarp();
            if (lane == 0 && next < tile_end && s2 != -1.f) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(next);
            }
            __syncwarp();
        };
        meanclip_pixel<NB, NLO, SYM>(y, fp, a, (int64_t)(uint32_t)(pix0 + tile * WT + lane), rearm);
    }
}

// The same pipeline for uint16 frames (apgpu_stack_reduce_u16).  A 128-byte row of the stage now holds 64
// pixels (narrower rows lose bandwidth in proportion: a fixed number of requests in flight per SM), so a
// warp tile is 64 pixels x N frames and the warp reduces it in two halves of 32 pixels -- the same
// shared-memory footprint and the same bytes per copy as the float32 kernel, twice the pixels.  A raw
// sample becomes the float 2^23 + value with two ALU-pipe instructions (PRMT picks the byte order and puts
// the exponent byte on top, LOP3 folds FITS' BZERO): the pivot shift y = x - pivot that the algorithm does
// anyway then removes the 2^23 again, exactly, so the integer-to-float conversion costs no conversion
// instruction (I2F runs at a quarter of the rate).
constexpr int WT16 = 64;

template <int NB, int NLO, bool SYM>
__global__ void __launch_bounds__(TPB, meanclip_min_blocks(NB))
stack_meanclip_tmap16_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FramePtrs<NB, uint16_t> fp,
                             const __grid_constant__ StackArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint16_t* stage = reinterpret_cast<uint16_t*>(smem_raw) + (size_t)warp * NB * WT16;   // [NB][64]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)(TPB / 32) * NB * WT16 * sizeof(uint16_t)) + warp;
    const int N = a.N;
    if (lane == 0) mbar_init(bar, 1);
    for (int i = N * WT16 + lane; i < NB * WT16; i += 32) stage[i] = 0;       // padding rows: never copied into
    __syncwarp();
    const int pix0 = (int)a.pix0;
    auto issue = [&](int t) {
        mbar_expect_tx(bar, (uint32_t)a.N * WT16 * sizeof(uint16_t));
        tma_load_2d(stage, &tmap, pix0 + t * WT16, 0, bar, l2_evict_first_policy());
    };
    const int ntiles = (int)(a.npix / WT16);                                  // full warp tiles (host launches the tail)
    const int run = (TPB / 32) * a.tiles_per_warp;
    int tile = (int)blockIdx.x * run + warp;
    const int tile_end = min((int)(blockIdx.x + 1) * run, ntiles);
    constexpr int nwarps = TPB / 32;
    uint32_t parity = 0;
    if (tile < tile_end && lane == 0) issue(tile);
    for (; tile < tile_end; tile += nwarps) {
        while (!mbar_try_wait(bar, parity)) {}
        parity ^= 1u;
        const int next = tile + nwarps;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const uint16_t* col = stage + half * 32 + lane;
            float2 y[NB / 2];
#pragma unroll
            for (int j = 0; j < NB / 2; ++j) {
                y[j].x = staged_sample_biased(col + (2 * j) * WT16, a);
                y[j].y = staged_sample_biased(col + (2 * j + 1) * WT16, a);
            }
            // the stage is re-armed once the second half's samples have been consumed
            auto rearm = [&](float s2) {
                __syncwarp();
                if (half == 1 && lane == 0 && next < tile_end && s2 != -1.f) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(next);
                }
                __syncwarp();
            };
            meanclip_pixel<NB, NLO, SYM, decltype(rearm), 1, FramePtrs<NB, uint16_t>>(
                y, fp, a, (int64_t)(uint32_t)(pix0 + tile * WT16 + half * 32 + lane), rearm);
        }
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
    // CTA runs of 4 * tiles_per_warp consecutive warp tiles, balanced by the hardware CTA scheduler
    // (see the tensor-map kernel: a static persistent assignment left SMs idle)
    const int64_t ntiles = a.npix / WT;                                       // full warp tiles (host launches the tail)
    const int64_t run = (int64_t)(TPB / 32) * a.tiles_per_warp;
    int64_t tile = (int64_t)blockIdx.x * run + warp;
    const int64_t tile_end = ((int64_t)(blockIdx.x + 1) * run < ntiles) ? (int64_t)(blockIdx.x + 1) * run : ntiles;
    constexpr int64_t nwarps = TPB / 32;
    if (tile < tile_end) issue_warp_tile(ptab, N, a.pix0 + tile * WT, stage, lane);
    for (; tile < tile_end; tile += nwarps) {
        cp_async_wait_all();
        __syncwarp();                                                         // every lane's copies are visible
        float2 y[NB / 2];
#pragma unroll
        for (int j = 0; j < NB / 2; ++j) {
            y[j].x = stage[(2 * j) * WT + lane];
            y[j].y = stage[(2 * j + 1) * WT + lane];
        }
        const int64_t next = tile + nwarps;
        // re-arm the stage once every staged sample has been consumed (the sums exist)
        auto rearm = [&](float s2) {
            __syncwarp();
            if (next < tile_end && s2 != -1.f) issue_warp_tile(ptab, N, a.pix0 + next * WT, stage, lane);
        };
        meanclip_pixel<NB, NLO, SYM>(y, fp, a, a.pix0 + tile * WT + lane, rearm);
    }
}

template <int NB, int NLO, bool SYM, typename T>
int launch_meanclip_sym(const FramePtrs<NB, T>& fp, const StackArgs& a, int staging, cudaStream_t st) {
    // staging: 0 direct global loads, 1 CTA-wide TMA bulk copies, 2 warp-granular cp.async pipeline,
    //          3 warp-granular tensor-map TMA pipeline (equally spaced frames only)
    // (uint16 frames: 3 or 0 only)
    StackArgs rest = a;
    if constexpr (sizeof(T) == 2) {
        if (staging == 3) {
            const int64_t ntiles = a.npix / WT16;
            CUtensorMap tmap;
            const int64_t stride = (const char*)fp.p[1] - (const char*)fp.p[0];
            if (ntiles > 0 && encode_stack_tensor_map(&tmap, fp.p[0], (uint64_t)(a.pix0 + a.npix), a.N, (uint64_t)stride, WT16)) {
                const size_t smem = (size_t)(TPB / 32) * NB * WT16 * sizeof(uint16_t) + (TPB / 32) * sizeof(uint64_t);
                APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_tmap16_kernel<NB, NLO, SYM>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                StackArgs at = a;
                at.tiles_per_warp = stack_tmap_tiles_per_warp();
                const int64_t run = (int64_t)(TPB / 32) * at.tiles_per_warp;
                const int64_t grid = (ntiles + run - 1) / run;
                stack_meanclip_tmap16_kernel<NB, NLO, SYM><<<(unsigned)grid, TPB, smem, st>>>(tmap, fp, at);
                APGPU_LAUNCH_CHECK("stack_meanclip_tmap16_kernel");
                rest.pix0 = a.pix0 + ntiles * WT16;          // the < 64-pixel tail goes through the direct kernel
                rest.npix = a.npix - ntiles * WT16;
            } else {
                stack_note_staging(0);
            }
        }
    } else {
    if (staging == 3) {
        const int64_t ntiles = a.npix / WT;
        CUtensorMap tmap;
        const int64_t stride = (const char*)fp.p[1] - (const char*)fp.p[0];
        if (ntiles > 0 && encode_stack_tensor_map(&tmap, fp.p[0], (uint64_t)(a.pix0 + a.npix), a.N, (uint64_t)stride, WT)) {
            const size_t smem = (size_t)(TPB / 32) * NB * WT * sizeof(float) + (TPB / 32) * sizeof(uint64_t);
            APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_tmap_kernel<NB, NLO, SYM>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            StackArgs at = a;
            at.tiles_per_warp = stack_tmap_tiles_per_warp();
            const int64_t run = (int64_t)(TPB / 32) * at.tiles_per_warp;
            const int64_t grid = (ntiles + run - 1) / run;
            stack_meanclip_tmap_kernel<NB, NLO, SYM><<<(unsigned)grid, TPB, smem, st>>>(tmap, fp, at);
            APGPU_LAUNCH_CHECK("stack_meanclip_tmap_kernel");
            rest.pix0 = a.pix0 + ntiles * WT;            // the < 32-pixel tail goes through the direct kernel
            rest.npix = a.npix - ntiles * WT;
        } else {
            stack_note_staging(0);
        }
    }
    if (staging == 2) {
        const int64_t ntiles = a.npix / WT;
        if (ntiles > 0) {
            const size_t smem = (size_t)NB * sizeof(float*) + (size_t)(TPB / 32) * NB * WT * sizeof(float);
            APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_cpasync_kernel<NB, NLO, SYM>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            StackArgs ac = a;
            ac.tiles_per_warp = stack_tmap_tiles_per_warp();
            const int64_t run = (int64_t)(TPB / 32) * ac.tiles_per_warp;
            const int64_t grid = (ntiles + run - 1) / run;
            stack_meanclip_cpasync_kernel<NB, NLO, SYM><<<(unsigned)grid, TPB, smem, st>>>(fp, ac);
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
    }
    if (rest.npix > 0) {
        int64_t blocks = (rest.npix + TPB - 1) / TPB;
        stack_meanclip_kernel<NB, NLO, SYM, T><<<(unsigned)blocks, TPB, 0, st>>>(fp, rest);
        APGPU_LAUNCH_CHECK("stack_meanclip_kernel");
    }
    return APGPU_OK;
}

template <int NB, int NLO, typename T>
int launch_meanclip(const T* const* frames, const StackArgs& a_in, cudaStream_t st, int flags) {
    static_assert(NB - NLO <= MEANCLIP_MAX_TAIL, "tail mask too small");
    FramePtrs<NB, T> fp;
    for (int i = 0; i < NB; ++i) fp.p[i] = i < a_in.N ? frames[i] : frames[0];   // padding: loaded, then masked
    StackArgs a = a_in;
    for (int i = NLO; i < NB; ++i) a.tailmask[i - NLO] = i < a.N ? 1.f : 0.f;
    // staging: 0 direct loads, 1 CTA-wide bulk copies, 2 warp-granular cp.async, 3 warp-granular tensor-map TMA.
    // Default (measured, bench.py variants / tools/time_variant.py): the tensor-map TMA pipeline wherever the
    // frames are equally spaced; otherwise cp.async for the medium stacks (85 % at N=64) and direct loads else.
    int staging = 3;
    if (flags & APGPU_STACK_USE_TMA) staging = 1;
    if (flags & APGPU_STACK_DIRECT_LOADS) staging = 0;
    if (flags & APGPU_STACK_USE_CPASYNC) staging = 2;
    if (flags & APGPU_STACK_USE_TENSORMAP) staging = 3;
    if (sizeof(T) == 2 && staging != 3) staging = 0;     // uint16 frames: tensor-map TMA or direct loads
    // (a TMA box must start on a 16-byte boundary: found by tests/test_gpu_stack.py::test_meanclip_row_band_on_cube)
    if (staging == 3 && (!stack_is_cube(frames, a.N, a.pix0 + a.npix) || (a.pix0 * (int64_t)sizeof(T)) % 16 != 0))
        staging = (sizeof(T) == 2 || (flags & APGPU_STACK_USE_TENSORMAP)) ? 0 : ((NB > 32 && NB <= 80) ? 2 : 0);   // measured (time_variant.py)
    // the per-warp stages must leave room for meanclip_min_blocks CTAs per SM
    const size_t smem_cta = (size_t)NB * sizeof(float*) + (size_t)(TPB / 32) * NB * WT * sizeof(float);
    if ((staging == 2 || staging == 3) && smem_cta * meanclip_min_blocks(NB) > (size_t)SMEM_MAX_BYTES) staging = 0;
    // asynchronous copies need 16-byte aligned sources: frame base + first pixel of the band
    if (staging == 1 || staging == 2)
        for (int i = 0; i < a.N; ++i)
            if (!apgpu_aligned(frames[i] + a.pix0, 16)) staging = 0;
    stack_note_staging(staging);
    if ((float)a.klo == (float)a.khi) return launch_meanclip_sym<NB, NLO, true, T>(fp, a, staging, st);
    return launch_meanclip_sym<NB, NLO, false, T>(fp, a, staging, st);
}

#define MC_CASE(NB_, NLO_) if (nb == NB_) return launch_meanclip<NB_, NLO_, MC_T>(frames, a, st, flags);

}  // namespace apgpu_stack

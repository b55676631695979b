// Frame-stack reducer: lane-cooperative median, median + uncertainty and median/MAD clip for long stacks
// (200 < N <= 512) on equally spaced frames.
// See stack_common.cuh / stack_sorted.cuh / stack_meanclip_coop.cuh.
//
// One thread cannot hold more than ~200 samples, and round 1 sent every median beyond 200 frames to the generic
// float64 kernel (local-memory Shell sort: 1-2 % of the HBM roofline).  Here P = 4 or 8 lanes share a pixel:
//   1. the 32-pixel x N-frame tile arrives by tensor-map TMA with the 128-byte swizzle, exactly as in
//      stack_meanclip_coop.cuh (P warps share a tile, 128-byte rows, conflict-free column reads);
//   2. every lane sorts ITS NBL = ceil(N / P) samples in registers with the Batcher network (+inf beyond N) and
//      writes the sorted run back over the samples it loaded (each stage element belongs to exactly one lane);
//   3. the P lanes of a pixel find the element of rank k1 = (N-1)/2 of the union of their sorted runs: every
//      lane keeps a window [lo, hi) of its run that may still hold it; per round the lane with the widest
//      window proposes its middle element v (broadcast by shuffle), every lane counts by binary search in
//      shared memory how many elements of its window are < v and <= v, the group sums the counts (butterfly of
//      shuffles) and all windows shrink to one side of v -- or v is the answer.  Comparison-only, exact, 3-4
//      rounds with the interpolated proposal; for even N the successor (rank k1 + 1) is the smallest element above the answer over all runs.
//   4. (MEDUNC / MEDMAD1) the MAD is a second selection of the same kind over the 2P deviation lists the sorted
//      runs split into at the median, in float64 like the oracle; the clip bounds then cut every run by binary
//      search and the kept ranges are summed per lane and across the lanes.
// The median is therefore bit-exact like the single-thread network, rejection counts are the oracle's, clipped
// means within a few ulp (lane-partial float64 sums).  Pixels holding NaN / inf are marked for the generic routine (stack_marked_kernel), which
// owns the reference's non-finite semantics.
#pragma once
#include "stack_meanclip_coop.cuh"
#include "stack_sorted.cuh"

namespace apgpu_stack {

// the network without the CTA barriers of sort_regs (the groups of a CTA do not run in lock step here, and a
// run of <= 64 samples is ~1450 instructions: it fits the instruction cache)
#define SY_NONE()
template <int NB> __device__ __forceinline__ void sort_regs_nosync(float (&x)[NB], int one_, int mone_);
#define APGPU_DEF_SORT_NS(n) \
    template <> __device__ __forceinline__ void sort_regs_nosync<n>(float (&x)[n], int one_, int mone_) { APGPU_SORTNET_##n(CE_X, CE_Y, SY_NONE) }
APGPU_DEF_SORT_NS(40) APGPU_DEF_SORT_NS(48) APGPU_DEF_SORT_NS(56) APGPU_DEF_SORT_NS(64)

__host__ __device__ constexpr int medcoop_min_blocks(int NBL, int P) {
    // registers: NBL samples + ~40 (64 K per SM); shared memory: NBL * P * 128 B per CTA (one tile group each)
    return P == 8 ? (NBL <= 40 ? 3 : 2) : 4;
}

// double-precision group reductions (the float ones live in LaneGroup)
template <int P> __device__ __forceinline__ double group_min(double v) {
#pragma unroll
    for (int o = 32 / P; o < 32; o <<= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}

// first index in [lo, hi) at which pred(i) is false (pred is true on a prefix)
template <typename Pred> __device__ __forceinline__ int partition_point(int lo, int hi, Pred pred) {
    while (lo < hi) { const int m = (lo + hi) >> 1; if (pred(m)) lo = m + 1; else hi = m; }
    return lo;
}

// MODE_MED: plain median.  MODE_MEDUNC: median + 1.4826 * MAD / sqrt(N).  MODE_MEDMAD1: one median/MAD clip pass,
// then the mean of the kept samples (the reference's ApMasterCal setting).
template <int NBL, int P, int MODE>
__global__ void __launch_bounds__(coop_tpb(P), medcoop_min_blocks(NBL, P))
stack_median_coop_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CubeFrames cube,
                         const __grid_constant__ StackArgs a) {
    constexpr int PIXW = 32 / P;                        // pixels per warp
    constexpr int NB = NBL * P;                         // stage rows (frames)
    constexpr int G = (coop_tpb(P) / 32) / P;           // tile groups per CTA
    constexpr int RB = 8 / P;
    constexpr unsigned FULL = 0xffffffffu;
    using LG = LaneGroup<P>;
    using IMap = SwizzledFrames<P>;
    static_assert(NB % 8 == 0 && NBL % RB == 0, "swizzled stage: whole 8-row atoms");
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* const smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp / P, wi = warp % P;
    const int q = lane % PIXW, r = lane / PIXW;
    const int col = wi * PIXW + q;                      // pixel column of the 32-pixel tile
    const int c = col >> 2, cw = col & 3;               // its 16-byte chunk and word within the chunk
    unsigned char* const stage = smem_raw + (size_t)grp * NB * 128;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)G * NB * 128) + grp;
    int* const cnt = reinterpret_cast<int*>(smem_raw + (size_t)G * NB * 128 + G * sizeof(uint64_t)) + grp;
    if (wi == 0 && lane == 0) { mbar_init(full, 1); *cnt = 0; }
    __syncthreads();
    // element i of this lane's run sits in stage row 8 * (i / RB) + RB * r + i % RB, column col (swizzled):
    // byte offset (i / RB) * 1024 + off[i % RB] with the RB <= 2 offsets of this lane precomputed
    static_assert(RB <= 2, "P = 4 or 8 lanes per pixel");
    const int row80 = RB * r, row81 = RB * r + (RB - 1);
    const unsigned char* const base0 = stage + row80 * 128 + ((c ^ row80) & 7) * 16 + cw * 4;
    const unsigned char* const base1 = stage + row81 * 128 + ((c ^ row81) & 7) * 16 + cw * 4;
    auto elem = [&](int i) -> const float* {
        if (RB == 1) return reinterpret_cast<const float*>(base0 + (size_t)i * 1024);
        return reinterpret_cast<const float*>(((i & 1) ? base1 : base0) + (size_t)(i >> 1) * 1024);
    };
    const int N = a.N;
    const int pix0 = (int)a.pix0;
    const int ntiles = (int)(a.npix / 32);              // full tiles (the host finishes the tail)
    const int run = G * a.tiles_per_warp;
    int tile = (int)blockIdx.x * run + grp;
    const int tile_end = min((int)(blockIdx.x + 1) * run, ntiles);
    uint32_t parity = 0;
    auto issue = [&](int t) {
        mbar_expect_tx(full, (uint32_t)NB * 128);       // out-of-bounds rows (frames >= N) are zero-filled and counted
        const uint64_t policy = l2_evict_first_policy();
        const int box_rows = a.box_rows;
        for (int k = 0; k < a.nchunks; ++k)
            tma_load_2d(stage + (size_t)k * box_rows * 128, &tmap, pix0 + t * 32, k * box_rows, full, policy);
    };
    // the stage doubles as the parking area, so the NEXT tile's copy can only start when this one is finished:
    // its lines are pulled into L2 meanwhile (TMA prefetch), which shortens the exposed part of that copy
    auto prefetch_l2 = [&](int t) {
        const int box_rows = a.box_rows;
        for (int k = 0; k < a.nchunks; ++k)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                         ::"l"(&tmap), "r"(pix0 + t * 32), "r"(k * box_rows) : "memory");
    };
    if (tile < tile_end && wi == 0 && lane == 0) issue(tile);
    // real samples of this lane's run (frames beyond N are padding)
    int nreal = 0;
#pragma unroll
    for (int j = 0; j < NBL; ++j) nreal += (IMap::idx(j, r) < N) ? 1 : 0;
    const int k1 = (N - 1) >> 1;
    for (; tile < tile_end; tile += G) {
        while (!mbar_try_wait(full, parity)) {}
        parity ^= 1u;
        if (wi == 0 && lane == 0 && tile + G < tile_end) prefetch_l2(tile + G);
        float x[NBL];
        float z = 0.f;
#pragma unroll
        for (int j = 0; j < NBL; ++j) {
            const int row8 = RB * r + (j % RB);
            const float v = *reinterpret_cast<const float*>(stage + (size_t)(j / RB) * 1024 + row8 * 128 +
                                                            ((c ^ row8) & 7) * 16 + cw * 4);
            const bool real = IMap::idx(j, r) < N;
            z = fmaf(real ? v : 0.f, 0.f, z);                            // NaN iff a real sample is NaN / inf
            x[j] = real ? v : INFINITY;
        }
        sort_regs_nosync<NBL>(x, a.one, a.minus_one);
        // the sorted run goes back over this lane's own stage elements
#pragma unroll
        for (int j = 0; j < NBL; ++j) {
            const int row8 = RB * r + (j % RB);
            *reinterpret_cast<float*>(stage + (size_t)(j / RB) * 1024 + row8 * 128 + ((c ^ row8) & 7) * 16 + cw * 4) = x[j];
        }
        __syncwarp();
        const bool nonfinite = LG::any(z != z, FULL);
        // ---- rank-k1 element of the union of the P sorted runs of this pixel ----
        int lo = 0, hi = nreal, kk = k1;
        bool found = nonfinite;
        float ans = 0.f;
        for (int round = 0; round <= NB; ++round) {                      // (every round removes at least its pivot)
            if (!__any_sync(FULL, !found)) break;
            const int size = found ? 0 : hi - lo;
            // the lane of this pixel with the widest window proposes an element: the one whose position in that
            // window matches the wanted rank's position in the union of the windows (the runs are samples of one
            // distribution, so this interpolation lands within a few elements of the answer: 3-4 rounds)
            int key = (size << 4) | (15 - r);
#pragma unroll
            for (int o = PIXW; o < 32; o <<= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
            const int rstar = 15 - (key & 15);
            const int total = LG::sum(size, FULL);
            float v = 0.f;
            if (!found && r == rstar) {
                int pos = lo + ((2 * kk + 1) * size) / (2 * max(total, 1));
                pos = min(max(pos, lo), hi - 1);
                v = *elem(pos);
            }
            v = __shfl_sync(FULL, v, rstar * PIXW + q);
            // elements of my window below v / not above v (binary search in the parked run; ties are rare)
            int lt = lo, le = lo;
            if (!found) {
                int b0 = lo, b1 = hi;
                while (b0 < b1) { const int m = (b0 + b1) >> 1; if (*elem(m) < v) b0 = m + 1; else b1 = m; }
                lt = b0;
                while (b0 < hi && *elem(b0) <= v) ++b0;                  // ties: the proposing lane has one, others rarely
                le = b0;
            }
            const int t_lt = LG::sum(lt - lo, FULL), t_le = LG::sum(le - lo, FULL);
            if (!found) {
                if (kk < t_lt) hi = lt;                                  // the answer is below v
                else if (kk < t_le) { found = true; ans = v; }           // v itself has rank kk
                else { lo = le; kk -= t_le; }                            // the answer is above v
            }
        }
        // ---- even N: the element of rank k1 + 1 ----
        float second = ans;
        if (!(N & 1)) {
            int b0 = 0, b1 = nreal;
            if (!nonfinite) { while (b0 < b1) { const int m = (b0 + b1) >> 1; if (*elem(m) <= ans) b0 = m + 1; else b1 = m; } }
            const int n_le = LG::sum(b0, FULL);                          // elements <= ans over all runs
            const float nxt = LG::min((!nonfinite && b0 < nreal) ? *elem(b0) : INFINITY, FULL);
            second = (n_le > k1 + 1) ? ans : nxt;
        }
        const double med = (N & 1) ? (double)ans : __dmul_rn(__dadd_rn((double)ans, (double)second), 0.5);
        double out_val = med, out_unc = (double)NAN;
        int out_nrej = 0;
        if constexpr (MODE != MODE_MED) {
            // ---- MAD: the rank-k1 (and k1+1) smallest of |x - med| over the union of the runs.  Every run splits
            // at the median into two lists with ascending deviations: L[j] = med - run[s-1-j], R[j] = run[s+j] - med
            // (float64, rounded like the oracle's np.abs(x - med)); the same windowed selection runs over 2P lists.
            const int s0 = nonfinite ? 0 : partition_point(0, nreal, [&](int i) { return (double)*elem(i) < med; });
            const int nL = s0, nR = nonfinite ? 0 : nreal - s0;
            auto devL = [&](int j) { return fabs(__dsub_rn((double)*elem(s0 - 1 - j), med)); };
            auto devR = [&](int j) { return fabs(__dsub_rn((double)*elem(s0 + j), med)); };
            int loL = 0, hiL = nL, loR = 0, hiR = nR;
            kk = k1;
            bool fnd = nonfinite;
            double d1 = 0.0;
            for (int round = 0; round <= NB; ++round) {
                if (!__any_sync(FULL, !fnd)) break;
                const int szL = fnd ? 0 : hiL - loL, szR = fnd ? 0 : hiR - loR;
                const int side = szR > szL ? 1 : 0;
                int key = ((side ? szR : szL) << 5) | (side << 4) | (15 - r);
#pragma unroll
                for (int o = PIXW; o < 32; o <<= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
                const int rstar = 15 - (key & 15), sstar = (key >> 4) & 1;
                const int total = LG::sum(szL + szR, FULL);
                double v = 0.0;
                if (!fnd && r == rstar) {
                    const int wl = sstar ? loR : loL, wh = sstar ? hiR : hiL;
                    int pos = wl + ((2 * kk + 1) * (wh - wl)) / (2 * max(total, 1));
                    pos = min(max(pos, wl), wh - 1);
                    v = sstar ? devR(pos) : devL(pos);
                }
                v = __shfl_sync(FULL, v, rstar * PIXW + q);
                int ltL = loL, leL = loL, ltR = loR, leR = loR;
                if (!fnd) {
                    ltL = partition_point(loL, hiL, [&](int j) { return devL(j) < v; });
                    leL = ltL;
                    while (leL < hiL && devL(leL) <= v) ++leL;
                    ltR = partition_point(loR, hiR, [&](int j) { return devR(j) < v; });
                    leR = ltR;
                    while (leR < hiR && devR(leR) <= v) ++leR;
                }
                const int t_lt = LG::sum((ltL - loL) + (ltR - loR), FULL);
                const int t_le = LG::sum((leL - loL) + (leR - loR), FULL);
                if (!fnd) {
                    if (kk < t_lt) { hiL = ltL; hiR = ltR; }
                    else if (kk < t_le) { fnd = true; d1 = v; }
                    else { loL = leL; loR = leR; kk -= t_le; }
                }
            }
            double d2 = d1;
            if (!(N & 1)) {
                const int ubL = nonfinite ? 0 : partition_point(0, nL, [&](int j) { return devL(j) <= d1; });
                const int ubR = nonfinite ? 0 : partition_point(0, nR, [&](int j) { return devR(j) <= d1; });
                const int n_le = LG::sum(ubL + ubR, FULL);
                double nx = (double)INFINITY;
                if (!nonfinite && ubL < nL) nx = devL(ubL);
                if (!nonfinite && ubR < nR) { const double t = devR(ubR); nx = t < nx ? t : nx; }
                nx = group_min<P>(nx);
                d2 = (n_le > k1 + 1) ? d1 : nx;
            }
            const double mad = (N & 1) ? d1 : __dmul_rn(__dadd_rn(d1, d2), 0.5);
            const double sd = __dmul_rn(MAD_TO_STD, mad);
            if constexpr (MODE == MODE_MEDUNC) {
                out_unc = __ddiv_rn(sd, __dsqrt_rn((double)N));        // Combiner.median_combine: mad_std / sqrt(n)
            } else {
                // ---- one clip pass, then the mean of the kept range of every run (float64 partial sums per lane,
                // butterfly across the lanes: within a few ulp of the oracle's frame-order sum)
                const double lo_b = __dsub_rn(med, __dmul_rn(sd, a.klo));
                const double hi_b = __dadd_rn(med, __dmul_rn(sd, a.khi));
                const int sa = nonfinite ? 0 : partition_point(0, nreal, [&](int i) { return (double)*elem(i) < lo_b; });
                const int sb = nonfinite ? 0 : partition_point(sa, nreal, [&](int i) { return !((double)*elem(i) > hi_b); });
                const int nk = LG::sum(sb - sa, FULL);
                double acc = 0.0;
                for (int i = sa; i < sb; ++i) acc = __dadd_rn(acc, (double)*elem(i));
                acc = LG::sum(acc, FULL);
                const double mean = __ddiv_rn(acc, (double)(nk > 0 ? nk : 1));   // nk >= 1: the median itself survives
                out_val = mean;
                out_nrej = N - nk;
                if (a.uncert) {
                    double qq = 0.0;
                    for (int i = sa; i < sb; ++i) {
                        const double d = __dsub_rn((double)*elem(i), mean);
                        qq = __dadd_rn(qq, __dmul_rn(d, d));
                    }
                    qq = LG::sum(qq, FULL);
                    out_unc = __ddiv_rn(__dsqrt_rn(__ddiv_rn(qq, (double)(nk > 0 ? nk : 1))), __dsqrt_rn((double)(nk > 0 ? nk : 1)));
                }
            }
        }
        // this warp is done with the stage: the last warp of the group re-arms it for the group's next tile
        const int next = tile + G;
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int old = atomicAdd(cnt, 1);
            if (old == P - 1) {
                *cnt = 0;
                if (next < tile_end) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(next);
                }
            }
        }
        __syncwarp();
        if (r == 0) {
            const int64_t p = (int64_t)(uint32_t)(pix0 + tile * 32 + col);
            if (nonfinite) mark_pixel(a, p);
            else write_pixel(a, p, out_val, out_nrej, out_unc, 0);
        }
    }
}

template <int NBL, int P, int MODE>
int launch_median_coop(const float* const* frames, const StackArgs& a_in, cudaStream_t st, int64_t* done_pix) {
    constexpr int NB = NBL * P;
    constexpr int G = (coop_tpb(P) / 32) / P;
    StackArgs a = a_in;
    const int64_t ntiles = a.npix / 32;
    *done_pix = 0;
    if (ntiles == 0) return APGPU_OK;
    if (a.pix0 % 4 != 0) return APGPU_ERR_UNSUPPORTED;        // a TMA box must start on a 16-byte boundary
    a.box_rows = 8;
    for (int d = 8; d <= 256; d += 8)
        if (NB % d == 0) a.box_rows = d;
    a.nchunks = NB / a.box_rows;
    a.tiles_per_warp = stack_tmap_tiles_per_warp();
    const int64_t stride = (const char*)frames[1] - (const char*)frames[0];
    CUtensorMap tmap;
    if (!encode_stack_tensor_map(&tmap, frames[0], (uint64_t)(a.pix0 + a.npix), a.N, (uint64_t)stride, 32, a.box_rows,
                                 /*swizzle128=*/true))
        return APGPU_ERR_UNSUPPORTED;
    CubeFrames cube{(const char*)frames[0], stride};
    const size_t smem = (size_t)G * NB * 128 + G * (sizeof(uint64_t) + sizeof(int)) + 1024;   // + alignment slack
    APGPU_CUDA(cudaFuncSetAttribute(stack_median_coop_kernel<NBL, P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t run = (int64_t)G * a.tiles_per_warp;
    const int64_t grid = (ntiles + run - 1) / run;
    stack_median_coop_kernel<NBL, P, MODE><<<(unsigned)grid, coop_tpb(P), smem, st>>>(tmap, cube, a);
    APGPU_LAUNCH_CHECK("stack_median_coop_kernel");
    *done_pix = ntiles * 32;
    stack_note_staging(5);
    return APGPU_OK;
}

#define MEDCOOP_CASE(NBL_, NLO_, P_) \
    if (a.N > NLO_ && a.N <= NBL_ * P_) return launch_median_coop<NBL_, P_, MEDCOOP_MODE>(frames, a, st, done_pix);

}  // namespace apgpu_stack

// Frame-stack reducer: warp-cooperative kappa-sigma kernels for long stacks (128 < N <= 512) on
// equally spaced frames.  See stack_common.cuh / stack_meanclip.cuh.
//
// The register-resident meanclip kernel keeps all N samples of a pixel in one thread, which stops
// scaling at N ~ 100-128 (at N = 200 it needs 255 registers and runs at a quarter of the memory
// roofline).  Here P = 2 or 4 lanes share a pixel, each holding N/P samples, and every sum /
// extremum of the clipping iteration is completed by a log2(P)-step butterfly of warp shuffles.
//
// The tile keeps the shape that the memory system wants: 32 pixels x N frames, i.e. 128-byte
// rows (a B200 SM sustains a fixed number of memory requests in flight, so 64-byte rows halve the
// bandwidth whatever issues them -- measured with TMA boxes and with cp.async alike).  One tile is
// therefore shared by P WARPS: warp w of the group owns pixels [w*32/P, (w+1)*32/P) of it, lane
// (q, r) of that warp the frames i = 8*(j / RB) + RB*r + j % RB (RB = 8/P) of pixel q.  The tile is
// filled by tensor-map TMA copies with the 128-byte swizzle (the 16-byte chunk index is XORed
// with row % 8): P lanes reading the same pixel column from rows that differ in r then hit different
// banks, so the column reads stay conflict-free.  Pipeline per group: a "full" mbarrier completed
// by the TMA bytes; the last warp of the group to have consumed its samples (shared-memory counter)
// issues the copy for the group's next tile.  No CTA-wide barrier in the steady state.
#pragma once
#include "stack_meanclip.cuh"

namespace apgpu_stack {

// threads per CTA: at least the P warps of one tile group
__host__ __device__ constexpr int coop_tpb(int P) { return 32 * (P > 4 ? P : 4); }
// CTAs per SM: shorter per-lane sample arrays need fewer registers, so more warps stay resident
__host__ __device__ constexpr int coop_min_blocks(int NBL, int P) {
    return P == 8 ? (NBL <= 50 ? 3 : 2) : (NBL <= 40 ? 6 : (NBL <= 50 ? 5 : (NBL <= 100 ? 4 : 3))) * 128 / coop_tpb(P);
}

template <int NBL, int NLO, int P, bool SYM>
__global__ void __launch_bounds__(coop_tpb(P), coop_min_blocks(NBL, P))
stack_meanclip_coop_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CubeFrames cube,
                           const __grid_constant__ StackArgs a) {
    constexpr int PIXW = 32 / P;                        // pixels per warp
    constexpr int NB = NBL * P;                         // stage rows (frames)
    constexpr int G = (coop_tpb(P) / 32) / P;           // tile groups per CTA
    constexpr int RB = 8 / P;
    static_assert(NB % 8 == 0, "swizzled stage: whole 8-row atoms");
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // the 128-byte swizzle pattern is a function of the shared-memory ADDRESS: 1024-byte aligned stages
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
    // byte address of (row 8*J + RB*r + jl, column col) = J*1024 + lane_base[jl]
    const float* lane_base[RB];
#pragma unroll
    for (int jl = 0; jl < RB; ++jl) {
        const int row8 = RB * r + jl;
        lane_base[jl] = reinterpret_cast<const float*>(stage + row8 * 128 + ((c ^ row8) & 7) * 16 + cw * 4);
    }
    // (all pixel / tile indices fit 32 bits: the host only takes this path below 2^31 pixels)
    const int pix0 = (int)a.pix0;
    const int ntiles = (int)(a.npix / 32);              // full tiles (the host finishes the tail)
    const int run = G * a.tiles_per_warp;
    int tile = (int)blockIdx.x * run + grp;
    const int tile_end = min((int)(blockIdx.x + 1) * run, ntiles);
    uint32_t parity = 0;
    auto issue = [&](int t) {
        // out-of-bounds rows (frames >= N) are zero-filled and count towards the mbarrier's byte total
        mbar_expect_tx(full, (uint32_t)NB * 128);
        const uint64_t policy = l2_evict_first_policy();
        const int box_rows = a.box_rows;
        for (int k = 0; k < a.nchunks; ++k)
            tma_load_2d(stage + (size_t)k * box_rows * 128, &tmap, pix0 + t * 32, k * box_rows, full, policy);
    };
    if (tile < tile_end && wi == 0 && lane == 0) issue(tile);
    for (; tile < tile_end; tile += G) {
        while (!mbar_try_wait(full, parity)) {}
        parity ^= 1u;
        float2 y[NBL / 2];
#pragma unroll
        for (int j = 0; j < NBL / 2; ++j) {
            y[j].x = lane_base[(2 * j) % RB][((2 * j) / RB) * 256];
            y[j].y = lane_base[(2 * j + 1) % RB][((2 * j + 1) / RB) * 256];
        }
        // pivot of the pixel: median of its first three frames (rows 0, 1, 2 of column col)
        const float* s0 = reinterpret_cast<const float*>(stage + cw * 4);
        const float pivot = med3(s0[(c ^ 0) * 4], s0[32 + (c ^ 1) * 4], s0[64 + (c ^ 2) * 4]);
        const int next = tile + G;
        // the last warp of the group to have consumed its samples re-arms the stage (the sums depend on
        // every staged sample of this warp's instruction stream: the predicate carries that dependence)
        auto rearm = [&](float s2) {
            __syncwarp();
            if (lane == 0 && s2 != -1.f) {
                __threadfence_block();
                const int old = atomicAdd(cnt, 1);
                if (old == P - 1) {
                    *cnt = 0;
                    if (next < tile_end) issue(next);
                }
            }
            __syncwarp();
        };
        meanclip_pixel<NBL, NLO, SYM, decltype(rearm), P, CubeFrames, SwizzledFrames<P>>(
            y, cube, a, (int64_t)(uint32_t)(pix0 + tile * 32 + col), rearm, pivot, 0xffffffffu, r);
    }
}

template <int NBL, int NLO, int P, bool SYM>
int launch_meanclip_coop_sym(const float* const* frames, const StackArgs& a_in, cudaStream_t st, int64_t* done_pix) {
    constexpr int NB = NBL * P;
    constexpr int G = (coop_tpb(P) / 32) / P;
    StackArgs a = a_in;
    const int64_t ntiles = a.npix / 32;
    *done_pix = 0;
    if (ntiles == 0) return APGPU_OK;
    if (a.pix0 % 4 != 0) return APGPU_ERR_UNSUPPORTED;        // a TMA box must start on a 16-byte boundary
    // One tile = the fewest TMA boxes (<= 256 rows each, whole 8-row swizzle atoms) on one mbarrier;
    // smaller boxes were measured (APGPU_COOP_BOX_ROWS): no gain.
    a.box_rows = 8;
    for (int d = 8; d <= stack_coop_box_rows_max(); d += 8)
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
    APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_coop_kernel<NBL, NLO, P, SYM>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t run = (int64_t)G * a.tiles_per_warp;
    const int64_t grid = (ntiles + run - 1) / run;
    stack_meanclip_coop_kernel<NBL, NLO, P, SYM><<<(unsigned)grid, coop_tpb(P), smem, st>>>(tmap, cube, a);
    APGPU_LAUNCH_CHECK("stack_meanclip_coop_kernel");
    *done_pix = ntiles * 32;
    stack_note_staging(5);
    return APGPU_OK;
}

template <int NBL, int NLO, int P>
int launch_meanclip_coop(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    if ((float)a.klo == (float)a.khi) return launch_meanclip_coop_sym<NBL, NLO, P, true>(frames, a, st, done_pix);
    return launch_meanclip_coop_sym<NBL, NLO, P, false>(frames, a, st, done_pix);
}

#define COOP_CASE(NBL_, NLO_, P_) \
    if (a.N > NLO_ && a.N <= NBL_ * P_) return launch_meanclip_coop<NBL_, NLO_, P_>(frames, a, st, done_pix);

}  // namespace apgpu_stack

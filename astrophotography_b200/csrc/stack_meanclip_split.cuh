// Frame-stack reducer: lane-split kappa-sigma kernels for very long stacks (512 < N <= 1024) on equally
// spaced frames (shorter stacks use the warp-cooperative kernels of stack_meanclip_coop.cuh, which keep
// 128-byte rows and are 2-3x faster; this simpler variant covers what they do not).  See stack_common.cuh / stack_meanclip.cuh.
//
// The register-resident meanclip kernel keeps all N samples of a pixel in one thread, which stops
// scaling at N ~ 100 (128 registers, 4 warps per scheduler; at N = 200 it needs 255 registers and
// runs at a quarter of the memory roofline).  Here P = 2, 4 or 8 lanes of a warp share one pixel:
// lane = r * (32/P) + q holds the samples i = P*j + r (j < NBL) of pixel q of a (32/P)-pixel warp
// tile, so a thread never holds more than ~100 samples whatever N is, and every sum / extremum of
// the clipping iteration is completed by a log2(P)-step butterfly of warp shuffles.
// Staging: every warp is its own pipeline (private [N][32/P] stage, no CTA barrier).  The rows of a
// tile are only 128/P bytes long, and the TMA unit moves one box row per ~6.4 cycles per SM whatever
// its length (measured: 64-byte rows 2.9 TB/s, 32-byte rows 1.4 TB/s), so the tile is filled with
// 16-byte cp.async (LDGSTS) copies instead: one warp instruction copies the rows of 4*P frames, the
// source pointer advances by a constant stride.  Viewed as [NBL][32] floats the stage is read with
// exactly the conflict-free column pattern of the one-lane-per-pixel kernel (sample j of lane
// `lane` sits at stage[32*j + lane]).  Rows beyond N are zeroed once and masked to y = 0.
#pragma once
#include "stack_meanclip.cuh"

namespace apgpu_stack {

template <int NBL, int NLO, int P, bool SYM>
__global__ void __launch_bounds__(TPB, meanclip_min_blocks(NBL))
stack_meanclip_split_kernel(const __grid_constant__ CubeFrames cube, const __grid_constant__ StackArgs a) {
    constexpr int PIXW = 32 / P;                        // pixels per warp tile
    constexpr int SROWS = NBL * P;                      // stage rows
    constexpr int CPR = PIXW / 4;                       // 16-byte chunks per row
    constexpr int RPI = 32 / CPR;                       // rows copied by one warp instruction (= 4 P)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane % PIXW, r = lane / PIXW;
    unsigned gmask = 0;
#pragma unroll
    for (int k = 0; k < P; ++k) gmask |= 1u << (q + k * PIXW);
    float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * SROWS * PIXW;     // [SROWS][PIXW]
    const int N = a.N;
    for (int i = N * PIXW + lane; i < SROWS * PIXW; i += 32) stage[i] = 0.f;   // rows no copy ever writes
    __syncwarp();
    const int64_t ntiles = a.npix / PIXW;               // full warp tiles (the host finishes the tail)
    const int64_t run = (int64_t)(TPB / 32) * a.tiles_per_warp;
    int64_t tile = (int64_t)blockIdx.x * run + warp;
    const int64_t tile_end = (tile - warp + run < ntiles) ? tile - warp + run : ntiles;
    constexpr int64_t nwarps = TPB / 32;
    // this lane copies chunk `ch` of the rows rowoff, rowoff + RPI, ...
    const int ch = lane % CPR, rowoff = lane / CPR;
    const char* const src_lane = cube.base + (int64_t)rowoff * cube.stride + (a.pix0 + ch * 4) * (int64_t)sizeof(float);
    float* const dst_lane = stage + rowoff * PIXW + ch * 4;
    const int64_t src_step = (int64_t)RPI * cube.stride;
    auto issue = [&](int64_t t) {
        const char* s = src_lane + t * (int64_t)(PIXW * sizeof(float));
        float* d = dst_lane;
        int row = rowoff;
#pragma unroll 4
        for (; row < N; row += RPI) {
            cp_async16(d, s);
            s += src_step;
            d += RPI * PIXW;
        }
        cp_async_commit();
    };
    if (tile < tile_end) issue(tile);
    for (; tile < tile_end; tile += nwarps) {
        cp_async_wait_all();
        __syncwarp();                                   // every lane's copies are visible
        float2 y[NBL / 2];
#pragma unroll
        for (int j = 0; j < NBL / 2; ++j) {
            y[j].x = stage[(2 * j) * 32 + lane];
            y[j].y = stage[(2 * j + 1) * 32 + lane];
        }
        // pivot of pixel q: median of its first three frames (broadcast reads)
        const float pivot = med3(stage[q], stage[PIXW + q], stage[2 * PIXW + q]);
        const int64_t next = tile + nwarps;
        // re-arm the stage once the sums (which depend on every staged sample of every lane of
        // this warp instruction stream) exist: the predicate below carries that dependence
        auto rearm = [&](float s2) {
            __syncwarp();
            if (next < tile_end && s2 != -1.f) issue(next);
        };
        meanclip_pixel<NBL, NLO, SYM, decltype(rearm), P, CubeFrames>(y, cube, a, a.pix0 + tile * PIXW + q, rearm,
                                                                        pivot, gmask, r);
    }
}

template <int NBL, int NLO, int P, bool SYM>
int launch_meanclip_split_sym(const float* const* frames, const StackArgs& a_in, cudaStream_t st, int64_t* done_pix) {
    constexpr int PIXW = 32 / P;
    constexpr int SROWS = NBL * P;
    StackArgs a = a_in;
    const int64_t ntiles = a.npix / PIXW;
    *done_pix = 0;
    if (ntiles == 0) return APGPU_OK;
    a.tiles_per_warp = stack_tmap_tiles_per_warp();
    const int64_t stride = (const char*)frames[1] - (const char*)frames[0];
    if (a.pix0 % 4 != 0) return APGPU_ERR_UNSUPPORTED;          // 16-byte copies: the band must start 16-byte aligned
    CubeFrames cube{(const char*)frames[0], stride};
    const size_t smem = (size_t)(TPB / 32) * SROWS * PIXW * sizeof(float);
    APGPU_CUDA(cudaFuncSetAttribute(stack_meanclip_split_kernel<NBL, NLO, P, SYM>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t run = (int64_t)(TPB / 32) * a.tiles_per_warp;
    const int64_t grid = (ntiles + run - 1) / run;
    stack_meanclip_split_kernel<NBL, NLO, P, SYM><<<(unsigned)grid, TPB, smem, st>>>(cube, a);
    APGPU_LAUNCH_CHECK("stack_meanclip_split_kernel");
    *done_pix = ntiles * PIXW;
    stack_note_staging(4);
    return APGPU_OK;
}

template <int NBL, int NLO, int P>
int launch_meanclip_split(const float* const* frames, const StackArgs& a, cudaStream_t st, int64_t* done_pix) {
    if ((float)a.klo == (float)a.khi) return launch_meanclip_split_sym<NBL, NLO, P, true>(frames, a, st, done_pix);
    return launch_meanclip_split_sym<NBL, NLO, P, false>(frames, a, st, done_pix);
}

#define SPLIT_CASE(NBL_, NLO_, P_) \
    if (a.N > NLO_ && a.N <= NBL_ * P_) return launch_meanclip_split<NBL_, NLO_, P_>(frames, a, st, done_pix);

}  // namespace apgpu_stack

// Frame-stack reducer: meanclip_smem (shared-memory-resident kappa-sigma kernel).  See stack_common.cuh.
#include "stack_common.cuh"

namespace apgpu_stack {

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
    if (!(S2 <= FLT_MAX) || !(fabsf(S1) <= FLT_MAX)) { mark_pixel(a, p); return; }

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
    if (uncertain || nk == 0) { mark_pixel(a, p); return; }

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

template <int CAP>
int launch_meanclip_smem_cap(const float* const* frames, const StackArgs& a, cudaStream_t st) {
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

int stack_launch_meanclip_smem(const float* const* frames, const StackArgs& a, cudaStream_t st) {
    return a.N <= 128 ? launch_meanclip_smem_cap<128>(frames, a, st) : launch_meanclip_smem_cap<512>(frames, a, st);
}

}  // namespace apgpu_stack

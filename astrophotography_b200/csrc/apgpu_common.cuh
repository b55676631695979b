// Shared helpers for the apgpu kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "apgpu.h"

void apgpu_set_error(const char* fmt, ...);
void apgpu_count_launch();

#define APGPU_REQUIRE(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            apgpu_set_error(__VA_ARGS__);        \
            return APGPU_ERR_ARG;                \
        }                                        \
    } while (0)

#define APGPU_LAUNCH_CHECK(what)                                                   \
    do {                                                                           \
        apgpu_count_launch();                                                      \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            apgpu_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e__)); \
            return APGPU_ERR_CUDA;                                                 \
        }                                                                          \
    } while (0)

#define APGPU_CUDA(call)                                                           \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            apgpu_set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
            return APGPU_ERR_CUDA;                                                 \
        }                                                                          \
    } while (0)

static inline bool apgpu_aligned(const void* p, size_t a) {
    return (reinterpret_cast<uintptr_t>(p) % a) == 0;
}

// Streaming (read-once / write-once) global accesses: evict-first so the
// 126 MB L2 is not polluted by data that is never re-read.
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

constexpr int APGPU_NUM_SMS = 148;   // B200

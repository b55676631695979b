// Error state, ABI version and launch accounting for libapgpu.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "apgpu_common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void apgpu_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void apgpu_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" int apgpu_abi_version(void) { return APGPU_ABI_VERSION; }
extern "C" const char* apgpu_last_error(void) { return g_err; }
extern "C" uint64_t apgpu_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

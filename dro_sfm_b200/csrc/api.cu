// Version / error plumbing of the drosfm_b200 C ABI (include/drosfm_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace drosfm {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch errors only (no synchronisation): positive cudaError_t on failure.
int launch_status(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return DROSFM_OK;
}

}  // namespace drosfm

extern "C" {

int drosfm_version(void) { return DROSFM_ABI_VERSION; }

const char* drosfm_last_error(void) { return drosfm::g_err; }

unsigned long long drosfm_launch_count(void) { return drosfm::g_launches.load(std::memory_order_relaxed); }

size_t drosfm_ws_bytes(int slots) {
    return static_cast<size_t>(slots < 1 ? 1 : slots) * drosfm::kSub * sizeof(drosfm::Slot);
}

}  // extern "C"

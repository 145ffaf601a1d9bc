// Version / error plumbing of the drosfm_b200 C ABI (include/drosfm_b200.h).
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace drosfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch errors only (no synchronisation): positive cudaError_t on failure.
int launch_status(const char* what) {
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

size_t drosfm_ws_bytes(int slots) {
    return static_cast<size_t>(slots < 1 ? 1 : slots) * sizeof(drosfm::Slot);
}

}  // extern "C"

// api.cu -- error reporting and bookkeeping for the C ABI.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pb200 {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return PB200_ERR_CUDA;
    }
    count_launch();
    return PB200_OK;
}
}  // namespace pb200

extern "C" int pb200_abi_version(void) { return PB200_ABI_VERSION; }
extern "C" const char* pb200_last_error(void) { return pb200::g_err; }
extern "C" int64_t pb200_launch_count(void) { return pb200::g_launches.load(); }

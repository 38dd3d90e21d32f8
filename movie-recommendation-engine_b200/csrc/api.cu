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

// The walk kernel reads one random 32-byte bucket per step; with the default 64-byte L2 fetch
// granularity every miss moves two sectors from DRAM (ncu: dram bytes = 2 x the sectors requested).
extern "C" int pb200_set_l2_fetch_granularity(int bytes) {
    PB_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, "set_l2_fetch_granularity: 32, 64 or 128");
    PB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return PB200_OK;
}
extern "C" int pb200_get_l2_fetch_granularity(void) {
    size_t v = 0;
    if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return -1;
    return (int)v;
}

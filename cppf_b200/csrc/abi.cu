// Library-level entry points of the C ABI (include/cppf_b200.h).
#include "common.cuh"

#include "../../include/cppf_b200.h"

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

namespace cppf {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

thread_local bool t_workspace_prepared = false;

int raise_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int> limit;      // (device, kernel) -> largest limit set so far
    int dev = 0;
    CPPF_RETURN_IF(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int& have = limit[{dev, kernel}];
    if (bytes <= have) return 0;
    CPPF_RETURN_IF(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    have = bytes;
    return 0;
}
}  // namespace cppf

extern "C" int cppf_abi_version(void) { return 2; }
extern "C" const char* cppf_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
extern "C" uint64_t cppf_launch_count(void) { return cppf::g_launches.load(std::memory_order_relaxed); }

// Library-level entry points of the C ABI (include/cppf_b200.h).
#include "common.cuh"

#include "../../include/cppf_b200.h"

#include <atomic>

namespace cppf {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace cppf

extern "C" int cppf_abi_version(void) { return 2; }
extern "C" const char* cppf_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
extern "C" uint64_t cppf_launch_count(void) { return cppf::g_launches.load(std::memory_order_relaxed); }

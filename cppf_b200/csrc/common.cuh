// Shared device helpers for the cppf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cppf {

// process-wide launch counter behind cppf_launch_count() (defined in abi.cu)
void count_launch(int n = 1);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) takes the context lock and costs microseconds, and the object loop
// launches six kernels that need more than 48 KB: the limit of a kernel is raised once per device (and again only if a
// later launch needs more), not before every launch.  Defined in abi.cu.
int raise_dynamic_smem(const void* kernel, int bytes);
// True while cppf_pose_fused is enqueueing an object on this host thread: it clears every accumulator of its workspace
// (vote scratch, argmax keys, sphere counts, statistics) with ONE memset and initialises the global-max scratch in its
// geometry kernel, so the launchers below skip their own small memsets / init kernels (6 driver calls per object).
// Standalone calls through the C ABI never see it set.  Defined in abi.cu.
extern thread_local bool t_workspace_prepared;
struct PreparedWorkspaceScope {
    PreparedWorkspaceScope() { t_workspace_prepared = true; }
    ~PreparedWorkspaceScope() { t_workspace_prepared = false; }
};

#define CPPF_RETURN_IF(err_expr)                         \
    do {                                                 \
        cudaError_t _e = (err_expr);                     \
        if (_e != cudaSuccess) return (int)_e;           \
    } while (0)

#define CPPF_LAUNCH_CHECK()                              \
    do {                                                 \
        cppf::count_launch();                            \
        cudaError_t _e = cudaGetLastError();             \
        if (_e != cudaSuccess) return (int)_e;           \
    } while (0)

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// dot / length / cross / circle offset in the exact operation order nvcc 12.9 gives the reference's helper_math.cuh
// expressions inside models/voting.py (read off the SASS of oracle/_ref/ref_{ppf_voting,backvote,rot_voting}.cubin, which
// agree with each other): dot = fma(z, z, fma(x, x, y * y)); a cross component a.y * b.z - a.z * b.y = fma(a.y, b.z,
// -(a.z * b.y)); cos * x + sin * y = fma(x, cos, y * sin).  Written with explicit intrinsics so that no other contraction
// can be chosen here: decisions taken on these values (in / out of the grid, within tol of the centre) are bit-exact
// with the reference kernels.
__device__ __forceinline__ float dot3(f3 a, f3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y))); }
__device__ __forceinline__ float len3(f3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return {__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
            __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x))};
}
// c = a - ab * proj_len  (models/voting.py:23,89): one FFMA in the reference's SASS
__device__ __forceinline__ f3 foot_point(f3 a, f3 ab, float mu) {
    return {__fmaf_rn(mu, -ab.x, a.x), __fmaf_rn(mu, -ab.y, a.y), __fmaf_rn(mu, -ab.z, a.z)};
}
// x = co / (length(co) + 1e-7) * odist  (:28,94): the division is inside pair_frame, this is the single multiply
__device__ __forceinline__ f3 scale3(f3 e, float s) { return {__fmul_rn(e.x, s), __fmul_rn(e.y, s), __fmul_rn(e.z, s)}; }
// offset = cos(angle) * x + sin(angle) * y  (models/voting.py:34,100,141)
__device__ __forceinline__ f3 circle_offset(f3 x, f3 y, float c, float s) {
    return {__fmaf_rn(x.x, c, __fmul_rn(y.x, s)), __fmaf_rn(x.y, c, __fmul_rn(y.y, s)), __fmaf_rn(x.z, c, __fmul_rn(y.z, s))};
}
__device__ __forceinline__ f3 ld3(const float* __restrict__ p, int64_t i) {
    return {__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}

// Vote-grid geometry of one object when it is derived on the device (cppf_pose_fused): the kernels of the
// fused path read it from here instead of from their launch parameters, so the host never waits for it.
struct Geom {
    float corner[3];         // nocs/inference.py:194: pc.min(0)
    int gx, gy, gz, cells;   // :195: int((max - min) / res) + 1
    float hx, hy, hz;        // exact upper bounds of models/voting.py:36-39: float_ceil(dim - 1.01)
    float dhx, dhy, dhz;     // conservative upper bounds on (candidate - corner), before the division by res
    float bx, by, bz;        // float(dim - 1): bounds of the back-vote test (models/voting.py:104-107)
    int status;              // 0 ok, 1 = the grid exceeds the capacity the caller provided
    int mode;                // 0: one shared-memory grid per CTA (vote_private), 1: routed x-slabs (vote_routed)
    int planes_per_slab, n_slabs;   // mode 1
};

// (a, b) of pair p: from an int32/int64 index list, or row-major dense enumeration.  row0: dense enumeration of a ROW
// BLOCK of the pair matrix (one object split over several GPUs): pair p is (row0 + p / N, p % N).
template <bool IDX64>
__device__ __forceinline__ void pair_ab(const void* __restrict__ idx, int64_t p, int n_points, int& a, int& b, int row0 = 0) {
    if (idx == nullptr) {
        if (((uint64_t)p >> 32) == 0) {                 // N < 65 536: a 32-bit divide is a third of the 64-bit one
            const uint32_t q = (uint32_t)p / (uint32_t)n_points;
            a = (int)q + row0;
            b = (int)((uint32_t)p - q * (uint32_t)n_points);
        } else {
            const int64_t q = p / n_points;
            a = (int)q + row0;
            b = (int)(p - q * n_points);
        }
    } else if (IDX64) {
        const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(idx) + p);
        a = (int)v.x;
        b = (int)v.y;
    } else {
        const int2 v = __ldg(reinterpret_cast<const int2*>(idx) + p);
        a = v.x;
        b = v.y;
    }
}

// The frame every voting kernel of the reference builds first
// (models/voting.py:18-29, 84-94, 127-137): unit ab and the unit in-plane axis ex.
// The double-precision islands of the CUDA-C strings (`1e-7` literals) are kept:
// the degenerate test is a double compare, the normaliser a double sum rounded to float.
__device__ __forceinline__ bool pair_frame(f3 a, f3 b, f3& ab, f3& ex) {
    ab = {__fadd_rn(a.x, -b.x), __fadd_rn(a.y, -b.y), __fadd_rn(a.z, -b.z)};
    const float len = len3(ab);
    if ((double)len < 1e-7) return false;
    const float den = (float)((double)len + 1e-7);
    ab = {__fdiv_rn(ab.x, den), __fdiv_rn(ab.y, den), __fdiv_rn(ab.z, den)};
    // co = (0, -ab.z, ab.y); the reference evaluates length(co) twice and nvcc compiles the two differently: the
    // degenerate test on sqrt(z*z + y*y) (two rounded products, one add), the normaliser on sqrt(fma(y, y, z*z))
    const float yy = __fmul_rn(ab.y, ab.y);
    const float lt = sqrtf(__fadd_rn(__fmul_rn(ab.z, ab.z), yy));
    f3 co;
    float lc2;
    if ((double)lt < 1e-7) {
        co = {-ab.y, ab.x, 0.f};
        lc2 = __fmaf_rn(ab.x, ab.x, yy);
    } else {
        co = {0.f, -ab.z, ab.y};
        lc2 = __fmaf_rn(ab.y, ab.y, __fmul_rn(ab.z, ab.z));
    }
    const float dc = (float)((double)sqrtf(lc2) + 1e-7);
    ex = {__fdiv_rn(co.x, dc), __fdiv_rn(co.y, dc), __fdiv_rn(co.z, dc)};
    return true;
}

// angle_i = float(double(2 i) * pi / double(n)) -- models/voting.py:33,99,140
__device__ __forceinline__ float rot_angle(int i, int n) {
    return (float)((double)(i * 2) * 3.14159265358979323846264338327950288 / (double)n);
}

// adaptive rotation count -- models/voting.py:31,97: min(int(odist / res * (2*M_PI)), n_rots)
__device__ __forceinline__ int adaptive_rots(float odist, float res, int n_rots) {
    const int m = (int)((double)(odist / res) * (2 * 3.14159265358979323846264338327950288));
    return m < n_rots ? m : n_rots;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: 4 uniform words per (key, counter).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// uniform in [0,1) with 24 random bits
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

}  // namespace cppf

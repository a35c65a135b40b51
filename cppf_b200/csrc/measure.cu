// Measurement aids behind bench.py's roofline block (never on the pose path):
//   * cppf_vote_count  -- the ALGORITHMIC work of one centre-vote launch, counted on the device with the reference's own
//     acceptance test (models/voting.py:21-39): rotation steps walked and in-bounds candidates; 8 x the latter is the
//     number of trilinear atomic adds the reference issues (:56-63) and every vote kernel of this library performs;
//   * cppf_peak_shared_atomics -- what THIS GPU sustains, in the same process, on the bare pattern those adds make in
//     shared memory (32 lanes x 8 corners of a random cell each: the random-bank roof of any unsorted scatter) and on a
//     conflict-free pattern (lane l -> bank l: the hardware roof, 32 atomics per clock per SM);
//   * cppf_peak_global_red -- the same for fp32 reductions on a grid in global memory (the reference's own atomicAdd).
#include "common.cuh"
#include "vote_common.cuh"

#include "../../include/cppf_b200.h"

namespace cppf {

struct CountParams {
    const float* points;
    const float* mu_nu;
    const uint8_t* bins;
    const float* lut;
    const void* idx;
    const float* corner;
    float res, lo, hx, hy, hz;
    int n_points;
    long long n_pairs;
    int n_rots, adaptive;
    unsigned long long* out;       // [0] rotation steps walked, [1] in-bounds candidates, [2] non-degenerate pairs
};

template <bool IDX64>
__global__ void __launch_bounds__(256) vote_count_kernel(const CountParams prm) {
    const float cx = __ldg(prm.corner), cy = __ldg(prm.corner + 1), cz = __ldg(prm.corner + 2);
    unsigned long long steps = 0, inb = 0, live = 0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < prm.n_pairs;
         p += (long long)gridDim.x * blockDim.x) {
        int ia, ib;
        pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib);
        float mu, nu;
        if (prm.bins != nullptr) {
            const uchar4 bn = __ldg(reinterpret_cast<const uchar4*>(prm.bins) + p);
            mu = __ldg(prm.lut + bn.x);
            nu = __ldg(prm.lut + 32 + bn.y);
        } else {
            const float2 mn = __ldg(reinterpret_cast<const float2*>(prm.mu_nu) + p);
            mu = mn.x;
            nu = mn.y;
        }
        const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
        f3 ab, ex;
        if (!pair_frame(a, b, ab, ex)) continue;                              // models/voting.py:21
        ++live;
        const f3 c = foot_point(a, ab, mu);
        const f3 x = scale3(ex, nu);
        const f3 y = cross3(x, ab);
        int n = prm.n_rots;
        if (prm.adaptive) n = adaptive_rots(nu, prm.res, prm.n_rots);         // :31
        for (int i = 0; i < n; ++i) {
            const float ang = rot_angle(i, n);
            const f3 off = circle_offset(x, y, cosf(ang), sinf(ang));
            const float gx = (c.x + off.x - cx) / prm.res, gy = (c.y + off.y - cy) / prm.res, gz = (c.z + off.z - cz) / prm.res;
            ++steps;
            if (gx < prm.lo || gy < prm.lo || gz < prm.lo || gx >= prm.hx || gy >= prm.hy || gz >= prm.hz) continue;   // :36-39
            ++inb;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_xor_sync(0xffffffffu, steps, o);
        inb += __shfl_xor_sync(0xffffffffu, inb, o);
        live += __shfl_xor_sync(0xffffffffu, live, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(prm.out, steps);
        atomicAdd(prm.out + 1, inb);
        atomicAdd(prm.out + 2, live);
    }
}

__device__ __forceinline__ uint32_t lcg_next(uint32_t& s) {
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

// mode 0: the 8-corner splat of a uniformly random base cell per lane (what an unsorted vote does to the banks);
// mode 1: conflict-free -- lane l always hits bank l (8 different rows of its own bank column)
__global__ void __launch_bounds__(1024) shared_atomics_kernel(unsigned* __restrict__ sink, int cells, int iters, int mode, int gz,
                                                              int gyz) {
    extern __shared__ unsigned sa_grid[];
    for (int i = threadIdx.x; i < cells; i += blockDim.x) sa_grid[i] = 0u;
    __syncthreads();
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 777u;
    const int lane = threadIdx.x & 31;
    const int rows = cells / 32 - 8;
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = lcg_next(s);
        if (mode == 0) {
            unsigned* c = sa_grid + r % (cells - gyz - gz - 2);
            atomicAdd(c, 1u); atomicAdd(c + 1, 1u); atomicAdd(c + gz, 1u); atomicAdd(c + gz + 1, 1u);
            atomicAdd(c + gyz, 1u); atomicAdd(c + gyz + 1, 1u); atomicAdd(c + gyz + gz, 1u); atomicAdd(c + gyz + gz + 1, 1u);
        } else {
            unsigned* c = sa_grid + (r % rows) * 32 + lane;
            atomicAdd(c, 1u); atomicAdd(c + 32, 1u); atomicAdd(c + 64, 1u); atomicAdd(c + 96, 1u);
            atomicAdd(c + 128, 1u); atomicAdd(c + 160, 1u); atomicAdd(c + 192, 1u); atomicAdd(c + 224, 1u);
        }
    }
    __syncthreads();
    unsigned acc = 0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) acc += sa_grid[i];
    if (acc == 0xFFFFFFFFu) sink[0] = acc;                                    // keeps the atomics alive
}

__global__ void __launch_bounds__(256) global_red_kernel(float* __restrict__ grid, int cells, int iters, int gz, int gyz) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 12345u;
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = lcg_next(s);
        float* c = grid + r % (cells - gyz - gz - 2);
        atomicAdd(c, 1.f); atomicAdd(c + 1, 1.f); atomicAdd(c + gz, 1.f); atomicAdd(c + gz + 1, 1.f);
        atomicAdd(c + gyz, 1.f); atomicAdd(c + gyz + 1, 1.f); atomicAdd(c + gyz + gz, 1.f); atomicAdd(c + gyz + gz + 1, 1.f);
    }
}

template <typename F>
static int best_ms(F launch, int reps, cudaStream_t stream, float* out_ms) {
    cudaEvent_t a, b;
    CPPF_RETURN_IF(cudaEventCreate(&a));
    CPPF_RETURN_IF(cudaEventCreate(&b));
    launch();                                                                 // warm-up
    float best = 1e30f;
    int err = 0;
    for (int r = 0; r < reps && err == 0; ++r) {
        cudaEventRecord(a, stream);
        launch();
        cudaEventRecord(b, stream);
        if (cudaEventSynchronize(b) != cudaSuccess) err = (int)cudaGetLastError();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (err == 0) err = (int)cudaGetLastError();
    *out_ms = best;
    return err;
}

}  // namespace cppf

using namespace cppf;

extern "C" int cppf_vote_count(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, const void* idx,
                               int idx_is_64, const float* corner, float res, int n_points, int64_t n_pairs, int n_rots, int gx,
                               int gy, int gz, int adaptive, uint64_t* out3, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if ((mu_nu == nullptr) == (bins == nullptr) || (bins != nullptr && lut == nullptr) || out3 == nullptr)
        return (int)cudaErrorInvalidValue;
    if (idx == nullptr) n_pairs = (int64_t)n_points * n_points;
    CPPF_RETURN_IF(cudaMemsetAsync(out3, 0, 3 * sizeof(uint64_t), stream));
    CountParams prm{points, mu_nu, bins, lut, idx, corner, res, float_ceil_p(0.01), float_ceil_p((double)gx - 1.01),
                    float_ceil_p((double)gy - 1.01), float_ceil_p((double)gz - 1.01), n_points, (long long)n_pairs, n_rots,
                    adaptive, reinterpret_cast<unsigned long long*>(out3)};
    const int blocks = sm_count() * 8;
    if (idx_is_64) vote_count_kernel<true><<<blocks, 256, 0, stream>>>(prm);
    else vote_count_kernel<false><<<blocks, 256, 0, stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_peak_shared_atomics(int gx, int gy, int gz, int conflict_free, int reps, double* g_atomics_per_s,
                                        void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cells = gx * gy * gz;
    if (cells < 1024 || cells * 4 > 200 * 1024 || g_atomics_per_s == nullptr) return (int)cudaErrorInvalidValue;
    CPPF_RETURN_IF(cudaFuncSetAttribute(shared_atomics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    unsigned* sink = nullptr;
    CPPF_RETURN_IF(cudaMalloc(&sink, 256));
    const int iters = 2048, blocks = sm_count(), threads = 1024;
    float ms = 0.f;
    const int err = best_ms([&] {
        shared_atomics_kernel<<<blocks, threads, (size_t)cells * 4, stream>>>(sink, cells, iters, conflict_free ? 1 : 0, gz, gy * gz);
        count_launch();
    }, reps > 0 ? reps : 5, stream, &ms);
    cudaFree(sink);
    if (err) return err;
    *g_atomics_per_s = (double)blocks * threads * iters * 8.0 / ((double)ms * 1e-3) * 1e-9;
    return 0;
}

extern "C" int cppf_peak_global_red(int gx, int gy, int gz, int reps, double* g_atomics_per_s, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int cells = gx * gy * gz;
    if (cells < 1024 || g_atomics_per_s == nullptr) return (int)cudaErrorInvalidValue;
    float* grid = nullptr;
    CPPF_RETURN_IF(cudaMalloc(&grid, (size_t)cells * 4));
    CPPF_RETURN_IF(cudaMemsetAsync(grid, 0, (size_t)cells * 4, stream));
    const int iters = 96, blocks = sm_count() * 8, threads = 256;
    float ms = 0.f;
    const int err = best_ms([&] {
        global_red_kernel<<<blocks, threads, 0, stream>>>(grid, cells, iters, gz, gy * gz);
        count_launch();
    }, reps > 0 ? reps : 3, stream, &ms);
    cudaFree(grid);
    if (err) return err;
    *g_atomics_per_s = (double)blocks * threads * iters * 8.0 / ((double)ms * 1e-3) * 1e-9;
    return 0;
}

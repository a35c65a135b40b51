// Micro-benchmark behind the vote-kernel design: throughput of the
// candidate accumulation primitives on B200.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// mode 0: global RED.ADD.F32, addresses uniform in [0,cells)
// mode 1: global RED.ADD.F32, 8-corner trilinear pattern (base, +1, +gz, +gz+1, +gyz, ...) like the vote kernel
// mode 2: global, hot: 90% uniform, 10% into 27 cells
__global__ void k_global(float* grid, int cells, int iters, int mode, int gz, int gyz) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 12345u;
    for (int i = 0; i < iters; ++i) {
        uint32_t r = lcg(s);
        if (mode == 0) {
            atomicAdd(grid + r % cells, 1.0f);
        } else if (mode == 1) {
            float* c = grid + r % (cells - gyz - gz - 2);
            atomicAdd(c, 1.f); atomicAdd(c + 1, 1.f); atomicAdd(c + gz, 1.f); atomicAdd(c + gz + 1, 1.f);
            atomicAdd(c + gyz, 1.f); atomicAdd(c + gyz + 1, 1.f); atomicAdd(c + gyz + gz, 1.f); atomicAdd(c + gyz + gz + 1, 1.f);
            i += 7;
        } else {
            uint32_t a = (r % 10 == 0) ? (cells / 2 + (r >> 4) % 27) : r % cells;
            atomicAdd(grid + a, 1.0f);
        }
    }
}

// shared-memory privatised grid: mode 0 u32 ATOMS.ADD, mode 1 float CAS loop, mode 2 u32 8-corner pattern,
// mode 3 u32 hot cells
__global__ void k_shared(unsigned* gout, int cells, int iters, int mode, int gz, int gyz) {
    extern __shared__ unsigned sg[];
    for (int i = threadIdx.x; i < cells; i += blockDim.x) sg[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 777u;
    for (int i = 0; i < iters; ++i) {
        uint32_t r = lcg(s);
        if (mode == 0) atomicAdd(sg + r % cells, 3u);
        else if (mode == 1) atomicAdd(reinterpret_cast<float*>(sg) + r % cells, 1.0f);
        else if (mode == 2) {
            unsigned* c = sg + r % (cells - gyz - gz - 2);
            atomicAdd(c, 1u); atomicAdd(c + 1, 1u); atomicAdd(c + gz, 1u); atomicAdd(c + gz + 1, 1u);
            atomicAdd(c + gyz, 1u); atomicAdd(c + gyz + 1, 1u); atomicAdd(c + gyz + gz, 1u); atomicAdd(c + gyz + gz + 1, 1u);
            i += 7;
        } else {
            uint32_t a = (r % 10 == 0) ? (cells / 2 + (r >> 4) % 27) : r % cells;
            atomicAdd(sg + a, 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) if (sg[i]) atomicAdd(gout + i, sg[i]);
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int cells_small = 18 * 55 * 18, cells_big = 64 * 64 * 64;
    float* g; cudaMalloc(&g, cells_big * 4 + 1024); cudaMemset(g, 0, cells_big * 4);
    const int iters = 4096, threads = 256;
    printf("{\"sms\": %d", sms);
    for (int per_sm : {4, 8}) {
        const int blocks = sms * per_sm;
        const double n = (double)blocks * threads * iters;
        struct { const char* name; int cells, mode, gz, gyz; } cg[] = {
            {"global_f32_uniform_71KB", cells_small, 0, 18, 55 * 18}, {"global_f32_uniform_1MB", cells_big, 0, 64, 4096},
            {"global_f32_corners_71KB", cells_small, 1, 18, 55 * 18}, {"global_f32_corners_1MB", cells_big, 1, 64, 4096},
            {"global_f32_hot_71KB", cells_small, 2, 18, 55 * 18}, {"global_f32_hot_1MB", cells_big, 2, 64, 4096}};
        for (auto& c : cg) {
            float ms = time_ms([&] { k_global<<<blocks, threads>>>(g, c.cells, iters, c.mode, c.gz, c.gyz); });
            printf(", \"%s_occ%d_Gatom_s\": %.2f", c.name, per_sm, n / ms * 1e-6);
        }
    }
    cudaFuncSetAttribute(k_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int thr : {256, 512, 1024}) {
        const int blocks = sms * (thr == 256 ? 2 : 1) * ((cells_small * 4 * (thr == 256 ? 2 : 1)) <= 200 * 1024 ? 1 : 1);
        const double n = (double)blocks * thr * iters;
        const char* names[] = {"shared_u32_uniform", "shared_f32cas_uniform", "shared_u32_corners", "shared_u32_hot"};
        for (int mode = 0; mode < 4; ++mode) {
            float ms = time_ms([&] { k_shared<<<blocks, thr, cells_small * 4>>>((unsigned*)g, cells_small, iters, mode, 18, 55 * 18); });
            printf(", \"%s_71KB_t%d_b%d_Gatom_s\": %.2f", names[mode], thr, blocks / sms, n / ms * 1e-6);
        }
    }
    printf("}\n");
    return 0;
}

// Micro-benchmark for the large-grid vote design: throughput of u32 reductions into a vote grid that is
// distributed over the shared memories of a thread-block cluster (DSMEM), against the same reductions into
// the CTA's own shared memory and into global memory.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/dsmem_atomics_bench tools/dsmem_atomics_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ void red_cluster(uint32_t local_smem_addr, uint32_t rank, uint32_t v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

// slab_cells u32 cells per CTA; the cluster-wide grid has cluster_size * slab_cells cells.
// mode 0: uniform random cell; mode 1: 8-corner pattern on a 64x64 plane layout (x-slabs: 4 corners in plane fx,
// 4 in plane fx+1, each plane = 4096 cells), mode 2: like 0 but always the CTA's own slab (local baseline through
// the same instruction); mode 3: local slab with plain atomicAdd (ATOMS) baseline.
__global__ void k_dsmem(unsigned* gout, int slab_cells, int iters, int mode, int cluster_size) {
    extern __shared__ unsigned sg[];
    for (int i = threadIdx.x; i < slab_cells; i += blockDim.x) sg[i] = 0;
    uint32_t my_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(my_rank));
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sg);
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x + 4242u;
    const uint32_t total = (uint32_t)slab_cells * cluster_size;
    for (int i = 0; i < iters; ++i) {
        const uint32_t r = lcg(s);
        if (mode == 0) {
            const uint32_t c = r % total;
            red_cluster(base + (c % slab_cells) * 4, c / slab_cells, 3u);
        } else if (mode == 1) {
            const uint32_t planes = total / 4096;
            const uint32_t fx = r % (planes - 1), yz = (r >> 7) % (4096 - 66);
            const uint32_t ppc = slab_cells / 4096;              // planes per CTA
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const uint32_t x = fx + dx, rank = x / ppc, off = (x % ppc) * 4096 + yz;
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + off * 4), "r"(rank));
                asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
                asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0+4], %1;" ::"r"(remote), "r"(1u) : "memory");
                asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0+256], %1;" ::"r"(remote), "r"(1u) : "memory");
                asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0+260], %1;" ::"r"(remote), "r"(1u) : "memory");
            }
            i += 7;
        } else if (mode == 2) {
            red_cluster(base + (r % slab_cells) * 4, my_rank, 3u);
        } else {
            atomicAdd(sg + r % slab_cells, 3u);
        }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    unsigned acc = 0;
    for (int i = threadIdx.x; i < slab_cells; i += blockDim.x) acc += sg[i];
    if (acc == 0xFFFFFFFFu) gout[0] = acc;
}

static float run(int blocks, int threads, int cluster, int slab_cells, int iters, int mode, unsigned* g) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = (size_t)slab_cells * 4;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_dsmem, g, slab_cells, iters, mode, cluster);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) return -1.f;
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* g;
    cudaMalloc(&g, 1 << 20);
    cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4096, threads = 1024;
    printf("{\"sms\": %d", sms);
    const char* names[] = {"uniform", "corners64", "own_slab_via_red_cluster", "own_slab_atoms"};
    for (int cluster : {1, 2, 4, 8}) {
        const int slab_cells = 8 * 4096;                       // 128 KB per CTA: 8 x-planes of a 64x64 plane
        const int blocks = (sms / cluster) * cluster;
        const double n = (double)blocks * threads * iters;
        for (int mode = 0; mode < 4; ++mode) {
            if (mode == 1 && cluster == 1) continue;
            const float ms = run(blocks, threads, cluster, slab_cells, iters, mode, g);
            printf(", \"cluster%d_%s_Gatom_s\": %.2f", cluster, names[mode], ms > 0 ? n / ms * 1e-6 : -1.0);
        }
    }
    printf("}\n");
    return 0;
}

// tcgen05 probe: D[128 x N] = A[128 x K] . B[N x K]^T in 3xTF32 (hi*hi + lo*hi + hi*lo), operands in
// shared memory in the no-swizzle K-major canonical layout, accumulator in TMEM, read back with
// tcgen05.ld.32x32b.  Validates the descriptor encodings used by csrc/encode_tc.cu on real hardware.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tc_probe tools/tc_probe.cu && tools/_bin/tc_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    return d;          // layout_type 0 = no swizzle, base_offset 0
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

template <int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int mode, int reps, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int KC = K / 4;                     // 16-byte chunks along K
    constexpr int A_PLANE = 128 * 16;             // bytes per K-chunk plane of A
    constexpr int B_PLANE = N * 16;
    unsigned char* sAh = smem;
    unsigned char* sAl = sAh + KC * A_PLANE;
    unsigned char* sBh = sAl + KC * A_PLANE;
    unsigned char* sBl = sBh + KC * B_PLANE;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;

    // operands -> canonical layout, split into tf32 hi + residual lo
    for (int c = 0; c < KC; ++c) {
        float4 v = *reinterpret_cast<const float4*>(A + t * K + 4 * c);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(sAh + c * A_PLANE + t * 16) = h;
        *reinterpret_cast<float4*>(sAl + c * A_PLANE + t * 16) = l;
    }
    for (int n = t; n < N; n += 128)
        for (int c = 0; c < KC; ++c) {
            float4 v = *reinterpret_cast<const float4*>(B + n * K + 4 * c);
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
            *reinterpret_cast<float4*>(sBh + c * B_PLANE + n * 16) = h;
            *reinterpret_cast<float4*>(sBl + c * B_PLANE + n * 16) = l;
        }
    constexpr int TCOLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = clock64();
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {
        if (t == 0) {
            const int passes = mode == 0 ? 1 : 3;
            for (int ps = 0; ps < passes; ++ps) {
                const uint32_t a_base = smem_u32(ps == 1 ? sAl : sAh);
                const uint32_t b_base = smem_u32(ps == 2 ? sBl : sBh);
                for (int j = 0; j < K / 8; ++j) {
                    const uint64_t da = make_desc(a_base + 2 * j * A_PLANE, A_PLANE, 128);
                    const uint64_t db = make_desc(b_base + 2 * j * B_PLANE, B_PLANE, 128);
                    mma_tf32(tmem, da, db, idesc, (ps | j) != 0 ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    long long t1 = clock64();
    if (t == 0 && cycles) *cycles = t1 - t0;
    // read back: thread t <- lane t, N columns, 16 at a time
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[t * N + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
}

// pipe throughput: one thread issues n_mma m128nNk8 MMAs round-robin over n_acc independent accumulators, one commit
template <int N>
__global__ void __launch_bounds__(128) pace_kernel(int n_mma, int n_acc, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (2 * 2048 + 2 * N * 16) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(smem), 2048, 128);
    const uint64_t db = make_desc(smem_u32(smem + 4096), N * 16, 128);
    long long t0 = clock64();
    if (t == 0) {
        for (int i = 0; i < n_mma; ++i) mma_tf32(tmem + (uint32_t)((i % n_acc) * N), da, db, idesc, i >= n_acc ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    long long t1 = clock64();
    if (t == 0) *cycles = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N>
void pace(const char* name) {
    long long* dC;
    cudaMalloc(&dC, 8);
    const size_t smem = 2 * 2048 + 2 * N * 16;
    for (int n_acc : {1, 2, 4}) {
        if (n_acc * N > 512) continue;
        long long c1 = 0, c2 = 0;
        pace_kernel<N><<<1, 128, smem>>>(64, n_acc, dC);
        cudaDeviceSynchronize();
        cudaMemcpy(&c1, dC, 8, cudaMemcpyDeviceToHost);
        pace_kernel<N><<<1, 128, smem>>>(576, n_acc, dC);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&c2, dC, 8, cudaMemcpyDeviceToHost);
        printf("pace %s n_acc=%d: %.1f cycles per m128k8 tf32 MMA (%s)\n", name, n_acc, (c2 - c1) / 512.0, cudaGetErrorString(e));
    }
    cudaFree(dC);
}

template <int N, int K>
int run(const char* name) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1234 + N * 100 + K);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 2 * (K / 4) * (128 * 16) + 2 * (K / 4) * (N * 16);
    cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bad = 0;
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dD, 0, D.size() * 4);
        probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, mode, 1, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s mode %d: CUDA error %s\n", name, mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) {
                    float a = A[m * K + k], b = B[n * K + k];
                    if (mode == 0) {
                        uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
                        ua &= 0xFFFFE000u; ub &= 0xFFFFE000u; memcpy(&a, &ua, 4); memcpy(&b, &ub, 4);
                    }
                    ref += (double)a * (double)b;
                }
                max_err = fmax(max_err, fabs(ref - (double)D[m * N + n]));
                max_ref = fmax(max_ref, fabs(ref));
            }
        const double tol = (mode == 0 ? 4e-6 : 4e-6) * fmax(1.0, max_ref);   // the TMEM accumulator itself is not IEEE fp32
        printf("%s %s: max|err| = %.3e (max|ref| %.2f) %s\n", name, mode == 0 ? "1xTF32 vs tf32-truncated fp64" : "3xTF32 vs fp64",
               max_err, max_ref, max_err < tol ? "OK" : "FAIL");
        bad += max_err < tol ? 0 : 1;
    }
    // crude pacing: 200 back-to-back 3-pass layers, each followed by a commit + mbarrier wait
    probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, 1, 200, dC);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    printf("%s: %.1f cycles per 3-pass layer (issue + commit + wait round trip)\n", name, cyc / 200.0);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return bad;
}

int main() {
    int bad = 0;
    bad += run<32, 32>("N=32  K=32");
    bad += run<16, 16>("N=16  K=16");
    bad += run<112, 16>("N=112 K=16");
    bad += run<144, 16>("N=144 K=16");
    bad += run<64, 8>("N=64  K=8 ");
    pace<16>("N=16 ");
    pace<32>("N=32 ");
    pace<64>("N=64 ");
    pace<128>("N=128");
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}

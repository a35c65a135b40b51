// tcgen05 / TMEM building blocks shared by the tensor-core kernels (encode_tc.cu: pair encoder; point_encoder_tc.cu: SPRIN):
// 128-row tiles (M = 128, cta_group::1), 3xTF32 operands in the no-swizzle K-major canonical layout -- element (row r, k)
// of a [rows x K] operand at byte (k / 4) * rows * 16 + r * 16 + (k % 4) * 4 -- one 128-thread group per tile with
// thread = row = TMEM lane, one elected MMA-issuing lane per group, one mbarrier per group.
#pragma once
#include "common.cuh"

namespace cppf {
namespace tc {

constexpr int kTile = 128;
constexpr int kAPlane = kTile * 16;               // bytes of one K plane of an A operand
constexpr int kABytes = 8 * kAPlane;              // K = 32
constexpr int kGroupBytes = 2 * kABytes;          // hi + lo
constexpr int kOnesBytes = 2 * kAPlane;           // the shared ones operand: K = 8 (two planes)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major: LBO = byte stride between K planes,
// SBO = byte stride between 8-row groups (= 128: rows are contiguous 16-byte chunks)
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[128 x N] = A . B^T in 3xTF32: lo.hi + hi.lo + hi.hi (small terms first)
template <int N, int K, bool ACC_INIT = false>
__device__ __forceinline__ void issue3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
    constexpr uint32_t bplane = N * 16;
    constexpr uint32_t b_lo_off = N * K * 4;
#pragma unroll
    for (int ps = 0; ps < 3; ++ps) {
        const uint32_t a = ps == 0 ? a_lo : a_hi;
        const uint32_t b = ps == 1 ? b_hi + b_lo_off : b_hi;
#pragma unroll
        for (int j = 0; j < K / 8; ++j)
            mma_tf32(tmem_d, kdesc(a + j * 2 * kAPlane, kAPlane), kdesc(b + j * 2 * bplane, bplane), idesc,
                     (ACC_INIT || (ps | j) != 0) ? 1u : 0u);
    }
}

// D[128 x N] = ones[128 x 8] . V[N x 8]^T: every row of D becomes the vector packed in V (k = 0: hi, k = 4: lo; exact)
template <int N>
__device__ __forceinline__ void issue_bias(uint32_t tmem_d, uint32_t ones, uint32_t v) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
    mma_tf32(tmem_d, kdesc(ones, kAPlane), kdesc(v, N * 16), idesc, 0u);
}

// one lane of a converged warp (elect.sync): the thread that issues tcgen05.mma / commit for its group
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// TMEM -> registers: this thread's lane, consecutive columns (load + wait in one statement so no
// use of the destination registers can be scheduled before the wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// the same 32 columns as two 16-column loads: the second is in flight while the caller consumes the first
// (`tmem_ld32_second` waits for it and hands the registers over; the "+r" operands keep every use behind the wait)
__device__ __forceinline__ void tmem_ld32_first(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr + 16)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_second(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// this thread's row, K chunk c (4 consecutive k) -> hi / lo planes
__device__ __forceinline__ void st_chunk(unsigned char* a_hi, int c, int row, float x0, float x1, float x2, float x3) {
    const float h0 = tf32_hi(x0), h1 = tf32_hi(x1), h2 = tf32_hi(x2), h3 = tf32_hi(x3);
    *reinterpret_cast<float4*>(a_hi + c * kAPlane + row * 16) = make_float4(h0, h1, h2, h3);
    *reinterpret_cast<float4*>(a_hi + kABytes + c * kAPlane + row * 16) = make_float4(x0 - h0, x1 - h1, x2 - h2, x3 - h3);
}


// publish this group's A operand, let the leader issue `ISSUE`, wait until the accumulators are complete.
// Expects in scope: g (group index, warp-uniform), lead_warp (warp-uniform bool), bar (shared address of the group's
// mbarrier), phase (uint32_t, toggled here).
#define CPPF_TC_STEP(ISSUE)                                                                                 \
    do {                                                                                                    \
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                        \
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                    \
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");                                          \
        if (lead_warp) {                                                                                    \
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                 \
            if (elect_one()) {                                                                              \
                ISSUE;                                                                                      \
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) \
                             : "memory");                                                                   \
            }                                                                                               \
        }                                                                                                   \
        __syncwarp();                                                                                       \
        mbar_wait(bar, phase);                                                                              \
        phase ^= 1u;                                                                                        \
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                     \
    } while (0)

}  // namespace tc
}  // namespace cppf

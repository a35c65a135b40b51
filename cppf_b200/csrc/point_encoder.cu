// Point encoder: exact k-nearest-neighbour selection and the SPRIN
// rotation-invariant convolution of models/model.py:46-77 + models/sprin.py:40-107, one warp per point.
//
//   cppf_knn            : for each point the k smallest exact squared distances (self included, like
//                         torch.topk(dist, k, largest=False) at models/model.py:47), one warp per query, no
//                         N x N matrix: two sweeps over the cloud (1024-bin histogram of the top bits of d^2, then
//                         collection + ranking of the boundary bin); a 4 x 8-bit radix select on the float bits
//                         (5 sweeps) is the fallback when the boundary bin overflows.
//   cppf_point_encode   : models/model.py:63-77 for one neighbour list: neighbour features
//                         [|p_k - p|, n_k . n] (:49-53), rifeat invariants (sprin.py:40-60), the kernel MLP
//                         6 -> 32 -> 64 -> 32 -> 32 -> 32 with LayerNorm + ReLU between layers (sprin.py:63-71)
//                         on 32-row warp tiles (register-tiled FFMA, activations in the warp's private
//                         shared-memory buffers), the rank-32 contraction over the k neighbours (:98), the
//                         64 -> 32 output Linear + LayerNorm (:99-101), GlobalInfoProp's 32 -> 8 Linear and the
//                         max over all points (sprin.py:74-83, ordered-int atomic max: exact and order-free).
//   cppf_point_glob     : writes the 8 global-max columns into every row of feat[N, 40].
#include "encode.cuh"
#include "tc_common.cuh"

#include "../../include/cppf_b200.h"

#include <math.h>
#include <stdlib.h>

namespace cppf {
namespace pe {

// ---- weight blob (floats); Linear matrices are K-MAJOR with columns permuted for the warp tile
// (column j = og + 8c stored at og*NO + c, NO = cols/8; mlp_layout.h), biases permuted the same way
constexpr int kW1 = 0;                       // [6][32]
constexpr int kB1 = kW1 + 6 * 32;
constexpr int kG1 = kB1 + 32;                // LayerNorm gamma / beta in natural order
constexpr int kE1 = kG1 + 32;
constexpr int kW2 = kE1 + 32;                // [32][64]
constexpr int kB2 = kW2 + 32 * 64;
constexpr int kG2 = kB2 + 64;
constexpr int kE2 = kG2 + 64;
constexpr int kW3 = kE2 + 64;                // [64][32]
constexpr int kB3 = kW3 + 64 * 32;
constexpr int kG3 = kB3 + 32;
constexpr int kE3 = kG3 + 32;
constexpr int kW4 = kE3 + 32;                // [32][32]
constexpr int kB4 = kW4 + 32 * 32;
constexpr int kG4 = kB4 + 32;
constexpr int kE4 = kG4 + 32;
constexpr int kW5 = kE4 + 32;                // [32][32]
constexpr int kB5 = kW5 + 32 * 32;
constexpr int kWo = kB5 + 32;                // outnet [64][32] k-major, natural columns (k = rank*2 + nbr feature)
constexpr int kBo = kWo + 64 * 32;
constexpr int kGo = kBo + 32;
constexpr int kEo = kGo + 32;
constexpr int kWa = kEo + 32;                // GlobalInfoProp linear [32][8] k-major
constexpr int kBa = kWa + 32 * 8;
constexpr int kBlobFloats = kBa + 8;

#ifndef CPPF_PE_LOCAL_MAX
#define CPPF_PE_LOCAL_MAX 1
#endif
#ifndef CPPF_PE_WARPS
#define CPPF_PE_WARPS 8
#endif
constexpr int kWarps = CPPF_PE_WARPS;
constexpr int kWarpFloats = 32 * XS + 64 * XS + 64 + 64;      // P[32][XS], Q[64][XS], nbr feats [2][32], contraction [64]
constexpr float kLnEps = 1e-5f;                               // torch.nn.LayerNorm default

__device__ __forceinline__ void load8(const float* w, float (&v)[8]) {
    const float4 t0 = *reinterpret_cast<const float4*>(w);
    const float4 t1 = *reinterpret_cast<const float4*>(w + 4);
    v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
}

// acc[i][c] += sum_k X[k][i] * W[k][c] for a 64-column layer (NO = 8)
template <int K>
__device__ __forceinline__ void tile_gemm8(const float* __restrict__ X, const float* __restrict__ W, float (&acc)[8][8]) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(X + k * XS);
        const float4 a1 = *reinterpret_cast<const float4*>(X + k * XS + 4);
        float w[8];
        load8(W + k * 64, w);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(a[i], w[c], acc[i][c]);
    }
}

// LayerNorm over the NF features of each row (lane = row) followed by ReLU, in place -- sprin.py:67-68
template <int NF>
__device__ __forceinline__ void layer_norm_relu(float* __restrict__ buf, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, int lane) {
    float v[NF];
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        v[j] = buf[j * XS + lane];
        mean += v[j];
    }
    mean *= 1.f / NF;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        const float d = v[j] - mean;
        var = fmaf(d, d, var);
    }
    const float rstd = rsqrtf(var * (1.f / NF) + kLnEps);
#pragma unroll
    for (int j = 0; j < NF; ++j) buf[j * XS + lane] = fmaxf(fmaf((v[j] - mean) * rstd, gamma[j], beta[j]), 0.f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

struct Params {
    const float* pc;
    const float* nrm;
    const long long* nbrs;     // [N][k] int64 (torch.topk indices)
    const float* blob;
    float* feat;               // [N][40]: columns 0:32 written here
    float* glob;               // [8], pre-set to -inf
    int n_points;
    int k;
};

__global__ void __launch_bounds__(kWarps * 32, 1) point_encode_kernel(const Params prm) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* P = smem + kBlobFloats + warp * kWarpFloats;
    float* Q = P + 32 * XS;
    float* NF = Q + 64 * XS;                  // [2][32]
    float* CT = NF + 64;                      // [64]
    {
        const float4* src = reinterpret_cast<const float4*>(prm.blob);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < kBlobFloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int og = lane & 7, pb = (lane >> 3) * 8;
    const float inv_k = 1.f / (float)prm.k;
    float tmax = -INFINITY;                   // running max of this warp's GlobalInfoProp columns (lanes 0..7)

    for (int n = blockIdx.x * kWarps + warp; n < prm.n_points; n += gridDim.x * kWarps) {
        const f3 ctr = ld3(prm.pc, n), nc = ld3(prm.nrm, n);
        // neighbours lane and 32 + lane of this point (models/model.py:48,52: absolute coordinates, normals)
        long long id[2];
        f3 nb[2], nn[2];
        bool ok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            ok[h] = h * 32 + lane < prm.k;
            id[h] = ok[h] ? __ldg(prm.nbrs + (long long)n * prm.k + h * 32 + lane) : 0;
            nb[h] = ld3(prm.pc, id[h]);
            nn[h] = ld3(prm.nrm, id[h]);
        }
        f3 mean;                              // sprin.py:51 torch.mean over the k neighbours
        mean.x = warp_sum((ok[0] ? nb[0].x : 0.f) + (ok[1] ? nb[1].x : 0.f)) * inv_k;
        mean.y = warp_sum((ok[0] ? nb[0].y : 0.f) + (ok[1] ? nb[1].y : 0.f)) * inv_k;
        mean.z = warp_sum((ok[0] ? nb[0].z : 0.f) + (ok[1] ? nb[1].z : 0.f)) * inv_k;
        float c0 = 0.f, c1 = 0.f;             // contraction accumulators of rank `lane`
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            if (h * 32 >= prm.k) break;
            {   // rifeat (sprin.py:52-60) and the two neighbour features (models/model.py:49-53), lane = row
                const f3 l1 = mean - nb[h], l2 = nb[h] - ctr, l3 = ctr - mean;
                const float n1 = len3(l1), n2 = len3(l2), n3 = len3(l3);
                float ri[6] = {n1, n2, n3, dot3(l1, l2) / (n1 * n2 + 1e-7f), dot3(l2, l3) / (n2 * n3 + 1e-7f),
                               dot3(l3, l1) / (n3 * n1 + 1e-7f)};
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 6; ++q) Q[q * XS + lane] = ok[h] ? ri[q] : 0.f;
                NF[lane] = ok[h] ? n2 : 0.f;                          // |p_k - p|   (rows beyond k contribute 0)
                NF[32 + lane] = ok[h] ? dot3(nn[h], nc) : 0.f;        // n_k . n
            }
            __syncwarp();
            {   // Linear(6, 32) -> LN -> ReLU
                float acc[8][4];
                zero_acc(acc);
                tile_gemm<6, 4>(Q + pb, sW + kW1 + og * 4, 32, acc);
                float bias[4];
                WVec<4>::load(sW + kB1 + og * 4, bias);
                tile_store<4>(P + pb, og, acc, bias, nullptr, 0, 0);
                __syncwarp();
                layer_norm_relu<32>(P, sW + kG1, sW + kE1, lane);
                __syncwarp();
            }
            {   // Linear(32, 64) -> LN -> ReLU
                float acc[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
                tile_gemm8<32>(P + pb, sW + kW2 + og * 8, acc);
                float bias[8];
                load8(sW + kB2 + og * 8, bias);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int j = og + 8 * c;
                    *reinterpret_cast<float4*>(Q + pb + j * XS) =
                        make_float4(acc[0][c] + bias[c], acc[1][c] + bias[c], acc[2][c] + bias[c], acc[3][c] + bias[c]);
                    *reinterpret_cast<float4*>(Q + pb + j * XS + 4) =
                        make_float4(acc[4][c] + bias[c], acc[5][c] + bias[c], acc[6][c] + bias[c], acc[7][c] + bias[c]);
                }
                __syncwarp();
                layer_norm_relu<64>(Q, sW + kG2, sW + kE2, lane);
                __syncwarp();
            }
            {   // Linear(64, 32) -> LN -> ReLU
                float acc[8][4];
                zero_acc(acc);
                tile_gemm<64, 4>(Q + pb, sW + kW3 + og * 4, 32, acc);
                float bias[4];
                WVec<4>::load(sW + kB3 + og * 4, bias);
                tile_store<4>(P + pb, og, acc, bias, nullptr, 0, 0);
                __syncwarp();
                layer_norm_relu<32>(P, sW + kG3, sW + kE3, lane);
                __syncwarp();
            }
            {   // Linear(32, 32) -> LN -> ReLU
                float acc[8][4];
                zero_acc(acc);
                tile_gemm<32, 4>(P + pb, sW + kW4 + og * 4, 32, acc);
                float bias[4];
                WVec<4>::load(sW + kB4 + og * 4, bias);
                tile_store<4>(Q + pb, og, acc, bias, nullptr, 0, 0);
                __syncwarp();
                layer_norm_relu<32>(Q, sW + kG4, sW + kE4, lane);
                __syncwarp();
            }
            {   // Linear(32, 32): the rank-32 kernel of each neighbour
                float acc[8][4];
                zero_acc(acc);
                tile_gemm<32, 4>(Q + pb, sW + kW5 + og * 4, 32, acc);
                float bias[4];
                WVec<4>::load(sW + kB5 + og * 4, bias);
                tile_store<4>(P + pb, og, acc, bias, nullptr, 0, 0);
                __syncwarp();
            }
            // einsum('bnkr,bnki->bnri') (sprin.py:98): lane = rank r, sum over this tile's rows
#pragma unroll 8
            for (int row = 0; row < 32; ++row) {
                const float kv = P[lane * XS + row];
                c0 = fmaf(kv, NF[row], c0);
                c1 = fmaf(kv, NF[32 + row], c1);
            }
        }
        __syncwarp();
        CT[2 * lane] = c0;                    // flatten(-2): index r*2 + i
        CT[2 * lane + 1] = c1;
        __syncwarp();
        float o = sW[kBo + lane];             // outnet Linear(64, 32), lane = output column   (sprin.py:99)
#pragma unroll 8
        for (int k = 0; k < 64; ++k) o = fmaf(CT[k], sW[kWo + k * 32 + lane], o);
        const float m = warp_sum(o) * (1.f / 32.f);          // LayerNorm(32) across the lanes   (sprin.py:100-101)
        const float d = o - m;
        const float var = warp_sum(d * d) * (1.f / 32.f);
        const float y = fmaf(d * rsqrtf(var + kLnEps), sW[kGo + lane], sW[kEo + lane]);
        prm.feat[(long long)n * 40 + lane] = y;
        // GlobalInfoProp (sprin.py:80-82): 8 columns, max over all points
        float t = lane < 8 ? sW[kBa + lane] : 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float yk = __shfl_sync(0xffffffffu, y, k);
            if (lane < 8) t = fmaf(yk, sW[kWa + k * 8 + lane], t);
        }
#if CPPF_PE_LOCAL_MAX
        tmax = (t > tmax || t != t) ? t : tmax;               // NaN sticks, like torch.max
    }
    if (lane < 8) atomic_max_float(prm.glob + lane, tmax);    // one set of atomics per warp instead of one per point
#else
        if (lane < 8) atomic_max_float(prm.glob + lane, t);
    }
    (void)tmax;
#endif
}

// ---------------------------------------------------------------------------------------------------------
// The same convolution on the 5th-generation tensor cores (tcgen05 + TMEM, building blocks of tc_common.cuh).
//
//   * a tile is 128 rows = TWO points x 64 neighbour slots (k <= 64; slots >= k are zero rows); thread t of a 128-thread
//     group owns row t = TMEM lane t through the whole kernel MLP, so LayerNorm (a per-row statistic) needs no shuffle;
//   * every Linear is D[128 x N] (TMEM, fp32) = A[128 x K] . W^T in 3xTF32 (hi.hi + hi.lo + lo.hi), its bias put into the
//     accumulators by a ones-operand MMA; the epilogue of a layer (TMEM -> registers, LayerNorm, ReLU, tf32 split) writes the
//     next layer's A operand.  Linear(64, 32) runs as two K = 32 halves through the same 32 KB A buffer, so that four
//     groups (four tiles in flight per SM, one group's epilogue under the others' MMAs) fit beside the weights;
//   * the MMA operands (hi / lo split, canonical layout) are packed by the host behind the FFMA section of the blob
//     (cppf_b200/model.py: pack_pe_weights) and staged with one straight copy;
//   * the rank-32 contraction over a point's neighbours (sprin.py:98) goes through the group's A buffer, free again after
//     the last MMA: every row parks its 32 kernel values there and 128 threads sum (rank, feature) columns over 64 rows;
//     the 64 -> 32 output Linear + LayerNorm + GlobalInfoProp run on one warp per point as in the FFMA kernel.
// 245 760 rows (N = 4096, k = 60) -> 1920 tiles, 67 MMAs of shape 128 x N x 8 each.
namespace tcpe {
using namespace tc;
constexpr int kGroups = 4;
constexpr int kThreads = kGroups * kTile;
constexpr int kSlots = 64;                       // neighbour slots per point in a tile
// shared-memory operand section (floats): B operands [K/4][N][4], hi block then lo block
constexpr int oW1 = 0;                           // N = 32, K = 8 (k = 6, 7 zero)
constexpr int oW2 = oW1 + 2 * 32 * 8;            // N = 64, K = 32
constexpr int oW3a = oW2 + 2 * 64 * 32;          // N = 32, K = 32: input columns 0..31 of Linear(64, 32)
constexpr int oW3b = oW3a + 2 * 32 * 32;         //                 input columns 32..63
constexpr int oW4 = oW3b + 2 * 32 * 32;
constexpr int oW5 = oW4 + 2 * 32 * 32;
constexpr int oV1 = oW5 + 2 * 32 * 32;           // bias operands [N x 8]: k = 0 hi, k = 4 lo
constexpr int oV2 = oV1 + 32 * 8;
constexpr int oV3 = oV2 + 64 * 8;
constexpr int oV4 = oV3 + 32 * 8;
constexpr int oV5 = oV4 + 32 * 8;
constexpr int oLN = oV5 + 32 * 8;                // gamma1 beta1 | gamma2 beta2 | gamma3 beta3 | gamma4 beta4
constexpr int oG1 = oLN, oE1 = oG1 + 32, oG2 = oE1 + 32, oE2 = oG2 + 64, oG3 = oE2 + 64, oE3 = oG3 + 32, oG4 = oE3 + 32,
              oE4 = oG4 + 32;
constexpr int oOut = oE4 + 32;                   // outnet ... GlobalInfoProp, verbatim from the blob (kWo .. kBlobFloats)
constexpr int kOutFloats = kBlobFloats - kWo;
constexpr int kSmemFloats = oOut + kOutFloats;
constexpr int kSmemBytes = kSmemFloats * 4 + kOnesBytes + kGroups * kGroupBytes;
constexpr int kTmemColsPerGroup = 64;
// scratch inside a group's A buffer (floats), used between the last MMA of a tile and the first store of the next
constexpr int kSS = 33;                          // row stride of the parked kernel values: (row + rank) mod 32 banks
constexpr int sNF = kTile * kSS;                 // [128][2] neighbour features
constexpr int sCT = 6144;                        // [2][64] contraction, in plane 4 of the lo half: the first two K planes of
                                                 // either half are the ones the next tile's step 0 writes
static_assert(sNF + 2 * kTile <= sCT && (sCT + 128) * 4 <= kGroupBytes, "scratch fits the A buffer");
static_assert(sCT * 4 >= kABytes + 2 * kAPlane, "contraction result clear of the next tile's K = 8 operand");
static_assert(kSmemBytes <= 227 * 1024, "shared memory");
static_assert(kBlobFloats % 4 == 0 && kSmemFloats % 4 == 0, "float4 staging");

// LayerNorm + ReLU of NF values of this thread's row (sprin.py:67-68), then the row's K chunks of the next A operand
template <int NF>
__device__ __forceinline__ void ln_relu(float (&v)[NF], const float* __restrict__ gamma, const float* __restrict__ beta) {
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < NF; ++j) mean += v[j];
    mean *= 1.f / NF;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        v[j] -= mean;
        var = fmaf(v[j], v[j], var);
    }
    const float rstd = rsqrtf(var * (1.f / NF) + kLnEps);
#pragma unroll
    for (int j = 0; j < NF; j += 4) {
        const float4 gm = *reinterpret_cast<const float4*>(gamma + j), bt = *reinterpret_cast<const float4*>(beta + j);
        v[j] = fmaxf(fmaf(v[j] * rstd, gm.x, bt.x), 0.f);
        v[j + 1] = fmaxf(fmaf(v[j + 1] * rstd, gm.y, bt.y), 0.f);
        v[j + 2] = fmaxf(fmaf(v[j + 2] * rstd, gm.z, bt.z), 0.f);
        v[j + 3] = fmaxf(fmaf(v[j + 3] * rstd, gm.w, bt.w), 0.f);
    }
}

__global__ void __launch_bounds__(kThreads, 1) point_encode_tc_kernel(const Params prm) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar[kGroups];
    __shared__ uint32_t s_tmem;
    __shared__ float s_red[kGroups][4][3];
    float* sWf = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction (see encode_tc.cu)
    const int lane = tid & 31;
    const int g = warp >> 2, tg = tid & 127, wg = warp & 3;
    unsigned char* s_ones = smem + kSmemFloats * 4;
    unsigned char* a_hi = s_ones + kOnesBytes + g * kGroupBytes;
    float* scratch = reinterpret_cast<float*>(a_hi);

    {   // operands -> shared memory (packed by the host behind the FFMA section of the blob), TMEM allocation, barriers
        const float4* src = reinterpret_cast<const float4*>(prm.blob + kBlobFloats);
        float4* dst = reinterpret_cast<float4*>(sWf);
        for (int i = tid; i < kSmemFloats / 4; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < kOnesBytes / 16; i += kThreads)                 // k = 0 and k = 4 of every row are 1
            reinterpret_cast<float4*>(s_ones)[i] = make_float4(1.f, 0.f, 0.f, 0.f);
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                         "r"(kGroups * kTmemColsPerGroup)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (tid < kGroups) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[tid])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t tm = s_tmem + g * kTmemColsPerGroup;
    const uint32_t tml = tm + ((uint32_t)(wg * 32) << 16);                    // this warp's lane quarter for tcgen05.ld
    const uint32_t bar = smem_u32(&s_bar[g]);
    const uint32_t sA = smem_u32(a_hi), sAl = sA + kABytes;
    const uint32_t sW = smem_u32(sWf);
    const uint32_t sOnes = smem_u32(s_ones);
    const bool lead_warp = wg == 0;
    uint32_t phase = 0;
    const float* sOutW = sWf + oOut;                                          // offsets below relative to kWo
    const int ps = tg >> 6, slot = tg & (kSlots - 1);
    const float inv_k = 1.f / (float)prm.k;
    const int n_tiles = (prm.n_points + 1) / 2;
    float tmax = -INFINITY;                                                   // warps 0, 1 of a group, lanes 0..7

    // The neighbour data of a tile is fetched one tile ahead (index early in the previous tile, the two gathers it feeds in
    // the middle of it): an SM configured with 200 KB of shared memory has next to no L1, and the index -> point chain is
    // two L2 round trips at the head of every tile otherwise.
    const int t_stride = gridDim.x * kGroups;
    long long id_n = 0;
    f3 ctr_n = {0.f, 0.f, 0.f}, ncn_n = ctr_n, nb_n = ctr_n, nn_n = ctr_n;
    auto fetch_id = [&](int tile) {
        const int n = 2 * tile + ps;
        id_n = (tile < n_tiles && n < prm.n_points && slot < prm.k) ? __ldg(prm.nbrs + (long long)n * prm.k + slot) : 0;
    };
    auto fetch_pts = [&](int tile) {
        const int n = 2 * tile + ps;
        const int nc_ = (tile < n_tiles && n < prm.n_points) ? n : 0;
        ctr_n = ld3(prm.pc, nc_); ncn_n = ld3(prm.nrm, nc_);
        nb_n = ld3(prm.pc, id_n); nn_n = ld3(prm.nrm, id_n);
    };
    int tile = blockIdx.x * kGroups + g;
    fetch_id(tile);
    fetch_pts(tile);
    for (; tile < n_tiles; tile += t_stride) {
        const int n = 2 * tile + ps;
        const bool have_pt = n < prm.n_points;
        const bool ok = have_pt && slot < prm.k;
        const f3 ctr = ctr_n, ncn = ncn_n, nb = nb_n, nn = nn_n;
        fetch_id(tile + t_stride);
        {   // mean of the point's k neighbours (sprin.py:51): two warps per point
            const float sx = warp_sum(ok ? nb.x : 0.f), sy = warp_sum(ok ? nb.y : 0.f), sz = warp_sum(ok ? nb.z : 0.f);
            if (lane == 0) {
                s_red[g][wg][0] = sx; s_red[g][wg][1] = sy; s_red[g][wg][2] = sz;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        }
        f3 mean;
        mean.x = (s_red[g][2 * ps][0] + s_red[g][2 * ps + 1][0]) * inv_k;
        mean.y = (s_red[g][2 * ps][1] + s_red[g][2 * ps + 1][1]) * inv_k;
        mean.z = (s_red[g][2 * ps][2] + s_red[g][2 * ps + 1][2]) * inv_k;
        float nf0 = 0.f, nf1 = 0.f;
        {   // rifeat (sprin.py:52-60) and the two neighbour features (models/model.py:49-53)
            const f3 l1 = mean - nb, l2 = nb - ctr, l3 = ctr - mean;
            const float n1 = len3(l1), n2 = len3(l2), n3 = len3(l3);
            const float r3 = dot3(l1, l2) / (n1 * n2 + 1e-7f), r4 = dot3(l2, l3) / (n2 * n3 + 1e-7f),
                        r5 = dot3(l3, l1) / (n3 * n1 + 1e-7f);
            if (ok) {
                st_chunk(a_hi, 0, tg, n1, n2, n3, r3);
                st_chunk(a_hi, 1, tg, r4, r5, 0.f, 0.f);
                nf0 = n2;                                                     // |p_k - p|
                nf1 = dot3(nn, ncn);                                          // n_k . n
            } else {
                st_chunk(a_hi, 0, tg, 0.f, 0.f, 0.f, 0.f);
                st_chunk(a_hi, 1, tg, 0.f, 0.f, 0.f, 0.f);
            }
        }
        float x[32], y[32];
        // ---- Linear(6, 32) -> LN -> ReLU
        CPPF_TC_STEP((issue_bias<32>(tm, sOnes, sW + oV1 * 4), issue3<32, 8, true>(tm, sA, sAl, sW + oW1 * 4)));
        tmem_ld32(tml, x);
        ln_relu<32>(x, sWf + oG1, sWf + oE1);
#pragma unroll
        for (int q = 0; q < 8; ++q) st_chunk(a_hi, q, tg, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        // ---- Linear(32, 64) -> LN -> ReLU
        CPPF_TC_STEP((issue_bias<64>(tm, sOnes, sW + oV2 * 4), issue3<64, 32, true>(tm, sA, sAl, sW + oW2 * 4)));
        {
            float z[64];
            tmem_ld32(tml, x);
            tmem_ld32(tml + 32, y);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                z[q] = x[q];
                z[32 + q] = y[q];
            }
            ln_relu<64>(z, sWf + oG2, sWf + oE2);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                x[q] = z[q];
                y[q] = z[32 + q];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) st_chunk(a_hi, q, tg, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        // ---- Linear(64, 32) in two K = 32 halves -> LN -> ReLU
        CPPF_TC_STEP((issue_bias<32>(tm, sOnes, sW + oV3 * 4), issue3<32, 32, true>(tm, sA, sAl, sW + oW3a * 4)));
#pragma unroll
        for (int q = 0; q < 8; ++q) st_chunk(a_hi, q, tg, y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
        fetch_pts(tile + t_stride);
        CPPF_TC_STEP((issue3<32, 32, true>(tm, sA, sAl, sW + oW3b * 4)));
        tmem_ld32(tml, x);
        ln_relu<32>(x, sWf + oG3, sWf + oE3);
#pragma unroll
        for (int q = 0; q < 8; ++q) st_chunk(a_hi, q, tg, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        // ---- Linear(32, 32) -> LN -> ReLU
        CPPF_TC_STEP((issue_bias<32>(tm, sOnes, sW + oV4 * 4), issue3<32, 32, true>(tm, sA, sAl, sW + oW4 * 4)));
        tmem_ld32(tml, x);
        ln_relu<32>(x, sWf + oG4, sWf + oE4);
#pragma unroll
        for (int q = 0; q < 8; ++q) st_chunk(a_hi, q, tg, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        // ---- Linear(32, 32): the rank-32 kernel of each neighbour
        CPPF_TC_STEP((issue_bias<32>(tm, sOnes, sW + oV5 * 4), issue3<32, 32, true>(tm, sA, sAl, sW + oW5 * 4)));
        tmem_ld32(tml, x);
        // ---- einsum('bnkr,bnki->bnri') (sprin.py:98) through the A buffer (every MMA that read it has completed)
#pragma unroll
        for (int r = 0; r < 32; ++r) scratch[tg * kSS + r] = x[r];
        *reinterpret_cast<float2*>(scratch + sNF + 2 * tg) = make_float2(nf0, nf1);
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        {
            const int r = slot >> 1, i = slot & 1;                            // this thread's (rank, neighbour feature) of point ps
            const float* kv = scratch + (ps * kSlots) * kSS + r;
            const float* nf = scratch + sNF + 2 * (ps * kSlots) + i;
            float c = 0.f;
#pragma unroll 8
            for (int row = 0; row < kSlots; ++row) c = fmaf(kv[row * kSS], nf[2 * row], c);
            scratch[sCT + ps * 64 + slot] = c;                                // flatten(-2): index r * 2 + i
        }
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
        if (wg < 2) {                        // warp wg of the group finishes point wg of the tile: lane = output column
            const int np = 2 * tile + wg;
            const float* ct = scratch + sCT + wg * 64;
            float o = sOutW[(kBo - kWo) + lane];                              // outnet Linear(64, 32)   (sprin.py:99)
#pragma unroll 8
            for (int k = 0; k < 64; ++k) o = fmaf(ct[k], sOutW[k * 32 + lane], o);
            const float m = warp_sum(o) * (1.f / 32.f);                       // LayerNorm(32) across the lanes (sprin.py:100-101)
            const float d = o - m;
            const float var = warp_sum(d * d) * (1.f / 32.f);
            const float yv = fmaf(d * rsqrtf(var + kLnEps), sOutW[(kGo - kWo) + lane], sOutW[(kEo - kWo) + lane]);
            if (np < prm.n_points) prm.feat[(long long)np * 40 + lane] = yv;
            float t = lane < 8 ? sOutW[(kBa - kWo) + lane] : 0.f;             // GlobalInfoProp (sprin.py:80-82)
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float yk = __shfl_sync(0xffffffffu, yv, k);
                if (lane < 8) t = fmaf(yk, sOutW[(kWa - kWo) + k * 8 + lane], t);
            }
            if (np < prm.n_points) tmax = (t > tmax || t != t) ? t : tmax;    // NaN sticks, like torch.max
        }
    }
    if (wg < 2 && lane < 8 && tmax != -INFINITY) atomic_max_float(prm.glob + lane, tmax);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(kGroups * kTmemColsPerGroup)
                     : "memory");
}
}  // namespace tcpe

__global__ void __launch_bounds__(256) point_glob_kernel(float* __restrict__ feat, const float* __restrict__ glob, int n_points) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_points * 8) feat[(long long)(i >> 3) * 40 + 32 + (i & 7)] = glob[i & 7];
}

__global__ void glob_init_kernel(float* glob) {
    if (threadIdx.x < 8) glob[threadIdx.x] = -INFINITY;
}

// ---------------------------------------------------------------------------------------------------------
// k nearest neighbours by radix select on the bits of the exact squared distance (>= 0, so the float order is
// the unsigned-integer order of the bits).  One warp per query; every sweep recomputes the distances with the
// same expression, so the passes agree bit for bit.
__device__ __forceinline__ unsigned d2_bits(f3 q, const float* __restrict__ pc, int j) {
    const f3 p = ld3(pc, j);
    const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    return __float_as_uint(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
}

#ifndef CPPF_KNN_WARPS
#define CPPF_KNN_WARPS 14     // one CTA per SM is resident: 4096 queries / 14 = 293 CTAs = 1.98 waves of 148
#endif
constexpr int kKnnWarps = CPPF_KNN_WARPS;
constexpr int kKnnBins = 1024;           // coarse histogram: sign (0) + exponent + 2 mantissa bits of d^2 = bits >> 21
constexpr int kKnnCand = 256;            // candidates of the boundary bin kept in shared memory

// One query, generic path: radix select on all 32 bits (4 sweeps of 8 bits + the collection sweep).
__device__ __noinline__ void knn_query_radix(const float* __restrict__ pc, int n_points, int k, int qi, unsigned* hist,
                                                int lane, long long* __restrict__ out) {
    const unsigned lt_mask = (1u << lane) - 1u;
    const f3 q = ld3(pc, qi);
    unsigned prefix = 0;                  // the bits of the k-th smallest distance found so far
    int need = k;                         // rank of the wanted element among those matching the prefix
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
#pragma unroll
        for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0u;
        __syncwarp();
        for (int j = lane; j < n_points; j += 32) {
            const unsigned b = d2_bits(q, pc, j);
            if (pass == 0 || (b >> (shift + 8)) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncwarp();
        unsigned cnt[8], mine = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            cnt[i] = hist[lane * 8 + i];
            mine += cnt[i];
        }
        unsigned incl = mine;             // inclusive scan of the per-lane totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned excl = incl - mine;
        const bool here = (unsigned)need > excl && (unsigned)need <= incl;   // exactly one lane
        int digit = 0, below = 0;
        if (here) {
            unsigned run = excl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((unsigned)need > run && (unsigned)need <= run + cnt[i]) {
                    digit = lane * 8 + i;
                    below = (int)run;
                }
                run += cnt[i];
            }
        }
        const unsigned src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
        digit = __shfl_sync(0xffffffffu, digit, src);
        below = __shfl_sync(0xffffffffu, below, src);
        prefix = (prefix << 8) | (unsigned)digit;
        need -= below;
        __syncwarp();
    }
    // collection sweep: everything strictly below the threshold, then the first `need` ties in index order
    const unsigned thr = prefix;
    int n_lt = 0, n_eq = 0;
    const int base_eq = k - need;         // ties are written after the strictly-smaller ones
    for (int j0 = 0; j0 < n_points; j0 += 32) {
        const int j = j0 + lane;
        const unsigned b = j < n_points ? d2_bits(q, pc, j) : 0xFFFFFFFFu;
        const bool lt = b < thr, eq = b == thr;
        const unsigned m_lt = __ballot_sync(0xffffffffu, lt), m_eq = __ballot_sync(0xffffffffu, eq);
        if (lt) out[n_lt + __popc(m_lt & lt_mask)] = j;
        const int e = n_eq + __popc(m_eq & lt_mask);
        if (eq && e < need) out[base_eq + e] = j;
        n_lt += __popc(m_lt);
        n_eq += __popc(m_eq);
    }
}

// One warp per query.  Fast path, two sweeps: (1) a 1024-bin histogram of the top bits of d^2 locates the bin that
// holds the k-th neighbour (19 % wide in d^2, so it holds a handful of points); (2) the collection sweep writes
// everything in lower bins and parks the boundary bin's points in shared memory, where the `need` smallest are
// picked by (d^2 bits, index) rank -- the same set, ties to the lowest indices, as the full radix select, which
// remains the fallback when the boundary bin overflows (e.g. many coincident points).
__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(const float* __restrict__ pc, int n_points, int k,
                                                              long long* __restrict__ out_idx) {
    extern __shared__ unsigned knn_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* hist = knn_smem + warp * (kKnnBins + 2 * kKnnCand);
    unsigned* cand_key = hist + kKnnBins;
    unsigned* cand_idx = cand_key + kKnnCand;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int qi = blockIdx.x * kKnnWarps + warp; qi < n_points; qi += gridDim.x * kKnnWarps) {
        const f3 q = ld3(pc, qi);
        long long* out = out_idx + (long long)qi * k;
        for (int i = lane; i < kKnnBins; i += 32) hist[i] = 0u;
        __syncwarp();
        for (int j = lane; j < n_points; j += 32) atomicAdd(&hist[d2_bits(q, pc, j) >> 21], 1u);
        __syncwarp();
        // bin of the k-th smallest: lane l owns bins [32 l, 32 l + 32)
        unsigned mine = 0;
        for (int i = 0; i < 32; ++i) mine += hist[lane * 32 + ((i + lane) & 31)];
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned excl = incl - mine;
        const bool here = (unsigned)k > excl && (unsigned)k <= incl;
        int bin = 0, below = 0, in_bin = 0;
        if (here) {
            unsigned run = excl;
            for (int i = 0; i < 32; ++i) {
                const unsigned c = hist[lane * 32 + i];
                if ((unsigned)k > run && (unsigned)k <= run + c) {
                    bin = lane * 32 + i;
                    below = (int)run;
                    in_bin = (int)c;
                }
                run += c;
            }
        }
        const unsigned src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
        bin = __shfl_sync(0xffffffffu, bin, src);
        below = __shfl_sync(0xffffffffu, below, src);
        in_bin = __shfl_sync(0xffffffffu, in_bin, src);
        __syncwarp();
        if (in_bin > kKnnCand) {                      // boundary bin too full for the candidate list
            knn_query_radix(pc, n_points, k, qi, hist, lane, out);
            __syncwarp();
            continue;
        }
        const int need = k - below;                   // how many of the boundary bin's points belong to the k nearest
        int n_lt = 0, n_c = 0;
        for (int j0 = 0; j0 < n_points; j0 += 32) {
            const int j = j0 + lane;
            const unsigned b = j < n_points ? d2_bits(q, pc, j) : 0xFFFFFFFFu;
            const unsigned bj = b >> 21;
            const bool lt = j < n_points && bj < (unsigned)bin, eq = j < n_points && bj == (unsigned)bin;
            const unsigned m_lt = __ballot_sync(0xffffffffu, lt), m_eq = __ballot_sync(0xffffffffu, eq);
            if (lt) out[n_lt + __popc(m_lt & lt_mask)] = j;
            if (eq) {
                const int e = n_c + __popc(m_eq & lt_mask);
                cand_key[e] = b;
                cand_idx[e] = (unsigned)j;
            }
            n_lt += __popc(m_lt);
            n_c += __popc(m_eq);
        }
        __syncwarp();
        // rank of every candidate by (key, index); the `need` smallest follow the strictly-lower bins
        for (int c0 = 0; c0 < n_c; c0 += 32) {
            const int c = c0 + lane;
            if (c < n_c) {
                const unsigned key = cand_key[c], idx = cand_idx[c];
                int rank = 0;
                for (int o = 0; o < n_c; ++o) {
                    const unsigned ko = cand_key[o], io = cand_idx[o];
                    rank += (ko < key || (ko == key && io < idx)) ? 1 : 0;
                }
                if (rank < need) out[below + rank] = (long long)idx;
            }
        }
        __syncwarp();
    }
}

}  // namespace pe
}  // namespace cppf

using namespace cppf;

extern "C" int cppf_pe_blob_floats(void) { return pe::kBlobFloats + pe::tcpe::kSmemFloats; }

extern "C" int cppf_knn(const float* pc, int n_points, int k, int64_t* out_idx, void* stream) {
    if (n_points <= 0) return 0;
    if (k <= 0 || k > n_points) return (int)cudaErrorInvalidValue;
    int blocks = (n_points + pe::kKnnWarps - 1) / pe::kKnnWarps;
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)pe::kKnnWarps * (pe::kKnnBins + 2 * pe::kKnnCand) * sizeof(unsigned);
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)pe::knn_kernel, (int)smem));
    pe::knn_kernel<<<blocks, pe::kKnnWarps * 32, smem, (cudaStream_t)stream>>>(pc, n_points, k,
                                                                                reinterpret_cast<long long*>(out_idx));
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_point_encode(const float* pc, const float* nrm, const int64_t* nbrs, const float* pe_blob, float* feat,
                                 float* glob_scratch, int n_points, int k, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_points <= 0) return 0;
    if (k <= 0 || k > 64) return (int)cudaErrorInvalidValue;
    if (!t_workspace_prepared) {                  // cppf_pose_fused: the geometry kernel wrote the -inf already
        pe::glob_init_kernel<<<1, 32, 0, stream>>>(glob_scratch);
        CPPF_LAUNCH_CHECK();
    }
    pe::Params prm{pc, nrm, reinterpret_cast<const long long*>(nbrs), pe_blob, feat, glob_scratch, n_points, k};
    // CPPF_PE_IMPL=simt selects the FFMA kernel (kept as the cross-check of the tensor-core one)
    const char* impl_env = getenv("CPPF_PE_IMPL");                // read per call so that a test can compare the two kernels
    const bool simt = impl_env != nullptr && impl_env[0] == 's';
    if (!simt) {
        const long long n_tiles = ((long long)n_points + 1) / 2;
        const long long per_round = (long long)sm_count() * pe::tcpe::kGroups;
        const long long rounds = (n_tiles + per_round - 1) / per_round;
        long long ctas = (n_tiles + rounds * pe::tcpe::kGroups - 1) / (rounds * pe::tcpe::kGroups);   // fewest CTAs for that many rounds
        if (ctas > sm_count()) ctas = sm_count();
        CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)pe::tcpe::point_encode_tc_kernel, pe::tcpe::kSmemBytes));
        pe::tcpe::point_encode_tc_kernel<<<(int)ctas, pe::tcpe::kThreads, pe::tcpe::kSmemBytes, stream>>>(prm);
        CPPF_LAUNCH_CHECK();
    } else {
        const size_t smem = sizeof(float) * ((size_t)pe::kBlobFloats + (size_t)pe::kWarps * pe::kWarpFloats);
        int blocks = (n_points + pe::kWarps - 1) / pe::kWarps;
        if (blocks > sm_count()) blocks = sm_count();
        CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)pe::point_encode_kernel, (int)smem));
        pe::point_encode_kernel<<<blocks, pe::kWarps * 32, smem, stream>>>(prm);
        CPPF_LAUNCH_CHECK();
    }
    pe::point_glob_kernel<<<(n_points * 8 + 255) / 256, 256, 0, stream>>>(feat, glob_scratch, n_points);
    CPPF_LAUNCH_CHECK();
    return 0;
}

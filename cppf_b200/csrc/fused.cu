// Fused pair encoder + categorical sampling ("mode F"): logits never leave the SM.
//
// For every pair: PPF tuple -> ResLayer stack (encode.cuh) -> the `final` layer evaluated
// head by head into a shared-memory tile -> softmax + inverse-CDF draw per head -> 4 bin
// bytes (mu, nu, up, right) + the 5 tail logits (aux up, aux right, 3 log-scales).
// Replaces models/model.py:117-137 + nocs/inference.py:183-188 and :238-256 in one pass:
// the reference's second encoder call (:236) re-evaluates the same pairs with the same
// inputs, so its heads are produced here once (redundancy ii).
//
// HBM traffic per pair: 8/16 B index read (0 when dense) + 4 B bins + 20 B tail written,
// against 564 B for materialised logits.
#include "encode.cuh"

#include "../../include/cppf_b200.h"

namespace cppf {

// head section layout (floats), all matrices k-major [16][cols], columns permuted for NO
constexpr int kTrBins = 32, kRotBins = 36;
constexpr int kHeadMu = 0;                         // W[16][32] + b[32]
constexpr int kHeadNu = kHeadMu + 16 * 32 + 32;
constexpr int kHeadUp = kHeadNu + 16 * 32 + 32;    // Wa[16][32] + Wb[16][8] + ba[32] + bb[8]
constexpr int kHeadRt = kHeadUp + 16 * 32 + 16 * 8 + 32 + 8;
constexpr int kHeadTail = kHeadRt + 16 * 32 + 16 * 8 + 32 + 8;   // W[16][8] + b[8]
constexpr int kHeadFloats = kHeadTail + 16 * 8 + 8;

constexpr int kFusedWarps = 16;
constexpr int kFusedWarpFloats = 2 * kActFloats + 16 * XS;

struct FusedParams {
    const float* pc;
    const float* nrm;
    const float* table;
    const float* blob;
    const float* hblob;
    const void* idx;
    const float* uniforms;      // optional [n_pairs][4]
    unsigned long long seed;
    uint8_t* bins;              // [n_pairs][4]
    float* tail;                // [5][n_pairs] (plane-major)
    int n_points;
    long long n_pairs;
    int heads;                  // bit0 mu/nu, bit1 up, bit2 right, bit3 tail
};

// One categorical draw from NB logits held column-major in shared memory (lane = pair):
// e_k = 2^((l_k - max) * log2 e); bin = #{k : cumsum(e)_k <= u * sum(e)}, clamped.
template <int NB>
__device__ __forceinline__ int sample_head(const float* __restrict__ LG, float u) {
    float e[NB];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        e[k] = LG[k * XS];
        m = fmaxf(m, e[k]);
    }
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        e[k] = exp2f((e[k] - m) * 1.4426950408889634f);
        tot += e[k];
    }
    const float t = u * tot;
    float acc = 0.f;
    int bin = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        acc += e[k];
        bin += acc <= t ? 1 : 0;
    }
    return bin < NB - 1 ? bin : NB - 1;
}

// 32-column head: LG[j][pair] = X3 . W[:, j] + b[j]
__device__ __forceinline__ void head32(const float* __restrict__ Xl, const float* __restrict__ W, float* __restrict__ LGl,
                                       int og) {
    float acc[8][4];
    zero_acc(acc);
    tile_gemm<16, 4>(Xl, W + og * 4, 32, acc);
    float bias[4];
    WVec<4>::load(W + 16 * 32 + og * 4, bias);
    tile_store<4>(LGl, og, acc, bias, nullptr, 0, 0);
}

// columns 32..39 of a 36(+4 pad)-column head: one column per output group
__device__ __forceinline__ void head8(const float* __restrict__ Xl, const float* __restrict__ W, const float* __restrict__ b,
                                      float (&acc)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(Xl + k * XS);
        const float4 a1 = *reinterpret_cast<const float4*>(Xl + k * XS + 4);
        const float w = W[k * 8];
        acc[0] = fmaf(a0.x, w, acc[0]); acc[1] = fmaf(a0.y, w, acc[1]);
        acc[2] = fmaf(a0.z, w, acc[2]); acc[3] = fmaf(a0.w, w, acc[3]);
        acc[4] = fmaf(a1.x, w, acc[4]); acc[5] = fmaf(a1.y, w, acc[5]);
        acc[6] = fmaf(a1.z, w, acc[6]); acc[7] = fmaf(a1.w, w, acc[7]);
    }
    const float bb = b[0];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += bb;
}

template <bool IDX64>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) encode_sample_kernel(const FusedParams prm) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                       // pair section up to (not including) WF
    float* sH = smem + kOffWF;              // head section
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + kOffWF + kHeadFloats + warp * kFusedWarpFloats;
    float* H = wbase;
    float* R = wbase + kActFloats;
    float* X3 = wbase + 2 * kActFloats;
    float* LG = wbase;                      // head logits tile [<=40][XS], aliases H|R after the chain

    {
        const float4* src = reinterpret_cast<const float4*>(prm.blob + kOffPair);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < kOffWF / 4; i += blockDim.x) dst[i] = __ldg(src + i);
        const float4* hsrc = reinterpret_cast<const float4*>(prm.hblob);
        float4* hdst = reinterpret_cast<float4*>(sH);
        for (int i = threadIdx.x; i < kHeadFloats / 4; i += blockDim.x) hdst[i] = __ldg(hsrc + i);
    }
    __syncthreads();

    const int og = lane & 7, pg = lane >> 3;
    const long long n_tiles = (prm.n_pairs + WP - 1) / WP;
    for (long long tile = (long long)blockIdx.x * kFusedWarps + warp; tile < n_tiles;
         tile += (long long)gridDim.x * kFusedWarps) {
        const long long p0 = tile * WP;
        const long long p = p0 + lane;
        const bool valid = p < prm.n_pairs;
        int a = 0, b = 0;
        if (valid) pair_ab<IDX64>(prm.idx, p, prm.n_points, a, b);
        float ppf[4];
        {
            const f3 pa = ld3(prm.pc, a), pb = ld3(prm.pc, b);
            const f3 na = ld3(prm.nrm, a), nb = ld3(prm.nrm, b);
            ppf_tuple(pa, pb, na, nb, -1.f, ppf);
        }
        float4 u4;
        if (prm.uniforms != nullptr) {
            u4 = valid ? __ldg(reinterpret_cast<const float4*>(prm.uniforms) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const uint4 w = philox4x32_10(make_uint4((uint32_t)p, (uint32_t)((unsigned long long)p >> 32), 0u, 0u),
                                          make_uint2((uint32_t)prm.seed, (uint32_t)(prm.seed >> 32)));
            u4 = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }
        __syncwarp();                       // previous tile's sampling reads of LG (= H|R) are done
        layer0_front(prm.table, sW + kOffWppf, a, b, ppf, H, R, lane);
        reslayers(sW, H, R, X3, lane);

        const float* Xl = X3 + pg * 8;
        float* LGl = LG + pg * 8;
        uchar4 bins = make_uchar4(0, 0, 0, 0);
        if (prm.heads & 1) {
            head32(Xl, sH + kHeadMu, LGl, og);
            __syncwarp();
            bins.x = (unsigned char)sample_head<kTrBins>(LG + lane, u4.x);
            __syncwarp();
            head32(Xl, sH + kHeadNu, LGl, og);
            __syncwarp();
            bins.y = (unsigned char)sample_head<kTrBins>(LG + lane, u4.y);
            __syncwarp();
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!(prm.heads & (2 << h))) continue;
            const float* Wh = sH + (h == 0 ? kHeadUp : kHeadRt);
            head32(Xl, Wh, LGl, og);        // bias of the 32-col part sits right after the [16][32] matrix
            float acc[8];
            head8(Xl, Wh + 16 * 32 + 32 + og, Wh + 16 * 32 + 32 + 16 * 8 + og, acc);
            if (og < kRotBins - 32) {
                *reinterpret_cast<float4*>(LGl + (32 + og) * XS) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                *reinterpret_cast<float4*>(LGl + (32 + og) * XS + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
            __syncwarp();
            const int bn = sample_head<kRotBins>(LG + lane, h == 0 ? u4.z : u4.w);
            if (h == 0) bins.z = (unsigned char)bn;
            else bins.w = (unsigned char)bn;
            __syncwarp();
        }
        if (valid) reinterpret_cast<uchar4*>(prm.bins)[p] = bins;
        if (prm.heads & 8) {
            float acc[8];
            head8(Xl, sH + kHeadTail + og, sH + kHeadTail + 16 * 8 + og, acc);
            if (og < 5) {
                float* dst = prm.tail + (long long)og * prm.n_pairs + p0 + pg * 8;
                if (p0 + WP <= prm.n_pairs && (prm.n_pairs & 3) == 0) {
                    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (p0 + pg * 8 + i < prm.n_pairs) dst[i] = acc[i];
                }
            }
        }
    }
}

}  // namespace cppf

using namespace cppf;

extern "C" int cppf_head_blob_floats(void) { return kHeadFloats; }

extern "C" int cppf_encode_sample(const float* pc, const float* nrm, const float* table, const float* blob,
                                  const float* head_blob, const void* idx, int idx_is_64, int n_points, int64_t n_pairs,
                                  const float* uniforms, uint64_t seed, int heads, uint8_t* bins, float* tail,
                                  void* stream) {
    if (n_pairs <= 0) return 0;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    if ((heads & 8) && tail == nullptr) return (int)cudaErrorInvalidValue;
    FusedParams prm{pc, nrm, table, blob, head_blob, idx, uniforms, seed, bins, tail, n_points, (long long)n_pairs, heads};
    const size_t smem = sizeof(float) * ((size_t)kOffWF + kHeadFloats + (size_t)kFusedWarps * kFusedWarpFloats);
    const long long n_tiles = (n_pairs + WP - 1) / WP;
    long long ctas = (n_tiles + kFusedWarps - 1) / kFusedWarps;
    if (ctas > sm_count()) ctas = sm_count();
    auto kern = idx_is_64 ? encode_sample_kernel<true> : encode_sample_kernel<false>;
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)kern, (int)smem));
    kern<<<(int)ctas, kFusedWarps * 32, smem, (cudaStream_t)stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}

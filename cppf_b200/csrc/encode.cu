// Materialised-logits pair encoder ("mode M"): the drop-in for
// PPFEncoder.forward_with_idx (reference models/model.py:117-137) and the dense
// branch of PPFEncoder.forward (:92-115).
//
// Persistent kernel, one 256-thread CTA per SM, each warp loops over 32-pair tiles:
//   gather (lane = pair) -> warp-tile ResLayer chain (encode.cuh) -> `final` layer into
//   a shared-memory staging tile laid out exactly like the output rows -> one TMA bulk
//   store (cp.async.bulk.global.shared::cta) of the contiguous 32 x out_dim block.
// Algorithmic HBM traffic: 8/16 B of indices in (0 in dense mode) + 4*out_dim B out
// per pair; the per-point inputs (pc, normals, 128-float pre-projected table) are
// L2-resident.
#include "encode.cuh"

#include "../../include/cppf_b200.h"

namespace cppf {

struct EncodeParams {
    const float* pc;
    const float* nrm;
    const float* table;
    const float* blob;
    const void* idx;
    const float* dist;
    float* out;
    int n_points;
    long long n_pairs;
    int out_dim;
    int col_begin;
    int col_count;
};

__device__ __forceinline__ void bulk_store_tile(float* gdst, const float* ssrc, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int kEncWarps = 8;

template <bool IDX64>
__global__ void __launch_bounds__(kEncWarps * 32, 1) ppf_encode_kernel(const EncodeParams prm) {
    extern __shared__ __align__(16) float smem[];
    const int outp = padded_out(prm.out_dim);
    const int wfloats = pair_section_floats(prm.out_dim);
    const int ncols = prm.col_count;
    const int stage_floats = WP * ncols;
    const int region = (stage_floats > 2 * kActFloats ? stage_floats : 2 * kActFloats);   // staging aliases H|R
    float* sW = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = smem + wfloats + warp * (region + 16 * XS);
    float* H = wbase;
    float* R = wbase + kActFloats;
    float* S = wbase;                       // staging [pair][ncols], valid only after the chain
    float* X3 = wbase + region;

    // stage the pair section of the weight blob (coalesced 128-bit copies)
    {
        const float4* src = reinterpret_cast<const float4*>(prm.blob + kOffPair);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = threadIdx.x; i < wfloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int og = lane & 7, pg = lane >> 3;
    const long long n_tiles = (prm.n_pairs + WP - 1) / WP;
    const int ch0 = prm.col_begin / kFinalChunk;
    const int ch1 = (prm.col_begin + ncols + kFinalChunk - 1) / kFinalChunk;
    bool store_pending = false;

    for (long long tile = (long long)blockIdx.x * kEncWarps + warp; tile < n_tiles;
         tile += (long long)gridDim.x * kEncWarps) {
        const long long p0 = tile * WP;
        const long long p = p0 + lane;
        const bool valid = p < prm.n_pairs;
        int a = 0, b = 0;
        if (valid) pair_ab<IDX64>(prm.idx, p, prm.n_points, a, b);
        float ppf[4];
        {
            const f3 pa = ld3(prm.pc, a), pb = ld3(prm.pc, b);
            const f3 na = ld3(prm.nrm, a), nb = ld3(prm.nrm, b);
            const float dab = (prm.dist != nullptr && valid) ? __ldg(prm.dist + (long long)a * prm.n_points + b) : -1.f;
            ppf_tuple(pa, pb, na, nb, dab, ppf);
        }
        if (store_pending) {               // staging of the previous tile aliases H|R
            if (lane == 0) bulk_wait_read();
            __syncwarp();
            store_pending = false;
        }
        layer0_front(prm.table, sW + kOffWppf, a, b, ppf, H, R, lane);
        reslayers(sW, H, R, X3, lane);

        // final: logits[:, col] = X3 . WF[:, col] + BF[col]      (models/model.py:137)
        const float* Xl = X3 + pg * 8;
        for (int ch = ch0; ch < ch1; ++ch) {
            float acc[8][6];
            zero_acc(acc);
            tile_gemm<16, 6>(Xl, sW + kOffWF + ch * kFinalChunk + og * 6, outp, acc);
            float bias[6];
            WVec<6>::load(sW + kOffWF + 16 * outp + ch * kFinalChunk + og * 6, bias);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int col = ch * kFinalChunk + og + 8 * c - prm.col_begin;
                if (col >= 0 && col < ncols) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) S[(pg * 8 + i) * ncols + col] = acc[i][c] + bias[c];
                }
            }
        }
        __syncwarp();
        float* gdst = prm.out + p0 * ncols;
        const long long rows = prm.n_pairs - p0 < WP ? prm.n_pairs - p0 : WP;
        if (rows == WP) {
            if (lane == 0) bulk_store_tile(gdst, S, (uint32_t)(WP * ncols * sizeof(float)));
            store_pending = true;
        } else {                            // ragged last tile: plain coalesced stores
            for (int i = lane; i < (int)rows * ncols; i += 32) gdst[i] = S[i];
            __syncwarp();
        }
    }
    if (lane == 0) bulk_wait_all();
}

// Per-point pre-projection of layer 0's feature columns: table[n][c] for c in
// [0,64) = PRE_WA^T feat[n] + PRE_BIAS, [64,128) = PRE_WB^T feat[n].
__global__ void __launch_bounds__(128) preproject_kernel(const float* __restrict__ feat, const float* __restrict__ blob,
                                                         float* __restrict__ table, int n_points) {
    __shared__ float f[kFeat];
    const int n = blockIdx.x;
    if (threadIdx.x < kFeat) f[threadIdx.x] = feat[(int64_t)n * kFeat + threadIdx.x];
    __syncthreads();
    const int c = threadIdx.x;          // 0..127
    const float* W = blob + (c < 64 ? kOffPreWA + c : kOffPreWB + (c - 64));
    float acc = c < 64 ? blob[kOffPreBias + c] : 0.f;
#pragma unroll 8
    for (int k = 0; k < kFeat; ++k) acc = fmaf(f[k], __ldg(W + k * 64), acc);
    table[(int64_t)n * kTable + c] = acc;
}

}  // namespace cppf

using namespace cppf;

extern "C" int cppf_ppf_blob_floats(int out_dim) { return blob_floats(out_dim); }
extern "C" int cppf_ppf_feat_dim(void) { return kFeat; }

extern "C" int cppf_ppf_preproject(const float* feat, const float* blob, float* table, int n_points, void* stream) {
    if (n_points <= 0) return 0;
    preproject_kernel<<<n_points, 128, 0, (cudaStream_t)stream>>>(feat, blob, table, n_points);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_ppf_encode(const float* pc, const float* nrm, const float* table, const float* blob,
                               const void* idx, int idx_is_64, const float* dist, float* out, int n_points,
                               int64_t n_pairs, int out_dim, int col_begin, int col_count, void* stream) {
    if (n_pairs <= 0) return 0;
    if (out_dim <= 0 || col_begin < 0 || col_count <= 0 || col_begin + col_count > out_dim)
        return (int)cudaErrorInvalidValue;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    EncodeParams prm{pc, nrm, table, blob, idx, dist, out, n_points, (long long)n_pairs, out_dim, col_begin, col_count};
    const int stage_floats = WP * col_count;
    const int region = stage_floats > 2 * kActFloats ? stage_floats : 2 * kActFloats;
    const size_t smem = sizeof(float) * ((size_t)pair_section_floats(out_dim) + (size_t)kEncWarps * (region + 16 * XS));
    if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
    const long long n_tiles = (n_pairs + WP - 1) / WP;
    long long ctas = (n_tiles + kEncWarps - 1) / kEncWarps;
    if (ctas > sm_count()) ctas = sm_count();
    auto kern = idx_is_64 ? ppf_encode_kernel<true> : ppf_encode_kernel<false>;
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)kern, (int)smem));
    kern<<<(int)ctas, kEncWarps * 32, smem, (cudaStream_t)stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}

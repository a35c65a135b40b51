// Warp-tile pair MLP: device pieces shared by the materialised-logits kernel
// (encode.cu) and the fused encode->sample->vote kernel (fused.cu).
//
// One warp owns a tile of WP = 32 point pairs end to end.  Activations live in the
// warp's private shared-memory buffers as [feature][pair] (row stride XS floats), so
// consecutive layers only need __syncwarp() -- no CTA barrier anywhere in the chain.
// Each layer is a [32 pairs x NOUT] x K register-tiled GEMM: lane (og = lane&7,
// pg = lane>>3) accumulates 8 pairs (8*pg .. 8*pg+7) x NO outputs (j = og + 8c) in
// registers; per k it issues two 128-bit activation loads (broadcast across the 8
// lanes of an output group) and one weight load (broadcast across the 4 pair groups),
// then 8*NO FFMAs.  Weight matrices are k-major with columns pre-permuted by the host
// packer (mlp_layout.h) so each lane's NO weights are contiguous.
//
// Reference semantics: models/model.py:26-31 (ResLayer), :124-137 (PPF tuple + stack).
#pragma once
#include "common.cuh"
#include "mlp_layout.h"

namespace cppf {

constexpr int WP = 32;        // pairs per warp tile
constexpr int XS = WP + 4;    // activation row stride: +4 floats keeps 128-bit row stores conflict-free
constexpr int kActFloats = 32 * XS;   // one [32][XS] activation buffer

template <int NO>
struct WVec;
template <>
struct WVec<4> {
    static __device__ __forceinline__ void load(const float* w, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(w);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <>
struct WVec<2> {
    static __device__ __forceinline__ void load(const float* w, float (&v)[2]) {
        const float2 t = *reinterpret_cast<const float2*>(w);
        v[0] = t.x; v[1] = t.y;
    }
};
template <>
struct WVec<6> {
    static __device__ __forceinline__ void load(const float* w, float (&v)[6]) {
        const float2 t0 = *reinterpret_cast<const float2*>(w);
        const float2 t1 = *reinterpret_cast<const float2*>(w + 2);
        const float2 t2 = *reinterpret_cast<const float2*>(w + 4);
        v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y; v[4] = t2.x; v[5] = t2.y;
    }
};

// acc[i][c] += sum_k X[k][i] * W[k][c]   (X already offset to the lane's 8 pairs,
// W already offset to the lane's NO columns; wstride = floats per weight row)
template <int K, int NO>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ X, const float* __restrict__ W, int wstride,
                                          float (&acc)[8][NO]) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(X + k * XS);
        const float4 a1 = *reinterpret_cast<const float4*>(X + k * XS + 4);
        float w[NO];
        WVec<NO>::load(W + k * wstride, w);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < NO; ++c) acc[i][c] = fmaf(a[i], w[c], acc[i][c]);
    }
}

template <int NO>
__device__ __forceinline__ void zero_acc(float (&acc)[8][NO]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < NO; ++c) acc[i][c] = 0.f;
}

// out[j][8pg..8pg+7] = acc[.][c] (+ bias[c]) (+ residual row) with optional ReLU on the
// first `relu_cols` of the lane's NO columns.  `O`/`Rsd` are offset to the lane's pairs.
template <int NO>
__device__ __forceinline__ void tile_store(float* __restrict__ O, int og, const float (&acc)[8][NO],
                                           const float* __restrict__ bias, const float* __restrict__ Rsd,
                                           int rsd_row0, int relu_cols) {
#pragma unroll
    for (int c = 0; c < NO; ++c) {
        const int j = og + 8 * c;
        float v[8];
        const float bj = bias ? bias[c] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[i][c] + bj;
        if (Rsd) {
            const float4 r0 = *reinterpret_cast<const float4*>(Rsd + (rsd_row0 + j) * XS);
            const float4 r1 = *reinterpret_cast<const float4*>(Rsd + (rsd_row0 + j) * XS + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (c < relu_cols) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        *reinterpret_cast<float4*>(O + j * XS) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(O + j * XS + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// The PPF tuple of one pair -- models/model.py:120-129 (indexed) / :92-105 (dense).
// If dist_ab >= 0 it is the caller-supplied distance of the dense signature.
__device__ __forceinline__ void ppf_tuple(f3 pa, f3 pb, f3 na, f3 nb, float dist_ab, float (&ppf)[4]) {
    const f3 d = pa - pb;
    const float dn = dist_ab >= 0.f ? dist_ab : sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    const float inv = dn + 1e-7f;
    const f3 dh = {d.x / inv, d.y / inv, d.z / inv};
    ppf[0] = na.x * dh.x + na.y * dh.y + na.z * dh.z;
    ppf[1] = nb.x * dh.x + nb.y * dh.y + nb.z * dh.z;
    ppf[2] = na.x * nb.x + na.y * nb.y + na.z * nb.z;
    ppf[3] = dn;
}

// Layer-0 front end, one lane per pair: gathers the pre-projected rows of a and b,
// adds the PPF columns, and writes  H = relu(fc1_0(x))  and  R = fc0_0(x) + fc2_0.b
// (models/model.py:27-28 with the feature columns pre-multiplied per point).
__device__ __forceinline__ void layer0_front(const float* __restrict__ table, const float* __restrict__ sWppf,
                                             int a, int b, const float (&ppf)[4], float* __restrict__ H,
                                             float* __restrict__ R, int lane) {
    const float4* TA = reinterpret_cast<const float4*>(table + (int64_t)a * kTable);
    const float4* TB = reinterpret_cast<const float4*>(table + (int64_t)b * kTable + 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 ta = __ldg(TA + q), tb = __ldg(TB + q);
        float v[4] = {ta.x + tb.x, ta.y + tb.y, ta.z + tb.z, ta.w + tb.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float4 w = *reinterpret_cast<const float4*>(sWppf + t * 64 + q * 4);
            v[0] = fmaf(ppf[t], w.x, v[0]);
            v[1] = fmaf(ppf[t], w.y, v[1]);
            v[2] = fmaf(ppf[t], w.z, v[2]);
            v[3] = fmaf(ppf[t], w.w, v[3]);
        }
        if (q < 8) {
#pragma unroll
            for (int e = 0; e < 4; ++e) H[(q * 4 + e) * XS + lane] = fmaxf(v[e], 0.f);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) R[((q - 8) * 4 + e) * XS + lane] = v[e];
        }
    }
}

// ResLayer stack after the front end.  In: H = relu(fc1_0), R = fc0_0 branch.
// Out: X3[16][XS] = output of the third ResLayer (input of `final`).
// sW = shared-memory copy of the blob's pair section.
__device__ __forceinline__ void reslayers(const float* __restrict__ sW, float* __restrict__ H,
                                          float* __restrict__ R, float* __restrict__ X3, int lane) {
    const int og = lane & 7, pb = (lane >> 3) * 8;
    float* Hl = H + pb;
    float* Rl = R + pb;
    float* Xl = X3 + pb;
    {   // x1 = fc2_0(h) + r0           (models/model.py:30-31, layer 0)
        float acc[8][4];
        zero_acc(acc);
        __syncwarp();
        tile_gemm<32, 4>(Hl, sW + kOffW2_0 + og * 4, 32, acc);
        tile_store<4>(Rl, og, acc, nullptr, Rl, 0, 0);
    }
    {   // u = relu(fc1_1(x1))
        float acc[8][4];
        zero_acc(acc);
        __syncwarp();
        tile_gemm<32, 4>(Rl, sW + kOffW1_1 + og * 4, 32, acc);
        float bias[4];
        WVec<4>::load(sW + kOffB1_1 + og * 4, bias);
        tile_store<4>(Hl, og, acc, bias, nullptr, 0, 4);
    }
    {   // x2 = fc2_1(u) + x1            (identity skip: dim_in == dim_out, :23-25)
        float acc[8][4];
        zero_acc(acc);
        __syncwarp();
        tile_gemm<32, 4>(Hl, sW + kOffW2_1 + og * 4, 32, acc);
        float bias[4];
        WVec<4>::load(sW + kOffB2_1 + og * 4, bias);
        tile_store<4>(Rl, og, acc, bias, Rl, 0, 0);
    }
    {   // [u ; r] = [relu(fc1_2(x2)) ; fc0_2(x2) + fc2_2.b]   -> H rows 0:16 / 16:32
        float acc[8][4];
        zero_acc(acc);
        __syncwarp();
        tile_gemm<32, 4>(Rl, sW + kOffW10_2 + og * 4, 32, acc);
        float bias[4];
        WVec<4>::load(sW + kOffB10_2 + og * 4, bias);
        tile_store<4>(Hl, og, acc, bias, nullptr, 0, 2);     // columns c=0,1 are j<16 -> ReLU
    }
    {   // x3 = fc2_2(u) + r
        float acc[8][2];
        zero_acc(acc);
        __syncwarp();
        tile_gemm<16, 2>(Hl, sW + kOffW2_2 + og * 2, 16, acc);
        tile_store<2>(Xl, og, acc, nullptr, Hl, 16, 0);
    }
    __syncwarp();
}

}  // namespace cppf

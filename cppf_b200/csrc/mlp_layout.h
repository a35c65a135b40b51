// Packed weight blob of the pair MLP (PPFEncoder, reference models/model.py:80-137,
// architecture ppffcs = [2*F+4, 32, 32, 16], nocs/inference.py:83).
//
// All matrices are stored K-MAJOR ([k][column]) so that a layer's weights for one
// input feature are contiguous, and each layer's output columns are PERMUTED for the
// warp tile of csrc/encode.cuh: a thread with output-group `og` (0..7) owns the NO
// outputs j = og + 8*c (c = 0..NO-1) and finds them contiguous at column og*NO + c.
//
//   section      rows x cols   contents (nn.Linear.weight is [out][in])
//   PRE_WA       40 x 64       col j<32: fc1_0.W[j][k]      ; col 32+j: fc0_0.W[j][k]        (k = feat[a] column)
//   PRE_WB       40 x 64       col j<32: fc1_0.W[j][40+k]   ; col 32+j: fc0_0.W[j][40+k]     (k = feat[b] column)
//   PRE_BIAS     64            [fc1_0.b ; fc0_0.b + fc2_0.b]
//   W_PPF        4 x 64        col j<32: fc1_0.W[j][80+q]   ; col 32+j: fc0_0.W[j][80+q]     (natural column order)
//   W2_0         32 x 32       fc2_0.W^T, permuted columns (NO = 4)
//   W1_1         32 x 32       fc1_1.W^T, permuted ; B1_1 32 permuted
//   W2_1         32 x 32       fc2_1.W^T, permuted ; B2_1 32 permuted
//   W10_2        32 x 32       logical cols 0:16 = fc1_2.W^T, 16:32 = fc0_2.W^T, permuted (NO = 4)
//   B10_2        32            [fc1_2.b ; fc0_2.b + fc2_2.b] permuted
//   W2_2         16 x 16       fc2_2.W^T, permuted (NO = 2)
//   WF           16 x OUTP     final.W^T, zero padded to OUTP = ceil(out_dim/48)*48, permuted per 48-chunk (NO = 6)
//   BF           OUTP          final.b, same permutation
//
// The Python packer (cppf_b200/model.py: pack_ppf_weights) mirrors these offsets.
#pragma once

namespace cppf {

constexpr int kFeat = 40;            // per-point feature width (32 conv + 8 global)
constexpr int kH1 = 32, kH2 = 32, kH3 = 16;
constexpr int kTable = 128;          // pre-projected table row: [A_fc1(32) A_fc0(32) B_fc1(32) B_fc0(32)]
constexpr int kFinalChunk = 48;      // final-layer columns per pass (8 output groups x 6)

constexpr int kOffPreWA = 0;
constexpr int kOffPreWB = kOffPreWA + kFeat * 64;
constexpr int kOffPreBias = kOffPreWB + kFeat * 64;
constexpr int kOffPair = kOffPreBias + 64;          // everything from here is staged in shared memory
// offsets relative to kOffPair
constexpr int kOffWppf = 0;                          // 4 x 64
constexpr int kOffW2_0 = kOffWppf + 4 * 64;         // 32 x 32
constexpr int kOffW1_1 = kOffW2_0 + 32 * 32;
constexpr int kOffB1_1 = kOffW1_1 + 32 * 32;
constexpr int kOffW2_1 = kOffB1_1 + 32;
constexpr int kOffB2_1 = kOffW2_1 + 32 * 32;
constexpr int kOffW10_2 = kOffB2_1 + 32;
constexpr int kOffB10_2 = kOffW10_2 + 32 * 32;
constexpr int kOffW2_2 = kOffB10_2 + 32;            // 16 x 16
constexpr int kOffWF = kOffW2_2 + 16 * 16;          // 16 x OUTP, then BF (OUTP)

__host__ __device__ constexpr int padded_out(int out_dim) {
    return (out_dim + kFinalChunk - 1) / kFinalChunk * kFinalChunk;
}
__host__ __device__ constexpr int pair_section_floats(int out_dim) {
    return kOffWF + 17 * padded_out(out_dim);
}
__host__ __device__ constexpr int blob_floats(int out_dim) {
    return kOffPair + pair_section_floats(out_dim);
}

}  // namespace cppf

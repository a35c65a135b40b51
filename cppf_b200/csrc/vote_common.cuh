// Helpers shared by the shared-memory voting kernels (vote_private.cu, vote_routed.cu).
#pragma once
#include "common.cuh"

#include <math.h>

namespace cppf {

constexpr int kMaxRotsP = 72;
constexpr int kRotTabP = kMaxRotsP * (kMaxRotsP + 1) / 2;
const float2* rot_table_device(cudaStream_t stream, int* err);   // vote.cu: (cos, sin) of angle(i, n), row n at n(n-1)/2

// x-slabs of the routed vote (vote_routed.cu): as many planes per slab as fit `cap` u32 cells, minus the overlap
// plane of the x+1 corners.  Shared by the host plan and the device-side geometry kernel (pose.cu).
constexpr int kMaxSlabs = 8;
// Strides of a slab in shared memory.  The slab is private to its CTA, so its rows and planes can be padded: with the
// natural strides of a 64^3 grid (gz = 64, gy * gz = 4096, both multiples of the 32 banks) the bank of a cell is fz mod 32
// alone, the candidates a warp splats together are close in z, and an ATOMS took 5.3 wavefronts instead of the 3.6 of
// random banks.  Odd strides make every unit step in x or y move the bank by an odd amount.
__host__ __device__ inline int slab_row_stride(int gz) { return gz | 1; }
__host__ __device__ inline int slab_plane_stride(int gy, int gz) { return (gy * slab_row_stride(gz)) | 1; }
__host__ __device__ inline bool routed_plan_hd(int gx, int gy, int gz, long long cap, int* pps, int* n_slabs) {
    const long long gyz = (long long)slab_plane_stride(gy, gz);
    const long long planes = cap / gyz - 1;
    if (planes < 1 || gx > 1024) return false;
    *pps = (int)(planes < gx ? planes : gx);
    *n_slabs = (gx + *pps - 1) / *pps;
    return *n_slabs <= kMaxSlabs;
}

constexpr int kFixShift = 14;                      // fixed-point fraction bits of a vote weight
constexpr float kFixScale = 16384.f;
constexpr unsigned kFixBudget = 0xFFFFFFFFu >> kFixShift;   // whole votes a u32 cell can absorb between flushes

static inline float float_ceil_p(double d) {
    float f = (float)d;
    if ((double)f < d) f = nextafterf(f, INFINITY);
    return f;
}

// a / b with b fixed: q0 = a*y, r = a - b*q0 (exact in an FMA), q = q0 + r*y.  With y the correctly
// rounded reciprocal of b this returns the correctly rounded quotient (Markstein), i.e. the same
// bits as the reference's `/ res`, in 3 instructions instead of the ~8 of an IEEE divide.
__device__ __forceinline__ float div_by(float a, float b, float y) {
    const float q0 = a * y;
    const float r = fmaf(-b, q0, a);
    return fmaf(r, y, q0);
}

// trilinear splat of one in-bounds candidate at grid coordinates g -- models/voting.py:40-63 with
// prob == 1 (nocs/inference.py:201).  Each corner weight is rounded ONCE to 2^-14 units and leaves the multiplier
// already AS AN INTEGER: the x-y factor carries 2^-60 and the z factor 2^-75 (exact power-of-two scalings of normal
// floats), so the last product w_xy * w_z * 2^(14 - 149) lands in the denormal range, where an FMUL rounds the exact
// product to a multiple of 2^-149 (round-half-even) and the bit pattern of the result IS that multiple.  No F2I, no
// magic-number FFMA + subtraction: one FMUL per corner (fp32 denormals run at full rate in the FMA pipe; this
// translation unit must not be built with -ftz=true).  Bit-identical to round-half-even(w_xy * w_z * 2^14).
#ifndef CPPF_SPLAT_DENORM
#define CPPF_SPLAT_DENORM 1
#endif
#ifndef CPPF_SPLAT_I2F
#define CPPF_SPLAT_I2F 1
#endif
__device__ __forceinline__ void red_shared_u32(unsigned saddr, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// grid_saddr: 32-bit shared-window address of the grid (cvta once per kernel, not per splat)
__device__ __forceinline__ void splat_fixed(unsigned grid_saddr, float gxf, float gyf, float gzf, int gyz, int gz) {
    const int fx = (int)gxf, fy = (int)gyf, fz = (int)gzf;                     // :40
#if CPPF_SPLAT_I2F
    // in-bounds coordinates are >= 0.01, so truncation is the floor of :42 and the integer converts back exactly: one
    // conversion on the XU pipe per axis (F2I) instead of two (F2I + FRND)
    const float rx = gxf - (float)fx, ry = gyf - (float)fy, rz = gzf - (float)fz;
#else
    const float rx = gxf - floorf(gxf), ry = gyf - floorf(gyf), rz = gzf - floorf(gzf);
#endif
    const unsigned cell = grid_saddr + 4u * (unsigned)(fx * gyz + fy * gz + fz);
    const unsigned sy = 4u * (unsigned)gz, sx = 4u * (unsigned)gyz;
#if CPPF_SPLAT_DENORM
    constexpr float kSx = 8.67361737988403547e-19f;                            // 2^-60
    constexpr float kSz = 2.64697796016968855e-23f;                            // 2^-75
    const float wx0 = fmaf(-rx, kSx, kSx), wx1 = rx * kSx;                      // (1 - rx) 2^-60, rx 2^-60: exact
    const float wy0 = 1.f - ry;
    const float z0 = fmaf(-rz, kSz, kSz), z1 = rz * kSz;                        // (1 - rz) 2^-75, rz 2^-75: exact
    const float w00 = wx0 * wy0, w01 = wx0 * ry, w10 = wx1 * wy0, w11 = wx1 * ry;
    red_shared_u32(cell, __float_as_uint(w00 * z0));
    red_shared_u32(cell + 4u, __float_as_uint(w00 * z1));
    red_shared_u32(cell + sy, __float_as_uint(w01 * z0));
    red_shared_u32(cell + sy + 4u, __float_as_uint(w01 * z1));
    red_shared_u32(cell + sx, __float_as_uint(w10 * z0));
    red_shared_u32(cell + sx + 4u, __float_as_uint(w10 * z1));
    red_shared_u32(cell + sx + sy, __float_as_uint(w11 * z0));
    red_shared_u32(cell + sx + sy + 4u, __float_as_uint(w11 * z1));
#else
    const float wx0 = 1.f - rx, wy0 = 1.f - ry;
    const float z1 = rz * kFixScale, z0 = (1.f - rz) * kFixScale;
    const float w00 = wx0 * wy0, w01 = wx0 * ry, w10 = rx * wy0, w11 = rx * ry;
    constexpr float kMagic = 8388608.f;                                        // 2^23
    constexpr unsigned kMagicBits = 0x4B000000u;
    red_shared_u32(cell, __float_as_uint(fmaf(w00, z0, kMagic)) - kMagicBits);
    red_shared_u32(cell + 4u, __float_as_uint(fmaf(w00, z1, kMagic)) - kMagicBits);
    red_shared_u32(cell + sy, __float_as_uint(fmaf(w01, z0, kMagic)) - kMagicBits);
    red_shared_u32(cell + sy + 4u, __float_as_uint(fmaf(w01, z1, kMagic)) - kMagicBits);
    red_shared_u32(cell + sx, __float_as_uint(fmaf(w10, z0, kMagic)) - kMagicBits);
    red_shared_u32(cell + sx + 4u, __float_as_uint(fmaf(w10, z1, kMagic)) - kMagicBits);
    red_shared_u32(cell + sx + sy, __float_as_uint(fmaf(w11, z0, kMagic)) - kMagicBits);
    red_shared_u32(cell + sx + sy + 4u, __float_as_uint(fmaf(w11, z1, kMagic)) - kMagicBits);
#endif
}

__device__ __forceinline__ void st_shared_f4(unsigned addr, float x, float y, float z) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

}  // namespace cppf

// Voting kernels: centre voting, grid argmax, back-vote filter, pair compaction,
// orientation candidates, sphere-bin count, peak finder, categorical sampling.
// Reference semantics: models/voting.py (CUDA-C strings) and the host glue of
// nocs/inference.py:183-284; citations at each kernel.
#include "common.cuh"

#include "../../include/cppf_b200.h"

#include <math.h>

#include <mutex>

namespace cppf {

// ---------------------------------------------------------------------------
// (cos, sin) of every angle the reference kernels can form: angle(i, n) =
// float(double(2i) * pi / double(n)), 0 <= i < n <= kMaxRots.  Row n starts at
// n(n-1)/2.  Filled once per device with the accurate cosf/sinf the reference
// strings call (models/voting.py:33-34), so looking a value up is bit-identical to
// recomputing it in the loop.
constexpr int kMaxRots = 72;
constexpr int kRotTableSize = kMaxRots * (kMaxRots + 1) / 2;
__device__ float2 g_rot_table[kRotTableSize];

__global__ void rot_table_init_kernel() {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kRotTableSize) return;
    int n = 1;
    while ((n + 1) * n / 2 <= t) ++n;          // row n holds entries [n(n-1)/2, n(n+1)/2)
    const int i = t - n * (n - 1) / 2;
    const float ang = rot_angle(i, n);
    g_rot_table[t] = make_float2(cosf(ang), sinf(ang));
}

// Once per device.  The first call waits for the table (unless the stream is being captured, in which case the
// initialisation is simply repeated by later calls), so a concurrent first call on another stream cannot run ahead of it.
static int ensure_rot_table(cudaStream_t stream) {
    static std::mutex mu;
    static bool ready[64] = {false};
    int dev = 0;
    CPPF_RETURN_IF(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && ready[dev]) return 0;
    rot_table_init_kernel<<<(kRotTableSize + 255) / 256, 256, 0, stream>>>();
    CPPF_LAUNCH_CHECK();
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CPPF_RETURN_IF(cudaStreamIsCapturing(stream, &cap));
    if (cap == cudaStreamCaptureStatusNone) {
        CPPF_RETURN_IF(cudaStreamSynchronize(stream));
        if (dev < 64) ready[dev] = true;
    }
    return 0;
}

// For kernels in other translation units (no relocatable device code): device address of the
// initialised table, or nullptr with *err set.
const float2* rot_table_device(cudaStream_t stream, int* err) {
    *err = ensure_rot_table(stream);
    if (*err) return nullptr;
    void* ptr = nullptr;
    const cudaError_t e = cudaGetSymbolAddress(&ptr, g_rot_table);
    if (e != cudaSuccess) {
        *err = (int)e;
        return nullptr;
    }
    return reinterpret_cast<const float2*>(ptr);
}

// Smallest float f with (double)f >= d: turns the reference's double-precision bound
// tests `(double)g < d` / `(double)g >= d` into exact float compares `g < f` / `g >= f`.
static float float_ceil(double d) {
    float f = (float)d;
    if ((double)f < d) f = nextafterf(f, INFINITY);
    return f;
}

struct VoteParams {
    const float* points;
    const float* mu_nu;
    const float* probs;
    const void* idx;
    float* grid;
    const float* corner;
    float res;
    float lo;               // float_ceil(0.01)
    float hx, hy, hz;       // float_ceil(dim - 1.01)
    int n_points;
    long long n_pairs;
    int n_rots, gx, gy, gz, adaptive;
};

// ---------------------------------------------------------------------------
// Centre voting -- models/voting.py:8-66.  One lane per pair; the circle of candidate
// centres is walked with table look-ups instead of per-iteration double-precision angle
// arithmetic + cos/sin; in-bounds candidates are splatted with 8 fire-and-forget
// fp32 reductions (RED.ADD.F32) exactly like the reference's atomicAdd.
template <bool IDX64, bool TABLE>
__global__ void __launch_bounds__(256) ppf_vote_kernel(const VoteParams prm) {
    __shared__ float2 s_tab[TABLE ? kRotTableSize : 1];
    if (TABLE) {
        for (int i = threadIdx.x; i < kRotTableSize; i += blockDim.x) s_tab[i] = g_rot_table[i];
        __syncthreads();
    }
    const int gyz = prm.gy * prm.gz;
    const float cx = __ldg(prm.corner), cy = __ldg(prm.corner + 1), cz = __ldg(prm.corner + 2);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < prm.n_pairs;
         p += (long long)gridDim.x * blockDim.x) {
        int ia, ib;
        pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib);
        const float2 mn = __ldg(reinterpret_cast<const float2*>(prm.mu_nu) + p);
        const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
        f3 ab, ex;
        if (!pair_frame(a, b, ab, ex)) continue;                           // :21
        const f3 c = foot_point(a, ab, mn.x);                                        // :23
        const float prob = prm.probs ? fmaxf(__ldg(prm.probs + ia), __ldg(prm.probs + ib)) : 1.f;   // :25
        const f3 x = scale3(ex, mn.y);                                            // :28
        const f3 y = cross3(x, ab);                                        // :29
        int n = prm.n_rots;
        if (prm.adaptive) n = adaptive_rots(mn.y, prm.res, prm.n_rots);    // :31
        const float2* tab = TABLE ? s_tab + (n > 0 ? n * (n - 1) / 2 : 0) : nullptr;
        for (int i = 0; i < n; ++i) {
            float ca, sa;
            if (TABLE) {
                const float2 cs = tab[i];
                ca = cs.x;
                sa = cs.y;
            } else {
                const float ang = rot_angle(i, n);                         // :33
                ca = cosf(ang);
                sa = sinf(ang);
            }
            const f3 off = circle_offset(x, y, ca, sa);                                // :34
            const f3 g = {(c.x + off.x - cx) / prm.res, (c.y + off.y - cy) / prm.res,
                          (c.z + off.z - cz) / prm.res};                   // :35
            if (g.x < prm.lo || g.y < prm.lo || g.z < prm.lo || g.x >= prm.hx || g.y >= prm.hy || g.z >= prm.hz)
                continue;                                                  // :36-39
            const int fx = (int)g.x, fy = (int)g.y, fz = (int)g.z;         // :40
            const float rx = g.x - floorf(g.x), ry = g.y - floorf(g.y), rz = g.z - floorf(g.z);   // :42
            const float wx0 = 1.f - rx, wy0 = 1.f - ry, wz0 = 1.f - rz;
            float* cell = prm.grid + (long long)fx * gyz + fy * prm.gz + fz;
            atomicAdd(cell, wx0 * wy0 * wz0 * prob);                       // :56-63
            atomicAdd(cell + 1, wx0 * wy0 * rz * prob);
            atomicAdd(cell + prm.gz, wx0 * ry * wz0 * prob);
            atomicAdd(cell + prm.gz + 1, wx0 * ry * rz * prob);
            atomicAdd(cell + gyz, rx * wy0 * wz0 * prob);
            atomicAdd(cell + gyz + 1, rx * wy0 * rz * prob);
            atomicAdd(cell + gyz + prm.gz, rx * ry * wz0 * prob);
            atomicAdd(cell + gyz + prm.gz + 1, rx * ry * rz * prob);
        }
    }
}

// ---------------------------------------------------------------------------
// Grid argmax -- nocs/inference.py:207-208 (np.argmax: first maximum in C order).
// key = (order-preserving float bits << 32) | (0xFFFFFFFF - index): atomicMax picks the
// largest value and, among equals, the lowest index.
__device__ __forceinline__ unsigned long long argmax_key(float v, uint32_t i) {
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - i);
}

__global__ void __launch_bounds__(256) grid_argmax_kernel(const float* __restrict__ grid, long long n,
                                                          unsigned long long* __restrict__ key_out,
                                                          const int* __restrict__ n_dev) {
    if (n_dev != nullptr) n = *n_dev < n ? *n_dev : n;      // cell count derived on the device (cppf_pose_fused)
    unsigned long long best = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = argmax_key(__ldg(grid + i), (uint32_t)i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    __shared__ unsigned long long s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = s[w] > best ? s[w] : best;
        atomicMax(key_out, best);
    }
}

__global__ void grid_argmax_finish_kernel(long long* out_index, float* out_value) {
    const unsigned long long k = *reinterpret_cast<unsigned long long*>(out_index);
    uint32_t u = (uint32_t)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    *out_index = (long long)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
    if (out_value) *out_value = __uint_as_float(u);
}

// ---------------------------------------------------------------------------
// Back-vote filter -- models/voting.py:74-112.
struct BackvoteParams {
    const float* points;
    const float* mu_nu;
    float* out_offsets;
    uint8_t* out_mask;
    const void* idx;
    const float* corner;
    const float* centre;
    float res, tol;
    float hx, hy, hz;       // float(dim - 1): the reference compares against int expressions (:104-107)
    int n_points;
    long long n_pairs;
    int n_rots;
};

template <bool IDX64, bool TABLE>
__global__ void __launch_bounds__(256) backvote_kernel(const BackvoteParams prm) {
    __shared__ float2 s_tab[TABLE ? kRotTableSize : 1];
    if (TABLE) {
        for (int i = threadIdx.x; i < kRotTableSize; i += blockDim.x) s_tab[i] = g_rot_table[i];
        __syncthreads();
    }
    const float cx = __ldg(prm.corner), cy = __ldg(prm.corner + 1), cz = __ldg(prm.corner + 2);
    const float tx = __ldg(prm.centre), ty = __ldg(prm.centre + 1), tz = __ldg(prm.centre + 2);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < prm.n_pairs;
         p += (long long)gridDim.x * blockDim.x) {
        int ia, ib;
        pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib);
        const float2 mn = __ldg(reinterpret_cast<const float2*>(prm.mu_nu) + p);
        const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
        f3 ab, ex;
        bool hit = false;
        f3 res_off = {0.f, 0.f, 0.f};
        const bool live = pair_frame(a, b, ab, ex);                        // :87 degenerate rows are not written
        if (live) {
            const f3 c = foot_point(a, ab, mn.x);
            const f3 x = scale3(ex, mn.y);
            const f3 y = cross3(x, ab);
            const int n = adaptive_rots(mn.y, prm.res, prm.n_rots);        // :97 always adaptive
            const float2* tab = TABLE ? s_tab + (n > 0 ? n * (n - 1) / 2 : 0) : nullptr;
            for (int i = 0; i < n; ++i) {
                float ca, sa;
                if (TABLE) {
                    const float2 cs = tab[i];
                    ca = cs.x;
                    sa = cs.y;
                } else {
                    const float ang = rot_angle(i, n);
                    ca = cosf(ang);
                    sa = sinf(ang);
                }
                const f3 off = circle_offset(x, y, ca, sa);
                const f3 pc = c + off;                                     // :101
                const f3 dlt = {pc.x - tx, pc.y - ty, pc.z - tz};
                if (len3(dlt) > prm.tol) continue;                         // :102
                const f3 g = {(pc.x - cx) / prm.res, (pc.y - cy) / prm.res, (pc.z - cz) / prm.res};
                if (g.x < 0.f || g.y < 0.f || g.z < 0.f || g.x >= prm.hx || g.y >= prm.hy || g.z >= prm.hz) continue;
                res_off = {-off.x, -off.y, -off.z};                        // :108
                hit = true;
                break;
            }
            if (prm.out_offsets) {
                prm.out_offsets[3 * p] = res_off.x;
                prm.out_offsets[3 * p + 1] = res_off.y;
                prm.out_offsets[3 * p + 2] = res_off.z;
            }
        }
        if (prm.out_mask)      // nocs/inference.py:230  np.any(oc != 0, -1)
            prm.out_mask[p] = (hit && (res_off.x != 0.f || res_off.y != 0.f || res_off.z != 0.f)) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// Order-preserving compaction of surviving pairs -- nocs/inference.py:230-231.
constexpr int kCompactBlock = 256;
constexpr int kCompactItems = 8;      // 2048 pairs per block

__global__ void __launch_bounds__(kCompactBlock) compact_count_kernel(const uint8_t* __restrict__ mask, long long n,
                                                                      int* __restrict__ block_counts) {
    const long long base = (long long)blockIdx.x * kCompactBlock * kCompactItems;
    int cnt = 0;
    for (int k = 0; k < kCompactItems; ++k) {
        const long long i = base + (long long)k * kCompactBlock + threadIdx.x;
        if (i < n && mask[i]) ++cnt;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ int s[kCompactBlock / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kCompactBlock / 32; ++w) t += s[w];
        block_counts[blockIdx.x] = t;
    }
}

// single-block exclusive scan of the per-block counts (64-bit running total)
__global__ void __launch_bounds__(1024) compact_scan_kernel(const int* __restrict__ block_counts, int n_blocks,
                                                            long long* __restrict__ block_offsets,
                                                            long long* __restrict__ total) {
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const long long v = i < n_blocks ? block_counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const long long warp_excl = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const long long carry = s_carry;
        if (i < n_blocks) block_offsets[i] = carry + warp_excl + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

template <bool IDX64>
__global__ void __launch_bounds__(kCompactBlock) compact_scatter_kernel(const uint8_t* __restrict__ mask,
                                                                        const void* __restrict__ idx, int n_points,
                                                                        long long n,
                                                                        const long long* __restrict__ block_offsets,
                                                                        int* __restrict__ out_idx,
                                                                        long long* __restrict__ out_pos) {
    // items are taken in contiguous runs per thread so that output order == input order
    const long long base = (long long)blockIdx.x * kCompactBlock * kCompactItems + (long long)threadIdx.x * kCompactItems;
    uint32_t flags = 0;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k) {
        const long long i = base + k;
        if (i < n && mask[i]) {
            flags |= 1u << k;
            ++cnt;
        }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ int s_warp[kCompactBlock / 32];
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    int warp_excl = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_excl += s_warp[w];
    long long o = block_offsets[blockIdx.x] + warp_excl + incl - cnt;
#pragma unroll
    for (int k = 0; k < kCompactItems; ++k) {
        if (flags & (1u << k)) {
            int a = 0, b = 0;
            if (out_idx) pair_ab<IDX64>(idx, base + k, n_points, a, b);
            if (out_idx) reinterpret_cast<int2*>(out_idx)[o] = make_int2(a, b);
            if (out_pos) out_pos[o] = base + k;
            ++o;
        }
    }
}

// ---------------------------------------------------------------------------
// Orientation candidates -- models/voting.py:119-147.  A 128-thread block takes 16
// pairs: 16 lanes build the pair frames into shared memory, then all threads emit the
// 16 x n_rots float3 rows with contiguous (coalesced) stores.
constexpr int kRotPairs = 16;

template <bool IDX64>
__global__ void __launch_bounds__(128) rot_vote_kernel(const float* __restrict__ points,
                                                       const float* __restrict__ preds_rot,
                                                       float* __restrict__ outputs_up, const void* __restrict__ idx,
                                                       long long n_pairs, int n_rots) {
    __shared__ float s_frame[kRotPairs][12];     // ab(3) x(3) y(3) tan live pad
    const long long p0 = (long long)blockIdx.x * kRotPairs;
    if (threadIdx.x < kRotPairs) {
        const long long p = p0 + threadIdx.x;
        float* fr = s_frame[threadIdx.x];
        fr[10] = 0.f;
        if (p < n_pairs) {
            int ia, ib;
            pair_ab<IDX64>(idx, p, 0, ia, ib);
            const f3 a = ld3(points, ia), b = ld3(points, ib);
            f3 ab, ex;
            if (pair_frame(a, b, ab, ex)) {                                 // :130 degenerate rows untouched
                const f3 y = cross3(ex, ab);                                // :137
                const float t = tanf(__ldg(preds_rot + p));                 // :142
                fr[0] = ab.x; fr[1] = ab.y; fr[2] = ab.z;
                fr[3] = ex.x; fr[4] = ex.y; fr[5] = ex.z;
                fr[6] = y.x; fr[7] = y.y; fr[8] = y.z;
                fr[9] = t;
                fr[10] = 1.f;
            }
        }
    }
    __syncthreads();
    const bool table = n_rots <= kMaxRots;
    const float2* tab = g_rot_table + n_rots * (n_rots - 1) / 2;
    const int total = kRotPairs * n_rots;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int lp = t / n_rots, i = t - lp * n_rots;
        const long long p = p0 + lp;
        const float* fr = s_frame[lp];
        if (p >= n_pairs || fr[10] == 0.f) continue;
        float ca, sa;
        if (table) {
            const float2 cs = __ldg(tab + i);
            ca = cs.x;
            sa = cs.y;
        } else {
            const float ang = rot_angle(i, n_rots);                        // :140
            ca = cosf(ang);
            sa = sinf(ang);
        }
        const f3 ab = {fr[0], fr[1], fr[2]}, x = {fr[3], fr[4], fr[5]}, y = {fr[6], fr[7], fr[8]};
        const float tn = fr[9];
        const f3 off = circle_offset(x, y, ca, sa);                                    // :141
        const f3 axis = tn > 0.f ? ab : f3{-ab.x, -ab.y, -ab.z};
        f3 up = off * tn + axis;                                           // :142
        up = up / (float)((double)len3(up) + 1e-7);                        // :143
        float* o = outputs_up + (p * n_rots + i) * 3;
        o[0] = up.x;
        o[1] = up.y;
        o[2] = up.z;
    }
}

// ---------------------------------------------------------------------------
// Sphere-bin count -- nocs/inference.py:282-283: counts[s] += #{c : cand[c] . sphere[s] > thr}.
// Thread s keeps its bin direction in registers; candidates stream through shared memory.
constexpr int kSphereChunk = 1024;

__global__ void __launch_bounds__(512) sphere_count_kernel(const float* __restrict__ cand, long long n_cand,
                                                           const float* __restrict__ sphere, int n_bins, float thr,
                                                           int* __restrict__ counts) {
    __shared__ float4 s_c[kSphereChunk];
    const long long c0 = (long long)blockIdx.x * kSphereChunk;
    const int m = (int)(n_cand - c0 < kSphereChunk ? n_cand - c0 : kSphereChunk);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const float* c = cand + (c0 + i) * 3;
        s_c[i] = make_float4(__ldg(c), __ldg(c + 1), __ldg(c + 2), 0.f);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < n_bins; s += blockDim.x) {
        const float sx = __ldg(sphere + 3 * s), sy = __ldg(sphere + 3 * s + 1), sz = __ldg(sphere + 3 * s + 2);
        int cnt = 0;
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const float4 c = s_c[i];
            const float d = fmaf(c.z, sz, fmaf(c.y, sy, c.x * sx));
            cnt += d > thr ? 1 : 0;
        }
        if (cnt) atomicAdd(counts + s, cnt);
    }
}

// ---------------------------------------------------------------------------
// findpeak -- models/voting.py:154-171 (literal: the y reads lose their x term, :165-166).
__global__ void __launch_bounds__(256) findpeak_kernel(const float* __restrict__ grid, float* __restrict__ out, int width,
                                                       int gx, int gy, int gz, int literal) {
    const long long n = (long long)gx * gy * gz;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int gyz = gy * gz;
    const int x = (int)(idx / gyz), yz = (int)(idx - (long long)x * gyz);
    const int y = yz / gz, z = yz - y * gz;
    const float g = grid[idx];
    const long long xo = (long long)x * gyz, xoy = literal ? 0 : xo;
    const float dx = g - grid[(long long)min(gx - 1, x + width) * gyz + y * gz + z] + g -
                     grid[(long long)max(0, x - width) * gyz + y * gz + z];
    const float dy = g - grid[xoy + min(gy - 1, y + width) * gz + z] + g - grid[xoy + max(0, y - width) * gz + z];
    const float dz = g - grid[xo + y * gz + min(gz - 1, z + width)] + g - grid[xo + y * gz + max(0, z - width)];
    out[idx] = dx + dy + dz;
}

// ---------------------------------------------------------------------------
// Categorical sampling + bin decode -- nocs/inference.py:185-188, 245-256.
struct SampleParams {
    const float* logits;
    long long n_rows;
    int row_stride, col0, n_bins, mode;
    const float* noise;
    unsigned long long seed;
    uint32_t stream_id;
    float div, mul_a, mul_b, sub;
    float* out_val;
    int out_stride;
    int* out_bin;
};

__global__ void __launch_bounds__(256) sample_bins_kernel(const SampleParams prm) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= prm.n_rows) return;
    const float* l = prm.logits + r * prm.row_stride + prm.col0;
    float m = -INFINITY;
    for (int k = 0; k < prm.n_bins; ++k) m = fmaxf(m, __ldg(l + k));
    float tot = 0.f;
    for (int k = 0; k < prm.n_bins; ++k) tot += expf(__ldg(l + k) - m);
    int bin = 0;
    if (prm.mode == 0) {                    // argmax(softmax(l) / q), first maximum
        const float* q = prm.noise + r * prm.n_bins;
        float best = -INFINITY;
        for (int k = 0; k < prm.n_bins; ++k) {
            const float pk = expf(__ldg(l + k) - m) / tot;
            const float v = pk / __ldg(q + k);
            if (v > best) {
                best = v;
                bin = k;
            }
        }
    } else {                                // inverse CDF, one uniform per row
        float u;
        if (prm.mode == 1) {
            u = __ldg(prm.noise + r);
        } else {
            const uint4 w = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)((unsigned long long)r >> 32), prm.stream_id, 0u),
                                          make_uint2((uint32_t)prm.seed, (uint32_t)(prm.seed >> 32)));
            u = u01(w.x);
        }
        const float t = u * tot;
        float acc = 0.f;
        bin = prm.n_bins - 1;
        for (int k = 0; k < prm.n_bins; ++k) {
            acc += expf(__ldg(l + k) - m);
            if (acc > t) {
                bin = k;
                break;
            }
        }
    }
    // nocs/inference.py:187-188,252: ((bin / (B-1)) * m_a) * m_b - s, every step rounded to fp32
    if (prm.out_val)
        prm.out_val[r * prm.out_stride] =
            __fsub_rn(__fmul_rn(__fmul_rn(__fdiv_rn((float)bin, prm.div), prm.mul_a), prm.mul_b), prm.sub);
    if (prm.out_bin) prm.out_bin[r] = bin;
}

static int blocks_for(long long n, int threads, int per_sm) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * per_sm;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace cppf

using namespace cppf;

extern "C" int cppf_ppf_vote(const float* points, const float* mu_nu, const float* probs, const void* idx,
                             int idx_is_64, float* grid, const float* corner, float res, int n_points,
                             int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive, void* stream_) {
    if (n_pairs <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    const bool table = n_rots <= kMaxRots;
    if (table) {
        const int e = ensure_rot_table(stream);
        if (e) return e;
    }
    VoteParams prm{points, mu_nu, probs, idx, grid, corner, res,
                   float_ceil(0.01), float_ceil((double)gx - 1.01), float_ceil((double)gy - 1.01),
                   float_ceil((double)gz - 1.01), n_points, (long long)n_pairs, n_rots, gx, gy, gz, adaptive};
    const int blocks = blocks_for(n_pairs, 256, 8);
    if (idx_is_64) {
        if (table) ppf_vote_kernel<true, true><<<blocks, 256, 0, stream>>>(prm);
        else ppf_vote_kernel<true, false><<<blocks, 256, 0, stream>>>(prm);
    } else {
        if (table) ppf_vote_kernel<false, true><<<blocks, 256, 0, stream>>>(prm);
        else ppf_vote_kernel<false, false><<<blocks, 256, 0, stream>>>(prm);
    }
    CPPF_LAUNCH_CHECK();
    return 0;
}

namespace cppf {
int grid_argmax_launch(const float* grid, int64_t n_cells, const int* n_cells_dev, int64_t* out_index, float* out_value,
                       cudaStream_t stream) {
    if (n_cells <= 0 || n_cells > 0xFFFFFFFFll) return (int)cudaErrorInvalidValue;
    if (!t_workspace_prepared) CPPF_RETURN_IF(cudaMemsetAsync(out_index, 0, sizeof(int64_t), stream));
    grid_argmax_kernel<<<blocks_for(n_cells, 256, 4), 256, 0, stream>>>(grid, (long long)n_cells,
                                                                       reinterpret_cast<unsigned long long*>(out_index),
                                                                       n_cells_dev);
    CPPF_LAUNCH_CHECK();
    grid_argmax_finish_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<long long*>(out_index), out_value);
    CPPF_LAUNCH_CHECK();
    return 0;
}
}  // namespace cppf

extern "C" int cppf_grid_argmax(const float* grid, int64_t n_cells, int64_t* out_index, float* out_value, void* stream_) {
    return grid_argmax_launch(grid, n_cells, nullptr, out_index, out_value, (cudaStream_t)stream_);
}

extern "C" int cppf_backvote(const float* points, const float* mu_nu, float* out_offsets, uint8_t* out_mask,
                             const void* idx, int idx_is_64, const float* corner, float res, int n_points,
                             int64_t n_pairs, int n_rots, int gx, int gy, int gz, const float* centre, float tol,
                             void* stream_) {
    if (n_pairs <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    const bool table = n_rots <= kMaxRots;
    if (table) {
        const int e = ensure_rot_table(stream);
        if (e) return e;
    }
    BackvoteParams prm{points, mu_nu, out_offsets, out_mask, idx, corner, centre, res, tol,
                       (float)(gx - 1), (float)(gy - 1), (float)(gz - 1), n_points, (long long)n_pairs, n_rots};
    const int blocks = blocks_for(n_pairs, 256, 8);
    if (idx_is_64) {
        if (table) backvote_kernel<true, true><<<blocks, 256, 0, stream>>>(prm);
        else backvote_kernel<true, false><<<blocks, 256, 0, stream>>>(prm);
    } else {
        if (table) backvote_kernel<false, true><<<blocks, 256, 0, stream>>>(prm);
        else backvote_kernel<false, false><<<blocks, 256, 0, stream>>>(prm);
    }
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t cppf_compact_scratch_bytes(int64_t n_pairs) {
    const int64_t nb = (n_pairs + kCompactBlock * kCompactItems - 1) / (kCompactBlock * kCompactItems);
    return nb * (int64_t)(sizeof(int) + sizeof(long long)) + 64;
}

// count + scan only: *out_count = number of set mask bytes, scratch[0 : nb] (int64) = exclusive survivor offsets of the
// 2048-pair blocks -- all the fused path needs to address "the r-th survivor" without materialising the list
extern "C" int cppf_compact_count(const uint8_t* mask, int64_t n_pairs, int64_t* out_count, void* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pairs <= 0) return (int)cudaMemsetAsync(out_count, 0, sizeof(int64_t), stream);
    const int64_t nb = (n_pairs + kCompactBlock * kCompactItems - 1) / (kCompactBlock * kCompactItems);
    if (nb > 0x7FFFFFFF) return (int)cudaErrorInvalidValue;
    long long* block_offsets = reinterpret_cast<long long*>(scratch);
    int* block_counts = reinterpret_cast<int*>(block_offsets + nb);
    compact_count_kernel<<<(int)nb, kCompactBlock, 0, stream>>>(mask, (long long)n_pairs, block_counts);
    CPPF_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, stream>>>(block_counts, (int)nb, block_offsets, reinterpret_cast<long long*>(out_count));
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_compact_pairs(const uint8_t* mask, const void* idx, int idx_is_64, int n_points, int64_t n_pairs,
                                  int32_t* out_idx, int64_t* out_pos, int64_t* out_count, void* scratch,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pairs <= 0) return (int)cudaMemsetAsync(out_count, 0, sizeof(int64_t), stream);
    const int64_t nb = (n_pairs + kCompactBlock * kCompactItems - 1) / (kCompactBlock * kCompactItems);
    if (nb > 0x7FFFFFFF) return (int)cudaErrorInvalidValue;
    long long* block_offsets = reinterpret_cast<long long*>(scratch);
    int* block_counts = reinterpret_cast<int*>(block_offsets + nb);
    compact_count_kernel<<<(int)nb, kCompactBlock, 0, stream>>>(mask, (long long)n_pairs, block_counts);
    CPPF_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, stream>>>(block_counts, (int)nb, block_offsets,
                                                reinterpret_cast<long long*>(out_count));
    CPPF_LAUNCH_CHECK();
    if (idx_is_64)
        compact_scatter_kernel<true><<<(int)nb, kCompactBlock, 0, stream>>>(mask, idx, n_points, (long long)n_pairs,
                                                                            block_offsets, out_idx,
                                                                            reinterpret_cast<long long*>(out_pos));
    else
        compact_scatter_kernel<false><<<(int)nb, kCompactBlock, 0, stream>>>(mask, idx, n_points, (long long)n_pairs,
                                                                             block_offsets, out_idx,
                                                                             reinterpret_cast<long long*>(out_pos));
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_rot_vote(const float* points, const float* preds_rot, float* outputs_up, const void* idx,
                             int idx_is_64, int64_t n_pairs, int n_rots, void* stream_) {
    if (n_pairs <= 0 || n_rots <= 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (idx == nullptr) return (int)cudaErrorInvalidValue;
    if (n_rots <= kMaxRots) {
        const int e = ensure_rot_table(stream);
        if (e) return e;
    }
    const int64_t nb = (n_pairs + kRotPairs - 1) / kRotPairs;
    if (nb > 0x7FFFFFFF) return (int)cudaErrorInvalidValue;
    if (idx_is_64) rot_vote_kernel<true><<<(int)nb, 128, 0, stream>>>(points, preds_rot, outputs_up, idx, n_pairs, n_rots);
    else rot_vote_kernel<false><<<(int)nb, 128, 0, stream>>>(points, preds_rot, outputs_up, idx, n_pairs, n_rots);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_sphere_count(const float* cand, int64_t n_cand, const float* sphere, int n_bins, float thr,
                                 int32_t* counts, void* stream_) {
    if (n_cand <= 0 || n_bins <= 0) return 0;
    const int64_t nb = (n_cand + kSphereChunk - 1) / kSphereChunk;
    if (nb > 0x7FFFFFFF) return (int)cudaErrorInvalidValue;
    sphere_count_kernel<<<(int)nb, 512, 0, (cudaStream_t)stream_>>>(cand, (long long)n_cand, sphere, n_bins, thr, counts);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_findpeak(const float* grid, float* out, int width, int gx, int gy, int gz, int literal, void* stream_) {
    const long long n = (long long)gx * gy * gz;
    if (n <= 0) return 0;
    findpeak_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(grid, out, width, gx, gy, gz, literal);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_sample_bins(const float* logits, int64_t n_rows, int row_stride, int col0, int n_bins, int mode,
                                const float* noise, uint64_t seed, uint32_t stream_id, float div, float mul_a,
                                float mul_b, float sub, float* out_val, int out_stride, int32_t* out_bin,
                                void* stream_) {
    if (n_rows <= 0) return 0;
    if (n_bins <= 0 || mode < 0 || mode > 2 || (mode < 2 && noise == nullptr)) return (int)cudaErrorInvalidValue;
    SampleParams prm{logits, (long long)n_rows, row_stride, col0, n_bins, mode, noise, seed, stream_id,
                     div, mul_a, mul_b, sub, out_val, out_stride, out_bin};
    sample_bins_kernel<<<(int)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}

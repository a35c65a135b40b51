// Fast-path voting kernels (the "mode F" pipeline after encode_sample):
//   * centre voting into a shared-memory-privatised fixed-point grid,
//   * back-vote filter driven by the sampled bins,
//   * fused orientation candidates + sphere histogram (no [P,72,3] dump),
//   * survivor statistics (scale mean, aux-sign sums).
// Semantics follow models/voting.py:8-66, 74-112, 119-147 and nocs/inference.py:276-302,335;
// the measurements behind the design are in profiles/r1_atomics_microbench.json:
// shared-memory u32 atomics sustain ~2000 G/s (500+ G/s on hot cells) against 90-180 G/s
// (16 G/s hot) for global fp32 reductions.
#include "common.cuh"
#include "vote_common.cuh"

#include <mutex>

#include "../../include/cppf_b200.h"

#include <math.h>

namespace cppf {

bool dense_rows_ok(const void* idx, int n_points, int64_t n_pairs, int row0);    // encode_tc.cu

constexpr int kSelectBlockBytes = 2048;  // = kCompactBlock * kCompactItems of vote.cu (cppf_compact_count)
constexpr int kSelectBlock = kSelectBlockBytes;

struct VotePParams {
    const float2* rot_tab;
    const float* points;
    const float* mu_nu;        // [P][2] floats, or
    const uint8_t* bins;       // [P][4] bins decoded through lut (mu: lut[0:32], nu: lut[32:64])
    const float* lut;
    const void* idx;
    unsigned long long* acc;   // [cells] fixed-point accumulator (global)
    const float* corner;
    float res, inv_res;
    float lo, hx, hy, hz;      // exact bounds on g = d / res (models/voting.py:36-39, double literals rounded up)
    float dlo, dhx, dhy, dhz;  // conservative bounds on d = candidate - corner (phase 1, before the division)
    int n_points;
    long long n_pairs;
    int n_rots, gx, gy, gz, adaptive;
    const Geom* geom;          // optional: device-side geometry overrides corner / dims / upper bounds
    int max_cells;             // capacity of the shared-memory grid of this launch
    int rep_stride;            // 0, or the offset (in cells) of a second replica of the grid used by the odd lanes
    int row0;                  // dense mode: first row of the block of the pair matrix (n_pairs / n_points rows) this launch covers
    int slab_planes;           // 0: every CTA holds the whole grid.  > 0 ("slab passes", grids of up to gridDim slabs): CTA b
                               // holds the x-slab (b % n_slabs) of slab_planes base planes (+ 1 overlap plane), walks the
                               // pairs of part (b / n_slabs) and keeps only the candidates whose floor(g.x) it owns --
                               // phase 1 is repeated per slab, but nothing is routed through memory
    int n_slabs;
};

#ifndef CPPF_VOTE_THREADS
#define CPPF_VOTE_THREADS 1024
#endif
#ifndef CPPF_VOTE_CONST_TAB
#define CPPF_VOTE_CONST_TAB 1     // rotation table read from constant memory (the lanes of a sorted chunk share the entry): 2.49 -> 2.46 ms
#endif
#ifndef CPPF_VOTE_UNROLL
#define CPPF_VOTE_UNROLL 4
#endif
constexpr int kVoteThreads = CPPF_VOTE_THREADS;
constexpr int kVoteUnroll = CPPF_VOTE_UNROLL;     // unroll factor of the rotation walk
constexpr int kVoteBatch = 2 * kVoteThreads;       // pairs sorted and voted between two block barriers
constexpr int kVoteQueue = 64;                     // per-warp ring of in-bounds candidates (float4 slots)
constexpr int kVoteKeys = kMaxRotsP + 1;           // sort key = rotation count of the pair (0..72)
#ifndef CPPF_VOTE_TILEB
#define CPPF_VOTE_TILEB 16
#endif
constexpr int kTileB = CPPF_VOTE_TILEB, kTileA = kVoteBatch / kTileB;          // dense mode: a batch is a kTileA x kTileB tile of the pair matrix
static_assert(kTileA * kTileB == kVoteBatch, "tile = batch");
// Overflow guard: after every batch each cell holding >= 2^30 units is flushed to the global u64 accumulator.
// A batch adds at most kVoteBatch * 72 candidates * 2^14 units = 2.42e9 < 2^32 - 2^30 to any one cell.
constexpr unsigned kFlushAt = 1u << 30;
static_assert((unsigned long long)kVoteBatch * kMaxRotsP * (1ull << kFixShift) < (1ull << 32) - kFlushAt, "guard");

// Per batch of 2048 pairs a CTA
//   (1) counting-sorts the pairs by their rotation count n (adaptive voting, models/voting.py:31: n depends
//       on nu, so a warp of unsorted pairs walks max(n) = 66 steps for a mean n of 49) -- largest n first;
//   (2) lets its warps pull 32 sorted pairs at a time from a shared counter (longest chunks first, so the
//       warps reach the batch barrier together);
// and every warp works in two phases decoupled by a shared-memory ring:
//   phase 1 (one lane = one pair, all lanes walk their circle together): candidate offset from the grid
//            corner and a conservative in-bounds test on it (no division); the ~45 % of candidates that
//            pass are appended to the warp's ring with a ballot + popc prefix;
//   phase 2 (whenever 32 candidates are queued): one lane = one queued candidate -> the reference's `/ res`
//            and exact in-bounds test (models/voting.py:35-39), then 8 shared-memory atomics.
// Without the ring the splat runs under the divergence of the in-bounds test (about a third of the lanes
// active); with it the atomics always issue from full warps.  Integer sums are order-independent, so neither
// the sort nor the dynamic chunk assignment changes the result.
#if CPPF_VOTE_CONST_TAB
__constant__ float2 c_rot_tab[kRotTabP];     // the rotation table in constant memory: lanes of a sorted chunk read the same entry
#endif

template <bool IDX64, bool BINS, bool SLABS>
__global__ void __launch_bounds__(kVoteThreads, 1) vote_private_kernel(const VotePParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
#if CPPF_VOTE_CONST_TAB
    float* s_lut = reinterpret_cast<float*>(smem_raw);          // the rotation table lives in constant memory
#else
    float2* s_tab = reinterpret_cast<float2*>(smem_raw);
    float* s_lut = reinterpret_cast<float*>(s_tab + kRotTabP);
#endif
    float4* s_queue = reinterpret_cast<float4*>(s_lut + 64);
    unsigned short* s_perm = reinterpret_cast<unsigned short*>(s_queue + (kVoteThreads / 32) * kVoteQueue);
    unsigned* s_grid = reinterpret_cast<unsigned*>(s_perm + kVoteBatch);
    __shared__ int s_hist[kVoteKeys + 3], s_start[kVoteKeys + 3], s_nlut[32];
    __shared__ int s_total, s_next;
    int gx = prm.gx, gy = prm.gy, gzd = prm.gz;
    float hx = prm.hx, hy = prm.hy, hz = prm.hz, dhx = prm.dhx, dhy = prm.dhy, dhz = prm.dhz;
    const float* corner = prm.corner;
    int pps = SLABS ? prm.slab_planes : 0, n_slabs = prm.n_slabs;
    if (prm.geom != nullptr) {
        const Geom g = *prm.geom;
        if (g.status != 0) return;
        if (pps == 0 ? (g.mode != 0 || g.cells > prm.max_cells) : g.mode != 1) return;
        gx = g.gx; gy = g.gy; gzd = g.gz;
        hx = g.hx; hy = g.hy; hz = g.hz; dhx = g.dhx; dhy = g.dhy; dhz = g.dhz;
        corner = prm.geom->corner;
        if (pps != 0) {                              // slab plan from the capacity of this launch
            pps = prm.max_cells / (gy * gzd) - 1;
            if (pps < 1) return;
            if (pps > gx) pps = gx;
            n_slabs = (gx + pps - 1) / pps;
        }
    }
    int x0 = 0, part = blockIdx.x, parts = gridDim.x;
    float dlx = prm.dlo;
    if (SLABS) {
        parts = gridDim.x / n_slabs;
        if (parts < 1 || (int)blockIdx.x >= parts * n_slabs) return;
        x0 = (blockIdx.x % n_slabs) * pps;
        part = blockIdx.x / n_slabs;
        // conservative bounds on d.x for floor(g.x) in [x0, x0 + pps): the division rounds within 2^-24
        if (x0 > 0) dlx = fmaxf(dlx, (float)x0 * prm.res * 0.999998f);
        dhx = fminf(dhx, (float)(x0 + pps) * prm.res * 1.000002f);
    }
    const int x_hi = SLABS ? x0 + pps : 0x7fffffff;              // owned base planes: [x0, x_hi)
    const int cells = (SLABS ? min(gx, x0 + pps + 1) - x0 : gx) * gy * gzd;
    const long long acc_off = SLABS ? (long long)x0 * gy * gzd : 0ll;
    const float fx0 = (float)x0;
#if !CPPF_VOTE_CONST_TAB
    for (int i = threadIdx.x; i < kRotTabP; i += blockDim.x) s_tab[i] = __ldg(prm.rot_tab + i);
#endif
    if (BINS && threadIdx.x < 64) s_lut[threadIdx.x] = __ldg(prm.lut + threadIdx.x);
    if (BINS && threadIdx.x < 32) {
        int n = prm.n_rots;
        if (prm.adaptive) n = adaptive_rots(__ldg(prm.lut + 32 + threadIdx.x), prm.res, prm.n_rots);   // :31
        s_nlut[threadIdx.x] = n < 0 ? 0 : n;
    }
    // Two replicas when they fit: a quarter of the candidates of a 32-wide splat share their base cell with another
    // lane (votes concentrate by design) and same-address shared-memory atomics serialise; odd lanes vote into the
    // second copy, 8 banks away (simulated and measured: 4.9 -> 4.0 wavefronts per ATOMS).
    const int rep = prm.rep_stride;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s_grid[i] = 0u;
    if (rep)
        for (int i = threadIdx.x; i < cells; i += blockDim.x) s_grid[rep + i] = 0u;
    if (threadIdx.x < kVoteKeys + 3) s_hist[threadIdx.x] = 0;
    __syncthreads();

    const int gyz = gy * gzd, gz = gzd;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned s_lane = (unsigned)__cvta_generic_to_shared(s_grid + (lane & 1) * rep);   // this lane's replica
    const unsigned q_addr = (unsigned)__cvta_generic_to_shared(s_queue + (threadIdx.x >> 5) * kVoteQueue);
    const float cx = __ldg(corner), cy = __ldg(corner + 1), cz = __ldg(corner + 2);
    // A batch is 2048 consecutive entries of the pair list, or -- dense mode -- a 128 x 16 tile of the pair matrix
    // (128 different points a): pairs that share a reach the vote peak in the same iterations, and mixing the a's of a
    // batch takes another 9 % off the same-cell replays of the splat (tools/sim_vote_banks.py: 4.0 -> 3.65 per ATOMS).
    const bool tiled = prm.idx == nullptr;
    const int tiles_x = (prm.n_points + kTileB - 1) / kTileB;
    const int n_rows = tiled ? (int)(prm.n_pairs / prm.n_points) : 0;      // rows [row0, row0 + n_rows) of the pair matrix
    const long long n_batches = tiled ? (long long)((n_rows + kTileA - 1) / kTileA) * tiles_x
                                      : (prm.n_pairs + kVoteBatch - 1) / kVoteBatch;
    auto pair_index = [&](long long batch, int local) -> long long {      // position in this launch's pair arrays; -1: none
        if (!tiled) {
            const long long q = batch * kVoteBatch + local;
            return q < prm.n_pairs ? q : -1;
        }
        const int tr = (int)(batch / tiles_x), tc = (int)(batch - (long long)tr * tiles_x);
        const int r = tr * kTileA + (local / kTileB), b = tc * kTileB + (local % kTileB);
        return (r < n_rows && b < prm.n_points) ? (long long)r * prm.n_points + b : -1;
    };

    for (long long batch = part; batch < n_batches; batch += parts) {
        // ---- (1a) keys + ranks: n of each of this thread's two pairs
        int key[2], rank[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const long long p = pair_index(batch, threadIdx.x + j * kVoteThreads);
            key[j] = -1;
            if (p >= 0) {
                int n;
                if (BINS) {
                    n = s_nlut[__ldg(prm.bins + 4 * p + 1) & 31];
                } else {
                    n = prm.n_rots;
                    if (prm.adaptive) n = adaptive_rots(__ldg(prm.mu_nu + 2 * p + 1), prm.res, prm.n_rots);
                    n = n < 0 ? 0 : n;
                }
                key[j] = n;
                rank[j] = atomicAdd(&s_hist[n], 1);
            }
        }
        // ---- overflow guard on the votes of the previous batches (nobody is voting now): four cells per load, the
        // per-cell flush only where one of them has reached the threshold
        for (int r = 0; r < (rep ? 2 : 1); ++r) {
            unsigned* gr = s_grid + r * rep;
            for (int i = 4 * threadIdx.x; i < cells; i += 4 * blockDim.x) {
                if (i + 4 <= cells) {
                    const uint4 v = *reinterpret_cast<const uint4*>(gr + i);
                    if (max(max(v.x, v.y), max(v.z, v.w)) < kFlushAt) continue;
                }
                for (int j = i; j < min(i + 4, cells); ++j) {
                    const unsigned v = gr[j];
                    if (v >= kFlushAt) {
                        atomicAdd(prm.acc + acc_off + j, (unsigned long long)v);
                        gr[j] = 0u;
                    }
                }
            }
        }
        __syncthreads();
        // ---- (1b) exclusive scan of the histogram, largest n first
        if (threadIdx.x < 32) {
            int v[3], sum = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int e = 3 * lane + k;                      // e = 72 - n
                v[k] = e < kVoteKeys ? s_hist[kMaxRotsP - e] : 0;
                sum += v[k];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int run = incl - sum;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int e = 3 * lane + k;
                if (e < kVoteKeys) {
                    s_start[kMaxRotsP - e] = run;
                    s_hist[kMaxRotsP - e] = 0;
                }
                run += v[k];
            }
            if (lane == 31) {
                s_total = incl;
                s_next = 0;
            }
        }
        __syncthreads();
        // ---- (1c) permutation
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (key[j] >= 0) s_perm[s_start[key[j]] + rank[j]] = (unsigned short)(threadIdx.x + j * kVoteThreads);
        __syncthreads();
        const int total = s_total;

        // ---- (2) warps pull chunks of 32 sorted pairs
        unsigned q_head = 0, q_tail = 0;                                       // warp-uniform
        while (true) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&s_next, 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk * 32 >= total) break;
            const int item = chunk * 32 + lane;
            int n = 0;
            f3 c = {0.f, 0.f, 0.f}, x = c, y = c;
            if (item < total) {
                const long long p = pair_index(batch, s_perm[item]);
                int ia, ib;
                pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib, prm.row0);
                float mu, nu;
                if (BINS) {
                    const uchar4 bn = __ldg(reinterpret_cast<const uchar4*>(prm.bins) + p);
                    mu = s_lut[bn.x];
                    nu = s_lut[32 + bn.y];
                    n = s_nlut[bn.y & 31];
                } else {
                    const float2 mn = __ldg(reinterpret_cast<const float2*>(prm.mu_nu) + p);
                    mu = mn.x;
                    nu = mn.y;
                    n = prm.n_rots;
                    if (prm.adaptive) n = adaptive_rots(nu, prm.res, prm.n_rots);  // :31
                    if (n < 0) n = 0;
                }
                const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
                f3 ab, ex;
                if (pair_frame(a, b, ab, ex)) {                                // voting.py:21
                    c = foot_point(a, ab, mu);                                           // :23
                    x = scale3(ex, nu);                                               // :28
                    y = cross3(x, ab);                                         // :29
                } else {
                    n = 0;
                }
            }
            // row n of the rotation table; reads past column n stay inside the table.  (Starting every lane at a different
            // phase of its circle would decorrelate the splat addresses -- the pairs of a warp share point a and reach the
            // vote peak in the same iterations -- and was measured: fewer ATOMS replays, but a slower kernel overall.)
            const int row = n > 0 ? n * (n - 1) / 2 : 0;
#if !CPPF_VOTE_CONST_TAB
            const float2* tab = s_tab + row;
#endif
            const int n_max = __reduce_max_sync(0xffffffffu, n);
#pragma unroll kVoteUnroll
            for (int i = 0; i < n_max; ++i) {
#if CPPF_VOTE_CONST_TAB
                const float2 cs = c_rot_tab[row + i];
#else
                const float2 cs = tab[i];
#endif
                const f3 off = circle_offset(x, y, cs.x, cs.y);                            // :34
                const float dx = c.x + off.x - cx, dy = c.y + off.y - cy, dz = c.z + off.z - cz;   // :35 before `/ res`
                // conservative: every candidate the exact test of phase 2 accepts passes here
                const bool inb = i < n && dx >= dlx && dy >= prm.dlo && dz >= prm.dlo && dx < dhx && dy < dhy && dz < dhz;
                const unsigned m = __ballot_sync(0xffffffffu, inb);
                if (inb) st_shared_f4(q_addr + (((q_tail + __popc(m & lt_mask)) & (kVoteQueue - 1)) << 4), dx, dy, dz);
                q_tail += __popc(m);
                if (q_tail - q_head >= 32u) {
                    __syncwarp();
                    const float4 d = ld_shared_f4(q_addr + (((q_head + lane) & (kVoteQueue - 1)) << 4));
                    const float gxf = div_by(d.x, prm.res, prm.inv_res);       // :35
                    const float gyf = div_by(d.y, prm.res, prm.inv_res);
                    const float gzf = div_by(d.z, prm.res, prm.inv_res);
                    const bool vote = !(gxf < prm.lo || gyf < prm.lo || gzf < prm.lo || gxf >= hx || gyf >= hy || gzf >= hz) &&
                                      (!SLABS || ((int)gxf >= x0 && (int)gxf < x_hi));
                    const float gxl = SLABS ? gxf - fx0 : gxf;
                    // (choosing the replica by the rank among the lanes that share a base cell -- __match_any_sync -- was
                    // simulated at 4.00 -> 3.66 wavefronts per ATOMS and measured: the MATCH costs 1 ms per launch)
                    if (vote) splat_fixed(s_lane, gxl, gyf, gzf, gyz, gz);                 // :36-63
                    q_head += 32u;
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        if (lane < (int)(q_tail - q_head)) {
            const float4 d = ld_shared_f4(q_addr + (((q_head + lane) & (kVoteQueue - 1)) << 4));
            const float gxf = div_by(d.x, prm.res, prm.inv_res);
            const float gyf = div_by(d.y, prm.res, prm.inv_res);
            const float gzf = div_by(d.z, prm.res, prm.inv_res);
            if (!(gxf < prm.lo || gyf < prm.lo || gzf < prm.lo || gxf >= hx || gyf >= hy || gzf >= hz) &&
                (!SLABS || ((int)gxf >= x0 && (int)gxf < x_hi)))
                splat_fixed(s_lane, SLABS ? gxf - fx0 : gxf, gyf, gzf, gyz, gz);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const unsigned long long v = (unsigned long long)s_grid[i] + (rep ? (unsigned long long)s_grid[rep + i] : 0ull);
        if (v) atomicAdd(prm.acc + acc_off + i, v);
    }
}

// grid[i] += acc[i] * 2^-14   (exact integer sum -> one rounding; deterministic run to run)
__global__ void __launch_bounds__(256) vote_finalize_kernel(const unsigned long long* __restrict__ acc,
                                                            float* __restrict__ grid, int cells, const Geom* geom,
                                                            int only_mode) {
    if (geom != nullptr) {
        if (only_mode >= 0 && geom->mode != only_mode) return;
        cells = geom->status == 0 ? min(geom->cells, cells) : 0;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) grid[i] += (float)((double)acc[i] * (1.0 / 16384.0));
}

// ---------------------------------------------------------------------------
// Back-vote from bins -- models/voting.py:74-112 with (mu, nu) decoded from the sampled bins.
struct BackvotePParams {
    const float2* rot_tab;
    const float* points;
    const uint8_t* bins;
    const float* lut;
    const void* idx;
    uint8_t* out_mask;
    const float* corner;
    const long long* argmax_flat;    // winning cell (device) -> centre = corner + cell * res
    float res, inv_res, tol;
    float hx, hy, hz;
    int n_points;
    long long n_pairs;
    int n_rots, gx, gy, gz;
    const Geom* geom;                // optional: device-side geometry overrides corner / dims / bounds
    int row0;                        // dense mode: first row of the block of the pair matrix this launch covers
    double res_host;                 // the resolution as the DOUBLE the host multiplies the winning cell with (nocs/inference.py:209:
                                     // `corners[0] + cand * cfg.res`, a Python float); (double)(float)res differs from it in the 9th digit
};

#ifndef CPPF_BV_THREADS
#define CPPF_BV_THREADS 256
#endif
#ifndef CPPF_BV_BLOCKS_PER_SM
#define CPPF_BV_BLOCKS_PER_SM 8
#endif
#ifndef CPPF_BV_MIN_BLOCKS
#define CPPF_BV_MIN_BLOCKS 1
#endif
#ifndef CPPF_STATS_BLOCKS_PER_SM
#define CPPF_STATS_BLOCKS_PER_SM 4      // one full wave of the 4 resident blocks that 64 registers allow
#endif
#ifndef CPPF_STATS_MIN_BLOCKS
#define CPPF_STATS_MIN_BLOCKS 4
#endif

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool IDX64>
__global__ void __launch_bounds__(CPPF_BV_THREADS, CPPF_BV_MIN_BLOCKS) backvote_bins_kernel(const BackvotePParams prm) {
    __shared__ float2 s_tab[kRotTabP];
    __shared__ float s_lut[64];
    for (int i = threadIdx.x; i < kRotTabP; i += blockDim.x) s_tab[i] = __ldg(prm.rot_tab + i);
    if (threadIdx.x < 64) s_lut[threadIdx.x] = __ldg(prm.lut + threadIdx.x);
    __shared__ int s_nlut[32];                                             // rotation count of every nu bin (:97)
    if (threadIdx.x < 32) s_nlut[threadIdx.x] = adaptive_rots(__ldg(prm.lut + 32 + threadIdx.x), prm.res, prm.n_rots);
    __syncthreads();
    int gy = prm.gy, gz = prm.gz;
    float hx = prm.hx, hy = prm.hy, hz = prm.hz;
    const float* corner = prm.corner;
    if (prm.geom != nullptr) {
        if (prm.geom->status != 0) return;
        gy = prm.geom->gy; gz = prm.geom->gz;
        hx = prm.geom->bx; hy = prm.geom->by; hz = prm.geom->bz;
        corner = prm.geom->corner;
    }
    const float cx = __ldg(corner), cy = __ldg(corner + 1), cz = __ldg(corner + 2);
    // nocs/inference.py:208-209: T = corner + cell * res in float64, cast to float32 for the kernel (:226)
    const long long flat = *prm.argmax_flat;
    const int gyz = gy * gz;
    const int ix = (int)(flat / gyz), iy = (int)((flat % gyz) / gz), iz = (int)(flat % gz);
    const float tx = (float)((double)cx + (double)ix * prm.res_host);
    const float ty = (float)((double)cy + (double)iy * prm.res_host);
    const float tz = (float)((double)cz + (double)iz * prm.res_host);
    // is the tolerance ball around the winning cell at least one cell away from every face of the grid?
    const int gxi = prm.geom != nullptr ? prm.geom->gx : prm.gx;
    const int marg = (int)(prm.tol * prm.inv_res) + 2;
    const bool interior = ix >= marg && iy >= marg && iz >= marg && ix < gxi - 1 - marg && iy < gy - 1 - marg && iz < gz - 1 - marg;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < prm.n_pairs;
         p += (long long)gridDim.x * blockDim.x) {
        int ia, ib;
        pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib, prm.row0);
        const uchar4 bn = __ldg(reinterpret_cast<const uchar4*>(prm.bins) + p);
        const float mu = s_lut[bn.x], nu = s_lut[32 + bn.y];
        const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
        f3 ab, ex;
        bool hit = false;
        if (pair_frame(a, b, ab, ex)) {
            const f3 c = foot_point(a, ab, mu);
            const f3 x = scale3(ex, nu);
            const f3 y = cross3(x, ab);
            const int n = s_nlut[bn.y & 31];                                   // :97
            const float2* tab = s_tab + (n > 0 ? n * (n - 1) / 2 : 0);
            // The reference walks i = 0..n-1 and stops at the first candidate within `tol` of the centre; only
            // whether such a candidate EXISTS is consumed (nocs/inference.py:229-230).  The candidates lie on a
            // circle (centre c, radius nu, plane normal ab), so the ones within tol of T form one arc around the
            // angle of T's in-plane projection: bound that arc analytically (with generous slack), then test only
            // its candidates -- each with the reference's own arithmetic, so the mask is unchanged.
            int i_lo = 0, i_cnt = n;
            // in the pair's own frame (ab, ex, ey orthonormal) the squared distance of candidate i from the centre T is
            //   nu^2 + |v|^2 - 2 nu (px cos_i + py sin_i),   v = T - c, px = v.ex, py = v.ey
            // so a candidate can only be within tol if  nu (px cos_i + py sin_i) >= K = (nu^2 + |v|^2 - tol^2) / 2.
            const f3 v = {tx - c.x, ty - c.y, tz - c.z};
            const f3 ey = cross3(ex, ab);
            const float dpl = dot3(v, ab), px = dot3(v, ex), py = dot3(v, ey);
            const float vv = dot3(v, v);
            const float tol2 = prm.tol * prm.tol;
            const float kq = 0.5f * (nu * nu + vv - tol2) - (0.02f * tol2 + 1e-6f * (nu * nu + vv));   // K minus rounding slack
            // the shortcuts below assume |ex| = |ey| = 1; the reference's `+ 1e-7` normaliser shrinks ex when ab is within
            // ~1e-3 rad of the x axis -- such pairs take the plain scan with the reference's arithmetic only
            const bool unit_frame = dot3(ex, ex) > 0.9999f;
            if (n > 12 && unit_frame) {
                const float rho = sqrt_approx(px * px + py * py);    // window placement only: every shortcut here has slack
                const float room = tol2 * 1.01f + 1e-12f - dpl * dpl - (nu - rho) * (nu - rho);   // >= 2 nu rho (1 - cos d) for a hit
                const float two_nr = 2.f * fabsf(nu) * rho;
                if (room < 0.f) {
                    i_cnt = 0;
                } else if (room < 1.9f * two_nr) {
                    // half-width of the arc: acos(1 - q) <= sqrt(2q) (1 + 0.22 q) on [0, 1.9] (no acosf, no atan2f: a
                    // polynomial atan2 good to 2e-4 rad only has to place the centre of the window)
                    const float q = __fdividef(room, two_nr);
                    const float dmax = sqrt_approx(2.f * q) * fmaf(0.22f, q, 1.f) + 1e-5f;
                    const float inv_step = (float)n * 0.159154943f;             // 1 / (2 pi / n)
                    const float apx = fabsf(px), apy = fabsf(py);
                    const float mx = fmaxf(apx, apy), mn = fminf(apx, apy);
                    const float t = __fdividef(mn, fmaxf(mx, 1e-30f)), t2 = t * t;
                    float th = fmaf(fmaf(fmaf(-0.0464964749f, t2, 0.15931422f), t2, -0.327622764f) * t2, t, t);
                    if (apy > apx) th = 1.57079633f - th;
                    if (px < 0.f) th = 3.14159265f - th;
                    if (py < 0.f) th = -th;
                    if (nu < 0.f) th += 3.14159265f;                             // x, y carry the sign of nu
                    if (th < 0.f) th += 6.2831853f;
                    const int w = (int)(dmax * inv_step + 0.52f) + 1;
                    if (2 * w + 1 < n) {
                        i_lo = (int)(th * inv_step + 0.5f) - w;
                        i_cnt = 2 * w + 1;
                    }
                }
            }
            // start of the window into [0, n) (i_lo is in [-w, n] with w < n / 2): one wrap test per candidate below
            i_lo = i_lo < 0 ? i_lo + n : (i_lo >= n ? i_lo - n : i_lo);
            for (int k = 0, i = i_lo; k < i_cnt; ++k, ++i) {
                if (i >= n) i -= n;
                const float2 cs = tab[i];
                if (unit_frame) {
                    const float qv = nu * fmaf(px, cs.x, py * cs.y);
                    if (qv < kq) continue;                                     // cannot be within tol (see above)
                    // well inside the tolerance ball of an interior centre: in bounds by construction, offset of length
                    // |nu| != 0 -- the reference's tests (:102-109) hold with a 10 % margin, no need to evaluate them
                    if (interior && nu != 0.f && fmaf(-2.f, qv, nu * nu + vv) <= 0.9f * tol2) {
                        hit = true;
                        break;
                    }
                }
                const f3 off = circle_offset(x, y, cs.x, cs.y);
                const f3 pc = c + off;
                const f3 dlt = {pc.x - tx, pc.y - ty, pc.z - tz};
                if (len3(dlt) > prm.tol) continue;                             // :102
                const float gxf = div_by(pc.x - cx, prm.res, prm.inv_res);
                const float gyf = div_by(pc.y - cy, prm.res, prm.inv_res);
                const float gzf = div_by(pc.z - cz, prm.res, prm.inv_res);
                if (gxf < 0.f || gyf < 0.f || gzf < 0.f || gxf >= hx || gyf >= hy || gzf >= hz) continue;
                hit = off.x != 0.f || off.y != 0.f || off.z != 0.f;            // inference.py:230 any(oc != 0)
                if (hit) break;
            }
        }
        prm.out_mask[p] = hit ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------
// Orientation voting on a sub-sample of the survivors, fused with the sphere histogram:
// models/voting.py:119-147 + nocs/inference.py:276-284.  Survivor j of the sub-sample is
// pos[(offset + j*stride) mod count] (stride coprime with count: a sample without
// replacement, like the reference's shuffle).
//
// The reference multiplies every candidate with every sphere bin ([720 000,3] x [3,480]) and counts
// dot > cos(1.5 deg).  A candidate c (|c| <= 1) and a bin s (|s| = 1) with c.s > thr are closer than
// d = sqrt(2 - 2 thr) (0.0262 for 1.5 deg), so their y coordinates differ by less than d; the Fibonacci
// sphere of utils/util.py:102-118 has y strictly decreasing in the bin index, so only the ~14 bins of the
// window |y_s - c_y| <= d can count.  Each thread takes candidates, binary-searches that window in shared
// memory and tests its bins with the same dot-product expression as the full scan -> identical counts at
// 1/35 of the work.  A sphere whose y is not monotone (any other `sphere` array) falls back to the full scan.
// positions of the sub-sampled survivors, one warp per sample: sample j is survivor r_j = (off + j * stride) mod count
// (the same bijection as rot_hist_kernel), found as the r_j-th set byte of the mask -- block by binary search over the
// exclusive block offsets, then the warp scans the block 512 bytes at a time (popc of 0/1 bytes + warp prefix)
// Sub-sample of the survivors (nocs/inference.py:277-281 shuffles them and keeps 10 000): sample j is survivor
// r_j = (off + j * stride) mod count with stride an odd number next to count / golden ratio that is coprime with count --
// a bijection of [0, count) whose first m terms are spread evenly over the whole survivor list for every m (a Kronecker
// sequence).  A fixed prime stride is not: once count exceeds it, the samples fall into count / stride short runs of
// adjacent survivors, i.e. into pairs anchored in a few patches of the cloud.
__device__ __forceinline__ unsigned long long subsample_stride(unsigned long long count) {
    if (count < 3ull) return 1ull;
    unsigned long long s = (unsigned long long)((double)count * 0.6180339887498949) | 1ull;
    while (true) {
        unsigned long long a = count, b = s;
        while (b != 0ull) {
            const unsigned long long t = a % b;
            a = b;
            b = t;
        }
        if (a == 1ull) return s;
        s += 2ull;
    }
}

__global__ void __launch_bounds__(256) select_samples_kernel(const uint8_t* __restrict__ mask,
                                                             const long long* __restrict__ block_offsets, int n_blocks,
                                                             long long n_pairs, const long long* __restrict__ count_ptr,
                                                             long long max_samples, unsigned long long offset_seed,
                                                             long long* __restrict__ out_pos) {
    const long long count = *count_ptr;
    const long long m = count < max_samples ? count : max_samples;
    const int lane = threadIdx.x & 31;
    const unsigned long long stride = m == count ? 1ull : subsample_stride((unsigned long long)count);
    const unsigned long long off = m == count ? 0ull : offset_seed % (unsigned long long)(count > 0 ? count : 1);
    for (long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < m;
         j += (long long)gridDim.x * (blockDim.x >> 5)) {
        const long long r = (long long)((off + (unsigned long long)j * stride) % (unsigned long long)count);
        int lo = 0, hi = n_blocks - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (block_offsets[mid] <= r) lo = mid; else hi = mid - 1;
        }
        const long long base = (long long)lo * kSelectBlockBytes;
        int need = (int)(r - block_offsets[lo]);
        long long found = n_pairs - 1;
        for (int c = 0; c < kSelectBlockBytes; c += 512) {
            const long long q = base + c + lane * 16;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (q + 16 <= n_pairs) {
                w = __ldg(reinterpret_cast<const uint4*>(mask + q));
            } else {
                unsigned char b[16];
                for (int t = 0; t < 16; ++t) b[t] = q + t < n_pairs ? mask[q + t] : 0;
                w = make_uint4(b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24, b[4] | b[5] << 8 | b[6] << 16 | b[7] << 24,
                               b[8] | b[9] << 8 | b[10] << 16 | b[11] << 24, b[12] | b[13] << 8 | b[14] << 16 | b[15] << 24);
            }
            const int cnt = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (need >= total) {
                need -= total;
                continue;
            }
            const int excl = incl - cnt;
            const bool mine = need >= excl && need < incl;              // exactly one lane
            if (mine) {
                int k = need - excl;
                const unsigned words[4] = {w.x, w.y, w.z, w.w};
                for (int t = 0; t < 16; ++t)
                    if ((words[t >> 2] >> ((t & 3) * 8)) & 1u) {
                        if (k-- == 0) {
                            found = q + t;
                            break;
                        }
                    }
            }
            const unsigned src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            found = __shfl_sync(0xffffffffu, found, src);
            break;
        }
        if (lane == 0) out_pos[j] = found;
    }
}

struct RotHistParams {
    const float2* rot_tab;
    const float* points;
    const uint8_t* bins;
    const float* lut;            // up angles lut[64:100], right angles lut[100:136]
    const void* idx;
    const long long* pos;        // compacted survivor pair positions, or nullptr: survivor r = r-th set byte of `mask`
    int pos_is_sample;           //   1: pos[j] already is the j-th sub-sampled survivor (select_samples_kernel)
    const uint8_t* mask;         //   found through the exclusive per-2048-pair block offsets of cppf_compact_count
    const long long* block_offsets;
    int n_blocks;
    long long n_pairs;
    const long long* count;      // number of survivors (device)
    const float* sphere;         // [n_bins][3]
    float* counts;               // [n_bins] float (exact integers)
    int n_points, n_rots, n_bins, which;   // which: 0 up, 1 right
    long long max_samples;
    unsigned long long offset_seed;
    float thr;
    float ywin;                  // half-width of the y window (>= 2: scan every bin)
    int row0;                    // dense mode: first row of the block of the pair matrix the pair positions refer to
};

constexpr int kRotHistPairs = 64;
constexpr int kRotHistThreads = 512;

// position of the r-th (0-based) surviving pair, in pair order: block by binary search over the exclusive block
// offsets, then a scan of that block's 0/1 mask bytes, 16 at a time
__device__ __forceinline__ long long select_survivor(const uint8_t* __restrict__ mask, const long long* __restrict__ block_offsets,
                                                     int n_blocks, long long n_pairs, long long r) {
    int lo = 0, hi = n_blocks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (block_offsets[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const long long base = (long long)lo * kSelectBlock;
    int need = (int)(r - block_offsets[lo]);
    for (int c = 0; c < kSelectBlock; c += 16) {
        const long long q = base + c;
        if (q + 16 <= n_pairs) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(mask + q));
            const int cnt = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
            if (need >= cnt) {
                need -= cnt;
                continue;
            }
        }
        for (int b = 0; b < 16 && q + b < n_pairs; ++b)
            if (mask[q + b] && need-- == 0) return q + b;
    }
    return n_pairs - 1;      // unreachable for r < count
}

template <bool IDX64>
__global__ void __launch_bounds__(kRotHistThreads) rot_hist_kernel(const RotHistParams prm) {
    extern __shared__ __align__(16) unsigned char rh_smem[];
    float4* s_sph = reinterpret_cast<float4*>(rh_smem);                       // [n_bins] (x, y, z, -)
    unsigned* s_cnt = reinterpret_cast<unsigned*>(s_sph + prm.n_bins);         // [n_bins]
    __shared__ float s_frame[kRotHistPairs][12];
    const long long count = *prm.count;
    const long long m = count < prm.max_samples ? count : prm.max_samples;
    const long long j0 = (long long)blockIdx.x * kRotHistPairs;
    if (j0 >= m) return;
    const unsigned long long stride = m == count ? 1ull : subsample_stride((unsigned long long)count);
    const unsigned long long off = m == count ? 0ull : prm.offset_seed % (unsigned long long)count;
    int mono = 1;
    for (int s = threadIdx.x; s < prm.n_bins; s += blockDim.x) {
        const float sy = __ldg(prm.sphere + 3 * s + 1);
        s_sph[s] = make_float4(__ldg(prm.sphere + 3 * s), sy, __ldg(prm.sphere + 3 * s + 2), 0.f);
        s_cnt[s] = 0u;
        if (s + 1 < prm.n_bins && !(__ldg(prm.sphere + 3 * (s + 1) + 1) < sy)) mono = 0;
    }
    if (threadIdx.x < kRotHistPairs) {
        float* fr = s_frame[threadIdx.x];
        fr[10] = 0.f;
        const long long j = j0 + threadIdx.x;
        if (j < m) {
            const long long r = (long long)((off + (unsigned long long)j * stride) % (unsigned long long)count);
            const long long p = prm.pos ? prm.pos[prm.pos_is_sample ? j : r]
                                        : select_survivor(prm.mask, prm.block_offsets, prm.n_blocks, prm.n_pairs, r);
            int ia, ib;
            pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib, prm.row0);
            const uchar4 bn = __ldg(reinterpret_cast<const uchar4*>(prm.bins) + p);
            const float rot = __ldg(prm.lut + (prm.which == 0 ? 64 + bn.z : 100 + bn.w));
            const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
            f3 ab, ex;
            if (pair_frame(a, b, ab, ex)) {
                const f3 y = cross3(ex, ab);
                fr[0] = ab.x; fr[1] = ab.y; fr[2] = ab.z;
                fr[3] = ex.x; fr[4] = ex.y; fr[5] = ex.z;
                fr[6] = y.x; fr[7] = y.y; fr[8] = y.z;
                fr[9] = tanf(rot);
                fr[10] = 1.f;
            }
        }
    }
    mono = __syncthreads_and(mono);
    const bool windowed = mono && prm.ywin < 2.f;
    const float2* tab = prm.rot_tab + prm.n_rots * (prm.n_rots - 1) / 2;
    const int total = kRotHistPairs * prm.n_rots;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int lp = t / prm.n_rots, i = t - lp * prm.n_rots;
        const float* fr = s_frame[lp];
        if (j0 + lp >= m) break;                           // lp is non-decreasing in t
        f3 up = {0.f, 0.f, 0.f};                           // degenerate pair: the zero vector, as rot_voting leaves it
        if (fr[10] != 0.f) {
            const float2 cs = __ldg(tab + i);
            const f3 ab = {fr[0], fr[1], fr[2]}, x = {fr[3], fr[4], fr[5]}, y = {fr[6], fr[7], fr[8]};
            const float tn = fr[9];
            const f3 o = circle_offset(x, y, cs.x, cs.y);
            const f3 axis = tn > 0.f ? ab : f3{-ab.x, -ab.y, -ab.z};
            up = o * tn + axis;
            up = up / (float)((double)len3(up) + 1e-7);
        }
        int lo = 0, hi = prm.n_bins;
        if (windowed) {
            const float y_hi = up.y + prm.ywin, y_lo = up.y - prm.ywin;
            int a0 = 0, a1 = prm.n_bins;                   // first s with y_s <= y_hi
            while (a0 < a1) {
                const int mid = (a0 + a1) >> 1;
                if (s_sph[mid].y > y_hi) a0 = mid + 1; else a1 = mid;
            }
            lo = a0;
            a1 = prm.n_bins;                               // first s with y_s < y_lo
            while (a0 < a1) {
                const int mid = (a0 + a1) >> 1;
                if (s_sph[mid].y >= y_lo) a0 = mid + 1; else a1 = mid;
            }
            hi = a0;
        }
        for (int s = lo; s < hi; ++s) {
            const float4 sp = s_sph[s];
            if (fmaf(up.z, sp.z, fmaf(up.y, sp.y, up.x * sp.x)) > prm.thr) atomicAdd(s_cnt + s, 1u);
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < prm.n_bins; s += blockDim.x) {
        const unsigned c = s_cnt[s];
        if (c) atomicAdd(prm.counts + s, (float)c);
    }
}

// ---------------------------------------------------------------------------
// Survivor statistics: out[0:3] = sum of log-scales, out[3] = count, out[4] = S_up, out[5] = S_right with
// S = sum_i aux_i * (2 t_i - 1), t_i = [flip(n_a) . best_dir > 0]  (nocs/inference.py:286-302,335).
// up_loss - down_loss of the reference's two BCE-with-logits means equals -2 S / count, so
// `down_loss < up_loss`  <=>  S < 0.
struct StatsParams {
    const float* points;
    const float* nrm;
    const float* tail;           // [5][n_pairs]
    const void* idx;
    const long long* pos;        // compacted survivor positions, or nullptr: every pair p with mask[p] != 0
    const uint8_t* mask;
    const long long* count;
    const float* sphere;
    const long long* best_up;    // argmax of the up histogram (device)
    const long long* best_right; // or nullptr
    double* out;                 // [6]
    int n_points;
    long long n_pairs;
    int vec4;                    // dense pairs through the mask with 16-byte aligned buffers and n_points % 4 == 0
    int row0;                    // dense mode: first row of the block of the pair matrix the pair positions refer to
};

template <bool IDX64>
__global__ void __launch_bounds__(256, CPPF_STATS_MIN_BLOCKS) survivor_stats_kernel(const StatsParams prm) {
    const long long count = prm.pos ? *prm.count : prm.n_pairs;      // mask mode walks every pair
    const long long bu = *prm.best_up;
    const f3 du = {__ldg(prm.sphere + 3 * bu), __ldg(prm.sphere + 3 * bu + 1), __ldg(prm.sphere + 3 * bu + 2)};
    f3 dr = {0.f, 0.f, 0.f};
    if (prm.best_right) {
        const long long br = *prm.best_right;
        dr = {__ldg(prm.sphere + 3 * br), __ldg(prm.sphere + 3 * br + 1), __ldg(prm.sphere + 3 * br + 2)};
    }
    // fp32 partial sums per thread (a thread sees <= a few hundred survivors), widened to double for the
    // cross-thread reduction; 4 survivors in flight per thread hide the pos -> tail gather latency
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long stride = (long long)gridDim.x * blockDim.x;
    // one survivor: nocs/inference.py:286-302 (flip the normal towards ab, sign of its component along the best axis)
    auto add_pair = [&](f3 a, f3 b, f3 n, float aux_u, float aux_r, float s0, float s1, float s2) {
        const f3 ab = a - b;
        const float inv = sqrtf(dot3(ab, ab)) + 1e-7f;                         // :288-289 (float32 numpy)
        const f3 abn = {ab.x / inv, ab.y / inv, ab.z / inv};
        if (dot3(n, abn) < 0.f) n = {-n.x, -n.y, -n.z};                        // :291-292
        acc[0] += s0; acc[1] += s1; acc[2] += s2;
        acc[3] += 1.f;
        acc[4] += aux_u * (dot3(n, du) > 0.f ? 1.f : -1.f);                    // :295
        if (prm.best_right) acc[5] += aux_r * (dot3(n, dr) > 0.f ? 1.f : -1.f);
    };
    if (prm.vec4) {
        // dense pairs addressed through the mask: 4 consecutive pairs per thread (same point a, 4 consecutive points b) so
        // that the mask, the five tail planes and the b coordinates arrive as 32-bit / 128-bit loads
        const long long n4 = prm.n_pairs >> 2;
        const float* t0 = prm.tail; const float* t1 = prm.tail + prm.n_pairs;
        const float* t2 = prm.tail + 2 * prm.n_pairs; const float* t3 = prm.tail + 3 * prm.n_pairs;
        const float* t4 = prm.tail + 4 * prm.n_pairs;
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
            const unsigned m4 = __ldg(reinterpret_cast<const unsigned*>(prm.mask) + q);
            if (m4 == 0u) continue;
            const long long p0 = q << 2;
            int ia, ib;
            pair_ab<IDX64>(nullptr, p0, prm.n_points, ia, ib, prm.row0);
            const f3 a = ld3(prm.points, ia), n = ld3(prm.nrm, ia);
            const float4* pb = reinterpret_cast<const float4*>(prm.points + 3 * (long long)ib);       // ib % 4 == 0: 16-byte aligned
            const float4 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
            const float4 au = __ldg(reinterpret_cast<const float4*>(t0 + p0));
            float4 ar = make_float4(0.f, 0.f, 0.f, 0.f);
            if (prm.best_right) ar = __ldg(reinterpret_cast<const float4*>(t1 + p0));
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(t2 + p0)), s1 = __ldg(reinterpret_cast<const float4*>(t3 + p0)),
                         s2 = __ldg(reinterpret_cast<const float4*>(t4 + p0));
            if (m4 & 0x000000ffu) add_pair(a, f3{b0.x, b0.y, b0.z}, n, au.x, ar.x, s0.x, s1.x, s2.x);
            if (m4 & 0x0000ff00u) add_pair(a, f3{b0.w, b1.x, b1.y}, n, au.y, ar.y, s0.y, s1.y, s2.y);
            if (m4 & 0x00ff0000u) add_pair(a, f3{b1.z, b1.w, b2.x}, n, au.z, ar.z, s0.z, s1.z, s2.z);
            if (m4 & 0xff000000u) add_pair(a, f3{b2.y, b2.z, b2.w}, n, au.w, ar.w, s0.w, s1.w, s2.w);
        }
    } else
    for (long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; j0 < count; j0 += 4 * stride) {
        long long p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long j = j0 + u * stride;
            p[u] = j < count ? (prm.pos ? prm.pos[j] : (prm.mask[j] ? j : -1)) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p[u] < 0) continue;
            int ia, ib;
            pair_ab<IDX64>(prm.idx, p[u], prm.n_points, ia, ib, prm.row0);
            const f3 a = ld3(prm.points, ia), b = ld3(prm.points, ib);
            const f3 ab = a - b;
            const float inv = sqrtf(dot3(ab, ab)) + 1e-7f;                     // :288-289 (float32 numpy)
            const f3 abn = {ab.x / inv, ab.y / inv, ab.z / inv};
            f3 n = ld3(prm.nrm, ia);
            if (dot3(n, abn) < 0.f) n = {-n.x, -n.y, -n.z};                    // :291-292
            const float tu = dot3(n, du) > 0.f ? 1.f : -1.f;                   // :295
            acc[0] += __ldg(prm.tail + 2 * prm.n_pairs + p[u]);
            acc[1] += __ldg(prm.tail + 3 * prm.n_pairs + p[u]);
            acc[2] += __ldg(prm.tail + 4 * prm.n_pairs + p[u]);
            acc[3] += 1.f;
            acc[4] += __ldg(prm.tail + p[u]) * tu;
            if (prm.best_right) acc[5] += __ldg(prm.tail + prm.n_pairs + p[u]) * (dot3(n, dr) > 0.f ? 1.f : -1.f);
        }
    }
    __shared__ double s[8][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double v = (double)acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += s[w][threadIdx.x];
        atomicAdd(prm.out + threadIdx.x, v);
    }
}

}  // namespace cppf

using namespace cppf;

extern "C" int64_t cppf_vote_scratch_bytes(int gx, int gy, int gz) { return (int64_t)gx * gy * gz * 8; }

static size_t vote_private_fixed_smem() {
    return (CPPF_VOTE_CONST_TAB ? 0 : (size_t)kRotTabP * 8) + 64 * 4 + (size_t)(kVoteThreads / 32) * kVoteQueue * 16 +
           (size_t)kVoteBatch * 2;
}

extern "C" int cppf_vote_private_max_cells(void) { return (int)((225 * 1024 - vote_private_fixed_smem() - 1024) / 4); }

// Host side of the conservative phase-1 bounds: d / res rounds to within 2^-24 of the quotient, so d below
// lo*res*(1 - 1e-6) can never reach lo (and likewise above the upper bounds).
static float vote_bound_below(float g, float res) { return nextafterf((float)((double)g * (double)res * (1.0 - 1e-6)), -INFINITY); }

namespace cppf {
// Launch of the privatised vote with the geometry either in the arguments (geom == nullptr) or in device
// memory (geom != nullptr: gx/gy/gz are ignored and max_cells sizes the shared-memory grid).
// slab_cells == 0: every CTA holds the whole grid (<= max_cells).  slab_cells > 0: "slab passes" for grids of up to
// sm_count() x-slabs -- every CTA holds one slab of <= slab_cells cells; scratch / grid then hold total_cells cells
// (host geometry: gx*gy*gz; device geometry: the caller's capacity) and the final conversion covers all of them.
int vote_fast_launch(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, const void* idx,
                     int idx_is_64, float* grid, void* scratch, const float* corner, float res, int n_points,
                     int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive, const Geom* geom, int max_cells,
                     cudaStream_t stream, int slab_cells, long long total_cells, int row0) {
    if (n_pairs <= 0) return 0;                     // empty pair list: nothing to vote (pointers may be null)
    const bool slabs = slab_cells > 0;
    long long cells = geom ? (long long)(slabs ? slab_cells : max_cells) : (long long)gx * gy * gz;
    int pps = 0, n_slabs = 1;
    if (slabs && !geom) {
        const long long gyz = (long long)gy * gz;
        pps = (int)(slab_cells / gyz) - 1;
        if (pps < 1) return (int)cudaErrorInvalidValue;
        if (pps > gx) pps = gx;
        n_slabs = (gx + pps - 1) / pps;
        if (n_slabs > sm_count()) return (int)cudaErrorInvalidValue;
        total_cells = cells;
        cells = (long long)(pps + 1 < gx ? pps + 1 : gx) * gyz;
    } else if (slabs) {
        pps = 1;                                     // recomputed in the kernel from the device geometry
    } else {
        total_cells = cells;
    }
    if (cells <= 0 || cells > cppf_vote_private_max_cells() || n_rots > kMaxRotsP || n_rots <= 0 || total_cells <= 0)
        return (int)cudaErrorInvalidValue;
    if ((mu_nu == nullptr) == (bins == nullptr)) return (int)cudaErrorInvalidValue;
    if (bins != nullptr && lut == nullptr) return (int)cudaErrorInvalidValue;
    if (!dense_rows_ok(idx, n_points, n_pairs, row0)) return (int)cudaErrorInvalidValue;
    if (n_pairs <= 0) return 0;
    int terr = 0;
    const float2* rot_tab = rot_table_device(stream, &terr);
    if (terr) return terr;
    if (!t_workspace_prepared) CPPF_RETURN_IF(cudaMemsetAsync(scratch, 0, (size_t)total_cells * 8, stream));
#if CPPF_VOTE_CONST_TAB
    {   // once per device: copy the table into this module's constant bank; the first call waits for the copy so that a
        // concurrent first call on another stream cannot run ahead of it
        static std::mutex mu;
        static bool ready[64] = {false};
        int dev = 0;
        CPPF_RETURN_IF(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !ready[dev]) {
            CPPF_RETURN_IF(cudaMemcpyToSymbolAsync(c_rot_tab, rot_tab, sizeof(float2) * kRotTabP, 0, cudaMemcpyDeviceToDevice, stream));
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            CPPF_RETURN_IF(cudaStreamIsCapturing(stream, &cap));
            if (cap == cudaStreamCaptureStatusNone) {
                CPPF_RETURN_IF(cudaStreamSynchronize(stream));
                if (dev < 64) ready[dev] = true;
            }
        }
    }
#endif
    const float lo = float_ceil_p(0.01);
    float hx = 0.f, hy = 0.f, hz = 0.f, dhx = 0.f, dhy = 0.f, dhz = 0.f;
    if (!geom) {
        hx = float_ceil_p((double)gx - 1.01), hy = float_ceil_p((double)gy - 1.01), hz = float_ceil_p((double)gz - 1.01);
        auto above = [&](float g) { return nextafterf((float)((double)g * (double)res * (1.0 + 1e-6)), INFINITY); };
        dhx = above(hx), dhy = above(hy), dhz = above(hz);
    }
    VotePParams prm{rot_tab, points, mu_nu, bins, lut, idx, reinterpret_cast<unsigned long long*>(scratch), corner, res,
                    (float)(1.0 / (double)res), lo, hx, hy, hz, vote_bound_below(lo, res), dhx, dhy, dhz, n_points,
                    (long long)n_pairs, n_rots, gx, gy, gz, adaptive, geom, (int)cells, 0, row0, pps, n_slabs};
    // second replica 8 banks away from the first, when both fit
    long long rep_stride = ((cells + 31) & ~31ll) + 8;
    if ((size_t)(rep_stride + cells) * 4 + vote_private_fixed_smem() + 1024 > (size_t)225 * 1024) rep_stride = 0;
    prm.rep_stride = (int)rep_stride;
    const size_t smem = vote_private_fixed_smem() + (size_t)(rep_stride ? rep_stride + cells : cells) * 4;
    const int threads = kVoteThreads;
    long long blocks = (n_pairs + kVoteBatch - 1) / kVoteBatch;
    if (blocks > sm_count()) blocks = sm_count();
    if (slabs) blocks = geom ? sm_count() : (long long)n_slabs * (sm_count() / n_slabs);
    void (*kern)(const VotePParams);
    if (slabs) {
        if (bins) kern = idx_is_64 ? vote_private_kernel<true, true, true> : vote_private_kernel<false, true, true>;
        else kern = idx_is_64 ? vote_private_kernel<true, false, true> : vote_private_kernel<false, false, true>;
    } else {
        if (bins) kern = idx_is_64 ? vote_private_kernel<true, true, false> : vote_private_kernel<false, true, false>;
        else kern = idx_is_64 ? vote_private_kernel<true, false, false> : vote_private_kernel<false, false, false>;
    }
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)kern, (int)smem));
    kern<<<(int)blocks, threads, smem, stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    vote_finalize_kernel<<<(int)((total_cells + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<unsigned long long*>(scratch), grid, (int)total_cells, geom, geom ? (slabs ? 1 : 0) : -1);
    CPPF_LAUNCH_CHECK();
    return 0;
}

int vote_finalize_launch(const unsigned long long* acc, float* grid, int cells, const Geom* geom, int only_mode,
                         cudaStream_t stream) {
    vote_finalize_kernel<<<(cells + 255) / 256, 256, 0, stream>>>(acc, grid, cells, geom, only_mode);
    CPPF_LAUNCH_CHECK();
    return 0;
}

int backvote_bins_launch(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                         uint8_t* out_mask, const float* corner, const int64_t* argmax_flat, float res, float tol,
                         int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, const Geom* geom,
                         cudaStream_t stream, double res_host, int row0) {
    if (n_pairs <= 0) return 0;
    if (n_rots > kMaxRotsP || n_rots <= 0) return (int)cudaErrorInvalidValue;
    if (!dense_rows_ok(idx, n_points, n_pairs, row0)) return (int)cudaErrorInvalidValue;
    int terr = 0;
    const float2* rot_tab = rot_table_device(stream, &terr);
    if (terr) return terr;
    BackvotePParams prm{rot_tab, points, bins, lut, idx, out_mask, corner, reinterpret_cast<const long long*>(argmax_flat), res,
                        (float)(1.0 / (double)res), tol, (float)(gx - 1), (float)(gy - 1), (float)(gz - 1), n_points,
                        (long long)n_pairs, n_rots, gx, gy, gz, geom, row0, res_host > 0.0 ? res_host : (double)res};
    long long blocks = (n_pairs + CPPF_BV_THREADS - 1) / CPPF_BV_THREADS;
    const long long cap = (long long)sm_count() * CPPF_BV_BLOCKS_PER_SM;
    if (blocks > cap) blocks = cap;
    if (idx_is_64) backvote_bins_kernel<true><<<(int)blocks, CPPF_BV_THREADS, 0, stream>>>(prm);
    else backvote_bins_kernel<false><<<(int)blocks, CPPF_BV_THREADS, 0, stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}
}  // namespace cppf

extern "C" int cppf_vote_fast(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut,
                              const void* idx, int idx_is_64, float* grid, void* scratch, const float* corner, float res,
                              int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive,
                              void* stream_) {
    if (n_pairs <= 0) return 0;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    return vote_fast_launch(points, mu_nu, bins, lut, idx, idx_is_64, grid, scratch, corner, res, n_points, n_pairs, n_rots,
                            gx, gy, gz, adaptive, nullptr, 0, (cudaStream_t)stream_, 0, 0, 0);
}

extern "C" int cppf_vote_fast_rows(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, int row0,
                                   float* grid, void* scratch, const float* corner, float res, int n_points, int64_t n_pairs,
                                   int n_rots, int gx, int gy, int gz, int adaptive, void* stream_) {
    return vote_fast_launch(points, mu_nu, bins, lut, nullptr, 0, grid, scratch, corner, res, n_points, n_pairs, n_rots, gx, gy,
                            gz, adaptive, nullptr, 0, (cudaStream_t)stream_, 0, 0, row0);
}

extern "C" int cppf_vote_finalize(const void* acc, float* grid, int64_t cells, void* stream_) {
    if (cells <= 0) return 0;
    if (acc == nullptr || grid == nullptr || cells > 0x7fffffffll) return (int)cudaErrorInvalidValue;
    return vote_finalize_launch(reinterpret_cast<const unsigned long long*>(acc), grid, (int)cells, nullptr, -1,
                                (cudaStream_t)stream_);
}

extern "C" int cppf_vote_slabs_supported(int gx, int gy, int gz) {
    const long long gyz = (long long)gy * gz;
    const long long pps = cppf_vote_private_max_cells() / gyz - 1;
    if (pps < 1) return 0;
    return (gx + pps - 1) / pps <= sm_count() ? 1 : 0;
}

extern "C" int cppf_vote_slabs(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut,
                               const void* idx, int idx_is_64, float* grid, void* scratch, const float* corner, float res,
                               int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive,
                               void* stream_) {
    if (n_pairs <= 0) return 0;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    return vote_fast_launch(points, mu_nu, bins, lut, idx, idx_is_64, grid, scratch, corner, res, n_points, n_pairs, n_rots,
                            gx, gy, gz, adaptive, nullptr, 0, (cudaStream_t)stream_, cppf_vote_private_max_cells(), 0, 0);
}

extern "C" int cppf_backvote_bins(const float* points, const uint8_t* bins, const float* lut, const void* idx,
                                  int idx_is_64, uint8_t* out_mask, const float* corner, const int64_t* argmax_flat,
                                  float res, float tol, double res_host, int n_points, int64_t n_pairs, int n_rots, int gx,
                                  int gy, int gz, void* stream_) {
    if (n_pairs <= 0) return 0;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    return backvote_bins_launch(points, bins, lut, idx, idx_is_64, out_mask, corner, argmax_flat, res, tol, n_points, n_pairs,
                                n_rots, gx, gy, gz, nullptr, (cudaStream_t)stream_, res_host, 0);
}

extern "C" int cppf_backvote_bins_rows(const float* points, const uint8_t* bins, const float* lut, int row0, uint8_t* out_mask,
                                       const float* corner, const int64_t* argmax_flat, float res, float tol, double res_host,
                                       int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, void* stream_) {
    return backvote_bins_launch(points, bins, lut, nullptr, 0, out_mask, corner, argmax_flat, res, tol, n_points, n_pairs, n_rots,
                                gx, gy, gz, nullptr, (cudaStream_t)stream_, res_host, row0);
}

namespace cppf {
int rot_hist_launch(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                    const int64_t* pos, const uint8_t* mask, const int64_t* block_offsets, int64_t n_pairs, const int64_t* count,
                    const float* sphere, float* counts, int n_points, int n_rots, int n_bins, int which, int64_t max_samples,
                    uint64_t offset_seed, float thr, cudaStream_t stream, int pos_is_sample = 0, int row0 = 0) {
    if (n_rots > kMaxRotsP || n_rots <= 0 || max_samples <= 0 || n_bins <= 0) return (int)cudaErrorInvalidValue;
    if (pos == nullptr && (mask == nullptr || block_offsets == nullptr)) return (int)cudaErrorInvalidValue;
    int terr = 0;
    const float2* rot_tab = rot_table_device(stream, &terr);
    if (terr) return terr;
    // half-width of the y window: |c - s|^2 < |c|^2 + |s|^2 - 2 thr with |c|, |s| <= 1 + 2e-6, plus slack
    const float ywin = thr > 0.f ? sqrtf(fmaxf(0.f, 2.00002f - 2.f * thr)) + 1e-4f : 4.f;
    const int n_blocks = (int)((n_pairs + kSelectBlock - 1) / kSelectBlock);
    RotHistParams prm{rot_tab, points, bins, lut, idx, reinterpret_cast<const long long*>(pos), pos_is_sample, mask,
                      reinterpret_cast<const long long*>(block_offsets), n_blocks, (long long)n_pairs,
                      reinterpret_cast<const long long*>(count), sphere, counts, n_points, n_rots, n_bins, which,
                      (long long)max_samples, offset_seed, thr, ywin, row0};
    const long long blocks = (max_samples + kRotHistPairs - 1) / kRotHistPairs;
    if (blocks > 0x7FFFFFFF) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)n_bins * 20;
    if (smem > 160 * 1024) return (int)cudaErrorInvalidValue;
    auto kern = idx_is_64 ? rot_hist_kernel<true> : rot_hist_kernel<false>;
    if (smem > 40 * 1024) CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)kern, (int)smem));
    kern<<<(int)blocks, kRotHistThreads, smem, stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}

int survivor_stats_launch(const float* points, const float* nrm, const float* tail, const void* idx, int idx_is_64,
                          const int64_t* pos, const uint8_t* mask, const int64_t* count, const float* sphere,
                          const int64_t* best_up, const int64_t* best_right, double* out, int n_points, int64_t n_pairs,
                          cudaStream_t stream, int row0 = 0) {
    if (pos == nullptr && mask == nullptr) return (int)cudaErrorInvalidValue;
    if (!t_workspace_prepared) CPPF_RETURN_IF(cudaMemsetAsync(out, 0, 6 * sizeof(double), stream));
    const bool aligned = (((uintptr_t)points | (uintptr_t)tail | (uintptr_t)mask) & 15u) == 0;
    const int vec4 = (pos == nullptr && idx == nullptr && (n_points & 3) == 0 && aligned) ? 1 : 0;
    StatsParams prm{points, nrm, tail, idx, reinterpret_cast<const long long*>(pos), mask,
                    reinterpret_cast<const long long*>(count), sphere, reinterpret_cast<const long long*>(best_up),
                    reinterpret_cast<const long long*>(best_right), out, n_points, (long long)n_pairs, vec4, row0};
    const int blocks = sm_count() * CPPF_STATS_BLOCKS_PER_SM;
    if (idx_is_64) survivor_stats_kernel<true><<<blocks, 256, 0, stream>>>(prm);
    else survivor_stats_kernel<false><<<blocks, 256, 0, stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}
}  // namespace cppf

extern "C" int cppf_rot_hist(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                             const int64_t* pos, const int64_t* count, const float* sphere, float* counts, int n_points,
                             int n_rots, int n_bins, int which, int64_t max_samples, uint64_t offset_seed, float thr,
                             void* stream_) {
    if (pos == nullptr) return (int)cudaErrorInvalidValue;
    return rot_hist_launch(points, bins, lut, idx, idx_is_64, pos, nullptr, nullptr, 0, count, sphere, counts, n_points, n_rots,
                           n_bins, which, max_samples, offset_seed, thr, (cudaStream_t)stream_);
}

extern "C" int cppf_rot_hist_rows(const float* points, const uint8_t* bins, const float* lut, int row0, const int64_t* pos,
                                  const int64_t* count, const float* sphere, float* counts, int n_points, int n_rots, int n_bins,
                                  int which, int64_t max_samples, uint64_t offset_seed, float thr, void* stream_) {
    if (pos == nullptr || row0 < 0 || row0 >= n_points) return (int)cudaErrorInvalidValue;
    return rot_hist_launch(points, bins, lut, nullptr, 0, pos, nullptr, nullptr, 0, count, sphere, counts, n_points, n_rots,
                           n_bins, which, max_samples, offset_seed, thr, (cudaStream_t)stream_, 0, row0);
}

extern "C" int cppf_rot_hist_mask(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                                  const uint8_t* mask, const int64_t* block_offsets, int64_t n_pairs, const int64_t* count,
                                  const float* sphere, float* counts, int n_points, int n_rots, int n_bins, int which,
                                  int64_t max_samples, uint64_t offset_seed, float thr, int64_t* sample_scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (mask == nullptr || block_offsets == nullptr || max_samples <= 0) return (int)cudaErrorInvalidValue;
    if (sample_scratch == nullptr)                  // no room for the sample list: every thread selects its own survivor
        return rot_hist_launch(points, bins, lut, idx, idx_is_64, nullptr, mask, block_offsets, n_pairs, count, sphere, counts,
                               n_points, n_rots, n_bins, which, max_samples, offset_seed, thr, stream);
    const int n_blocks = (int)((n_pairs + kSelectBlockBytes - 1) / kSelectBlockBytes);
    long long blocks = (max_samples + 7) / 8;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    select_samples_kernel<<<(int)blocks, 256, 0, stream>>>(mask, reinterpret_cast<const long long*>(block_offsets), n_blocks,
                                                          (long long)n_pairs, reinterpret_cast<const long long*>(count),
                                                          (long long)max_samples, offset_seed,
                                                          reinterpret_cast<long long*>(sample_scratch));
    CPPF_LAUNCH_CHECK();
    return rot_hist_launch(points, bins, lut, idx, idx_is_64, sample_scratch, nullptr, nullptr, n_pairs, count, sphere, counts,
                           n_points, n_rots, n_bins, which, max_samples, offset_seed, thr, stream, 1);
}

extern "C" int cppf_survivor_stats(const float* points, const float* nrm, const float* tail, const void* idx,
                                   int idx_is_64, const int64_t* pos, const int64_t* count, const float* sphere,
                                   const int64_t* best_up, const int64_t* best_right, double* out, int n_points,
                                   int64_t n_pairs, void* stream_) {
    if (pos == nullptr) return (int)cudaErrorInvalidValue;
    return survivor_stats_launch(points, nrm, tail, idx, idx_is_64, pos, nullptr, count, sphere, best_up, best_right, out,
                                 n_points, n_pairs, (cudaStream_t)stream_);
}

extern "C" int cppf_survivor_stats_rows(const float* points, const float* nrm, const float* tail, int row0, const int64_t* pos,
                                        const int64_t* count, const float* sphere, const int64_t* best_up,
                                        const int64_t* best_right, double* out, int n_points, int64_t n_pairs, void* stream_) {
    if (pos == nullptr || !dense_rows_ok(nullptr, n_points, n_pairs, row0)) return (int)cudaErrorInvalidValue;
    return survivor_stats_launch(points, nrm, tail, nullptr, 0, pos, nullptr, count, sphere, best_up, best_right, out, n_points,
                                 n_pairs, (cudaStream_t)stream_, row0);
}

extern "C" int cppf_survivor_stats_mask(const float* points, const float* nrm, const float* tail, const void* idx,
                                        int idx_is_64, const uint8_t* mask, const float* sphere, const int64_t* best_up,
                                        const int64_t* best_right, double* out, int n_points, int64_t n_pairs, void* stream_) {
    return survivor_stats_launch(points, nrm, tail, idx, idx_is_64, nullptr, mask, nullptr, sphere, best_up, best_right, out,
                                 n_points, n_pairs, (cudaStream_t)stream_);
}


// Centre voting (models/voting.py:8-66) for grids that do not fit one SM's shared memory but do fit eight:
// the 64^3 grid of BASELINE config 3 is 1 MB, a B200 SM holds ~176 KB of u32 cells.
//
// Measured alternatives on B200 (profiles/r1_atomics_microbench.json, profiles/r1c_dsmem_atomics_microbench.json):
// global fp32/u32 reductions sustain 90-180 G/s, reductions into a grid distributed over a cluster's shared
// memories (DSMEM) 50 G/s (cluster of 8) to 186 G/s (cluster of 2) -- both an order of magnitude below the
// ~2000 G/s of a CTA's own shared memory.  So the grid is cut into <= 8 x-slabs that each fit one SM, and the
// candidates are ROUTED to their slab through HBM instead of the atomics being routed to the grid:
//
//   route_kernel      phase 1 of vote_private_kernel (pairs sorted by rotation count, lane = pair): every
//                     in-bounds candidate (exact test of models/voting.py:35-39) is appended, as its grid
//                     coordinates (16 B), to the queue of the slab that owns floor(g.x); a warp stages 32
//                     candidates per slab in shared memory and writes them as one 512-byte row into 128-entry
//                     chunks it reserves from a global pool (one atomic per chunk);
//   slab_splat_kernel CTAs are dealt to the slabs in proportion to their chunk counts; a CTA keeps its slab
//                     (+ one overlap plane for the x+1 corners) as fixed-point u32 cells in shared memory, streams
//                     the chunks of its slab (coalesced) and does the 8 trilinear atomics per candidate there.
//
// HBM traffic: 32 B per in-bounds candidate (written once, read once) -- this is the one HBM-bound stage of the
// path.  Sums are exact integers, so the result is identical to vote_private_kernel's and deterministic.
#include "common.cuh"
#include "vote_common.cuh"

#include "../../include/cppf_b200.h"

namespace cppf {

int vote_finalize_launch(const unsigned long long* acc, float* grid, int cells, const Geom* geom, int only_mode,
                         cudaStream_t stream);

constexpr int kChunk = 512;                          // candidates per chunk (8 KB)
constexpr int kRouteWarps = 32;
constexpr int kRouteThreads = kRouteWarps * 32;
constexpr int kRouteBatch = 2 * kRouteThreads;       // pairs sorted between two block barriers
constexpr int kRouteKeys = kMaxRotsP + 1;
constexpr int kRTileA = 128, kRTileB = 16;           // dense mode: kRTileA * kRTileB == kRouteBatch
static_assert(kRTileA * kRTileB == kRouteBatch, "tile = batch");
constexpr int kStage = 64;                           // per-warp ring of in-bounds candidates (float4 slots)

struct RouteCounters {
    unsigned n_chunks;               // chunks reserved so far
    unsigned overflow;               // set when the pool ran out (cannot happen with the host's super-batch sizing)
    unsigned slab_chunks[kMaxSlabs];
};

struct RouteParams {
    const float2* rot_tab;
    const float* points;
    const float* mu_nu;
    const uint8_t* bins;
    const float* lut;
    const void* idx;
    const float* corner;
    float4* pool;                    // [max_chunks][kChunk]
    uint8_t* chunk_slab;             // [max_chunks]
    RouteCounters* counters;
    float res, inv_res;
    float lo, hx, hy, hz;            // exact bounds on g (models/voting.py:36-39)
    float dlo, dhx, dhy, dhz;        // conservative bounds on candidate - corner
    int n_points;
    long long batch_begin, batch_end;  // this pass, in 2048-pair batches (dense mode: 128 x 16 tiles of the pair matrix)
    long long n_pairs;
    int n_rots, adaptive;
    int gx, planes_per_slab, n_slabs;
    unsigned max_chunks;
    const Geom* geom;                // optional: device-side geometry (mode 1) overrides corner / dims / bounds / slabs
};

// per-warp, per-slab write cursor (shared memory): the open chunk of the warp for that slab
struct SlabState {
    unsigned base;                   // first pool entry of the open chunk; kNoChunk = none open (or pool exhausted)
    int used;                        // entries written into it
};
constexpr unsigned kNoChunk = 0xFFFFFFFFu;

template <bool IDX64, bool BINS>
__global__ void __launch_bounds__(kRouteThreads, 1) route_kernel(const RouteParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tab = reinterpret_cast<float2*>(smem_raw);
    float* s_lut = reinterpret_cast<float*>(s_tab + kRotTabP);
    float4* s_queue = reinterpret_cast<float4*>(s_lut + 64);                    // [warp][kStage]: in-bounds candidates (d)
    unsigned short* s_perm = reinterpret_cast<unsigned short*>(s_queue + kRouteWarps * kStage);
    __shared__ int s_hist[kRouteKeys + 3], s_start[kRouteKeys + 3], s_nlut[32];
    __shared__ int s_total, s_next;
    __shared__ unsigned char s_slab_of_x[1024];
    __shared__ SlabState s_state[kRouteWarps][kMaxSlabs];
    int pps = prm.planes_per_slab, n_slabs = prm.n_slabs;
    float hx = prm.hx, hy = prm.hy, hz = prm.hz, dhx = prm.dhx, dhy = prm.dhy, dhz = prm.dhz;
    const float* corner = prm.corner;
    if (prm.geom != nullptr) {
        const Geom g = *prm.geom;
        if (g.status != 0 || g.mode != 1) return;
        pps = g.planes_per_slab; n_slabs = g.n_slabs;
        hx = g.hx; hy = g.hy; hz = g.hz; dhx = g.dhx; dhy = g.dhy; dhz = g.dhz;
        corner = prm.geom->corner;
    }
    for (int i = threadIdx.x; i < kRotTabP; i += blockDim.x) s_tab[i] = __ldg(prm.rot_tab + i);
    if (BINS && threadIdx.x < 64) s_lut[threadIdx.x] = __ldg(prm.lut + threadIdx.x);
    if (BINS && threadIdx.x < 32) {
        int n = prm.n_rots;
        if (prm.adaptive) n = adaptive_rots(__ldg(prm.lut + 32 + threadIdx.x), prm.res, prm.n_rots);   // :31
        s_nlut[threadIdx.x] = n < 0 ? 0 : n;
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const int s = i / pps;
        s_slab_of_x[i] = (unsigned char)(s < n_slabs ? s : n_slabs - 1);
    }
    if (threadIdx.x < kRouteKeys + 3) s_hist[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < kRouteWarps * kMaxSlabs; i += blockDim.x) {
        SlabState* st = &s_state[0][0] + i;
        st->base = kNoChunk;
        st->used = 0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned q_addr = (unsigned)__cvta_generic_to_shared(s_queue + warp * kStage);
    SlabState* my_state = s_state[warp];
    const float cx = __ldg(corner), cy = __ldg(corner + 1), cz = __ldg(corner + 2);

    const float4 pad = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);      // NaN x = padding entry
    // close the warp's open chunk of slab s (warp-uniform s): NaN-pad its unused tail so that chunks are dense
    auto close_chunk = [&](int s) {
        const SlabState st = my_state[s];
        if (st.base != kNoChunk)
            for (int e = st.used + lane; e < kChunk; e += 32) prm.pool[(size_t)st.base + e] = pad;
        __syncwarp();
        if (lane == 0) my_state[s] = SlabState{kNoChunk, 0};
        __syncwarp();
    };
    // 32 queued candidates (d = candidate - corner): the reference's `/ res`, its exact in-bounds test, then every
    // survivor is written straight into the open chunk of the slab that owns floor(g.x) -- lanes of the same slab get
    // consecutive entries (match_any rank), so a call writes a few contiguous runs of 16-byte entries
    auto route32 = [&](unsigned q_head, int count) {
        bool ok = lane < count;
        float gxf = 0.f, gyf = 0.f, gzf = 0.f;
        if (ok) {
            const float4 d = ld_shared_f4(q_addr + (((q_head + lane) & (kStage - 1)) << 4));
            gxf = div_by(d.x, prm.res, prm.inv_res);                           // :35
            gyf = div_by(d.y, prm.res, prm.inv_res);
            gzf = div_by(d.z, prm.res, prm.inv_res);
            ok = !(gxf < prm.lo || gyf < prm.lo || gzf < prm.lo || gxf >= hx || gyf >= hy || gzf >= hz);   // :36-39
        }
        const int slab = ok ? (int)s_slab_of_x[(int)gxf & 1023] : (kMaxSlabs + lane);    // misses: private keys
        const unsigned peers = __match_any_sync(0xffffffffu, slab);
        const bool leader = ok && (peers & lt_mask) == 0u;
        if (leader && my_state[slab].base == kNoChunk) {                       // open a chunk for this slab
            const unsigned id = atomicAdd(&prm.counters->n_chunks, 1u);
            if (id < prm.max_chunks) {
                prm.chunk_slab[id] = (uint8_t)slab;
                atomicAdd(&prm.counters->slab_chunks[slab], 1u);
                my_state[slab] = SlabState{id * (unsigned)kChunk, 0};
            } else {
                prm.counters->overflow = 1u;                                   // pool exhausted: the votes are dropped
            }
        }
        __syncwarp();
        if (ok) {
            const SlabState st = my_state[slab];
            if (st.base != kNoChunk) prm.pool[(size_t)st.base + st.used + __popc(peers & lt_mask)] = make_float4(gxf, gyf, gzf, 0.f);
        }
        __syncwarp();
        if (leader) my_state[slab].used += __popc(peers);
        __syncwarp();
        // a chunk that cannot take another 32 entries is closed now, so the next call never has to split a run
        bool full = false;
        if (lane < n_slabs) full = my_state[lane].base != kNoChunk && my_state[lane].used > kChunk - 32;
        unsigned todo = __ballot_sync(0xffffffffu, full);
        while (todo) {
            const int s = __ffs(todo) - 1;
            todo &= todo - 1u;
            close_chunk(s);
        }
    };

    // a batch: 2048 consecutive entries of the pair list, or -- dense mode -- a 128 x 16 tile of the pair matrix, like
    // vote_private_kernel (mixing the points a of a batch decorrelates the lanes' candidates)
    const bool tiled = prm.idx == nullptr;
    const int tiles_x = (prm.n_points + kRTileB - 1) / kRTileB;
    auto pair_index = [&](long long batch, int local) -> long long {      // -1: no such pair
        if (!tiled) {
            const long long q = batch * kRouteBatch + local;
            return q < prm.n_pairs ? q : -1;
        }
        const int tr = (int)(batch / tiles_x), tc = (int)(batch - (long long)tr * tiles_x);
        const int a = tr * kRTileA + (local / kRTileB), b = tc * kRTileB + (local % kRTileB);
        return (a < prm.n_points && b < prm.n_points) ? (long long)a * prm.n_points + b : -1;
    };
    for (long long batch = prm.batch_begin + blockIdx.x; batch < prm.batch_end; batch += gridDim.x) {
        int key[2], rank[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const long long p = pair_index(batch, threadIdx.x + j * kRouteThreads);
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
        __syncthreads();
        if (threadIdx.x < 32) {                      // exclusive scan of the histogram, largest n first
            int v[3], sum = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int e = 3 * lane + k;
                v[k] = e < kRouteKeys ? s_hist[kMaxRotsP - e] : 0;
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
                if (e < kRouteKeys) {
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
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (key[j] >= 0) s_perm[s_start[key[j]] + rank[j]] = (unsigned short)(threadIdx.x + j * kRouteThreads);
        __syncthreads();
        const int total = s_total;

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
                pair_ab<IDX64>(prm.idx, p, prm.n_points, ia, ib);
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
            const float2* tab = s_tab + (n > 0 ? n * (n - 1) / 2 : 0);
            const int n_max = __reduce_max_sync(0xffffffffu, n);
            for (int i = 0; i < n_max; ++i) {
                const float2 cs = tab[i];
                const f3 off = circle_offset(x, y, cs.x, cs.y);                            // :34
                const float dx = c.x + off.x - cx, dy = c.y + off.y - cy, dz = c.z + off.z - cz;
                const bool inb = i < n && dx >= prm.dlo && dy >= prm.dlo && dz >= prm.dlo && dx < dhx && dy < dhy && dz < dhz;
                const unsigned m = __ballot_sync(0xffffffffu, inb);
                if (inb) st_shared_f4(q_addr + (((q_tail + __popc(m & lt_mask)) & (kStage - 1)) << 4), dx, dy, dz);
                q_tail += __popc(m);
                if (q_tail - q_head >= 32u) {
                    __syncwarp();
                    route32(q_head, 32);
                    q_head += 32u;
                }
            }
        }
        __syncwarp();
        if (q_tail != q_head) route32(q_head, (int)(q_tail - q_head));
        __syncthreads();
    }
    for (int s = 0; s < n_slabs; ++s) close_chunk(s);      // NaN-pad whatever is still open
}

struct SplatParams {
    const float4* pool;
    const uint8_t* chunk_slab;
    const RouteCounters* counters;
    unsigned long long* acc;         // [gx*gy*gz] fixed-point accumulator (global)
    int gx, gy, gz, planes_per_slab, n_slabs;
    unsigned max_chunks;
    const Geom* geom;
    int max_slab_cells;              // capacity of the shared-memory slab of this launch
};

constexpr int kSplatThreads = 1024;

__global__ void __launch_bounds__(kSplatThreads, 1) slab_splat_kernel(const SplatParams prm) {
    extern __shared__ __align__(16) unsigned s_grid[];
    __shared__ int s_slab, s_j, s_m;
    int gx = prm.gx, gy = prm.gy, gzd = prm.gz, pps = prm.planes_per_slab, n_slabs = prm.n_slabs;
    if (prm.geom != nullptr) {
        const Geom g = *prm.geom;
        if (g.status != 0 || g.mode != 1 || (g.planes_per_slab + 1) * slab_plane_stride(g.gy, g.gz) > prm.max_slab_cells) return;
        gx = g.gx; gy = g.gy; gzd = g.gz; pps = g.planes_per_slab; n_slabs = g.n_slabs;
    }
    if (threadIdx.x == 0) {
        // deal the CTAs to the non-empty slabs in proportion to their chunk counts (every CTA computes the same table)
        unsigned cnt[kMaxSlabs], tot = 0;
        int share[kMaxSlabs], used = 0, big = 0;
        for (int s = 0; s < kMaxSlabs; ++s) {
            cnt[s] = s < n_slabs ? prm.counters->slab_chunks[s] : 0u;
            tot += cnt[s];
        }
        const int G = (int)gridDim.x;
        for (int s = 0; s < kMaxSlabs; ++s) {
            share[s] = cnt[s] == 0u ? 0 : max(1, (int)((unsigned long long)cnt[s] * G / (tot ? tot : 1u)));
            used += share[s];
            if (cnt[s] > cnt[big]) big = s;
        }
        while (used > G) {                      // more non-empty slabs than spare CTAs: take from the largest shares
            int t = 0;
            for (int s = 1; s < kMaxSlabs; ++s)
                if (share[s] > share[t]) t = s;
            if (share[t] <= 1) break;
            --share[t];
            --used;
        }
        if (used < G) share[big] += G - used;
        int c = (int)blockIdx.x, slab = -1, j = 0;
        for (int s = 0; s < kMaxSlabs && slab < 0; ++s) {
            if (c < share[s]) {
                slab = s;
                j = c;
            }
            c -= share[s];
        }
        s_slab = tot == 0u ? -1 : slab;
        s_j = j;
        s_m = slab >= 0 ? share[slab] : 1;
    }
    __syncthreads();
    const int slab = s_slab, j = s_j, m = s_m;
    if (slab < 0) return;
    const int gyz = gy * gzd, gz = gzd;
    const int rs = slab_row_stride(gz), ps = slab_plane_stride(gy, gz);        // padded strides of the slab in shared memory
    const int x0 = slab * pps;
    const int x1 = min(gx, x0 + pps + 1);                          // + the overlap plane of the x+1 corners
    const int cells = (x1 - x0) * ps;                               // padded cells (the padding stays zero)
    const unsigned s_grid_addr = (unsigned)__cvta_generic_to_shared(s_grid);
    // padded slab index -> flat index of the cell in the global grid, or -1 for a padding word
    auto global_cell = [&](int i) -> long long {
        const int x = i / ps, r = i - x * ps;
        const int y = r / rs, z = r - y * rs;
        return (y < gy && z < gz) ? (long long)(x0 + x) * gyz + (long long)y * gz + z : -1ll;
    };
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s_grid[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned n_chunks = min(prm.counters->n_chunks, prm.max_chunks);
    const float fx0 = (float)x0;
    __shared__ unsigned s_list[kSplatThreads];
    __shared__ unsigned s_count;
    // a round = 1024 chunk tags per CTA: the chunks of this slab are compacted into a list, then the warps take them
    // round-robin (every chunk is the same amount of work).  Overflow guard: `pending` bounds the votes any cell can have
    // received since the last scan (every listed candidate could hit the same cell); before a round could push a cell past
    // 2^32, the cells holding >= 2^28 units are flushed to the global accumulator.
    constexpr unsigned kSplatFlushAt = 1u << 28;
    constexpr unsigned kHeadroomVotes = (0xFFFFFFFFu >> kFixShift) - 1024u;
    unsigned pending = 0;
    for (unsigned long long r0 = 0;; ++r0) {
        const unsigned long long first = (r0 * (unsigned long long)m + (unsigned long long)j) * 1024ull;
        if (first >= n_chunks) break;
        if (threadIdx.x == 0) s_count = 0u;
        __syncthreads();
        const unsigned long long k = first + threadIdx.x;
        const bool mine = k < n_chunks && __ldg(prm.chunk_slab + k) == (uint8_t)slab;
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        unsigned wbase = 0u;
        if (lane == 0 && bal) wbase = atomicAdd(&s_count, (unsigned)__popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (mine) s_list[wbase + __popc(bal & ((1u << lane) - 1u))] = threadIdx.x;
        __syncthreads();
        const unsigned count = s_count;
        for (unsigned c0 = 0; c0 < count; c0 += 256u) {                       // <= 256 chunks (131 072 votes) between guard checks
            const unsigned c1 = min(count, c0 + 256u);
            if (pending + (c1 - c0) * (unsigned)kChunk > kHeadroomVotes) {
                __syncthreads();
                for (int i = threadIdx.x; i < cells; i += blockDim.x) {
                    const unsigned v = s_grid[i];
                    if (v >= kSplatFlushAt) {
                        atomicAdd(prm.acc + global_cell(i), (unsigned long long)v);
                        s_grid[i] = 0u;
                    }
                }
                __syncthreads();
                pending = kSplatFlushAt >> kFixShift;                          // what an unflushed cell may still hold
            }
            pending += (c1 - c0) * (unsigned)kChunk;
            for (unsigned c = c0 + warp; c < c1; c += kSplatThreads / 32) {
                const float4* ch = prm.pool + (first + s_list[c]) * kChunk;
                float4 g[4], nx[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) g[q] = __ldg(ch + q * 32 + lane);
#pragma unroll 1
                for (int part = 0; part < kChunk / 128; ++part) {
                    if (part + 1 < kChunk / 128) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) nx[q] = __ldg(ch + (part + 1) * 128 + q * 32 + lane);
                    }
                    // Lane l takes its four entries in the order (q + l) mod 4: the 32 entries splatted by one instruction then
                    // come from four different 32-entry runs of the chunk (8 lanes each).  Consecutive entries of a chunk are
                    // candidates of ONE route32 call -- the same 32 pairs at the same rotation step, which reach the vote peak
                    // together -- and splatting them side by side cost 5.3 wavefronts per ATOMS (same-cell collisions).
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = (q + lane) & 3;
                        const float4 e = k == 0 ? g[0] : (k == 1 ? g[1] : (k == 2 ? g[2] : g[3]));
                        if (e.x == e.x) splat_fixed(s_grid_addr, e.x - fx0, e.y, e.z, ps, rs);                   // NaN = padding
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) g[q] = nx[q];
                }
            }
        }
        // every warp is done with s_list / s_count of this round before the next round rewrites them (racecheck, round 2:
        // a warp still between the barrier above and its read of s_count could see the next round's reset)
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const unsigned v = s_grid[i];
        if (v) atomicAdd(prm.acc + global_cell(i), (unsigned long long)v);
    }
}

static size_t route_smem() {
    return (size_t)kRotTabP * 8 + 64 * 4 + (size_t)kRouteWarps * kStage * 16 + (size_t)kRouteBatch * 2;
}

struct RoutedPlan {
    int planes_per_slab, n_slabs;
};

long long slab_cap_cells() { return ((long long)220 * 1024) / 4; }      // u32 cells of one slab CTA (+ 4 KB static)

static bool routed_plan(int gx, int gy, int gz, RoutedPlan* pl) {
    const long long gyz = (long long)gy * gz;
    const long long planes = slab_cap_cells() / slab_plane_stride(gy, gz) - 1;    // minus the overlap plane
    if (planes < 1 || gx > 1024) return false;
    pl->planes_per_slab = (int)(planes < gx ? planes : gx);
    pl->n_slabs = (gx + pl->planes_per_slab - 1) / pl->planes_per_slab;
    return pl->n_slabs <= kMaxSlabs;
}

static int64_t routed_slack_chunks() { return (int64_t)sm_count() * kRouteWarps * kMaxSlabs + 64; }   // open chunks
static int64_t routed_min_pairs() { return (int64_t)sm_count() * kRouteBatch; }

// bytes of counters + chunk tags + pool for `pairs` pairs per pass
int64_t routed_pool_bytes(int64_t n_pairs, int n_rots) {
    int64_t pairs = n_pairs < (4ll << 20) ? n_pairs : (4ll << 20);
    if (pairs < routed_min_pairs()) pairs = routed_min_pairs();
    const int64_t chunks = (pairs * n_rots + kChunk - 1) / kChunk + routed_slack_chunks() + 1024;
    return 256 + ((chunks + 255) & ~255ll) + chunks * kChunk * 16 + 4096;
}

// Routed vote into `acc` (u64 fixed point, zeroed by the caller); geometry from the arguments (geom == nullptr) or
// from device memory (geom != nullptr, mode 1; the kernels return at once for any other mode).
int vote_routed_launch(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, const void* idx,
                       int idx_is_64, unsigned long long* acc, void* pool_mem, int64_t pool_bytes, const float* corner,
                       float res, int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive,
                       const Geom* geom, cudaStream_t stream) {
    if (n_pairs <= 0) return 0;
    RoutedPlan pl{1, 1};
    if (!geom && !routed_plan(gx, gy, gz, &pl)) return (int)cudaErrorInvalidValue;
    if (n_rots > kMaxRotsP || n_rots <= 0) return (int)cudaErrorInvalidValue;
    if ((mu_nu == nullptr) == (bins == nullptr)) return (int)cudaErrorInvalidValue;
    if (bins != nullptr && lut == nullptr) return (int)cudaErrorInvalidValue;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    if (n_pairs <= 0) return 0;
    // carve: counters | chunk tags | pool
    const int64_t slack = routed_slack_chunks();
    int64_t max_chunks = (pool_bytes - 256 - 4096) / (kChunk * 16 + 1) - 256;
    if (max_chunks > 33000000ll) max_chunks = 33000000ll;      // pool entry indices stay below 2^32
    if (max_chunks < slack + (routed_min_pairs() * n_rots + kChunk - 1) / kChunk) return (int)cudaErrorInvalidValue;
    unsigned char* base = reinterpret_cast<unsigned char*>(pool_mem);
    RouteCounters* counters = reinterpret_cast<RouteCounters*>(base);
    uint8_t* chunk_slab = base + 256;
    float4* pool = reinterpret_cast<float4*>(base + 256 + ((max_chunks + 255) & ~255ll));
    int64_t batch_pairs = (max_chunks - slack) * kChunk / n_rots;
    batch_pairs -= batch_pairs % kRouteBatch;
    if (batch_pairs < routed_min_pairs()) return (int)cudaErrorInvalidValue;

    int terr = 0;
    const float2* rot_tab = rot_table_device(stream, &terr);
    if (terr) return terr;
    const float lo = float_ceil_p(0.01);
    float hx = 0.f, hy = 0.f, hz = 0.f, dhx = 0.f, dhy = 0.f, dhz = 0.f;
    if (!geom) {
        hx = float_ceil_p((double)gx - 1.01), hy = float_ceil_p((double)gy - 1.01), hz = float_ceil_p((double)gz - 1.01);
        auto above = [&](float g) { return nextafterf((float)((double)g * (double)res * (1.0 + 1e-6)), INFINITY); };
        dhx = above(hx), dhy = above(hy), dhz = above(hz);
    }
    const float dlo = nextafterf((float)((double)lo * (double)res * (1.0 - 1e-6)), -INFINITY);
    void (*route)(const RouteParams);
    if (bins) route = idx_is_64 ? route_kernel<true, true> : route_kernel<false, true>;
    else route = idx_is_64 ? route_kernel<true, false> : route_kernel<false, false>;
    const size_t rsmem = route_smem();
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)route, (int)rsmem));
    const long long slab_cells = geom ? slab_cap_cells() : (long long)(pl.planes_per_slab + 1) * slab_plane_stride(gy, gz);
    const size_t ssmem = (size_t)slab_cells * 4;
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)slab_splat_kernel, (int)ssmem));
    const int64_t total_batches = idx == nullptr
                                      ? (int64_t)((n_points + kRTileA - 1) / kRTileA) * ((n_points + kRTileB - 1) / kRTileB)
                                      : (n_pairs + kRouteBatch - 1) / kRouteBatch;
    const int64_t pass_batches = batch_pairs / kRouteBatch;
    for (int64_t p0 = 0; p0 < total_batches; p0 += pass_batches) {
        const int64_t p1 = p0 + pass_batches < total_batches ? p0 + pass_batches : total_batches;
        CPPF_RETURN_IF(cudaMemsetAsync(counters, 0, sizeof(RouteCounters), stream));
        RouteParams rp{rot_tab, points, mu_nu, bins, lut, idx, corner, pool, chunk_slab, counters, res,
                       (float)(1.0 / (double)res), lo, hx, hy, hz, dlo, dhx, dhy, dhz, n_points,
                       (long long)p0, (long long)p1, (long long)n_pairs, n_rots, adaptive, gx, pl.planes_per_slab, pl.n_slabs,
                       (unsigned)max_chunks, geom};
        long long blocks = p1 - p0;
        if (blocks > sm_count()) blocks = sm_count();
        route<<<(int)blocks, kRouteThreads, rsmem, stream>>>(rp);
        CPPF_LAUNCH_CHECK();
        SplatParams sp{pool, chunk_slab, counters, acc, gx, gy, gz, pl.planes_per_slab, pl.n_slabs, (unsigned)max_chunks, geom,
                       (int)slab_cells};
        slab_splat_kernel<<<sm_count(), kSplatThreads, ssmem, stream>>>(sp);
        CPPF_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace cppf

using namespace cppf;

extern "C" int cppf_vote_routed_supported(int gx, int gy, int gz) {
    RoutedPlan pl;
    return routed_plan(gx, gy, gz, &pl) ? 1 : 0;
}

// Recommended scratch: accumulator + room for every candidate of min(n_pairs, 4M) pairs per pass.
extern "C" int64_t cppf_vote_routed_scratch_bytes(int64_t n_pairs, int n_rots, int gx, int gy, int gz) {
    RoutedPlan pl;
    if (!routed_plan(gx, gy, gz, &pl)) return -1;
    const int64_t acc_bytes = ((int64_t)gx * gy * gz * 8 + 255) & ~255ll;
    return acc_bytes + routed_pool_bytes(n_pairs, n_rots);
}

extern "C" int cppf_vote_routed(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut,
                                const void* idx, int idx_is_64, float* grid, void* scratch, int64_t scratch_bytes,
                                const float* corner, float res, int n_points, int64_t n_pairs, int n_rots, int gx, int gy,
                                int gz, int adaptive, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RoutedPlan pl;
    if (!routed_plan(gx, gy, gz, &pl)) return (int)cudaErrorInvalidValue;
    if (n_pairs <= 0) return 0;
    const long long cells = (long long)gx * gy * gz;
    const int64_t acc_bytes = ((int64_t)cells * 8 + 255) & ~255ll;
    if (scratch_bytes <= acc_bytes) return (int)cudaErrorInvalidValue;
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(scratch);
    CPPF_RETURN_IF(cudaMemsetAsync(acc, 0, (size_t)cells * 8, stream));
    const int r = vote_routed_launch(points, mu_nu, bins, lut, idx, idx_is_64, acc, reinterpret_cast<unsigned char*>(scratch) + acc_bytes,
                                     scratch_bytes - acc_bytes, corner, res, n_points, n_pairs, n_rots, gx, gy, gz, adaptive,
                                     nullptr, stream);
    if (r != 0) return r;
    return vote_finalize_launch(acc, grid, (int)cells, nullptr, -1, stream);
}

// One call per object: the whole script body of nocs/inference.py:174-339 (sunrgbd/inference.py:142-287)
// between "cloud on the device" and "pose record", enqueued on one stream without a host round trip.
//
//   geometry (min / max / grid dims, :194-195)  ->  kNN + SPRIN point encoder (:180-181)
//   -> per-point projection -> pair MLP + sampling (:182-188, :236-256)  -> centre vote (:191-205)
//   -> argmax (:207-211) -> back-vote + compaction (:216-231) -> orientation histogram(s) (:258-284)
//   -> aux sign / scale sums (:286-302, :335) -> 16-double record.
//
// The vote-grid geometry is derived on the device (struct Geom) and read by the kernels that need it, so the
// host never waits for the grid dimensions and successive objects queue back to back; the Python mirror
// (cppf_b200/pipeline.py) turns the record into RT / scales exactly like nocs/inference.py:305-339.
#include "common.cuh"
#include "vote_common.cuh"

#include "../../include/cppf_b200.h"

#include <math.h>

#include <mutex>
#include <thread>
#include <vector>

namespace cppf {

int vote_fast_launch(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, const void* idx,
                     int idx_is_64, float* grid, void* scratch, const float* corner, float res, int n_points,
                     int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive, const Geom* geom, int max_cells,
                     cudaStream_t stream, int slab_cells = 0, long long total_cells = 0, int row0 = 0);
int backvote_bins_launch(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                         uint8_t* out_mask, const float* corner, const int64_t* argmax_flat, float res, float tol,
                         int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, const Geom* geom,
                         cudaStream_t stream, double res_host, int row0 = 0);
int vote_finalize_launch(const unsigned long long* acc, float* grid, int cells, const Geom* geom, int only_mode,
                         cudaStream_t stream);
int64_t routed_pool_bytes(int64_t n_pairs, int n_rots);
long long slab_cap_cells();
int vote_routed_launch(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, const void* idx,
                       int idx_is_64, unsigned long long* acc, void* pool_mem, int64_t pool_bytes, const float* corner,
                       float res, int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive,
                       const Geom* geom, cudaStream_t stream);
int grid_argmax_launch(const float* grid, int64_t n_cells, const int* n_cells_dev, int64_t* out_index, float* out_value,
                       cudaStream_t stream);

// ---- geometry: nocs/inference.py:194-195 in float32 like numpy (pc is float32, `res` a weak Python scalar)
__global__ void __launch_bounds__(1024) geom_kernel(const float* __restrict__ pc, int n_points, float res, int max_cells,
                                                    int routed_max_cells, int slab_cap_cells, Geom* __restrict__ out,
                                                    float* __restrict__ glob) {
    if (threadIdx.x < 8) glob[threadIdx.x] = -INFINITY;       // running global max of the point encoder (glob_init_kernel)
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n_points; i += blockDim.x) {
        const f3 p = ld3(pc, i);
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    __shared__ float s_lo[32][3], s_hi[32][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_lo[threadIdx.x >> 5][k] = lo[k];
            s_hi[threadIdx.x >> 5][k] = hi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int dims[3];
        for (int k = 0; k < 3; ++k) {
            float a = s_lo[0][k], b = s_hi[0][k];
            for (int w = 1; w < nw; ++w) {
                a = fminf(a, s_lo[w][k]);
                b = fmaxf(b, s_hi[w][k]);
            }
            out->corner[k] = a;
            dims[k] = (int)((b - a) / res) + 1;                                   // :195
        }
        out->gx = dims[0]; out->gy = dims[1]; out->gz = dims[2];
        const long long cells = (long long)dims[0] * dims[1] * dims[2];
        out->cells = cells > 0x7FFFFFFF ? 0x7FFFFFFF : (int)cells;
        // one shared-memory grid per CTA if it fits, else x-slabs routed through HBM (vote_routed.cu), else give up
        int pps = 1, n_slabs = 1;
        const bool fits = cells <= (long long)max_cells;
        const bool routed = !fits && cells <= (long long)routed_max_cells &&
                            routed_plan_hd(dims[0], dims[1], dims[2], slab_cap_cells, &pps, &n_slabs);
        out->status = (fits || routed) ? 0 : 1;
        out->mode = routed ? 1 : 0;
        out->planes_per_slab = pps;
        out->n_slabs = n_slabs;
        float h[3], dh[3];
        for (int k = 0; k < 3; ++k) {
            h[k] = __double2float_ru((double)dims[k] - 1.01);                      // models/voting.py:37-39
            dh[k] = nextafterf((float)((double)h[k] * (double)res * (1.0 + 1e-6)), INFINITY);
        }
        out->hx = h[0]; out->hy = h[1]; out->hz = h[2];
        out->dhx = dh[0]; out->dhy = dh[1]; out->dhz = dh[2];
        out->bx = (float)(dims[0] - 1); out->by = (float)(dims[1] - 1); out->bz = (float)(dims[2] - 1);
    }
}

// benchmark / test aid: overwrite the first `cols` bin columns of every pair (bins [n,4], src [n,cols])
__global__ void __launch_bounds__(256) inject_bins_kernel(uint8_t* __restrict__ bins, const uint8_t* __restrict__ src, int cols,
                                                          long long n) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        uchar4 b = reinterpret_cast<uchar4*>(bins)[p];
        const uint8_t* s = src + p * cols;
        if (cols > 0) b.x = s[0];
        if (cols > 1) b.y = s[1];
        if (cols > 2) b.z = s[2];
        if (cols > 3) b.w = s[3];
        reinterpret_cast<uchar4*>(bins)[p] = b;
    }
}

// nocs/inference.py:177: point_idxs = np.random.randint(0, N, (P, 2)) -- drawn on the device: pair p takes two words of
// Philox4x32-10(counter = (p, 0, 0x70616972 "pair", 0), key = seed), each mapped to [0, N) by a multiply-shift.
__global__ void __launch_bounds__(256) sample_pairs_kernel(int2* __restrict__ idx, long long n_pairs, int n_points,
                                                           unsigned long long seed) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (long long)gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(make_uint4((uint32_t)p, (uint32_t)((unsigned long long)p >> 32), 0x70616972u, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        idx[p] = make_int2((int)__umulhi(w.x, (uint32_t)n_points), (int)__umulhi(w.y, (uint32_t)n_points));
    }
}

// record[16] (double): 0 argmax flat | 1 best up bin | 2 best right bin (-1) | 3..5 sum log-scale | 6 survivors |
//                      7 S_up | 8 S_right | 9..11 corner | 12..14 grid dims | 15 status
__global__ void pack_record_kernel(const Geom* geom, const long long* flat, const long long* best, const double* stats,
                                   int n_dirs, double* rec) {
    rec[0] = (double)*flat;
    rec[1] = (double)best[0];
    rec[2] = n_dirs > 1 ? (double)best[1] : -1.0;
    rec[3] = stats[0]; rec[4] = stats[1]; rec[5] = stats[2];
    rec[6] = stats[3]; rec[7] = stats[4]; rec[8] = stats[5];
    rec[9] = geom->corner[0]; rec[10] = geom->corner[1]; rec[11] = geom->corner[2];
    rec[12] = geom->gx; rec[13] = geom->gy; rec[14] = geom->gz;
    rec[15] = geom->status;
}

// ---- workspace carving -------------------------------------------------------------------------------------
struct Carver {
    unsigned char* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Workspace {
    Geom* geom;
    long long* nbrs;
    float* feat;
    float* glob;
    float* table;
    uint8_t* bins;
    float* tail;
    uint8_t* mask;
    long long* pos;
    void* compact_scratch;
    float* grid;
    unsigned long long* acc;
    long long* flat;
    float* counts;
    long long* best;
    double* stats;
    long long* count;
    unsigned char* pool;
    size_t pool_bytes;
    int2* idx_gen;              // pairs drawn on the device (sample_pairs): [n_pairs] when the object is not dense
    size_t bytes;
    size_t zero_bytes;                 // grid .. count: the accumulators cleared at the start of an object
};

static Workspace carve(void* base, int n_points, int64_t n_pairs, int knn, int max_cells, int routed_max_cells, int n_rots,
                       int n_sphere, int64_t rot_subsample, bool dense) {
    const int cap_cells = max_cells > routed_max_cells ? max_cells : routed_max_cells;
    Carver c{reinterpret_cast<unsigned char*>(base)};
    Workspace w;
    w.geom = c.take<Geom>(1);
    w.nbrs = c.take<long long>((size_t)n_points * knn);
    w.feat = c.take<float>((size_t)n_points * 40);
    w.glob = c.take<float>(8);
    w.table = c.take<float>((size_t)n_points * cppf_tc_table_cols());
    w.bins = c.take<uint8_t>((size_t)n_pairs * 4);
    w.tail = c.take<float>((size_t)n_pairs * 5);
    w.mask = c.take<uint8_t>((size_t)n_pairs);
    // survivors are addressed through the mask; only the <= rot_subsample sampled positions are materialised
    w.pos = c.take<long long>((size_t)(rot_subsample > 0 && rot_subsample < n_pairs ? rot_subsample : n_pairs));
    w.compact_scratch = c.take<unsigned char>((size_t)cppf_compact_scratch_bytes(n_pairs));
    w.grid = c.take<float>((size_t)cap_cells);
    w.acc = c.take<unsigned long long>((size_t)cap_cells);
    w.flat = c.take<long long>(2);
    w.counts = c.take<float>((size_t)n_sphere * 2);
    w.best = c.take<long long>(2);
    w.stats = c.take<double>(6);
    w.count = c.take<long long>(1);
    w.zero_bytes = (size_t)(reinterpret_cast<unsigned char*>(w.count + 1) - reinterpret_cast<unsigned char*>(w.grid));
    w.pool_bytes = routed_max_cells > 0 ? (size_t)routed_pool_bytes(n_pairs, n_rots) : 0;
    w.pool = c.take<unsigned char>(w.pool_bytes);
    w.idx_gen = c.take<int2>(dense ? 0 : (size_t)n_pairs);
    w.bytes = (c.off + 255) & ~(size_t)255;
    return w;
}

// ---- stage timing (CUDA events on the launching stream) -------------------------------------------------------
static const char* kStageNames[] = {"geometry", "point_encoder", "preproject", "encode_sample", "vote", "argmax",
                                    "backvote", "compact", "rot_hist", "stats"};
constexpr int kStages = sizeof(kStageNames) / sizeof(kStageNames[0]);

struct Timing {
    std::vector<cudaEvent_t> pool;     // kStages + 1 events per recorded call
    size_t used = 0;
    cudaEvent_t next() {
        if (used == pool.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            pool.push_back(e);
        }
        return pool[used++];
    }
};

}  // namespace cppf

using namespace cppf;

extern "C" int cppf_pose_record_doubles(void) { return 16; }
extern "C" int cppf_pose_args_bytes(void) { return (int)sizeof(cppf_pose_args); }

extern "C" int64_t cppf_pose_workspace_bytes(int n_points, int64_t n_pairs, int knn, int max_cells, int routed_max_cells,
                                             int n_rots, int n_sphere, int64_t rot_subsample) {
    const bool dense = n_pairs <= 0 || n_pairs == (int64_t)n_points * n_points;
    if (dense) n_pairs = (int64_t)n_points * n_points;
    return (int64_t)carve(nullptr, n_points, n_pairs, knn, max_cells, routed_max_cells, n_rots, n_sphere, rot_subsample, dense).bytes;
}

extern "C" void* cppf_timing_create(void) { return new Timing(); }
extern "C" void cppf_timing_destroy(void* t) {
    if (!t) return;
    Timing* tm = reinterpret_cast<Timing*>(t);
    for (cudaEvent_t e : tm->pool) cudaEventDestroy(e);
    delete tm;
}
// pre-create the events of `n_calls` calls (cudaEventCreate is slow enough to show up when timing short objects)
extern "C" int cppf_timing_reserve(void* t, int n_calls) {
    Timing* tm = reinterpret_cast<Timing*>(t);
    const size_t want = (size_t)n_calls * (kStages + 1);
    while (tm->pool.size() < want) {
        cudaEvent_t e;
        const cudaError_t r = cudaEventCreate(&e);
        if (r != cudaSuccess) return (int)r;
        tm->pool.push_back(e);
    }
    return 0;
}
extern "C" int cppf_timing_stages(void) { return kStages; }
extern "C" const char* cppf_timing_stage_name(int i) { return (i >= 0 && i < kStages) ? kStageNames[i] : ""; }
// h_ms_sum[kStages] += elapsed per stage over every call recorded since the last collect; returns the number
// of calls (negative: CUDA error).  Synchronises on the last recorded event.
extern "C" int cppf_timing_collect(void* t, float* h_ms_sum) {
    Timing* tm = reinterpret_cast<Timing*>(t);
    const size_t per = kStages + 1;
    const size_t calls = tm->used / per;
    if (calls == 0) return 0;
    if (cudaEventSynchronize(tm->pool[tm->used - 1]) != cudaSuccess) return -1;
    for (size_t c = 0; c < calls; ++c)
        for (int s = 0; s < kStages; ++s) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, tm->pool[c * per + s], tm->pool[c * per + s + 1]) != cudaSuccess) return -1;
            h_ms_sum[s] += ms;
        }
    tm->used = 0;
    return (int)calls;
}

extern "C" int cppf_pose_fused(const cppf_pose_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (a == nullptr || a->struct_bytes != sizeof(cppf_pose_args)) return (int)cudaErrorInvalidValue;
    const int n = a->n_points;
    const bool sampled = a->sample_pairs != 0;
    if (sampled && a->idx != nullptr) return (int)cudaErrorInvalidValue;
    const bool dense = a->idx == nullptr && !sampled;
    const int64_t n_pairs = dense ? (int64_t)n * n : a->n_pairs;
    if (n <= 0 || n_pairs <= 0 || a->knn <= 0 || a->knn > 64 || a->knn > n) return (int)cudaErrorInvalidValue;
    if (a->max_cells <= 0 || a->max_cells > cppf_vote_private_max_cells()) return (int)cudaErrorInvalidValue;
    if (a->n_sphere <= 0 || a->record == nullptr || a->workspace == nullptr) return (int)cudaErrorInvalidValue;
    if (a->routed_max_cells < 0 || a->routed_max_cells > kMaxSlabs * slab_cap_cells()) return (int)cudaErrorInvalidValue;
    const Workspace w = carve(a->workspace, n, n_pairs, a->knn, a->max_cells, a->routed_max_cells, a->n_rots, a->n_sphere,
                              a->rot_subsample, dense);
    const void* idx = sampled ? (const void*)w.idx_gen : a->idx;
    const int idx_is_64 = sampled ? 0 : a->idx_is_64;
    const int cap_cells = a->max_cells > a->routed_max_cells ? a->max_cells : a->routed_max_cells;
    if ((int64_t)w.bytes > a->workspace_bytes) return (int)cudaErrorInvalidValue;
    Timing* tm = reinterpret_cast<Timing*>(a->timing);
    auto mark = [&]() -> int {
        if (!tm) return 0;
        cudaEvent_t e = tm->next();
        if (!e) return (int)cudaErrorMemoryAllocation;
        return (int)cudaEventRecord(e, stream);
    };
#define CPPF_TRY(expr)          \
    do {                        \
        const int _r = (expr);  \
        if (_r != 0) return _r; \
    } while (0)

    CPPF_TRY(mark());
    if (sampled) {                                                              // nocs/inference.py:177
        sample_pairs_kernel<<<(int)((n_pairs + 1023) / 1024 < 2048 ? (n_pairs + 1023) / 1024 : 2048), 256, 0, stream>>>(
            w.idx_gen, (long long)n_pairs, n, (unsigned long long)a->seed);
        CPPF_LAUNCH_CHECK();
    }
    geom_kernel<<<1, 1024, 0, stream>>>(a->pc, n, a->res, a->max_cells, a->routed_max_cells, (int)slab_cap_cells(),
                                        w.geom, w.glob);
    CPPF_LAUNCH_CHECK();
    // every accumulator of the object (vote grid, its u64 scratch, argmax keys, sphere counts, statistics, survivor count)
    // lies in one contiguous tail of the workspace: one memset instead of six
    CPPF_RETURN_IF(cudaMemsetAsync(w.grid, 0, w.zero_bytes, stream));
    const PreparedWorkspaceScope prepared;
    CPPF_TRY(mark());
    CPPF_TRY(cppf_knn(a->pc, n, a->knn, reinterpret_cast<int64_t*>(w.nbrs), stream));
    CPPF_TRY(cppf_point_encode(a->pc, a->nrm, reinterpret_cast<const int64_t*>(w.nbrs), a->pe_blob, w.feat, w.glob, n, a->knn,
                               stream));
    CPPF_TRY(mark());
    CPPF_TRY(cppf_tc_preproject(w.feat, a->tc_blob, w.table, n, stream));
    CPPF_TRY(mark());
    const int n_dirs = a->regress_right ? 2 : 1;
    const int heads = 1 | 2 | 8 | (a->regress_right ? 4 : 0);
    CPPF_TRY(cppf_encode_sample_tc(a->pc, a->nrm, w.table, a->tc_blob, idx, idx_is_64, n, n_pairs, a->uniforms, a->seed,
                                   heads, w.bins, w.tail, nullptr, stream));
    if (a->inject_bins != nullptr && a->inject_cols > 0) {
        inject_bins_kernel<<<sm_count() * 8, 256, 0, stream>>>(w.bins, a->inject_bins, a->inject_cols, (long long)n_pairs);
        CPPF_LAUNCH_CHECK();
    }
    CPPF_TRY(mark());
    CPPF_TRY(vote_fast_launch(a->pc, nullptr, w.bins, a->lut, idx, idx_is_64, w.grid, w.acc, nullptr, a->res, n, n_pairs,
                              a->n_rots, 0, 0, 0, a->adaptive, w.geom, a->max_cells, stream));
    if (a->routed_max_cells > 0) {      // grids of up to 8 shared-memory slabs: the kernels return at once unless geom->mode == 1
        CPPF_RETURN_IF(cudaMemsetAsync(w.acc, 0, (size_t)cap_cells * 8, stream));
        CPPF_TRY(vote_routed_launch(a->pc, nullptr, w.bins, a->lut, idx, idx_is_64, w.acc, w.pool, (int64_t)w.pool_bytes,
                                    nullptr, a->res, n, n_pairs, a->n_rots, 0, 0, 0, a->adaptive, w.geom, stream));
        CPPF_TRY(vote_finalize_launch(w.acc, w.grid, cap_cells, w.geom, 1, stream));
    }
    CPPF_TRY(mark());
    CPPF_TRY(grid_argmax_launch(w.grid, cap_cells, &w.geom->cells, reinterpret_cast<int64_t*>(w.flat), nullptr, stream));
    CPPF_TRY(mark());
    CPPF_TRY(backvote_bins_launch(a->pc, w.bins, a->lut, idx, idx_is_64, w.mask, nullptr,
                                  reinterpret_cast<const int64_t*>(w.flat), a->res, a->tol, n, n_pairs, a->n_rots, 0, 0, 0,
                                  w.geom, stream, a->res_host));
    CPPF_TRY(mark());
    // survivors stay addressed through the mask: per-block counts + scan instead of a materialised list (:230-231)
    CPPF_TRY(cppf_compact_count(w.mask, n_pairs, reinterpret_cast<int64_t*>(w.count), w.compact_scratch, stream));
    CPPF_TRY(mark());
    const int64_t max_samples = a->rot_subsample > 0 ? a->rot_subsample : n_pairs;
    for (int j = 0; j < n_dirs; ++j) {
        CPPF_TRY(cppf_rot_hist_mask(a->pc, w.bins, a->lut, idx, idx_is_64, w.mask,
                                    reinterpret_cast<const int64_t*>(w.compact_scratch), n_pairs,
                                    reinterpret_cast<const int64_t*>(w.count), a->sphere, w.counts + (size_t)j * a->n_sphere, n,
                                    a->n_rots, a->n_sphere, j, max_samples < n_pairs ? max_samples : n_pairs,
                                    a->seed * 7919ull + (uint64_t)j, a->cos_thr, reinterpret_cast<int64_t*>(w.pos), stream));
        CPPF_TRY(grid_argmax_launch(w.counts + (size_t)j * a->n_sphere, a->n_sphere, nullptr,
                                    reinterpret_cast<int64_t*>(w.best + j), nullptr, stream));
    }
    CPPF_TRY(mark());
    CPPF_TRY(cppf_survivor_stats_mask(a->pc, a->nrm, w.tail, idx, idx_is_64, w.mask, a->sphere,
                                      reinterpret_cast<const int64_t*>(w.best),
                                      n_dirs > 1 ? reinterpret_cast<const int64_t*>(w.best + 1) : nullptr, w.stats, n, n_pairs,
                                      stream));
    pack_record_kernel<<<1, 1, 0, stream>>>(w.geom, w.flat, w.best, w.stats, n_dirs, a->record);
    CPPF_LAUNCH_CHECK();
    CPPF_TRY(mark());
#undef CPPF_TRY
    return 0;
}

// ---- the object loop as one call -------------------------------------------------------------------------------
namespace {
struct BatchPool {
    std::mutex mu;                       // one cppf_pose_batch at a time per device
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> done;
    cudaEvent_t fork = nullptr;
};
BatchPool g_pools[64];

int pool_reserve(BatchPool& p, int n_streams) {
    if (p.fork == nullptr) CPPF_RETURN_IF(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    while ((int)p.streams.size() < n_streams) {
        cudaStream_t s;
        cudaEvent_t e;
        CPPF_RETURN_IF(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        CPPF_RETURN_IF(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p.streams.push_back(s);
        p.done.push_back(e);
    }
    return 0;
}
}  // namespace

extern "C" int cppf_pose_batch(const cppf_pose_args* args, int n_objects, int n_streams, int n_threads, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_objects < 0 || (n_objects > 0 && args == nullptr) || n_streams < 1 || n_streams > 16 || n_threads < 1 ||
        n_threads > n_streams)
        return (int)cudaErrorInvalidValue;
    if (n_objects == 0) return 0;
    for (int i = 0; i < n_objects; ++i)
        if (args[i].struct_bytes != sizeof(cppf_pose_args) || args[i].timing != nullptr) return (int)cudaErrorInvalidValue;
    int dev = 0;
    CPPF_RETURN_IF(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    BatchPool& pool = g_pools[dev];
    std::lock_guard<std::mutex> lock(pool.mu);
    if (n_streams > n_objects) n_streams = n_objects;
    if (n_threads > n_streams) n_threads = n_streams;
    const int r = pool_reserve(pool, n_streams);
    if (r != 0) return r;
    // fork: every worker stream starts after what the caller enqueued on `stream`
    CPPF_RETURN_IF(cudaEventRecord(pool.fork, stream));
    for (int s = 0; s < n_streams; ++s) CPPF_RETURN_IF(cudaStreamWaitEvent(pool.streams[s], pool.fork, 0));
    std::vector<int> errs((size_t)n_threads, 0);
    auto work = [&](int t) {
        if (cudaSetDevice(dev) != cudaSuccess) {
            errs[t] = (int)cudaErrorInvalidDevice;
            return;
        }
        // thread t owns worker streams t, t + T, ...; the objects of a stream are enqueued in order
        for (int i = 0; i < n_objects && errs[t] == 0; ++i) {
            const int s = i % n_streams;
            if (s % n_threads != t) continue;
            errs[t] = cppf_pose_fused(&args[i], pool.streams[s]);
        }
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        th.reserve((size_t)n_threads - 1);
        for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    // join: `stream` continues after every worker stream (also after an error: whatever was enqueued is waited for)
    int err = 0;
    for (int s = 0; s < n_streams; ++s) {
        const cudaError_t e1 = cudaEventRecord(pool.done[s], pool.streams[s]);
        const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(stream, pool.done[s], 0) : e1;
        if (e2 != cudaSuccess && err == 0) err = (int)e2;
    }
    for (int t = 0; t < n_threads; ++t)
        if (errs[t] != 0) return errs[t];
    return err;
}

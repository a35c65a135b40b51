/* cppf_b200 -- C ABI of the B200 (sm_100a) implementation of CPPF's per-object hot path.
 *
 * Drop-in boundary.  The reference reaches this path through two
 * Python modules; each entry point below names the reference interface it replaces
 * (paths relative to the reference tree):
 *
 *   models/model.py:117-137  PPFEncoder.forward_with_idx  -> cppf_ppf_encode (idx != NULL)
 *   models/model.py:89-115   PPFEncoder.forward (dense)    -> cppf_ppf_encode (idx == NULL)
 *   models/voting.py:4-67    ppf_kernel      (RawKernel)   -> cppf_ppf_vote
 *   models/voting.py:70-113  backvote_kernel (RawKernel)   -> cppf_backvote
 *   models/voting.py:115-148 rot_voting_kernel             -> cppf_rot_vote
 *   models/voting.py:150-172 findpeak_kernel               -> cppf_findpeak
 *   nocs/inference.py:185-188 softmax+multinomial+decode   -> cppf_sample_bins
 *   nocs/inference.py:207-208 grid.get()+np.argmax         -> cppf_grid_argmax
 *   nocs/inference.py:229-231 mask + compaction            -> cppf_compact_pairs
 *   nocs/inference.py:276-284 candidates.mm(sphere)>thr    -> cppf_sphere_count
 *
 * Conventions: plain pointers and sizes, no torch types.  Every pointer is a DEVICE
 * pointer unless its name starts with `h_`.  The caller owns all buffers.  Every
 * call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 * and returns a cudaError_t-style int (0 = success); nothing throws.  Degenerate
 * pairs and out-of-grid votes are silently dropped exactly as the reference
 * kernels do (models/voting.py:21,36-39).
 */
#ifndef CPPF_B200_H
#define CPPF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library info -------------------------------------------------------- */
int cppf_abi_version(void);                 /* bumped on any signature change */
const char* cppf_error_string(int code);    /* cudaGetErrorString passthrough */
/* Number of CUDA kernels this library has launched since load (for bench.py's
 * `gpu_launches`).  Process-wide, monotonically increasing. */
uint64_t cppf_launch_count(void);

/* ---- pair MLP weights ------------------------------------------------------
 * The pair MLP is specialised to the reference architecture
 * ppffcs=[2*40+4, 32, 32, 16] (nocs/inference.py:83); out_dim is free (141 for
 * NOCS/SUN RGB-D, 9 for the zero-shot regression head).
 * cppf_ppf_blob_floats(out_dim) gives the size of the packed weight blob; the host
 * side packs a PPFEncoder state_dict into it (cppf_b200/model.py:pack_ppf_weights)
 * following the layout documented in cppf_b200/csrc/mlp_layout.h. */
int cppf_ppf_blob_floats(int out_dim);
int cppf_ppf_feat_dim(void);                /* 40 */

/* Per-point pre-projection of ResLayer-0's feature columns (layer-0 algebra,
 * SURVEY.md section 7): table[n, 0:128].  Must run before cppf_ppf_encode / fused calls
 * whenever feat or the weights change.  feat [n_points, 40], table [n_points, 128]. */
int cppf_ppf_preproject(const float* feat, const float* blob, float* table, int n_points, void* stream);

/* replaces PPFEncoder.forward_with_idx (models/model.py:117-137) when idx != NULL:
 *   out[p, :] = final(ResLayers([feat[a], feat[b], ppf(a,b)])),  (a,b) = idx[p]
 * and PPFEncoder.forward's dense branch (models/model.py:92-115) when idx == NULL:
 *   all n_points^2 ordered pairs, row-major (p = a*n_points + b), n_pairs must be
 *   n_points^2; `dist` (optional, [n_points, n_points]) is the caller-supplied
 *   distance matrix of the reference signature -- NULL means exact norms.
 * pc, nrm [n_points,3]; table from cppf_ppf_preproject; idx [n_pairs,2] int64 or
 * int32 (idx_is_64); out [n_pairs, out_dim] fp32.
 * col_begin/col_count restrict the `final` layer to a column window (e.g. 0,64 for the
 * translation heads); out then has col_count columns. */
int cppf_ppf_encode(const float* pc, const float* nrm, const float* table, const float* blob,
                    const void* idx, int idx_is_64, const float* dist,
                    float* out, int n_points, int64_t n_pairs, int out_dim,
                    int col_begin, int col_count, void* stream);

/* replaces softmax + torch.multinomial + bin decode (nocs/inference.py:185-188,245-256).
 * logits [n_rows, row_stride], one categorical draw per row over columns
 * [col0, col0+n_bins).  mode 0: exponential race argmax(p/q) with injected noise
 * q[n_rows, n_bins] (reproduces torch.multinomial under the same generator);
 * mode 1: inverse CDF with injected uniforms u[n_rows]; mode 2: inverse CDF with
 * Philox4x32-10 uniforms keyed by (seed, row, stream_id).
 * value = ((bin / div) * mul_a) * mul_b - sub, each step rounded to fp32 like the torch
 * expression it replaces, written to out_val[row*out_stride]; out_bin (optional)
 * receives the int32 bin. */
int cppf_sample_bins(const float* logits, int64_t n_rows, int row_stride, int col0, int n_bins,
                     int mode, const float* noise, uint64_t seed, uint32_t stream_id,
                     float div, float mul_a, float mul_b, float sub,
                     float* out_val, int out_stride, int32_t* out_bin, void* stream);

/* replaces ppf_kernel / ppf_voting (models/voting.py:8-66).  grid [gx,gy,gz] is
 * accumulated in place (caller zeroes it).  idx int32 [n_pairs,2] like the reference
 * (nocs/inference.py:202) or int64.  idx == NULL enumerates all n_points^2 pairs. */
int cppf_ppf_vote(const float* points, const float* mu_nu, const float* probs, const void* idx, int idx_is_64,
                  float* grid, const float* corner, float res, int n_points, int64_t n_pairs, int n_rots,
                  int gx, int gy, int gz, int adaptive, void* stream);

/* replaces grid.get() + np.argmax (nocs/inference.py:207-208): first maximal flat
 * index in C order -> *out_index (int64, device), optional *out_value. */
int cppf_grid_argmax(const float* grid, int64_t n_cells, int64_t* out_index, float* out_value, void* stream);

/* replaces backvote_kernel (models/voting.py:74-112).  out_offsets [n_pairs,3] follows
 * the reference (written only for non-degenerate pairs); out_mask (optional, uint8
 * [n_pairs]) = any(out_offsets != 0), the only thing the caller consumes
 * (nocs/inference.py:229-230).  Either output may be NULL. */
int cppf_backvote(const float* points, const float* mu_nu, float* out_offsets, uint8_t* out_mask,
                  const void* idx, int idx_is_64, const float* corner, float res, int n_points, int64_t n_pairs,
                  int n_rots, int gx, int gy, int gz, const float* centre, float tol, void* stream);

/* replaces point_idxs[mask] (nocs/inference.py:230-231): order-preserving stream
 * compaction of the surviving pairs.  out_idx (optional) [<= n_pairs, 2] int32 (a,b); out_pos
 * (optional) the source pair position; *out_count (int64, device) the survivor count.
 * scratch must hold cppf_compact_scratch_bytes(n_pairs) bytes. */
int64_t cppf_compact_scratch_bytes(int64_t n_pairs);
int cppf_compact_pairs(const uint8_t* mask, const void* idx, int idx_is_64, int n_points, int64_t n_pairs,
                       int32_t* out_idx, int64_t* out_pos, int64_t* out_count, void* scratch, void* stream);

/* The first two steps of cppf_compact_pairs only: *out_count and, in scratch[0 : ceil(n_pairs / 2048)] (int64), the
 * exclusive survivor offsets of the 2048-pair blocks -- enough for cppf_rot_hist_mask / cppf_survivor_stats_mask to
 * address the survivors without the materialised list. */
int cppf_compact_count(const uint8_t* mask, int64_t n_pairs, int64_t* out_count, void* scratch, void* stream);

/* replaces rot_voting_kernel (models/voting.py:119-147): outputs_up [n_pairs, n_rots, 3]. */
int cppf_rot_vote(const float* points, const float* preds_rot, float* outputs_up, const void* idx, int idx_is_64,
                  int64_t n_pairs, int n_rots, void* stream);

/* replaces candidates.mm(sphere) > thr, sum(0) (nocs/inference.py:282-283).
 * cand [n_cand,3], sphere [n_bins,3] fp32, counts int32 [n_bins] accumulated in place. */
int cppf_sphere_count(const float* cand, int64_t n_cand, const float* sphere, int n_bins, float thr,
                      int32_t* counts, void* stream);

/* replaces findpeak_kernel (models/voting.py:154-171).  literal != 0 reproduces the
 * string as shipped (comma operator at :165-166 drops the x term of the y reads);
 * literal == 0 is the intended 6-neighbour second difference. */
int cppf_findpeak(const float* grid, float* out, int width, int gx, int gy, int gz, int literal, void* stream);

/* ---- point encoder (models/model.py:34-77 PointEncoder, models/sprin.py:40-107) -------------------
 * cppf_knn replaces torch.cdist + torch.topk(dist, k, largest=False, sorted=False) (nocs/inference.py:180,
 * models/model.py:47): out_idx [n_points, k] int64, the k points with the smallest exact squared distance
 * (self included), unordered; ties at the k-th distance resolve to the lowest indices.
 * cppf_point_encode replaces PointEncoder.forward_nbrs (models/model.py:63-77) for the reference
 * configuration (spfcs=[32,64,32,32], rank 32, 2 neighbour features, out_dim 32, k <= 64): feat
 * [n_points, 40] = [SPRIN conv + LayerNorm (32) | global max of the 8-column GlobalInfoProp linear].
 * The SPRIN MLP runs on the tensor cores (tcgen05, 3xTF32: fp32-grade results); the environment variable
 * CPPF_PE_IMPL=simt, read per call, selects the fp32 FFMA kernel kept as its cross-check.
 * pe_blob: cppf_pe_blob_floats() floats (cppf_b200/model.py:pack_pe_weights: the FFMA section, then the same layers as
 * tcgen05 operands); 16-byte aligned; glob_scratch: 8 floats. */
int cppf_pe_blob_floats(void);
int cppf_knn(const float* pc, int n_points, int k, int64_t* out_idx, void* stream);
int cppf_point_encode(const float* pc, const float* nrm, const int64_t* nbrs, const float* pe_blob, float* feat,
                      float* glob_scratch, int n_points, int k, void* stream);

/* ==== fused per-object path ==================================================
 * The same reference lines, re-cut so that logits, (mu,nu) floats and the [P,72,3]
 * candidate dump never reach HBM.  Per pair the path keeps 4 bin bytes and 5 tail floats.
 *
 * Decode table `lut` (device, 136 floats): [0:32] mu of each translation bin, [32:64] nu,
 * [64:100] up angle, [100:136] right angle -- the fp32 values of nocs/inference.py:187-188,
 * 252,256 computed once on the host.  These entry points are specialised to the reference
 * head layout out_dim = 2*32 + 2*36 + 2 + 3 (config/config.yaml:7-8). */

int cppf_head_blob_floats(void);

/* models/model.py:117-137 + nocs/inference.py:183-188 + :236-256 in one pass: pair MLP with
 * the `final` layer evaluated head by head in shared memory, one inverse-CDF categorical
 * draw per head (uniforms: optional [n_pairs,4] = (mu,nu,up,right); NULL -> Philox4x32-10
 * keyed by (seed, pair index)).  heads: bit0 mu+nu, bit1 up, bit2 right, bit3 tail.
 * bins [n_pairs,4] uint8; tail [5, n_pairs] fp32 = (aux_up, aux_right, log-scale x3). */
int cppf_encode_sample(const float* pc, const float* nrm, const float* table, const float* blob,
                       const float* head_blob, const void* idx, int idx_is_64, int n_points, int64_t n_pairs,
                       const float* uniforms, uint64_t seed, int heads, uint8_t* bins, float* tail, void* stream);

/* The same contract as cppf_encode_sample with the dense layers on the 5th-generation tensor cores
 * (tcgen05.mma kind::tf32, accumulators in TMEM, 3xTF32 operand splitting for fp32-grade logits;
 * csrc/encode_tc.cu).  Adjacent linear maps of models/model.py:26-31,134-137 are composed on the host
 * (cppf_b200/model.py:pack_tc_weights -> tc_blob, cppf_tc_blob_floats() floats), so `table` here is the
 * per-point projection written by cppf_tc_preproject (feat [n_points,40]): cppf_tc_table_cols() columns per
 * point, stored planar as [cols/4][n_points][4] floats so dense-mode gathers are coalesced.
 * dbg_t (optional, [n_pairs,32]) receives [fc1_2(x2) ; fc0_2(x2) + fc2_2.b], the last pre-activation. */
int cppf_tc_blob_floats(void);
int cppf_tc_table_cols(void);
int cppf_tc_preproject(const float* feat, const float* tc_blob, float* table, int n_points, void* stream);
int cppf_encode_sample_tc(const float* pc, const float* nrm, const float* table, const float* tc_blob,
                          const void* idx, int idx_is_64, int n_points, int64_t n_pairs, const float* uniforms,
                          uint64_t seed, int heads, uint8_t* bins, float* tail, float* dbg_t, void* stream);

/* models/voting.py:8-66 with prob == 1 (nocs/inference.py:201): votes accumulate in a
 * shared-memory-privatised fixed-point grid (weights rounded to 2^-14, exact integer sums,
 * deterministic) flushed into `scratch` (cppf_vote_scratch_bytes) and added to `grid`.
 * Exactly one of mu_nu ([n_pairs,2] fp32) / bins (+lut) is given.  Needs
 * gx*gy*gz <= cppf_vote_private_max_cells() and n_rots <= 72; cudaErrorInvalidValue otherwise
 * (use cppf_ppf_vote). */
int64_t cppf_vote_scratch_bytes(int gx, int gy, int gz);
int cppf_vote_private_max_cells(void);
int cppf_vote_fast(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut,
                   const void* idx, int idx_is_64, float* grid, void* scratch, const float* corner, float res,
                   int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive, void* stream);

/* On return from cppf_vote_fast / cppf_vote_slabs / cppf_vote_routed the first gx*gy*gz 64-bit words of `scratch` hold the exact vote
 * sums in units of 2^-14 (unsigned).  A caller that splits the pairs of ONE object over several GPUs (SURVEY.md
 * section 8e, second axis) sums those integer grids across ranks and converts once:
 *     grid[i] += (float)((double)acc[i] * 2^-14)
 * -- the same single rounding, hence bit for bit the same grid, as one call over all pairs. */
int cppf_vote_finalize(const void* acc, float* grid, int64_t cells, void* stream);

/* The same contract for grids of up to eight shared-memory slabs (e.g. the 64^3 grid of BASELINE config 3, 1 MB):
 * candidates are routed through HBM to the x-slab that owns them (32 B per in-bounds candidate) and accumulated
 * in that slab's shared-memory copy (csrc/vote_routed.cu).  Same fixed-point sums, hence the same grid as
 * cppf_vote_fast would produce.  scratch: >= cppf_vote_routed_scratch_bytes(...) is recommended (room for every
 * candidate of 4M pairs per pass); smaller scratch means more passes, too small returns cudaErrorInvalidValue. */
int cppf_vote_routed_supported(int gx, int gy, int gz);
int64_t cppf_vote_routed_scratch_bytes(int64_t n_pairs, int n_rots, int gx, int gy, int gz);
int cppf_vote_routed(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut,
                     const void* idx, int idx_is_64, float* grid, void* scratch, int64_t scratch_bytes,
                     const float* corner, float res, int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz,
                     int adaptive, void* stream);

/* A third way for large grids, "slab passes": the same shared-memory kernel as cppf_vote_fast, every CTA holding one
 * x-slab of the grid and keeping only the candidates whose floor(g.x) it owns.  Phase 1 (candidate generation) is
 * repeated once per slab but nothing is routed through memory; same exact sums.  scratch: gx*gy*gz*8 bytes. */
int cppf_vote_slabs_supported(int gx, int gy, int gz);
int cppf_vote_slabs(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut,
                    const void* idx, int idx_is_64, float* grid, void* scratch, const float* corner, float res,
                    int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, int adaptive, void* stream);

/* models/voting.py:74-112 + nocs/inference.py:207-211,229-230 from bins: the winning cell is
 * read from device memory (*argmax_flat), centre = float32(corner + cell * res_host) as at :209,:226 -- the host
 * multiplies in float64 with the Python float `cfg.res`, so res_host is that DOUBLE (0: use (double)res; the two differ in
 * the 9th digit, which moves the float32 centre by one ulp for some cells and flips candidates on the tolerance sphere);
 * `res` is the float32 the kernels divide by (cp.float32(cfg.res) at :225).  out_mask[p] = any(out_offsets[p] != 0). */
int cppf_backvote_bins(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                       uint8_t* out_mask, const float* corner, const int64_t* argmax_flat, float res, float tol,
                       double res_host, int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, void* stream);

/* models/voting.py:119-147 + nocs/inference.py:276-284 fused: orientation candidates of a
 * sub-sample (without replacement, <= max_samples) of the survivors pos[0:*count] are counted
 * against the sphere bins in shared memory; counts [n_bins] fp32 (exact integers) accumulate
 * in place.  which: 0 = up head, 1 = right head. */
int cppf_rot_hist(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                  const int64_t* pos, const int64_t* count, const float* sphere, float* counts, int n_points,
                  int n_rots, int n_bins, int which, int64_t max_samples, uint64_t offset_seed, float thr,
                  void* stream);

/* cppf_rot_hist / cppf_survivor_stats with the survivors given as the back-vote mask itself (uint8 [n_pairs]) instead of
 * the compacted list: survivor r is the r-th set byte (block_offsets from cppf_compact_count); same sample, same sums.
 * sample_scratch: int64 [max_samples] for the positions of the sub-sampled survivors (NULL: slower in-kernel selection). */
int cppf_rot_hist_mask(const float* points, const uint8_t* bins, const float* lut, const void* idx, int idx_is_64,
                       const uint8_t* mask, const int64_t* block_offsets, int64_t n_pairs, const int64_t* count,
                       const float* sphere, float* counts, int n_points, int n_rots, int n_bins, int which,
                       int64_t max_samples, uint64_t offset_seed, float thr, int64_t* sample_scratch, void* stream);
int cppf_survivor_stats_mask(const float* points, const float* nrm, const float* tail, const void* idx, int idx_is_64,
                             const uint8_t* mask, const float* sphere, const int64_t* best_up, const int64_t* best_right,
                             double* out, int n_points, int64_t n_pairs, void* stream);

/* nocs/inference.py:286-302,335 over the survivors: out[0:3] = sum of log-scales, out[3] =
 * count, out[4] = S_up, out[5] = S_right, S = sum aux*(2t-1); down_loss < up_loss <=> S < 0. */
int cppf_survivor_stats(const float* points, const float* nrm, const float* tail, const void* idx, int idx_is_64,
                        const int64_t* pos, const int64_t* count, const float* sphere, const int64_t* best_up,
                        const int64_t* best_right, double* out, int n_points, int64_t n_pairs, void* stream);

/* ==== one call per object ======================================================
 * The whole script body of nocs/inference.py:174-339 (sunrgbd/inference.py:142-287) between "cloud on the
 * device" and "pose record": vote-grid geometry (:194-195, derived on the device), kNN + SPRIN point encoder
 * (:180-181), pair MLP + sampling (:182-188,:236-256), centre vote (:191-205), argmax (:207-211), back-vote +
 * compaction (:216-231), orientation histogram(s) (:258-284), aux sign / scale sums (:286-302,:335).
 * Everything is enqueued on `stream` without a host round trip; `record` (device, cppf_pose_record_doubles()
 * doubles) receives
 *   [0] argmax flat index | [1] best up bin | [2] best right bin (-1) | [3..5] sum of log-scales |
 *   [6] survivor count | [7] S_up | [8] S_right | [9..11] grid corner | [12..14] grid dims | [15] status
 * status 1 = the vote grid exceeds max_cells and routed_max_cells (nothing was voted; use the staged entry points).
 * The host tail (Gram-Schmidt, scale, RT: nocs/inference.py:305-339) is cppf_b200/pipeline.py.
 * struct_bytes must be sizeof(cppf_pose_args).  All pointers are device pointers. */
typedef struct cppf_pose_args {
    int64_t struct_bytes;
    const float* pc;             /* [n_points,3] */
    const float* nrm;            /* [n_points,3] */
    const void* idx;             /* [n_pairs,2] int32/int64 pairs (nocs/inference.py:177), or NULL: all n_points^2 ordered pairs */
    const float* pe_blob;        /* packed PointEncoder weights (cppf_pe_blob_floats) */
    const float* tc_blob;        /* packed PPFEncoder weights (cppf_tc_blob_floats) */
    const float* lut;            /* [136] bin decode table (see "fused per-object path") */
    const float* sphere;         /* [n_sphere,3] orientation bins (utils/util.py:102-118) */
    const float* uniforms;       /* optional [n_pairs,4] sampling uniforms; NULL -> Philox keyed by (seed, pair) */
    const uint8_t* inject_bins;  /* optional [n_pairs,inject_cols]: overwrite the sampled bins (benchmark/test aid) */
    void* workspace;             /* cppf_pose_workspace_bytes(...) bytes */
    double* record;              /* out */
    void* timing;                /* optional cppf_timing_create() handle: CUDA events around every stage */
    int64_t n_pairs;             /* ignored when idx == NULL */
    int64_t workspace_bytes;
    int64_t rot_subsample;       /* nocs/inference.py:279-281 (10000); 0 = every survivor */
    uint64_t seed;
    double res_host;             /* cfg.res as the float64 the host multiplies the winning cell with (nocs/inference.py:209);
                                    0 = (double)res */
    int n_points;
    int idx_is_64;
    int knn;                     /* config/config.yaml:21 (60) */
    int n_rots;                  /* 72 */
    int adaptive;
    int regress_right;
    int n_sphere;
    int inject_cols;
    int max_cells;               /* capacity of the shared-memory vote grid: <= cppf_vote_private_max_cells() */
    int routed_max_cells;        /* 0, or the capacity for larger grids voted through routed x-slabs
                                    (cppf_vote_routed): <= 8 * 56320 */
    int sample_pairs;            /* 1 (with idx == NULL): draw the n_pairs random point pairs of nocs/inference.py:177 on the
                                    device (Philox keyed by seed, uniform over [0, n_points)^2) instead of enumerating all
                                    n_points^2 pairs */
    float res;
    float tol;                   /* back-vote tolerance, float32(3 * res) at nocs/inference.py:226 */
    float cos_thr;               /* float32(cos(angle_prec)) at :283 */
} cppf_pose_args;
int cppf_pose_record_doubles(void);
int cppf_pose_args_bytes(void);            /* sizeof(cppf_pose_args) as compiled: lets a binding check its mirror of the struct */
int64_t cppf_pose_workspace_bytes(int n_points, int64_t n_pairs, int knn, int max_cells, int routed_max_cells, int n_rots,
                                  int n_sphere, int64_t rot_subsample);
int cppf_pose_fused(const cppf_pose_args* args, void* stream);

/* The object loop of nocs/inference.py:120-129 (sunrgbd/inference.py:115) as ONE call: args[0..n_objects) are enqueued
 * like n_objects cppf_pose_fused calls, object i on internal worker stream (i % n_streams), the worker streams forked from
 * and joined back into `stream` with events (work enqueued on `stream` before the call is visible to every object; work
 * enqueued after it sees every record).  Small objects (the reference's P = 100 000 regime fills ~1.3 waves of the GPU per
 * kernel) overlap across the worker streams, and the ~25 launches per object are issued by n_threads host threads, so
 * neither the GPU nor the launching thread is the per-object bottleneck.
 * Objects that share a worker stream may share a workspace; objects on different worker streams must not.  Records are
 * identical to the per-object calls (same kernels, same arguments).  args[i].timing must be NULL.
 * n_streams in [1, 16], n_threads in [1, n_streams].  Returns the first error. */
int cppf_pose_batch(const cppf_pose_args* args, int n_objects, int n_streams, int n_threads, void* stream);

/* Stage timing for cppf_pose_fused: CUDA events recorded on the launching stream around every stage.
 * cppf_timing_collect adds the elapsed milliseconds per stage of every call recorded since the last collect
 * to h_ms_sum[cppf_timing_stages()] (HOST array), returns the number of calls and resets the handle. */
void* cppf_timing_create(void);
void cppf_timing_destroy(void* timing);
int cppf_timing_reserve(void* timing, int n_calls);    /* pre-create the events of n_calls calls */
int cppf_timing_stages(void);
const char* cppf_timing_stage_name(int stage);
int cppf_timing_collect(void* timing, float* h_ms_sum);

/* ==== per-object pre-processing =====================
 * What nocs/inference.py:131-142 does on the host with numpy / MinkowskiEngine / open3d before the hot path,
 * on the device.  Coordinates stay float64 until after the voxel quantisation, where the reference casts (:141).
 *
 * cppf_backproject replaces utils/util.py:598-631 `backproject` + nocs/inference.py:132,136-137: for every pixel
 * with mask != 0 and depth > 0, in row-major (np.where) order, out_pts[k] = ((Kinv @ (u, v, 1)) * z / w) / depth_scale
 * (float64 [<= H*W, 3]; the two axis flips of the reference cancel), out_pix[k] = v * width + u, *out_count = number
 * of points.  depth: uint16 (depth_is_u16, the NOCS png format) or float32 [H, W]; h_intrinsics_inv: 9 doubles on the
 * HOST (np.linalg.inv(intrinsics), utils/util.py:599).  scratch: cppf_backproject_scratch_bytes(H, W).
 *
 * cppf_voxel_first stands in for ME.utils.sparse_quantize(pc, return_index=True, quantization_size=res)[1]
 * (nocs/inference.py:140): voxel = floor(p / voxel) per axis; the FIRST point of every occupied voxel is kept, indices
 * increasing (MinkowskiEngine's own pick/order inside a voxel is an implementation detail of its hash map).
 * pts float64 [n_max, 3], *count (optional device int64) limits the valid prefix; out_index int64 [<= n_max],
 * out_pc (optional) float32 [<= n_max, 3] = float32(pts[out_index]) (:141).
 *
 * cppf_normals_pca stands in for open3d estimate_normals(KDTreeSearchParamKNN(knn)) (utils/util.py:61-65): covariance
 * of the k nearest neighbours (self included, cppf_knn) in float64, unit eigenvector of the smallest eigenvalue.
 * orient 0: raw sign (open3d leaves it unspecified); 1: towards the camera at the origin.  nbrs_scratch: int64
 * [n_points, k]. */
int64_t cppf_backproject_scratch_bytes(int height, int width);
int cppf_backproject(const void* depth, int depth_is_u16, const uint8_t* mask, int height, int width,
                     const double* h_intrinsics_inv, double depth_scale, double* out_pts, int64_t* out_pix,
                     int64_t* out_count, void* scratch, void* stream);
int64_t cppf_voxel_scratch_bytes(int64_t n_max);
int cppf_voxel_first(const double* pts, const int64_t* count, int64_t n_max, double voxel, float* out_pc,
                     int64_t* out_index, int64_t* out_count, void* scratch, void* stream);
int cppf_normals_pca(const float* pc, int n_points, int k, int orient, int64_t* nbrs_scratch, float* normals, void* stream);

/* ==== scene-scale ("zero-shot") mode: nocs/zero_shot.ipynb ============
 * The notebook votes ~5 M random pairs of a whole depth frame into one scene-sized grid (cppf_ppf_vote), smooths it,
 * extracts several peaks greedily and refines each with the per-object entry points above.  Cell = code cell of the
 * notebook.
 *
 * cppf_pair_filter (cell 6): out_keep[p] = 0 for indistinguishable pairs (|n_a.n_b| > 0.9 and |ab.n_a| < 0.1 and
 * |ab.n_b| < 0.1, ab the unit vector a - b), else 1.
 * cppf_gaussian3d (cell 9): scipy.ndimage.gaussian_filter(grid, sigma, truncate=truncate) for a float32 [gx,gy,gz] grid:
 * separable, radius int(truncate*sigma + 0.5), mode 'reflect', axis 0 then 1 then 2, each pass accumulated in float64 and
 * rounded to float32 like scipy.  tmp: a second [gx,gy,gz] float32 buffer; out may not alias grid.
 * cppf_scene_proposals (cell 9): greedy peaks of `grid` (modified in place: accepted boxes are zeroed): argmax, contrast =
 * value - mean of the means over the 12 edges of the box loc +- margin; a proposal is kept while contrast > thresh; the
 * loop stops at contrast < thresh or contrast < rel_stop * first contrast (0.7 in the notebook), or at max_props.
 * h_out: HOST float [max_props][5] = loc x, y, z, value, contrast.  Returns the proposal count, or -(CUDA error).
 * scratch: 64 device bytes.  Synchronises the stream once per proposal. */
int cppf_pair_filter(const float* pc, const float* nrm, const void* idx, int idx_is_64, int n_points, int64_t n_pairs,
                     uint8_t* out_keep, void* stream);
int cppf_gaussian3d(const float* grid, float* out, float* tmp, int gx, int gy, int gz, double sigma, double truncate,
                    void* stream);
int cppf_scene_proposals(float* grid, int gx, int gy, int gz, float thresh, int margin, float rel_stop, int max_props,
                         float* h_out, void* scratch, void* stream);

/* ==== one object over several GPUs: a ROW BLOCK of the dense pair matrix (second axis) ==========
 * The `_rows` forms of the dense (idx == NULL) entry points enumerate the pairs (a, b) with a in [row0, row0 + n_pairs /
 * n_points) and every b, in row-major order: pair p of the launch is (row0 + p / n_points, p % n_points), and bins / tail /
 * mask / pos are indexed by that LOCAL p (n_pairs entries).  They keep every dense-mode shortcut (row-aligned MMA tiles with
 * the a-side table row as a row constant, tiled vote batches, no index list in HBM).  The Philox stream of
 * cppf_encode_sample_tc_rows is keyed by the pair's index in the WHOLE matrix, so the row blocks of all ranks together draw
 * exactly what one full-matrix launch draws.  cppf_vote_fast_rows leaves its exact integer sums in `scratch` like
 * cppf_vote_fast (all_reduce them, then cppf_vote_finalize). */
int cppf_encode_sample_tc_rows(const float* pc, const float* nrm, const float* table, const float* tc_blob,
                               const void* idx, int idx_is_64, int n_points, int64_t n_pairs, int row0,
                               const float* uniforms, uint64_t seed, int heads, uint8_t* bins, float* tail, float* dbg_t,
                               void* stream);
int cppf_vote_fast_rows(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, int row0,
                        float* grid, void* scratch, const float* corner, float res, int n_points, int64_t n_pairs,
                        int n_rots, int gx, int gy, int gz, int adaptive, void* stream);
int cppf_backvote_bins_rows(const float* points, const uint8_t* bins, const float* lut, int row0, uint8_t* out_mask,
                            const float* corner, const int64_t* argmax_flat, float res, float tol, double res_host,
                            int n_points, int64_t n_pairs, int n_rots, int gx, int gy, int gz, void* stream);
int cppf_rot_hist_rows(const float* points, const uint8_t* bins, const float* lut, int row0, const int64_t* pos,
                       const int64_t* count, const float* sphere, float* counts, int n_points, int n_rots, int n_bins,
                       int which, int64_t max_samples, uint64_t offset_seed, float thr, void* stream);
int cppf_survivor_stats_rows(const float* points, const float* nrm, const float* tail, int row0, const int64_t* pos,
                             const int64_t* count, const float* sphere, const int64_t* best_up, const int64_t* best_right,
                             double* out, int n_points, int64_t n_pairs, void* stream);

/* ==== measurement aids (bench.py's roofline block; never on the pose path) =====================================
 * cppf_vote_count: the ALGORITHMIC work of one centre-vote launch over these inputs, counted on the device with the
 * reference's own acceptance test (models/voting.py:21-39).  out3 (device, 3 x uint64): [0] rotation steps walked
 * (sum over non-degenerate pairs of adaptive_n_rots, :31-32), [1] candidates that pass the in-bounds test (:36-39) --
 * 8 x this is the number of trilinear atomic adds (:56-63) the reference issues and every vote kernel here performs --
 * [2] non-degenerate pairs (:21).  Exactly one of mu_nu / bins (+lut) is given; idx == NULL enumerates all N^2 pairs.
 * cppf_peak_shared_atomics: G atomic adds / s this GPU sustains (best of `reps` timed launches, CUDA events on `stream`,
 * synchronises) on a [gx,gy,gz] u32 grid in shared memory, one 1024-thread CTA per SM, nothing else in the loop:
 * conflict_free = 0: 32 lanes x the 8-corner splat of a uniformly random base cell each (the bank behaviour of an
 * unsorted vote); 1: lane l always in bank l (the hardware roof, one wavefront per instruction).
 * cppf_peak_global_red: the same 8-corner pattern as fp32 reductions on a grid in global memory (the reference's own
 * atomicAdd, models/voting.py:56-63). */
int cppf_vote_count(const float* points, const float* mu_nu, const uint8_t* bins, const float* lut, const void* idx,
                    int idx_is_64, const float* corner, float res, int n_points, int64_t n_pairs, int n_rots, int gx,
                    int gy, int gz, int adaptive, uint64_t* out3, void* stream);
int cppf_peak_shared_atomics(int gx, int gy, int gz, int conflict_free, int reps, double* g_atomics_per_s, void* stream);
int cppf_peak_global_red(int gx, int gy, int gz, int reps, double* g_atomics_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CPPF_B200_H */

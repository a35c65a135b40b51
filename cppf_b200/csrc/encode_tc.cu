// Pair encoder + categorical sampling on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as encode_sample_kernel (fused.cu): for every pair the PPF tuple, the ResLayer
// stack of models/model.py:117-137 and the `final` layer are evaluated on chip, one categorical
// draw per head is taken (nocs/inference.py:183-188, 245-256) and only 4 bin bytes + 5 tail
// logits per pair reach HBM.  Here the dense layers are true GEMMs on the tensor pipe:
//
//   * a tile is 128 pairs = the M dimension of one tcgen05.mma (cta_group::1, M = 128);
//     thread t of a 128-thread group owns pair (row) t = TMEM lane t for the whole chain, so
//     residuals and biases stay in that thread's registers and no activation ever crosses threads;
//   * every layer is  D[128 x N] (TMEM, fp32) = A[128 x K] (smem) . W[N x K]^T (smem)  in 3xTF32:
//     operands are split into a tf32 `hi` part and the fp32 remainder `lo`, and
//     lo.hi + hi.lo + hi.hi is accumulated in TMEM -- fp32-grade logits (|err| ~ 2^-21 per product)
//     from a 10-bit-mantissa pipe, which is what the 1e-4 parity tolerance needs;
//   * operands use the no-swizzle K-major canonical layout: element (row r, k) of a [rows x K]
//     operand sits at byte  (k/4) * rows*16 + r*16 + (k%4)*4 , i.e. one 16-byte chunk per row per
//     "K plane"; a thread writes its own row with conflict-free 128-bit stores;
//   * ResLayer 0 (84 -> 32) never sees the 80 feature columns: their products are pre-projected
//     per point (cppf_tc_preproject) and gathered; only the 4 PPF columns go through a K = 8 MMA;
//   * adjacent linear maps are composed on the host (no nonlinearity sits between fc2 of one
//     ResLayer and fc1/fc0 of the next, nor between fc2_2 and `final`), so a tile needs 4 dependent
//     MMA steps instead of 8 and 39 MMAs instead of 69 -- see "chain algebra" below;
//   * a CTA holds 4 independent 128-thread groups (4 tiles in flight per SM) sharing one copy of
//     the weights; each group has its own A buffers, mbarrier and 128 TMEM columns, and its own
//     elected MMA-issuing thread, so one group's SIMT epilogue overlaps the others' MMAs.
#include "common.cuh"
#include "tc_common.cuh"

#include "../../include/cppf_b200.h"

namespace cppf {
namespace tc {

constexpr int kGroups = 4;
constexpr int kThreads = kGroups * kTile;

// ---- chain algebra (models/model.py:26-31,134-137; W10_2 = [fc1_2 ; fc0_2], b10_2 = [b1_2 ; b0_2 + b2_2])
//   v1 = fc1_0(x)                     r  = fc0_0(x) + b2_0            h = relu(v1)        x1 = W2_0 h + r
//   u  = relu(W1_1 x1 + b1_1)         = relu((W1_1 W2_0) h + q1),     q1 = W1_1 r + b1_1
//   x2 = W2_1 u + b2_1 + x1
//   t  = W10_2 x2 + b10_2             = (W10_2 W2_1) u + (W10_2 W2_0) h + q2,   q2 = W10_2 (r + b2_1) + b10_2
//   u2 = relu(t[0:16]), r2 = t[16:32]  x3 = W2_2 u2 + r2
//   logits = Wf x3 + bf               = [Wf W2_2 | Wf] [u2 ; r2] + bf
// v1, q1, q2 are linear in (feat[a], feat[b], ppf): the feature parts are pre-projected per point into a
// [N, 192] table (A side 96 = v1|q1|q2 with the biases, B side 96), the ppf part is the K = 8 front MMA.
//   step 0: D[0:96]   = ppf . [P_v1 | P_q1 | P_q2]                         -> h  = relu(D[0:32] + TA + TB)
//   step 1: D[32:96] += h . [W1_1 W2_0 ; W10_2 W2_0]^T                      -> u  = relu(D[32:64] + TA + TB)
//   step 2: D[64:96] += u . (W10_2 W2_1)^T                                  -> [u2 ; r2] from D[64:96] + TA + TB
//   step 3: D[0:112]  = [u2 ; r2] . [Wf W2_2 | Wf]^T (mu | nu | up | tail)  -> softmax + inverse-CDF draws
//
// ---- packed blob (floats).  First the per-point projection (read by cppf_tc_preproject only), then the
// part staged in shared memory: every MMA operand is [K/4][N][4] (canonical K-major), hi block then lo block.
constexpr int kTabCols = 192;
constexpr int kOffPreW = 0;                       // [40][192] k-major
constexpr int kOffPreB = kOffPreW + 40 * kTabCols;   // [192] (A side carries the biases)
constexpr int kOffSmem = kOffPreB + kTabCols;
// offsets relative to kOffSmem
constexpr int kOffWp = 0;                         // N = 96,  K = 8 (k 4..7 zero)
constexpr int kOffWs1 = kOffWp + 2 * 96 * 8;      // N = 64,  K = 32
constexpr int kOffWs2 = kOffWs1 + 2 * 64 * 32;    // N = 32,  K = 32
constexpr int kOffWh = kOffWs2 + 2 * 32 * 32;     // N = 112, K = 32: mu 32 | nu 32 | up 36 | tail 5 | 0 x 7
constexpr int kOffWr = kOffWh + 2 * 112 * 32;     // N = 48,  K = 32: right 36 | 0 x 12
// Row-constant addends as MMA operands: B[N x 8] holds the hi part of a vector v in k = 0 and its lo part in k = 4; the A
// side is a constant "ones" operand (k = 0 and k = 4 of every row = 1), so ONE MMA D = ones . B^T puts v (exact, hi + lo)
// into every row of the accumulators.  Used for the head biases and, in dense mode, for the a-side table row of a tile.
constexpr int kOffBh = kOffWr + 2 * 48 * 32;      // N = 112, K = 8
constexpr int kOffBr = kOffBh + 112 * 8;          // N = 48,  K = 8
constexpr int kSmemFloats = kOffBr + 48 * 8;
constexpr int kBlobFloats = kOffSmem + kSmemFloats;

constexpr int kTaBytes = 2 * 96 * 16;              // per group: the a-side table row of the tile as a [96 x 8] B operand
constexpr int kSmemBytes = kSmemFloats * 4 + kOnesBytes + kGroups * (kGroupBytes + kTaBytes);
constexpr int kTmemColsPerGroup = 128;

struct Params {
    const float* pc;
    const float* nrm;
    const float* table;
    const float* blob;
    const void* idx;
    const float* uniforms;
    unsigned long long seed;
    uint8_t* bins;
    float* tail;
    float* dbg_t;               // optional [n_pairs][32]: t = [fc1_2(x2) ; fc0_2(x2) + fc2_2.b] (before the ReLU)
    int n_points;
    long long n_pairs;
    int heads;
    int row0;                   // dense mode: first row of the block of the pair matrix this launch covers (n_pairs / n_points rows)
};

#ifndef CPPF_TC_ST_CS
#define CPPF_TC_ST_CS 1      // tail logits leave with st.global.cs: 335 MB per object that should not push the bins out of L2
#endif
#ifndef CPPF_TC_LD_PIPE
#define CPPF_TC_LD_PIPE 2    // TMEM -> register loads of the step epilogues split in two, the second in flight while the first is consumed
#endif
#ifndef CPPF_TC_PREFETCH
#define CPPF_TC_PREFETCH 1   // dense mode: point data of the next tile is loaded one tile ahead (see the tile loop)
#endif
#ifndef CPPF_TC_TB_LDCG
#define CPPF_TC_TB_LDCG 1    // table rows bypass L1 (ld.global.cg): with 224 KB of shared memory the L1 is too small to keep them
#endif
__device__ __forceinline__ float4 ld_tab(const float4* p) {
#if CPPF_TC_TB_LDCG
    return __ldcg(p);
#else
    return __ldg(p);
#endif
}

// One categorical draw from NB logits in registers.  The logits arrive in log2 units (the host folds log2 e into
// the head weights), so e_k = 2^(l_k - max l) is one FADD + one MUFU.EX2; bin = #{k : cumsum(e)_k <= u * sum(e)}
// (same draw as the sequential scan of fused.cu / the oracle, with the partial sums associated as a binary tree:
// 31 adds build the tree, 5 compare-and-descend levels walk it).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sel4(float a, float b, float c, float d, bool lo, bool hi) {   // index = hi*2 + lo
    const float x = lo ? b : a, y = lo ? d : c;
    return hi ? y : x;
}

// e[0:32] -> (bin, total); t_in = u * total is applied by the caller through `t`
__device__ __forceinline__ int tree32(const float (&e)[32], const float (&s2)[16], const float (&s4)[8], const float (&s8)[4],
                                      float s16_0, float t) {
    const bool b4 = t >= s16_0;                                   // bin bit 4
    t -= b4 ? s16_0 : 0.f;
    const float l8 = b4 ? s8[2] : s8[0];
    const bool b3 = t >= l8;
    t -= b3 ? l8 : 0.f;
    const float l4 = sel4(s4[0], s4[2], s4[4], s4[6], b3, b4);
    const bool b2 = t >= l4;
    t -= b2 ? l4 : 0.f;
    const float l2 = b4 ? sel4(s2[8], s2[10], s2[12], s2[14], b2, b3) : sel4(s2[0], s2[2], s2[4], s2[6], b2, b3);
    const bool b1 = t >= l2;
    t -= b1 ? l2 : 0.f;
    const float q0 = sel4(e[0], e[2], e[4], e[6], b1, b2), q1 = sel4(e[8], e[10], e[12], e[14], b1, b2);
    const float q2 = sel4(e[16], e[18], e[20], e[22], b1, b2), q3 = sel4(e[24], e[26], e[28], e[30], b1, b2);
    const float l1 = sel4(q0, q1, q2, q3, b3, b4);
    const bool b0 = t >= l1;
    return (b4 ? 16 : 0) + (b3 ? 8 : 0) + (b2 ? 4 : 0) + (b1 ? 2 : 0) + (b0 ? 1 : 0);
}

template <int NB>
__device__ __forceinline__ int sample_regs(float (&l)[NB], float u) {
    static_assert(NB == 32 || NB == 36, "heads have 32 (translation) or 36 (rotation) bins");
    float m = l[0];
#pragma unroll
    for (int k = 1; k < NB; ++k) m = fmaxf(m, l[k]);
    float e[32], s2[16], s4[8], s8[4];
#pragma unroll
    for (int k = 0; k < 32; ++k) e[k] = ex2_approx(l[k] - m);
#pragma unroll
    for (int k = 0; k < 16; ++k) s2[k] = e[2 * k] + e[2 * k + 1];
#pragma unroll
    for (int k = 0; k < 8; ++k) s4[k] = s2[2 * k] + s2[2 * k + 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) s8[k] = s4[2 * k] + s4[2 * k + 1];
    const float s16_0 = s8[0] + s8[1], s16_1 = s8[2] + s8[3];
    const float tot32 = s16_0 + s16_1;
    if (NB == 32) return tree32(e, s2, s4, s8, s16_0, u * tot32);
    // 36 bins: the four extra bins sit behind the tree
    const float x0 = ex2_approx(l[NB - 4] - m), x1 = ex2_approx(l[NB - 3] - m), x2 = ex2_approx(l[NB - 2] - m),
                x3 = ex2_approx(l[NB - 1] - m);
    const float x01 = x0 + x1, tot = tot32 + (x01 + (x2 + x3));
    const float t = u * tot;
    const int in_tree = tree32(e, s2, s4, s8, s16_0, t);
    const float r = t - tot32;
    const int extra = 32 + (r >= x0 ? 1 : 0) + (r >= x01 ? 1 : 0) + (r >= x01 + x2 ? 1 : 0);
    return t >= tot32 ? extra : in_tree;
}

__device__ __forceinline__ void ppf_of(f3 pa, f3 pb, f3 na, f3 nb, float (&ppf)[4]) {   // models/model.py:120-129
    const f3 d = pa - pb;
    const float dn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    const float inv = dn + 1e-7f;
    const f3 dh = {d.x / inv, d.y / inv, d.z / inv};
    ppf[0] = na.x * dh.x + na.y * dh.y + na.z * dh.z;
    ppf[1] = nb.x * dh.x + nb.y * dh.y + nb.z * dh.z;
    ppf[2] = na.x * nb.x + na.y * nb.y + na.z * nb.z;
    ppf[3] = dn;
}

template <bool IDX64, bool DENSE>
__global__ void __launch_bounds__(kThreads, 1) encode_sample_tc_kernel(const Params prm) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar[kGroups];
    __shared__ uint32_t s_tmem;
    float* sblob = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x;
    // warp index through a lane-0 broadcast: the compiler then knows that everything derived from it (group, operand and
    // TMEM addresses, descriptors) is warp-uniform and keeps it in uniform registers -- tcgen05.mma takes its operands from
    // there, and with per-thread values every MMA was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~13
    // instructions per MMA on the critical path of the group's leader warp)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = warp >> 2, tg = tid & 127;            // group, row within the tile
    unsigned char* s_ones = smem + kSmemFloats * 4;
    unsigned char* s_ta = s_ones + kOnesBytes + g * kTaBytes;                 // this group's a-side row operand (dense mode)
    unsigned char* a_hi = s_ones + kOnesBytes + kGroups * kTaBytes + g * kGroupBytes;

    {   // weights -> shared memory (one copy per CTA), TMEM allocation, barriers
        const float4* src = reinterpret_cast<const float4*>(prm.blob + kOffSmem);
        float4* dst = reinterpret_cast<float4*>(sblob);
        for (int i = tid; i < kSmemFloats / 4; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < kOnesBytes / 16; i += kThreads)          // k = 0 and k = 4 of every row are 1
            reinterpret_cast<float4*>(s_ones)[i] = make_float4(1.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < kGroups * kTaBytes / 16; i += kThreads)  // only k = 0 / k = 4 of these rows are ever rewritten
            reinterpret_cast<float4*>(s_ones + kOnesBytes)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                         "r"(kGroups * kTmemColsPerGroup)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (tid < kGroups) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[tid])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t tm = s_tmem + g * kTmemColsPerGroup;                    // MMA destination (lane 0, column base)
    const uint32_t tml = tm + ((uint32_t)((warp & 3) * 32) << 16);         // this warp's lane quarter for tcgen05.ld
    const uint32_t bar = smem_u32(&s_bar[g]);
    const uint32_t sA = smem_u32(a_hi), sAl = sA + kABytes;
    const uint32_t sW = smem_u32(sblob);
    const uint32_t sOnes = smem_u32(s_ones), sTa = smem_u32(s_ta);
    const bool lead_warp = (warp & 3) == 0;              // warp-uniform; one elected lane of it issues the MMAs
    uint32_t phase = 0;

    // dense mode: a tile is 128 consecutive points b of ONE point a (ragged at the end of a row), so that the a-side table row
    // is a row constant of the tile and enters through the ones-operand MMA instead of 96 FADDs per pair
    const int tiles_per_row = (prm.n_points + kTile - 1) / kTile;
    const long long n_tiles = DENSE ? (prm.n_pairs / prm.n_points) * tiles_per_row : (prm.n_pairs + kTile - 1) / kTile;
    // The point data of a tile is fetched ONE TILE AHEAD (dense mode: both points' xyz + normal and this
    // thread's chunk of the a-side table row), so that the L2 round trip at the head of a tile -- an SM configured
    // with 224 KB of shared memory has next to no L1 -- overlaps the previous tile instead of the MMA chain's critical path.
    const long long t_stride = (long long)gridDim.x * kGroups;
    const long long ts = prm.n_points;
    auto coords = [&](long long tile, long long& p, bool& valid, int& a, int& b) {
        a = 0;
        b = 0;
        if (DENSE) {
            const int r = (int)(tile / tiles_per_row);                            // row within the block
            a = r + prm.row0;
            b = (int)(tile - (long long)r * tiles_per_row) * kTile + tg;
            valid = b < prm.n_points;
            p = (long long)r * prm.n_points + b;                                  // position in this launch's outputs
            if (!valid) b = 0;
        } else {
            p = tile * kTile + tg;
            valid = p < prm.n_pairs;
            if (valid) pair_ab<IDX64>(prm.idx, p, prm.n_points, a, b);
        }
    };
    long long n_p = 0;
    bool n_valid = false;
    int n_a = 0, n_b = 0;
    f3 n_pa = {0.f, 0.f, 0.f}, n_pb = n_pa, n_na = n_pa, n_nb = n_pa;
    float4 n_ta = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch = [&](long long tile) {
        coords(tile, n_p, n_valid, n_a, n_b);
#if CPPF_TC_PREFETCH
        if (DENSE) {
            n_pa = ld3(prm.pc, n_a); n_pb = ld3(prm.pc, n_b); n_na = ld3(prm.nrm, n_a); n_nb = ld3(prm.nrm, n_b);
            if (tg < 24) n_ta = __ldg(reinterpret_cast<const float4*>(prm.table) + n_a + tg * ts);
        }
#endif
    };
    long long tile = (long long)blockIdx.x * kGroups + g;
    if (DENSE && tile < n_tiles) fetch(tile);
    for (; tile < n_tiles; tile += t_stride) {
        if (!DENSE) fetch(tile);                // indexed pairs: no look-ahead (the ta rows already fill the register file)
        const long long p = n_p;
        const bool valid = n_valid;
        const int a = n_a, b = n_b;
        float ppf[4];
#if CPPF_TC_PREFETCH
        if (DENSE) ppf_of(n_pa, n_pb, n_na, n_nb, ppf);
        else
#endif
            ppf_of(ld3(prm.pc, a), ld3(prm.pc, b), ld3(prm.nrm, a), ld3(prm.nrm, b), ppf);
        const float4 ta_row = n_ta;
        float4 u4;
        if (prm.uniforms != nullptr) {
            u4 = valid ? __ldg(reinterpret_cast<const float4*>(prm.uniforms) + p) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            // the counter is the pair's index in the WHOLE pair matrix, so a row block draws what the full launch draws
            const long long pg = DENSE ? p + (long long)prm.row0 * prm.n_points : p;
            const uint4 w = philox4x32_10(make_uint4((uint32_t)pg, (uint32_t)((unsigned long long)pg >> 32), 0u, 0u),
                                          make_uint2((uint32_t)prm.seed, (uint32_t)(prm.seed >> 32)));
            u4 = make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
        }
        // planar table: float4 chunk j of point n sits at [(side*24 + j) * N + n], so the 32 lanes of a warp
        // (consecutive b in dense mode) read 512 contiguous bytes per load and the a side is a broadcast
        const float4* TA = reinterpret_cast<const float4*>(prm.table) + a;
        const float4* TB = reinterpret_cast<const float4*>(prm.table) + (long long)(kTabCols / 8) * prm.n_points + b;
        float4 ta[8], tb[8];
        float x[32];

        // ---- step 0: ppf columns of v1 | q1 | q2 (K = 8, N = 96) -> h = relu(fc1_0(x))            models/model.py:27,29
        st_chunk(a_hi, 0, tg, ppf[0], ppf[1], ppf[2], ppf[3]);
        st_chunk(a_hi, 1, tg, 0.f, 0.f, 0.f, 0.f);
        if (DENSE) {
            if (tg < 24) {                     // rows 4 tg .. 4 tg + 3 of the [96 x 8] operand: k = 0 <- hi, k = 4 <- lo
#if CPPF_TC_PREFETCH
                const float4 v = ta_row;
#else
                const float4 v = __ldg(TA + tg * ts);
#endif
                const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float h = tf32_hi(vs[j]);
                    *reinterpret_cast<float*>(s_ta + (4 * tg + j) * 16) = h;
                    *reinterpret_cast<float*>(s_ta + 96 * 16 + (4 * tg + j) * 16) = vs[j] - h;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ta[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                tb[q] = ld_tab(TB + q * ts);
            }
            if (tile + t_stride < n_tiles) fetch(tile + t_stride);          // next tile's point data: consumed one tile later
            CPPF_TC_STEP((issue_bias<96>(tm, sOnes, sTa), issue3<96, 8, true>(tm, sA, sAl, sW + kOffWp * 4)));
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ta[q] = ld_tab(TA + q * ts);
                tb[q] = ld_tab(TB + q * ts);
            }
            CPPF_TC_STEP((issue3<96, 8>(tm, sA, sAl, sW + kOffWp * 4)));
        }
#if CPPF_TC_LD_PIPE
        {
            uint32_t xr[32];
            tmem_ld32_first(tml, xr);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q == 4) tmem_ld32_second(xr);
                st_chunk(a_hi, q, tg, fmaxf((DENSE ? tb[q].x : ta[q].x + tb[q].x) + __uint_as_float(xr[4 * q]), 0.f),
                         fmaxf((DENSE ? tb[q].y : ta[q].y + tb[q].y) + __uint_as_float(xr[4 * q + 1]), 0.f),
                         fmaxf((DENSE ? tb[q].z : ta[q].z + tb[q].z) + __uint_as_float(xr[4 * q + 2]), 0.f),
                         fmaxf((DENSE ? tb[q].w : ta[q].w + tb[q].w) + __uint_as_float(xr[4 * q + 3]), 0.f));
            }
        }
#else
        tmem_ld32(tml, x);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            st_chunk(a_hi, q, tg, fmaxf((DENSE ? tb[q].x : ta[q].x + tb[q].x) + x[4 * q], 0.f), fmaxf((DENSE ? tb[q].y : ta[q].y + tb[q].y) + x[4 * q + 1], 0.f),
                     fmaxf((DENSE ? tb[q].z : ta[q].z + tb[q].z) + x[4 * q + 2], 0.f), fmaxf((DENSE ? tb[q].w : ta[q].w + tb[q].w) + x[4 * q + 3], 0.f));
#endif
        // ---- step 1: [W1_1 W2_0 ; W10_2 W2_0] h accumulated onto q1 | q2 -> u = relu(fc1_1(x1))
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (!DENSE) ta[q] = ld_tab(TA + (8 + q) * ts);
            tb[q] = ld_tab(TB + (8 + q) * ts);
        }
        CPPF_TC_STEP((issue3<64, 32, true>(tm + 32, sA, sAl, sW + kOffWs1 * 4)));
#if CPPF_TC_LD_PIPE
        {
            uint32_t xr[32];
            tmem_ld32_first(tml + 32, xr);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q == 4) tmem_ld32_second(xr);
                st_chunk(a_hi, q, tg, fmaxf((DENSE ? tb[q].x : ta[q].x + tb[q].x) + __uint_as_float(xr[4 * q]), 0.f),
                         fmaxf((DENSE ? tb[q].y : ta[q].y + tb[q].y) + __uint_as_float(xr[4 * q + 1]), 0.f),
                         fmaxf((DENSE ? tb[q].z : ta[q].z + tb[q].z) + __uint_as_float(xr[4 * q + 2]), 0.f),
                         fmaxf((DENSE ? tb[q].w : ta[q].w + tb[q].w) + __uint_as_float(xr[4 * q + 3]), 0.f));
            }
        }
#else
        tmem_ld32(tml + 32, x);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            st_chunk(a_hi, q, tg, fmaxf((DENSE ? tb[q].x : ta[q].x + tb[q].x) + x[4 * q], 0.f), fmaxf((DENSE ? tb[q].y : ta[q].y + tb[q].y) + x[4 * q + 1], 0.f),
                     fmaxf((DENSE ? tb[q].z : ta[q].z + tb[q].z) + x[4 * q + 2], 0.f), fmaxf((DENSE ? tb[q].w : ta[q].w + tb[q].w) + x[4 * q + 3], 0.f));
#endif
        // ---- step 2: (W10_2 W2_1) u accumulated onto q2 + (W10_2 W2_0) h -> t = [fc1_2(x2) ; fc0_2(x2) + fc2_2.b]
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (!DENSE) ta[q] = ld_tab(TA + (16 + q) * ts);
            tb[q] = ld_tab(TB + (16 + q) * ts);
        }
        CPPF_TC_STEP((issue3<32, 32, true>(tm + 64, sA, sAl, sW + kOffWs2 * 4)));
#if CPPF_TC_LD_PIPE >= 2
        {
            uint32_t xr[32];
            tmem_ld32_first(tml + 64, xr);
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = __uint_as_float(xr[q]);
            tmem_ld32_second(xr);
#pragma unroll
            for (int q = 16; q < 32; ++q) x[q] = __uint_as_float(xr[q]);
        }
#else
        tmem_ld32(tml + 64, x);
#endif
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            x[4 * q] += (DENSE ? tb[q].x : ta[q].x + tb[q].x);
            x[4 * q + 1] += (DENSE ? tb[q].y : ta[q].y + tb[q].y);
            x[4 * q + 2] += (DENSE ? tb[q].z : ta[q].z + tb[q].z);
            x[4 * q + 3] += (DENSE ? tb[q].w : ta[q].w + tb[q].w);
        }
        if (prm.dbg_t != nullptr && valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                reinterpret_cast<float4*>(prm.dbg_t + p * 32)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)           // u2 = relu(t[0:16])   (k 0..15 of the head operand)
            st_chunk(a_hi, q, tg, fmaxf(x[4 * q], 0.f), fmaxf(x[4 * q + 1], 0.f), fmaxf(x[4 * q + 2], 0.f), fmaxf(x[4 * q + 3], 0.f));
#pragma unroll
        for (int q = 4; q < 8; ++q)           // r2 = t[16:32]        (k 16..31)
            st_chunk(a_hi, q, tg, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        // ---- step 3: `final` o fc2_2 on [u2 ; r2] (models/model.py:31,137): mu | nu | up | tail -> columns 0:112
        CPPF_TC_STEP((issue_bias<112>(tm, sOnes, sW + kOffBh * 4), issue3<112, 32, true>(tm, sA, sAl, sW + kOffWh * 4)));
        uchar4 bins = make_uchar4(0, 0, 0, 0);
        if (prm.heads & 1) {
            tmem_ld32(tml, x);                 // the biases are already in the accumulators (issue_bias)
            bins.x = (unsigned char)sample_regs<32>(x, u4.x);
            tmem_ld32(tml + 32, x);
            bins.y = (unsigned char)sample_regs<32>(x, u4.y);
        }
        if (prm.heads & (2 | 8)) {
            float y[16];                       // columns 96..111: up bins 32..35, tail 0..4, padding
            tmem_ld16(tml + 96, y);
            if ((prm.heads & 8) && valid) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
#if CPPF_TC_ST_CS
                    __stcs(prm.tail + (long long)k * prm.n_pairs + p, y[4 + k]);
#else
                    prm.tail[(long long)k * prm.n_pairs + p] = y[4 + k];
#endif
                }
            }
            if (prm.heads & 2) {
                float l[36];
                {
                    float z[32];
                    tmem_ld32(tml + 64, z);
#pragma unroll
                    for (int q = 0; q < 32; ++q) l[q] = z[q];
                }
                l[32] = y[0]; l[33] = y[1]; l[34] = y[2]; l[35] = y[3];
                bins.z = (unsigned char)sample_regs<36>(l, u4.z);
            }
        }
        if (prm.heads & 4) {                   // right head: one more MMA into columns 0:48 (mu/nu already consumed)
            CPPF_TC_STEP((issue_bias<48>(tm, sOnes, sW + kOffBr * 4), issue3<48, 32, true>(tm, sA, sAl, sW + kOffWr * 4)));
            float l[36];
            {
                float z[32];
                tmem_ld32(tml, z);
#pragma unroll
                for (int q = 0; q < 32; ++q) l[q] = z[q];
                float y[16];
                tmem_ld16(tml + 32, y);
                l[32] = y[0]; l[33] = y[1]; l[34] = y[2]; l[35] = y[3];
            }
            bins.w = (unsigned char)sample_regs<36>(l, u4.w);
        }
        if (valid) reinterpret_cast<uchar4*>(prm.bins)[p] = bins;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(kGroups * kTmemColsPerGroup)
                     : "memory");
}

// Per-point projection of the feature columns of v1 | q1 | q2 (chain algebra above): column c of point n =
// feat[n] . M[:, c] + bias[c]; columns 0:96 are read by pairs whose point a is n, 96:192 by pairs whose point b
// is n.  Stored planar in float4 chunks: table4[(c/4) * N + n] (see the gather in the kernel).
__global__ void __launch_bounds__(kTabCols) tc_preproject_kernel(const float* __restrict__ feat, const float* __restrict__ blob,
                                                                 float* __restrict__ table, int n_points) {
    __shared__ float f[40];
    const int n = blockIdx.x;
    if (threadIdx.x < 40) f[threadIdx.x] = feat[(int64_t)n * 40 + threadIdx.x];
    __syncthreads();
    const int c = threadIdx.x;
    float acc = blob[kOffPreB + c];
#pragma unroll 8
    for (int k = 0; k < 40; ++k) acc = fmaf(f[k], __ldg(blob + kOffPreW + k * kTabCols + c), acc);
    table[((int64_t)(c >> 2) * n_points + n) * 4 + (c & 3)] = acc;
}

}  // namespace tc
}  // namespace cppf

using namespace cppf;

extern "C" int cppf_tc_blob_floats(void) { return tc::kBlobFloats; }
extern "C" int cppf_tc_table_cols(void) { return tc::kTabCols; }

extern "C" int cppf_tc_preproject(const float* feat, const float* tc_blob, float* table, int n_points, void* stream) {
    if (n_points <= 0) return 0;
    tc::tc_preproject_kernel<<<n_points, tc::kTabCols, 0, (cudaStream_t)stream>>>(feat, tc_blob, table, n_points);
    CPPF_LAUNCH_CHECK();
    return 0;
}

namespace cppf {
// dense pairs (idx == NULL) cover rows [row0, row0 + n_pairs / n_points) of the pair matrix: all of it, or one block
bool dense_rows_ok(const void* idx, int n_points, int64_t n_pairs, int row0) {
    if (idx != nullptr) return row0 == 0;
    if (n_points <= 0 || row0 < 0 || n_pairs % n_points != 0) return false;
    return row0 + n_pairs / n_points <= n_points;
}
}  // namespace cppf

extern "C" int cppf_encode_sample_tc_rows(const float* pc, const float* nrm, const float* table, const float* tc_blob,
                                          const void* idx, int idx_is_64, int n_points, int64_t n_pairs, int row0,
                                          const float* uniforms, uint64_t seed, int heads, uint8_t* bins, float* tail,
                                          float* dbg_t, void* stream) {
    if (n_pairs <= 0) return 0;
    if (!dense_rows_ok(idx, n_points, n_pairs, row0)) return (int)cudaErrorInvalidValue;
    if ((heads & 8) && tail == nullptr) return (int)cudaErrorInvalidValue;
    tc::Params prm{pc, nrm, table, tc_blob, idx, uniforms, seed, bins, tail, dbg_t, n_points, (long long)n_pairs, heads, row0};
    const bool dense = idx == nullptr;
    const long long n_tiles = dense ? (n_pairs / n_points) * ((n_points + tc::kTile - 1) / tc::kTile)
                                    : (n_pairs + tc::kTile - 1) / tc::kTile;
    long long ctas = (n_tiles + tc::kGroups - 1) / tc::kGroups;
    if (ctas > sm_count()) {
        // the kernel runs in rounds of kGroups tiles per CTA: the fewest CTAs that keep the round count of a full grid finish
        // at the same time and leave the other SMs to whatever else is in flight (the object loop overlaps several objects:
        // 100 000 pairs = 782 tiles = 2 rounds on 98 CTAs instead of 1.3 rounds on 148)
        const long long rounds = (n_tiles + (long long)sm_count() * tc::kGroups - 1) / ((long long)sm_count() * tc::kGroups);
        ctas = (n_tiles + rounds * tc::kGroups - 1) / (rounds * tc::kGroups);
        if (ctas > sm_count()) ctas = sm_count();
    }
    void (*kern)(const tc::Params);
    if (dense) kern = tc::encode_sample_tc_kernel<false, true>;
    else kern = idx_is_64 ? tc::encode_sample_tc_kernel<true, false> : tc::encode_sample_tc_kernel<false, false>;
    CPPF_RETURN_IF((cudaError_t)raise_dynamic_smem((const void*)kern, tc::kSmemBytes));
    kern<<<(int)ctas, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>(prm);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_encode_sample_tc(const float* pc, const float* nrm, const float* table, const float* tc_blob,
                                     const void* idx, int idx_is_64, int n_points, int64_t n_pairs, const float* uniforms,
                                     uint64_t seed, int heads, uint8_t* bins, float* tail, float* dbg_t, void* stream) {
    if (n_pairs <= 0) return 0;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    return cppf_encode_sample_tc_rows(pc, nrm, table, tc_blob, idx, idx_is_64, n_points, n_pairs, 0, uniforms, seed, heads, bins,
                                      tail, dbg_t, stream);
}

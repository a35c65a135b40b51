/* TEST INFRASTRUCTURE (oracle) -- NOT part of the shipped product path.
 *
 * Plain-C, CPU restatement of the reference's four voting kernels and of the host
 * glue that follows them.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / reference arm may load this library; the product (cppf_b200/)
 * never does.
 *
 * Parity pinning: the reference has no tests or golden vectors,
 * so this restatement is pinned against the reference's OWN kernel strings compiled
 * for the CPU (oracle/_ref/libref_voting_cpu.so, built by oracle/build_ref.py from
 * /root/reference/models/voting.py where it lies) in tests/test_oracle_vs_ref.py and
 * against fixtures minted from them (tests/golden/, oracle/make_golden.py).
 *
 * Every function cites the reference lines it follows.  Arithmetic is fp32 with the
 * same double-precision islands the CUDA-C strings have (un-suffixed literals 1e-7,
 * 0.01, 1.01 and M_PI promote the surrounding sub-expression to double).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_PI 3.14159265358979323846264338327950288

typedef struct { float x, y, z; } v3;

static inline v3 v3_sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 v3_scale(v3 a, float s) { v3 r = { a.x * s, a.y * s, a.z * s }; return r; }
static inline v3 v3_div(v3 a, float s) { v3 r = { a.x / s, a.y / s, a.z / s }; return r; }
/* helper_math.cuh:1245-1248, 1288-1291: length = sqrtf(x*x + y*y + z*z) */
static inline float v3_len(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* helper_math.cuh:1417-1420 */
static inline v3 v3_cross(v3 a, v3 b) {
    v3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
static inline v3 load3(const float* p, long i) { v3 r = { p[3 * i], p[3 * i + 1], p[3 * i + 2] }; return r; }

/* Shared pair frame: voting.py:18-29 (ppf_voting), :84-94 (backvote), :127-137 (rot_voting).
 * Returns 0 for a degenerate pair (voting.py:21), else fills unit ab, and the unit
 * in-plane axis `ex` (before scaling by odist). */
static int pair_frame(const float* points, const int32_t* idx, long i, v3* a_out, v3* ab_out, v3* ex_out) {
    v3 a = load3(points, idx[2 * i]);
    v3 b = load3(points, idx[2 * i + 1]);
    v3 ab = v3_sub(a, b);
    float len = v3_len(ab);
    if ((double)len < 1e-7) return 0;                       /* :21 double compare */
    ab = v3_div(ab, (float)((double)len + 1e-7));           /* :22 double sum, float divide */
    v3 co = { 0.f, -ab.z, ab.y };                           /* :26 */
    if ((double)v3_len(co) < 1e-7) { co.x = -ab.y; co.y = ab.x; co.z = 0.f; }   /* :27 */
    *ex_out = v3_div(co, (float)((double)v3_len(co) + 1e-7));                  /* :28 first half */
    *a_out = a;
    *ab_out = ab;
    return 1;
}

/* voting.py:8-66  ppf_voting: splat each pair's circle of candidate centres trilinearly. */
void oracle_ppf_voting(const float* points, const float* outputs, const float* probs, const int32_t* point_idxs,
                       float* grid_obj, const float* corner, float res, long n_ppfs, int n_rots,
                       int gx, int gy, int gz, int adaptive) {
    v3 cr = { corner[0], corner[1], corner[2] };
    for (long i = 0; i < n_ppfs; ++i) {
        float proj_len = outputs[2 * i], odist = outputs[2 * i + 1];
        v3 a, ab, ex;
        if (!pair_frame(points, point_idxs, i, &a, &ab, &ex)) continue;
        v3 c = v3_sub(a, v3_scale(ab, proj_len));                               /* :23 */
        float pa = probs[point_idxs[2 * i]], pb = probs[point_idxs[2 * i + 1]];
        float prob = pa > pb ? pa : pb;                                         /* :25 */
        v3 x = v3_scale(ex, odist);                                             /* :28 */
        v3 y = v3_cross(x, ab);                                                 /* :29 */
        int n = n_rots;
        if (adaptive) {                                                         /* :31 */
            int m = (int)((double)(odist / res) * (2 * ORACLE_PI));
            n = m < n_rots ? m : n_rots;
        }
        for (int r = 0; r < n; ++r) {
            float angle = (float)((double)(r * 2) * ORACLE_PI / (double)n);     /* :33 */
            float ca = cosf(angle), sa = sinf(angle);
            v3 off = v3_add(v3_scale(x, ca), v3_scale(y, sa));                  /* :34 */
            v3 g = v3_div(v3_sub(v3_add(c, off), cr), res);                     /* :35 */
            if ((double)g.x < 0.01 || (double)g.y < 0.01 || (double)g.z < 0.01 ||
                (double)g.x >= gx - 1.01 || (double)g.y >= gy - 1.01 || (double)g.z >= gz - 1.01)
                continue;                                                       /* :36-39 */
            int fx = (int)g.x, fy = (int)g.y, fz = (int)g.z;                    /* :40 truncation */
            float rx = g.x - floorf(g.x), ry = g.y - floorf(g.y), rz = g.z - floorf(g.z);   /* :42 */
            float wx[2] = { 1.f - rx, rx }, wy[2] = { 1.f - ry, ry }, wz[2] = { 1.f - rz, rz };
            for (int dx = 0; dx < 2; ++dx)
                for (int dy = 0; dy < 2; ++dy)
                    for (int dz = 0; dz < 2; ++dz) {                            /* :47-63 */
                        float w = wx[dx] * wy[dy] * wz[dz];
                        grid_obj[(long)(fx + dx) * gy * gz + (long)(fy + dy) * gz + (fz + dz)] += w * prob;
                    }
        }
    }
}

/* Same votes accumulated in double: an order-independent yardstick for the fp32
 * atomics of both the reference and the product kernels (no reference line; the
 * reference's own grid is run-to-run nondeterministic). */
void oracle_ppf_voting_f64(const float* points, const float* outputs, const float* probs, const int32_t* point_idxs,
                           double* grid_obj, const float* corner, float res, long n_ppfs, int n_rots,
                           int gx, int gy, int gz, int adaptive) {
    v3 cr = { corner[0], corner[1], corner[2] };
    for (long i = 0; i < n_ppfs; ++i) {
        float proj_len = outputs[2 * i], odist = outputs[2 * i + 1];
        v3 a, ab, ex;
        if (!pair_frame(points, point_idxs, i, &a, &ab, &ex)) continue;
        v3 c = v3_sub(a, v3_scale(ab, proj_len));
        float pa = probs[point_idxs[2 * i]], pb = probs[point_idxs[2 * i + 1]];
        float prob = pa > pb ? pa : pb;
        v3 x = v3_scale(ex, odist);
        v3 y = v3_cross(x, ab);
        int n = n_rots;
        if (adaptive) {
            int m = (int)((double)(odist / res) * (2 * ORACLE_PI));
            n = m < n_rots ? m : n_rots;
        }
        for (int r = 0; r < n; ++r) {
            float angle = (float)((double)(r * 2) * ORACLE_PI / (double)n);
            float ca = cosf(angle), sa = sinf(angle);
            v3 off = v3_add(v3_scale(x, ca), v3_scale(y, sa));
            v3 g = v3_div(v3_sub(v3_add(c, off), cr), res);
            if ((double)g.x < 0.01 || (double)g.y < 0.01 || (double)g.z < 0.01 ||
                (double)g.x >= gx - 1.01 || (double)g.y >= gy - 1.01 || (double)g.z >= gz - 1.01)
                continue;
            int fx = (int)g.x, fy = (int)g.y, fz = (int)g.z;
            float rx = g.x - floorf(g.x), ry = g.y - floorf(g.y), rz = g.z - floorf(g.z);
            float wx[2] = { 1.f - rx, rx }, wy[2] = { 1.f - ry, ry }, wz[2] = { 1.f - rz, rz };
            for (int dx = 0; dx < 2; ++dx)
                for (int dy = 0; dy < 2; ++dy)
                    for (int dz = 0; dz < 2; ++dz) {
                        float w = wx[dx] * wy[dy] * wz[dz];
                        grid_obj[(long)(fx + dx) * gy * gz + (long)(fy + dy) * gz + (fz + dz)] += (double)(w * prob);
                    }
        }
    }
}

/* voting.py:74-112  backvote: first candidate within tol of the winning centre and
 * inside [0, dim-1) writes -offset; otherwise the row is zero.  Always adaptive (:97). */
void oracle_backvote(const float* points, const float* outputs, float* out_offsets, const int32_t* point_idxs,
                     const float* corner, float res, long n_ppfs, int n_rots, int gx, int gy, int gz,
                     const float* gt_center, float tol) {
    v3 cr = { corner[0], corner[1], corner[2] };
    v3 gc = { gt_center[0], gt_center[1], gt_center[2] };
    for (long i = 0; i < n_ppfs; ++i) {
        float proj_len = outputs[2 * i], odist = outputs[2 * i + 1];
        v3 a, ab, ex;
        if (!pair_frame(points, point_idxs, i, &a, &ab, &ex)) continue;      /* :87 leaves the row untouched */
        v3 c = v3_sub(a, v3_scale(ab, proj_len));
        v3 x = v3_scale(ex, odist);
        v3 y = v3_cross(x, ab);
        out_offsets[3 * i] = out_offsets[3 * i + 1] = out_offsets[3 * i + 2] = 0.f;   /* :96 */
        int m = (int)((double)(odist / res) * (2 * ORACLE_PI));                       /* :97 */
        int n = m < n_rots ? m : n_rots;
        for (int r = 0; r < n; ++r) {
            float angle = (float)((double)(r * 2) * ORACLE_PI / (double)n);
            float ca = cosf(angle), sa = sinf(angle);
            v3 off = v3_add(v3_scale(x, ca), v3_scale(y, sa));
            v3 p = v3_add(c, off);                                                    /* :101 */
            if (v3_len(v3_sub(p, gc)) > tol) continue;                                /* :102 float compare */
            v3 g = v3_div(v3_sub(p, cr), res);                                        /* :103 */
            if (g.x < 0 || g.y < 0 || g.z < 0 || g.x >= gx - 1 || g.y >= gy - 1 || g.z >= gz - 1)
                continue;                                                             /* :104-107 int literals */
            out_offsets[3 * i] = -off.x; out_offsets[3 * i + 1] = -off.y; out_offsets[3 * i + 2] = -off.z;  /* :108 */
            break;
        }
    }
}

/* voting.py:119-147  rot_voting: n_rots candidate axis directions per pair. */
void oracle_rot_voting(const float* points, const float* preds_rot, float* outputs_up, const int32_t* point_idxs,
                       long n_ppfs, int n_rots) {
    for (long i = 0; i < n_ppfs; ++i) {
        float rot = preds_rot[i];
        v3 a, ab, ex;
        if (!pair_frame(points, point_idxs, i, &a, &ab, &ex)) continue;      /* :130 rows stay as given */
        v3 x = ex;                                                            /* :136 unit, not scaled */
        v3 y = v3_cross(x, ab);
        float t = tanf(rot);
        v3 axis = t > 0 ? ab : v3_scale(ab, -1.f);                            /* :142 */
        for (int r = 0; r < n_rots; ++r) {
            float angle = (float)((double)(r * 2) * ORACLE_PI / (double)n_rots);   /* :140 */
            float ca = cosf(angle), sa = sinf(angle);
            v3 off = v3_add(v3_scale(x, ca), v3_scale(y, sa));
            v3 up = v3_add(v3_scale(off, t), axis);
            up = v3_div(up, (float)((double)v3_len(up) + 1e-7));              /* :143 */
            float* o = outputs_up + 3 * (i * n_rots + r);
            o[0] = up.x; o[1] = up.y; o[2] = up.z;
        }
    }
}

/* voting.py:154-171  findpeak.  `literal` != 0 reproduces the string as shipped: the
 * comma operator at :165-166 drops the x term of the two y-neighbour reads.
 * literal == 0 is the evidently intended 6-neighbour second difference. */
void oracle_findpeak(const float* grids, float* outputs, int width, int gx, int gy, int gz, int literal) {
    long n = (long)gx * gy * gz;
    for (long idx = 0; idx < n; ++idx) {
        int x = (int)(idx / ((long)gy * gz));
        int yz = (int)(idx % ((long)gy * gz));
        int y = yz / gz, z = yz % gz;
        int xp = x + width < gx - 1 ? x + width : gx - 1, xm = x - width > 0 ? x - width : 0;
        int yp = y + width < gy - 1 ? y + width : gy - 1, ym = y - width > 0 ? y - width : 0;
        int zp = z + width < gz - 1 ? z + width : gz - 1, zm = z - width > 0 ? z - width : 0;
        long xoff = literal ? 0 : (long)x * gy * gz;
        float g = grids[idx];
        float dx = g - grids[(long)xp * gy * gz + y * gz + z] + g - grids[(long)xm * gy * gz + y * gz + z];
        float dy = g - grids[xoff + (long)yp * gz + z] + g - grids[xoff + (long)ym * gz + z];
        float dz = g - grids[(long)x * gy * gz + y * gz + zp] + g - grids[(long)x * gy * gz + y * gz + zm];
        outputs[idx] = dx + dy + dz;
    }
}

/* nocs/inference.py:208  np.argmax: first maximal flat index in C order. */
long oracle_grid_argmax(const float* grid, long n) {
    long best = 0;
    for (long i = 1; i < n; ++i)
        if (grid[i] > grid[best]) best = i;
    return best;
}

/* nocs/inference.py:282-283  counts[s] = #{candidates c : dot(c, sphere[s]) > thr};
 * fp32 dot accumulated in the order a row-times-column product takes (x, y, z). */
void oracle_sphere_count(const float* cand, long n_cand, const float* sphere, int n_bins, float thr, int64_t* counts) {
    memset(counts, 0, sizeof(int64_t) * n_bins);
    for (long c = 0; c < n_cand; ++c) {
        float cx = cand[3 * c], cy = cand[3 * c + 1], cz = cand[3 * c + 2];
        for (int s = 0; s < n_bins; ++s) {
            float d = cx * sphere[3 * s] + cy * sphere[3 * s + 1] + cz * sphere[3 * s + 2];
            if (d > thr) counts[s]++;
        }
    }
}

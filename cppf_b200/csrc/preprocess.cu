// Per-object pre-processing that feeds the hot path, on the device:
//   cppf_backproject   utils/util.py:598-631 + nocs/inference.py:131-137 (masked depth -> camera-frame points)
//   cppf_voxel_first   ME.utils.sparse_quantize(..., return_index=True)[1] (nocs/inference.py:140): one point per voxel
//   cppf_normals_pca   open3d estimate_normals(KDTreeSearchParamKNN(knn)) (utils/util.py:61-65): kNN covariance,
//                      eigenvector of the smallest eigenvalue
// MinkowskiEngine and open3d are third-party dependencies of the reference that are not in its tree; their
// published behaviour is restated (oracle/ref_preprocess.py states what is pinned and what is not).
// Coordinates stay float64 until after the voxel quantisation, exactly where the reference casts (:141).
#include "common.cuh"

#include "../../include/cppf_b200.h"

#include <math.h>

namespace cppf {
namespace prep {

struct Kinv {
    double m[9];
};

// flag[i] = mask[i] && depth[i] > 0   (utils/util.py:609-610)
template <typename D>
__global__ void __launch_bounds__(256) valid_kernel(const D* __restrict__ depth, const uint8_t* __restrict__ mask, long long n,
                                                    uint8_t* __restrict__ flag) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        flag[i] = (mask[i] != 0 && depth[i] > (D)0) ? 1 : 0;
}

// pts[k] = ((Kinv @ (u, v, 1)) * z / w) * scale for the k-th valid pixel in row-major order (np.where order, :612);
// the two axis flips of utils/util.py:630-631 and nocs/inference.py:136-137 cancel exactly (negation is exact).
template <typename D>
__global__ void __launch_bounds__(256) backproject_kernel(const D* __restrict__ depth, const long long* __restrict__ pos,
                                                          const long long* __restrict__ count, int width, Kinv k, double scale,
                                                          double* __restrict__ pts) {
    const long long m = *count;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
        const long long p = pos[j];
        const double v = (double)(p / width), u = (double)(p % width);
        const double x = k.m[0] * u + k.m[1] * v + k.m[2];                  // :623
        const double y = k.m[3] * u + k.m[4] * v + k.m[5];
        const double w = k.m[6] * u + k.m[7] * v + k.m[8];
        const double z = (double)depth[p];                                  // :626
        pts[3 * j] = (x * z / w) / scale;                                   // :629, nocs/inference.py:132
        pts[3 * j + 1] = (y * z / w) / scale;
        pts[3 * j + 2] = (w * z / w) / scale;
    }
}

// ---- voxel hash: key = packed (floor(x/res), floor(y/res), floor(z/res)), value = lowest point index --------------
constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ unsigned long long voxel_key(const double* __restrict__ p, double res) {
    const long long x = (long long)floor(p[0] / res), y = (long long)floor(p[1] / res), z = (long long)floor(p[2] / res);
    return ((unsigned long long)((x + (1 << 20)) & 0x1FFFFF) << 42) | ((unsigned long long)((y + (1 << 20)) & 0x1FFFFF) << 21) |
           (unsigned long long)((z + (1 << 20)) & 0x1FFFFF);
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}

__global__ void __launch_bounds__(256) voxel_insert_kernel(const double* __restrict__ pts, const long long* __restrict__ count,
                                                           long long n_max, double res, unsigned long long* __restrict__ keys,
                                                           int* __restrict__ vals, unsigned long long cap_mask) {
    const long long n = count ? (*count < n_max ? *count : n_max) : n_max;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long key = voxel_key(pts + 3 * i, res);
        unsigned long long slot = mix64(key) & cap_mask;
        while (true) {
            const unsigned long long prev = atomicCAS(keys + slot, kEmpty, key);
            if (prev == kEmpty || prev == key) {
                atomicMin(vals + slot, (int)i);
                break;
            }
            slot = (slot + 1) & cap_mask;
        }
    }
}

__global__ void __launch_bounds__(256) voxel_first_kernel(const double* __restrict__ pts, const long long* __restrict__ count,
                                                          long long n_max, double res, const unsigned long long* __restrict__ keys,
                                                          const int* __restrict__ vals, unsigned long long cap_mask,
                                                          uint8_t* __restrict__ flag) {
    const long long n = count ? (*count < n_max ? *count : n_max) : n_max;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_max; i += (long long)gridDim.x * blockDim.x) {
        uint8_t f = 0;
        if (i < n) {
            const unsigned long long key = voxel_key(pts + 3 * i, res);
            unsigned long long slot = mix64(key) & cap_mask;
            while (keys[slot] != key) slot = (slot + 1) & cap_mask;
            f = vals[slot] == (int)i ? 1 : 0;
        }
        flag[i] = f;
    }
}

// out[j] = float32(pts[pos[j]])   (nocs/inference.py:141)
__global__ void __launch_bounds__(256) gather_f32_kernel(const double* __restrict__ pts, const long long* __restrict__ pos,
                                                         const long long* __restrict__ count, long long cap,
                                                         float* __restrict__ out) {
    const long long m = *count < cap ? *count : cap;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
        const long long p = pos[j];
        out[3 * j] = (float)pts[3 * p];
        out[3 * j + 1] = (float)pts[3 * p + 1];
        out[3 * j + 2] = (float)pts[3 * p + 2];
    }
}

// ---- normals: covariance of the k nearest neighbours (double), eigenvector of the smallest eigenvalue -------------
// Closed-form symmetric 3x3 eigen-decomposition (trigonometric eigenvalues; eigenvector = the largest cross product
// of two rows of A - lambda I), the same family of method as open3d's FastEigen3x3.
__device__ __forceinline__ void smallest_eigvec(double a00, double a01, double a02, double a11, double a12, double a22,
                                                double* nx, double* ny, double* nz) {
    const double scale = fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fmax(fabs(a01), fmax(fabs(a02), fabs(a12)))));
    if (!(scale > 0.0)) {                       // all neighbours coincide: open3d falls back to (0, 0, 1)
        *nx = 0.0; *ny = 0.0; *nz = 1.0;
        return;
    }
    const double inv = 1.0 / scale;
    a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
    const double norm = a01 * a01 + a02 * a02 + a12 * a12;
    double lam;
    if (norm > 0.0) {
        const double q = (a00 + a11 + a22) / 3.0;
        const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * norm) / 6.0);
        const double c00 = b11 * b22 - a12 * a12, c01 = a01 * b22 - a12 * a02, c02 = a01 * a12 - b11 * a02;
        const double det = (b00 * c00 - a01 * c01 + a02 * c02) / (p * p * p);
        const double half = fmin(fmax(det * 0.5, -1.0), 1.0);
        const double angle = acos(half) / 3.0;
        lam = q + p * cos(angle + 2.0943951023931953) * 2.0;      // smallest eigenvalue (angle + 2 pi / 3)
    } else {
        lam = fmin(a00, fmin(a11, a22));
    }
    const double r0x = a00 - lam, r0y = a01, r0z = a02;
    const double r1x = a01, r1y = a11 - lam, r1z = a12;
    const double r2x = a02, r2y = a12, r2z = a22 - lam;
    const double c0x = r0y * r1z - r0z * r1y, c0y = r0z * r1x - r0x * r1z, c0z = r0x * r1y - r0y * r1x;
    const double c1x = r0y * r2z - r0z * r2y, c1y = r0z * r2x - r0x * r2z, c1z = r0x * r2y - r0y * r2x;
    const double c2x = r1y * r2z - r1z * r2y, c2y = r1z * r2x - r1x * r2z, c2z = r1x * r2y - r1y * r2x;
    const double d0 = c0x * c0x + c0y * c0y + c0z * c0z, d1 = c1x * c1x + c1y * c1y + c1z * c1z,
                 d2 = c2x * c2x + c2y * c2y + c2z * c2z;
    double vx = c0x, vy = c0y, vz = c0z, d = d0;
    if (d1 > d) { vx = c1x; vy = c1y; vz = c1z; d = d1; }
    if (d2 > d) { vx = c2x; vy = c2y; vz = c2z; d = d2; }
    if (!(d > 0.0)) {                           // rank <= 1 (collinear neighbours): any vector orthogonal to the line
        *nx = 0.0; *ny = 0.0; *nz = 1.0;
        return;
    }
    const double s = rsqrt(d);
    *nx = vx * s; *ny = vy * s; *nz = vz * s;
}

// orient: 0 = raw eigenvector sign (open3d leaves it unspecified), 1 = towards the camera at the origin (n . p <= 0)
__global__ void __launch_bounds__(128) normals_kernel(const float* __restrict__ pc, const long long* __restrict__ nbrs,
                                                      int n_points, int k, int orient, float* __restrict__ normals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const long long* nb = nbrs + (long long)i * k;
    double mx = 0, my = 0, mz = 0;
    for (int j = 0; j < k; ++j) {
        const f3 p = ld3(pc, nb[j]);
        mx += p.x; my += p.y; mz += p.z;
    }
    mx /= k; my /= k; mz /= k;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int j = 0; j < k; ++j) {
        const f3 p = ld3(pc, nb[j]);
        const double x = p.x - mx, y = p.y - my, z = p.z - mz;
        a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z;
    }
    double nx, ny, nz;
    smallest_eigvec(a00 / k, a01 / k, a02 / k, a11 / k, a12 / k, a22 / k, &nx, &ny, &nz);
    if (orient == 1) {
        const f3 p = ld3(pc, i);
        if (nx * p.x + ny * p.y + nz * p.z > 0.0) { nx = -nx; ny = -ny; nz = -nz; }
    }
    normals[3 * i] = (float)nx;
    normals[3 * i + 1] = (float)ny;
    normals[3 * i + 2] = (float)nz;
}

static int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace prep
}  // namespace cppf

using namespace cppf;

extern "C" int64_t cppf_backproject_scratch_bytes(int height, int width) {
    const int64_t n = (int64_t)height * width;
    return ((n + 255) & ~255ll) + cppf_compact_scratch_bytes(n) + 256;
}

extern "C" int cppf_backproject(const void* depth, int depth_is_u16, const uint8_t* mask, int height, int width,
                                const double* h_intrinsics_inv, double depth_scale, double* out_pts, int64_t* out_pix,
                                int64_t* out_count, void* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n = (long long)height * width;
    if (n <= 0 || n > 0x7FFFFFFFll || depth_scale == 0.0) return (int)cudaErrorInvalidValue;
    uint8_t* flag = reinterpret_cast<uint8_t*>(scratch);
    void* cscratch = flag + ((n + 255) & ~255ll);
    prep::Kinv k;
    for (int i = 0; i < 9; ++i) k.m[i] = h_intrinsics_inv[i];
    if (depth_is_u16) prep::valid_kernel<uint16_t><<<prep::grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(depth), mask, n, flag);
    else prep::valid_kernel<float><<<prep::grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<const float*>(depth), mask, n, flag);
    CPPF_LAUNCH_CHECK();
    const int r = cppf_compact_pairs(flag, nullptr, 0, 1, n, nullptr, out_pix, out_count, cscratch, stream);
    if (r != 0) return r;
    if (depth_is_u16)
        prep::backproject_kernel<uint16_t><<<prep::grid_for(n, 256), 256, 0, stream>>>(
            reinterpret_cast<const uint16_t*>(depth), reinterpret_cast<const long long*>(out_pix),
            reinterpret_cast<const long long*>(out_count), width, k, depth_scale, out_pts);
    else
        prep::backproject_kernel<float><<<prep::grid_for(n, 256), 256, 0, stream>>>(
            reinterpret_cast<const float*>(depth), reinterpret_cast<const long long*>(out_pix),
            reinterpret_cast<const long long*>(out_count), width, k, depth_scale, out_pts);
    CPPF_LAUNCH_CHECK();
    return 0;
}

static long long voxel_capacity(long long n) {
    long long cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

extern "C" int64_t cppf_voxel_scratch_bytes(int64_t n_max) {
    const long long cap = voxel_capacity(n_max);
    return cap * 12 + ((n_max + 255) & ~255ll) + cppf_compact_scratch_bytes(n_max) + 512;
}

extern "C" int cppf_voxel_first(const double* pts, const int64_t* count, int64_t n_max, double voxel, float* out_pc,
                                int64_t* out_index, int64_t* out_count, void* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_max <= 0 || n_max > 0x7FFFFFFFll || !(voxel > 0.0)) return (int)cudaErrorInvalidValue;
    const long long cap = voxel_capacity(n_max);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(scratch);
    int* vals = reinterpret_cast<int*>(keys + cap);
    uint8_t* flag = reinterpret_cast<uint8_t*>(vals + cap);
    void* cscratch = flag + ((n_max + 255) & ~255ll);
    CPPF_RETURN_IF(cudaMemsetAsync(keys, 0xFF, (size_t)cap * 8, stream));
    CPPF_RETURN_IF(cudaMemsetAsync(vals, 0x7F, (size_t)cap * 4, stream));
    const long long* cnt = reinterpret_cast<const long long*>(count);
    prep::voxel_insert_kernel<<<prep::grid_for(n_max, 256), 256, 0, stream>>>(pts, cnt, n_max, voxel, keys, vals,
                                                                             (unsigned long long)(cap - 1));
    CPPF_LAUNCH_CHECK();
    prep::voxel_first_kernel<<<prep::grid_for(n_max, 256), 256, 0, stream>>>(pts, cnt, n_max, voxel, keys, vals,
                                                                            (unsigned long long)(cap - 1), flag);
    CPPF_LAUNCH_CHECK();
    const int r = cppf_compact_pairs(flag, nullptr, 0, 1, n_max, nullptr, out_index, out_count, cscratch, stream);
    if (r != 0) return r;
    if (out_pc) {
        prep::gather_f32_kernel<<<prep::grid_for(n_max, 256), 256, 0, stream>>>(
            pts, reinterpret_cast<const long long*>(out_index), reinterpret_cast<const long long*>(out_count), n_max, out_pc);
        CPPF_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int cppf_normals_pca(const float* pc, int n_points, int k, int orient, int64_t* nbrs_scratch, float* normals,
                                void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_points <= 0) return 0;
    if (k <= 0 || k > n_points) return (int)cudaErrorInvalidValue;
    const int r = cppf_knn(pc, n_points, k, nbrs_scratch, stream);
    if (r != 0) return r;
    prep::normals_kernel<<<(n_points + 127) / 128, 128, 0, stream>>>(pc, reinterpret_cast<const long long*>(nbrs_scratch),
                                                                     n_points, k, orient, normals);
    CPPF_LAUNCH_CHECK();
    return 0;
}

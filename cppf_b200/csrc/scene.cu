// Scene-scale ("zero-shot") mode of the reference, nocs/zero_shot.ipynb: the parts
// that are not already covered by the per-object entry points.
//   cppf_pair_filter      cell 6: drop indistinguishable pairs (|n1.n2| > 0.9, |ab.n1| < 0.1, |ab.n2| < 0.1)
//   cppf_gaussian3d       cell 9: scipy.ndimage.gaussian_filter(grid, sigma) (separable, radius int(4 sigma + 0.5),
//                         mode 'reflect'), one pass per axis, float32 in / out with float64 accumulation like scipy
//   cppf_scene_proposals  cell 9: greedy multi-peak proposals (argmax, contrast against the 12 edges of a +-margin
//                         box, suppression of the box) -- the one place the reference needs a peak finder
// Cell numbers refer to the notebook's code cells; the notebook is JSON, so citations are by cell.
#include "common.cuh"

#include "../../include/cppf_b200.h"

#include <math.h>

#include <vector>

namespace cppf {

int grid_argmax_launch(const float* grid, int64_t n_cells, const int* n_cells_dev, int64_t* out_index, float* out_value,
                       cudaStream_t stream);

namespace scene {

// ---- cell 6 ---------------------------------------------------------------------------------------------------
template <bool IDX64>
__global__ void __launch_bounds__(256) pair_filter_kernel(const float* __restrict__ pc, const float* __restrict__ nrm,
                                                          const void* __restrict__ idx, long long n_pairs, int n_points,
                                                          uint8_t* __restrict__ keep) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (long long)gridDim.x * blockDim.x) {
        int a, b;
        pair_ab<IDX64>(idx, p, n_points, a, b);
        const f3 n1 = ld3(nrm, a), n2 = ld3(nrm, b);
        f3 ab = ld3(pc, a) - ld3(pc, b);
        const float inv = sqrtf(dot3(ab, ab)) + 1e-7f;
        ab = {ab.x / inv, ab.y / inv, ab.z / inv};
        const bool drop = fabsf(dot3(n1, n2)) > 0.9f && fabsf(dot3(ab, n1)) < 0.1f && fabsf(dot3(ab, n2)) < 0.1f;
        keep[p] = drop ? 0 : 1;
    }
}

// ---- gaussian filter, one axis -----------------------------------------------------------------------------------
constexpr int kMaxRadius = 32;
struct Taps {
    double w[2 * kMaxRadius + 1];
    int radius;
};

// scipy 'reflect' (half-sample symmetric): ... c b a | a b c ... c b a | a b c ...
__device__ __forceinline__ int reflect(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

__global__ void __launch_bounds__(256) gauss_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int gx, int gy,
                                                         int gz, int axis, Taps taps) {
    const long long n = (long long)gx * gy * gz;
    const int len = axis == 0 ? gx : (axis == 1 ? gy : gz);
    const long long stride = axis == 0 ? (long long)gy * gz : (axis == 1 ? gz : 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / stride) % len);
        const long long base = i - (long long)c * stride;
        double acc = 0.0;
        for (int t = -taps.radius; t <= taps.radius; ++t)
            acc += taps.w[t + taps.radius] * (double)__ldg(in + base + (long long)reflect(c + t, len) * stride);
        out[i] = (float)acc;
    }
}

// ---- proposals ---------------------------------------------------------------------------------------------------
// out[0:3] = loc, out[3] = value at loc, out[4] = value - mean of the 12 box-edge means (cell 9)
__global__ void proposal_contrast_kernel(const float* __restrict__ grid, const long long* __restrict__ flat, int gx, int gy,
                                         int gz, int margin, float* __restrict__ out) {
    const long long f = *flat;
    const long long gyz = (long long)gy * gz;
    const int loc[3] = {(int)(f / gyz), (int)((f % gyz) / gz), (int)(f % gz)};
    const int dims[3] = {gx, gy, gz};
    int l[3], r[3];
    for (int k = 0; k < 3; ++k) {
        l[k] = max(0, loc[k] - margin);
        r[k] = min(dims[k] - 1, loc[k] + margin);
    }
    auto at = [&](int x, int y, int z) { return grid[(long long)x * gyz + (long long)y * gz + z]; };
    float total = 0.f;
    // the 12 edges: for every axis, the 4 combinations of the other two axes' (l, r); slices exclude r like numpy
    for (int axis = 0; axis < 3; ++axis) {
        const int o1 = (axis + 1) % 3, o2 = (axis + 2) % 3;
        for (int c1 = 0; c1 < 2; ++c1)
            for (int c2 = 0; c2 < 2; ++c2) {
                int q[3];
                q[o1] = c1 ? r[o1] : l[o1];
                q[o2] = c2 ? r[o2] : l[o2];
                float s = 0.f;
                for (int t = l[axis]; t < r[axis]; ++t) {
                    q[axis] = t;
                    s += at(q[0], q[1], q[2]);
                }
                total += s / (float)(r[axis] - l[axis]);       // empty slice: nan, as numpy's mean of an empty slice
            }
    }
    const float v = at(loc[0], loc[1], loc[2]);
    out[0] = (float)loc[0]; out[1] = (float)loc[1]; out[2] = (float)loc[2];
    out[3] = v;
    out[4] = v - total / 12.f;
}

__global__ void __launch_bounds__(256) suppress_box_kernel(float* __restrict__ grid, int gy, int gz, int lx, int ly, int lz, int nx,
                                                           int ny, int nz) {
    const long long n = (long long)nx * ny * nz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i % nz), y = (int)((i / nz) % ny), x = (int)(i / ((long long)nz * ny));
        grid[((long long)(lx + x) * gy + (ly + y)) * gz + (lz + z)] = 0.f;
    }
}

static int blocks_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

}  // namespace scene
}  // namespace cppf

using namespace cppf;

extern "C" int cppf_pair_filter(const float* pc, const float* nrm, const void* idx, int idx_is_64, int n_points, int64_t n_pairs,
                                uint8_t* out_keep, void* stream) {
    if (n_pairs <= 0) return 0;
    if (idx == nullptr && n_pairs != (int64_t)n_points * n_points) return (int)cudaErrorInvalidValue;
    const int blocks = scene::blocks_for(n_pairs, 256);
    if (idx_is_64) scene::pair_filter_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(pc, nrm, idx, n_pairs, n_points, out_keep);
    else scene::pair_filter_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(pc, nrm, idx, n_pairs, n_points, out_keep);
    CPPF_LAUNCH_CHECK();
    return 0;
}

extern "C" int cppf_gaussian3d(const float* grid, float* out, float* tmp, int gx, int gy, int gz, double sigma, double truncate,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (gx <= 0 || gy <= 0 || gz <= 0 || !(sigma > 0.0)) return (int)cudaErrorInvalidValue;
    scene::Taps taps;
    taps.radius = (int)(truncate * sigma + 0.5);                                  // scipy: lw = int(truncate * sd + 0.5)
    if (taps.radius > scene::kMaxRadius) return (int)cudaErrorInvalidValue;
    double sum = 0.0;
    for (int t = -taps.radius; t <= taps.radius; ++t) {
        taps.w[t + taps.radius] = exp(-0.5 / (sigma * sigma) * (double)t * (double)t);   // scipy _gaussian_kernel1d
        sum += taps.w[t + taps.radius];
    }
    for (int t = 0; t <= 2 * taps.radius; ++t) taps.w[t] /= sum;
    const long long n = (long long)gx * gy * gz;
    const int blocks = scene::blocks_for(n, 256);
    // scipy filters axis 0, then 1, then 2, each pass rounding to the output dtype
    scene::gauss_axis_kernel<<<blocks, 256, 0, stream>>>(grid, out, gx, gy, gz, 0, taps);
    CPPF_LAUNCH_CHECK();
    scene::gauss_axis_kernel<<<blocks, 256, 0, stream>>>(out, tmp, gx, gy, gz, 1, taps);
    CPPF_LAUNCH_CHECK();
    scene::gauss_axis_kernel<<<blocks, 256, 0, stream>>>(tmp, out, gx, gy, gz, 2, taps);
    CPPF_LAUNCH_CHECK();
    return 0;
}

// Greedy proposals on `grid` (MODIFIED in place: accepted boxes are zeroed like the notebook's smoothed_grid).
// h_out [max_props][5] on the HOST: loc x, y, z, value, contrast.  Returns the number of proposals (>= 0) or a
// negative CUDA error code.  Synchronises the stream once per proposal (the loop is data dependent).
extern "C" int cppf_scene_proposals(float* grid, int gx, int gy, int gz, float thresh, int margin, float rel_stop,
                                    int max_props, float* h_out, void* scratch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (gx <= 0 || gy <= 0 || gz <= 0 || max_props <= 0) return -(int)cudaErrorInvalidValue;
    long long* flat = reinterpret_cast<long long*>(scratch);
    float* d_out = reinterpret_cast<float*>(flat + 2);
    const long long n = (long long)gx * gy * gz;
    int count = 0;
    float max_val = 0.f;
    bool have_max = false;
    while (count < max_props) {
        int r = grid_argmax_launch(grid, n, nullptr, reinterpret_cast<int64_t*>(flat), nullptr, stream);
        if (r != 0) return -r;
        scene::proposal_contrast_kernel<<<1, 1, 0, stream>>>(grid, flat, gx, gy, gz, margin, d_out);
        count_launch();
        float h[5];
        cudaError_t e = cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return -(int)e;
        const float diff = h[4];
        if (diff > thresh) {
            if (!have_max) {
                max_val = diff;
                have_max = true;
            }
            for (int k = 0; k < 5; ++k) h_out[count * 5 + k] = h[k];
            ++count;
        }
        if (!(diff >= thresh) || (have_max && diff < max_val * rel_stop)) break;     // cell 9: diff < thresh or diff < 0.7 max
        const int loc[3] = {(int)h[0], (int)h[1], (int)h[2]}, dims[3] = {gx, gy, gz};
        int l[3], w[3];
        for (int k = 0; k < 3; ++k) {
            l[k] = loc[k] - margin > 0 ? loc[k] - margin : 0;
            const int rr = loc[k] + margin < dims[k] - 1 ? loc[k] + margin : dims[k] - 1;
            w[k] = rr - l[k];
        }
        if (w[0] > 0 && w[1] > 0 && w[2] > 0) {
            scene::suppress_box_kernel<<<scene::blocks_for((long long)w[0] * w[1] * w[2], 256), 256, 0, stream>>>(
                grid, gy, gz, l[0], l[1], l[2], w[0], w[1], w[2]);
            count_launch();
        } else {
            break;      // a degenerate box cannot suppress its own peak: the notebook would loop forever here
        }
    }
    return count;
}

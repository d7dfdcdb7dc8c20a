// dtfill_k6_outlier.cuh -- K6: KITTI outlier filter (data_read.py:103-128)
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

// ------------------------------------------------------------------------------------------------------
// K6: KITTI outlier filter, data_read.py:103-128: sum and count over the 7 x 7 diamond
// (cv2.filter2D, default border BORDER_REFLECT_101), average = sum / (count + 1e-5) in float64 (the reference's
// valid_pixels array is float64), a point more than 1.0 m FARTHER than the local average is dropped.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// The filter's verdict for one pixel with value x at (y, x_) of a frame: true = dropped.  BORDER: the window crosses
// the frame edge (reflected indices); kept out of line, it is rare.
template <bool BORDER>
__device__ __forceinline__ bool k6_verdict(const float* __restrict__ frame, int H, int W, int y, int x_, float x)
{
    constexpr int R = 3;
    float sum = 0.f;                       // float32 like cv2.filter2D(sparse_lidar, -1, ...)
    int cnt = 0;                           // the float64 filter of valid_pixels (np.float) counts exactly
#pragma unroll
    for (int dy = -R; dy <= R; ++dy) {
        const int w = R - (dy < 0 ? -dy : dy);
        const float* row = frame + (long)(BORDER ? reflect101(y + dy, H) : y + dy) * W;
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
            if (dx < -w || dx > w) continue;
            const float v = __ldg(row + (BORDER ? reflect101(x_ + dx, W) : x_ + dx));
            sum += v;
            cnt += v > 0.1f ? 1 : 0;       // data_read.py:116
        }
    }
    const double aveg = (double)sum / ((double)cnt + 0.00001);         // data_read.py:123
    return ((double)x - aveg) > 1.0;                                   // :125
}
__device__ __noinline__ bool k6_verdict_border(const float* __restrict__ frame, int H, int W, int y, int x_, float x) {
    return k6_verdict<true>(frame, H, W, y, x_, x);
}
__device__ __forceinline__ bool k6_is_outlier(const float* __restrict__ frame, int H, int W, int y, int x_, float x) {
    if (y >= 3 && y < H - 3 && x_ >= 3 && x_ < W - 3) return k6_verdict<false>(frame, H, W, y, x_, x);
    return k6_verdict_border(frame, H, W, y, x_, x);
}

// A pixel holding +-0 comes out as it went in whatever its neighbourhood holds (x * (1 - outlier), :128), and a
// LiDAR frame is ~95 % zeros: a thread streams 4 pixels (128-bit load and store) and only evaluates the 25-tap
// diamond -- straight from global memory through L1 -- for its non-zero ones, one after the other (one copy of the
// window code; a warp loops as often as its busiest lane has non-zero pixels).  VEC: W % 4 == 0, 16-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(256) k6_outlier_removal(const float* __restrict__ in, int H, int W, long ngroups,
                                                           float* __restrict__ out)
{
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    if (VEC) {
        const int gpr = W >> 2;                                        // groups per row
        const long row = g / gpr;
        const int x0 = (int)(g - row * gpr) * 4;
        const long frame = row / H;
        const int y = (int)(row - frame * H);
        const float* fr = in + frame * H * W;
        float4 v = *reinterpret_cast<const float4*>(in + g * 4);
        uint32_t todo = ((__float_as_uint(v.x) << 1) != 0u) | (((__float_as_uint(v.y) << 1) != 0u) << 1) |
                        (((__float_as_uint(v.z) << 1) != 0u) << 2) | (((__float_as_uint(v.w) << 1) != 0u) << 3);
        uint32_t drop = 0;
#pragma unroll 1
        while (todo) {
            const int e = __ffs(todo) - 1;
            todo &= todo - 1;
            const float x = e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w;
            if (k6_is_outlier(fr, H, W, y, x0 + e, x)) drop |= 1u << e;
        }
        if (drop & 1u) v.x = 0.0f;
        if (drop & 2u) v.y = 0.0f;
        if (drop & 4u) v.z = 0.0f;
        if (drop & 8u) v.w = 0.0f;
        st_stream_v4(out + g * 4, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
    } else {
        const long row = g / W;
        const int x0 = (int)(g - row * W);
        const long frame = row / H;
        const int y = (int)(row - frame * H);
        const float x = in[g];
        out[g] = ((__float_as_uint(x) << 1) != 0u && k6_is_outlier(in + frame * H * W, H, W, y, x0, x)) ? 0.0f : x;
    }
}

}  // namespace dtfill

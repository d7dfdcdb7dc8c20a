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

__global__ void __launch_bounds__(256) k6_outlier_removal(const float* __restrict__ in, int H, int W,
                                                           float* __restrict__ out)
{
    constexpr int R = 3;
    __shared__ float sd[K5_TH + 2 * R][K5_TW + 2 * R + 1];
    const long fpx = (long)blockIdx.z * H * W;
    const int x0 = blockIdx.x * K5_TW, y0 = blockIdx.y * K5_TH;
    const int tw = K5_TW + 2 * R, th = K5_TH + 2 * R;
    for (int i = threadIdx.x; i < tw * th; i += 256) {
        const int ly = i / tw, lx = i - ly * tw;
        const int gy = reflect101(y0 + ly - R, H), gx = reflect101(x0 + lx - R, W);
        sd[ly][lx] = in[fpx + (long)gy * W + gx];
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    float sum = 0.f;                       // float32 like cv2.filter2D(sparse_lidar, -1, ...)
    int cnt = 0;                           // the float64 filter of valid_pixels (np.float) counts exactly
#pragma unroll
    for (int dy = -R; dy <= R; ++dy) {
        const int w = R - (dy < 0 ? -dy : dy);
#pragma unroll
        for (int dx = -R; dx <= R; ++dx) {
            if (dx < -w || dx > w) continue;
            const float v = sd[ty + R + dy][tx + R + dx];
            sum += v;
            cnt += v > 0.1f ? 1 : 0;       // data_read.py:116
        }
    }
    const float x = sd[ty + R][tx + R];
    const double aveg = (double)sum / ((double)cnt + 0.00001);         // data_read.py:123
    const bool outlier = ((double)x - aveg) > 1.0;                     // :125
    out[fpx + (long)gy * W + gx] = outlier ? 0.0f : x;                 // :128  x * (1 - outlier)
}

}  // namespace dtfill

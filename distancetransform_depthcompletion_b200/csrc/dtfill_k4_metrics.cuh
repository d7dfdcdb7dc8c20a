// dtfill_k4_metrics.cuh -- K4: evaluation.py:82-123, 196-239
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

// ------------------------------------------------------------------------------------------------------
// K4: evaluation metrics (evaluation.py:82-123 Result.evaluate, :196-239 Result_NYU.evaluate)
// Stage 1: per (frame, chunk) partial sums in double, fixed order.  Stage 2: per-frame metrics + column sums.
// acc: 0 sum d^2, 1 sum d, 2 sum dinv^2, 3 sum dinv, 4 count, 5 sum d/t, 6..8 delta counts
// ------------------------------------------------------------------------------------------------------
constexpr int ACC = 9;

template <typename GT, int MODE>
__device__ __forceinline__ void metric_accumulate(float o, GT t, double* a)
{
    const bool valid = (o > 0.01f) && (t > (GT)0.01);                  // evaluation.py:85-87 / :199-201
    if (!valid) return;
    if (MODE == 0) {
        const float o_mm = 1e3f * o;                                   // :89 float32 product
        const GT t_mm = (GT)1e3 * t;                                   // :90
        const GT d = o_mm > t_mm ? (GT)o_mm - t_mm : t_mm - (GT)o_mm;  // :92
        const GT d2 = d * d;                                           // :94 np.power(.,2)
        const float io = 1.0f / (1e-3f * o);                           // :116
        const GT it = (GT)1.0 / ((GT)1e-3 * t);                        // :117
        const GT di = (GT)io > it ? (GT)io - it : it - (GT)io;         // :118
        const GT di2 = di * di;
        a[0] += (double)d2; a[1] += (double)d; a[2] += (double)di2; a[3] += (double)di; a[4] += 1.0;
    } else {
        const GT og = (GT)o;
        const GT d = og > t ? og - t : t - og;                         // :206
        const GT d2 = d * d;                                           // :208
        const GT rel = d / t;                                          // :210
        const GT r1 = og / t, r2 = t / og;                             // :217
        const GT mr = r1 > r2 ? r1 : r2;
        const GT io = (GT)(1.0f / o), it = (GT)1.0 / t;                // :232-233 (output ** (-1) stays float32)
        const GT di = io > it ? io - it : it - io;
        const GT di2 = di * di;
        a[0] += (double)d2; a[1] += (double)d; a[2] += (double)di2; a[3] += (double)di; a[4] += 1.0;
        a[5] += (double)rel;
        a[6] += mr < (GT)1.25 ? 1.0 : 0.0;                             // :218
        a[7] += mr < (GT)1.5625 ? 1.0 : 0.0;                           // :219  1.25**2
        a[8] += mr < (GT)1.953125 ? 1.0 : 0.0;                         // :220  1.25**3
    }
}

template <typename GT, int MODE>
__global__ void __launch_bounds__(256) k4_metrics_partial(const float* __restrict__ pred, const GT* __restrict__ gt,
                                                           long npx, int chunks, double* __restrict__ partial)
{
    __shared__ double sm[8][ACC];
    const int b = blockIdx.y, ch = blockIdx.x;
    const long per = (npx + chunks - 1) / chunks;
    const long i0 = ch * per, i1 = min(npx, i0 + per);
    const float* p = pred + (long)b * npx;
    const GT* g = gt + (long)b * npx;
    double a[ACC];
#pragma unroll
    for (int k = 0; k < ACC; ++k) a[k] = 0.0;
    if (((npx | i0) & 3) == 0 && ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 31) == 0) {
        // 4 pixels per thread and iteration: one 128-bit load of the prediction, one or two of the ground truth
        for (long i = i0 + 4 * (long)threadIdx.x; i < i1; i += 4 * 256) {
            const float4 o = __ldg(reinterpret_cast<const float4*>(p + i));
            GT t[4];
            if (sizeof(GT) == 8) {
                const double2 t0 = __ldg(reinterpret_cast<const double2*>(g + i));
                const double2 t1 = __ldg(reinterpret_cast<const double2*>(g + i + 2));
                t[0] = (GT)t0.x; t[1] = (GT)t0.y; t[2] = (GT)t1.x; t[3] = (GT)t1.y;
            } else {
                const float4 tf = __ldg(reinterpret_cast<const float4*>(g + i));
                t[0] = (GT)tf.x; t[1] = (GT)tf.y; t[2] = (GT)tf.z; t[3] = (GT)tf.w;
            }
            metric_accumulate<GT, MODE>(o.x, t[0], a);
            metric_accumulate<GT, MODE>(o.y, t[1], a);
            metric_accumulate<GT, MODE>(o.z, t[2], a);
            metric_accumulate<GT, MODE>(o.w, t[3], a);
        }
    } else {
        for (long i = i0 + threadIdx.x; i < i1; i += 256) metric_accumulate<GT, MODE>(p[i], g[i], a);
    }
#pragma unroll
    for (int k = 0; k < ACC; ++k) {
        double v = a[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        a[k] = v;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < ACC; ++k) sm[wid][k] = a[k];
    __syncthreads();
    if (threadIdx.x < ACC) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        partial[((long)b * chunks + ch) * ACC + threadIdx.x] = v;
    }
}

// one block; thread t handles frames t, t+blockDim, ...; then a fixed-order column sum
__global__ void __launch_bounds__(256) k4_metrics_final(const double* __restrict__ partial, int B, int chunks, int mode,
                                                         double* __restrict__ per_frame /*[B][9]*/,
                                                         double* __restrict__ sums /*[10]*/, int accumulate)
{
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double a[ACC];
        for (int k = 0; k < ACC; ++k) a[k] = 0.0;
        for (int c = 0; c < chunks; ++c)
            for (int k = 0; k < ACC; ++k) a[k] += partial[((long)b * chunks + c) * ACC + k];
        const double n = a[4];
        double* o = per_frame + (long)b * 9;
        const double mse = a[0] / n;
        o[0] = mse;
        o[1] = sqrt(mse);
        o[2] = (mode == 0 ? a[1] : a[5]) / n;
        o[3] = sqrt(a[2] / n);
        o[4] = a[3] / n;
        o[5] = mode == 0 ? 0.0 : a[6] / n;
        o[6] = mode == 0 ? 0.0 : a[7] / n;
        o[7] = mode == 0 ? 0.0 : a[8] / n;
        o[8] = n;
    }
    __syncthreads();
    if (sums && threadIdx.x < 10) {
        double s = 0.0;
        if (threadIdx.x < 9)
            for (int b = 0; b < B; ++b) s += per_frame[(long)b * 9 + threadIdx.x];
        else
            s = (double)B;
        // accumulate: running totals of an evaluation loop (eval.py:212-232), one fixed-order addition per batch
        sums[threadIdx.x] = accumulate ? sums[threadIdx.x] + s : s;
    }
}

}  // namespace dtfill

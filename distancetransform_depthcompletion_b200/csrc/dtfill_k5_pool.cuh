// dtfill_k5_pool.cuh -- K5: DT pooling of the CNN input stage (net.py:71-123)
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

// ------------------------------------------------------------------------------------------------------
// K5: one level of the CNN input stage's "DT pooling" (net.py:83-123 generate_multi_channel).
// For every pixel: among the pixels of its T x T window (zero padded) whose mask is set, those with the largest
// weight T - |dy| - |dx| (net.py:71-81), i.e. the city-block-nearest ones, are averaged:
// out = sum(data[sel]) / (1e-6 + |sel|)  (net.py:93).  With no masked pixel in the window all T*T positions tie at
// weight 0 and the result is sum(window) / (1e-6 + T*T).  mask == nullptr means mask = data > 0.001 (net.py:95).
// One thread per pixel, 32 x 8 tile + halo in shared memory, rings of growing city-block distance.
// ------------------------------------------------------------------------------------------------------
constexpr int K5_TW = 32, K5_TH = 8, K5_MAXR = 7;      // table_size <= 15

__global__ void __launch_bounds__(256) k5_dt_pool_level(const float* __restrict__ data, const float* __restrict__ mask,
                                                         int H, int W, int T, float* __restrict__ out, uint8_t* __restrict__ out_mask)
{
    __shared__ float sd[K5_TH + 2 * K5_MAXR][K5_TW + 2 * K5_MAXR + 1];
    __shared__ unsigned long long smk[K5_TH + 2 * K5_MAXR];      // one mask bit per tile column (<= 46 columns)
    const int R = T / 2;
    const long fpx = (long)blockIdx.z * H * W;
    const int x0 = blockIdx.x * K5_TW, y0 = blockIdx.y * K5_TH;
    const int tw = K5_TW + 2 * R, th = K5_TH + 2 * R;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // tile + halo: one warp per tile row, two ballots build the row's mask word
    for (int ly = wid; ly < th; ly += 8) {
        const int gy = y0 + ly - R;
        unsigned long long word = 0;
        for (int l0 = 0; l0 < tw; l0 += 32) {
            const int lx = l0 + lane;
            const int gx = x0 + lx - R;
            float v = 0.f;
            bool m = false;
            if (lx < tw && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                v = data[fpx + (long)gy * W + gx];
                m = mask ? (mask[fpx + (long)gy * W + gx] != 0.f) : (v > 0.001f);   // mask * weight > 0 <=> mask != 0
            }
            if (lx < tw) sd[ly][lx] = v;
            word |= (unsigned long long)__ballot_sync(0xffffffffu, m) << l0;
        }
        if (lane == 0) smk[ly] = word;
    }
    __syncthreads();
    const int tx = lane, ty = wid;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    const int cx = tx + R, cy = ty + R;
    const uint32_t fieldmask = (1u << T) - 1u, lowmask = (1u << R) - 1u;
    int best = 1 << 20;                  // smallest city-block distance to a masked pixel of the window
    float sum = 0.f, cnt = 0.f;
    for (int dy = -R; dy <= R; ++dy) {
        // the T mask bits of this window row, centre at bit R
        const uint32_t f = (uint32_t)(smk[cy + dy] >> tx) & fieldmask;
        if (!f) continue;
        const int ady = dy < 0 ? -dy : dy;
        int dxr = 1 << 20, dxl = 1 << 20;
        if ((f >> R) & 1u) dxr = dxl = 0;
        else {
            const uint32_t right = f >> (R + 1), left = f & lowmask;
            if (right) dxr = __ffs(right);
            if (left) dxl = R - (31 - __clz(left));
        }
        const int dx = min(dxr, dxl), d = ady + dx;
        if (d > best) continue;
        if (d < best) { best = d; sum = 0.f; cnt = 0.f; }
        if (dx == 0) { sum += sd[cy + dy][cx]; cnt += 1.f; }
        else {
            if (dxl == dx) { sum += sd[cy + dy][cx - dx]; cnt += 1.f; }
            if (dxr == dx) { sum += sd[cy + dy][cx + dx]; cnt += 1.f; }
        }
    }
    if (cnt == 0.f) {                    // nothing masked: every window position ties at weight 0
        for (int dy = -R; dy <= R; ++dy)
            for (int dx = -R; dx <= R; ++dx) sum += sd[cy + dy][cx + dx];
        cnt = (float)(T * T);
    }
    const float res = sum / (0.000001f + cnt);
        out[fpx + (long)gy * W + gx] = res;
        if (out_mask) out_mask[fpx + (long)gy * W + gx] = res > 0.001f;          // net.py:95 mask of the next level
}

// ------------------------------------------------------------------------------------------------------
// K5t: the same level for window sizes 3, 5, 7, 9 (R = 1..4), restructured so that the work per pixel is a few
// dozen instructions: a 64 x 32 tile (+ halo) in shared memory; phase 1 computes, once per tile row and output
// column, the row's nearest masked offset, the sum of the data there (left and right when they tie) and its
// count; phase 2 lets a thread walk 8 output rows of one column with those row records in registers: the window
// minimum of |dy| + dx, then the records at that distance.  Windows without any masked pixel sum all T*T values
// (net.py:91-93: every weight ties at 0); a bit per row record says whether that sum can be anything but +0.
// ------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) k5_dt_pool_tile(const float* __restrict__ data, const float* __restrict__ mask,
                                                        int H, int W, float* __restrict__ out, uint8_t* __restrict__ out_mask)
{
    constexpr int T = 2 * R + 1, TW = 64, TH = 32, SW = TW + 2 * R, SH = TH + 2 * R, NONE = 15, SR = 8;
    constexpr uint32_t FM = (1u << T) - 1u, LOW = (1u << R) - 1u;
    __shared__ float sd[SH][SW + 1];             // data, zero outside the frame
    __shared__ uint32_t smk[SH][4];              // mask bits of a tile row (SW <= 96) + a spare word
    __shared__ uint32_t snz[SH][4];              // bit = the value is not +0.0f
    __shared__ float rs[SH][TW];                 // row record: sum of the nearest masked values of the row
    __shared__ uint8_t ri[SH][TW];               // row record: dx (0..R, NONE) | count << 4 | "row part not all +0" << 7
    const long fpx = (long)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int ly = wid; ly < SH; ly += 8) {
        const int gy = y0 + ly - R;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int lx = c * 32 + lane, gx = x0 + lx - R;
            float v = 0.f;
            bool m = false;
            if (lx < SW && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                v = data[fpx + (long)gy * W + gx];
                m = mask ? (mask[fpx + (long)gy * W + gx] != 0.f) : (v > 0.001f);      // mask * weight > 0 <=> mask != 0
            }
            if (lx < SW) sd[ly][lx] = v;
            const uint32_t bm = __ballot_sync(0xffffffffu, m);
            const uint32_t bn = __ballot_sync(0xffffffffu, __float_as_uint(v) != 0u);
            if (lane == 0) { smk[ly][c] = bm; snz[ly][c] = bn; }
        }
        if (lane == 0) { smk[ly][3] = 0u; snz[ly][3] = 0u; }
    }
    __syncthreads();
    // phase 1: row records
    for (int i = tid; i < SH * TW; i += 256) {
        const int ly = i >> 6, tx = i & 63, w = tx >> 5, sh = tx & 31;
        const uint32_t f = __funnelshift_r(smk[ly][w], smk[ly][w + 1], sh) & FM;        // window columns, centre = bit R
        const uint32_t nzf = __funnelshift_r(snz[ly][w], snz[ly][w + 1], sh) & FM;
        const uint32_t right = f >> (R + 1), left = f & LOW;
        const int dxr = right ? __ffs(right) : NONE;
        const int dxl = left ? R - (31 - __clz(left)) : NONE;
        const bool centre = (f >> R) & 1u;
        const int dx = centre ? 0 : min(dxl, dxr);
        const int a = dx == NONE ? 0 : dx;
        const bool tl = !centre && dxl == dx && dx != NONE, tr = !centre && dxr == dx && dx != NONE;
        const float vl = sd[ly][tx + R - a], vr = sd[ly][tx + R + a];
        rs[ly][tx] = centre ? vl : ((tl ? vl : 0.f) + (tr ? vr : 0.f));
        ri[ly][tx] = (uint8_t)(dx | ((centre ? 1 : (int)tl + (int)tr) << 4) | (nzf ? 0x80 : 0));
    }
    __syncthreads();
    // phase 2: a thread owns column tx of SR consecutive output rows
    const int tx = tid & 63, ty0 = (tid >> 6) * SR;
    const int gx = x0 + tx;
    if (gx >= W) return;
    uint32_t e[SR + 2 * R];
    float v[SR + 2 * R];
#pragma unroll
    for (int k = 0; k < SR + 2 * R; ++k) { e[k] = ri[ty0 + k][tx]; v[k] = rs[ty0 + k][tx]; }
#pragma unroll
    for (int j = 0; j < SR; ++j) {
        const int gy = y0 + ty0 + j;
        if (gy >= H) break;
        int best = 2 * NONE;
        uint32_t anynz = 0;
#pragma unroll
        for (int k = 0; k < T; ++k) {
            best = min(best, (k < R ? R - k : k - R) + (int)(e[j + k] & 15u));
            anynz |= e[j + k];
        }
        float sum = 0.f, cnt = 0.f;
        if (best < NONE) {
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const bool sel = (k < R ? R - k : k - R) + (int)(e[j + k] & 15u) == best;
                sum += sel ? v[j + k] : 0.f;
                cnt += sel ? (float)((e[j + k] >> 4) & 3u) : 0.f;
            }
        } else {                             // nothing masked: every window position ties at weight 0
            if (anynz & 0x80u)
                for (int dy = 0; dy < T; ++dy)
                    for (int dx = 0; dx < T; ++dx) sum += sd[ty0 + j + dy][tx + dx];
            cnt = (float)(T * T);
        }
        const float res = sum / (0.000001f + cnt);
        out[fpx + (long)gy * W + gx] = res;
        if (out_mask) out_mask[fpx + (long)gy * W + gx] = res > 0.001f;          // net.py:95 mask of the next level
    }
}

// ------------------------------------------------------------------------------------------------------
// K5w: the same level for window sizes 3, 5, 7 (R = 1..3), the sizes whose T x T mask window fits one 64-bit word
// (row k of the window in bits 8k .. 8k+T-1).  A thread owns one column of 16 output rows and slides the window
// word down (shift by one row, OR the new row's T-bit field in); the city-block rings around the centre are
// compile-time masks, so "nearest masked pixels" is a 3-step search over the cumulative ring masks plus one AND,
// their number a population count, and only the selected pixels (1-3 as a rule) are read and added, in row-major
// order.  Tiles that hold neither a masked pixel nor a value other than +0 are written as zeros straight away.
// ------------------------------------------------------------------------------------------------------
template <int R>
struct PoolRings {
    uint64_t ring[2 * R + 2], disk[2 * R + 2];
    constexpr PoolRings() : ring{}, disk{} {
        constexpr int T = 2 * R + 1;
        for (int d = 0; d <= 2 * R + 1; ++d) {
            uint64_t m = 0;
            for (int k = 0; k < T; ++k)
                for (int c = 0; c < T; ++c)
                    if ((k < R ? R - k : k - R) + (c < R ? R - c : c - R) == d) m |= 1ull << (8 * k + c);
            ring[d] = m;
            disk[d] = m | (d ? disk[d - 1] : 0ull);
        }
    }
};

template <int R, bool VEC>
__global__ void __launch_bounds__(256) k5_dt_pool_win(const float* __restrict__ data, const float* __restrict__ mask,
                                                       int H, int W, float* __restrict__ out, uint8_t* __restrict__ out_mask)
{
    // tile columns start at x0 - 4 (16-byte aligned for 128-bit loads); output column tx reads from column tx + OFF
    constexpr int T = 2 * R + 1, TW = 64, TH = 64, OFF = 4 - R, SW = 72, SH = TH + 2 * R, SR = TH / 4, PITCH = SW;
    constexpr uint32_t FM = (1u << T) - 1u;
    constexpr PoolRings<R> rings{};
    __shared__ __align__(16) float sd[SH][PITCH];   // data, zero outside the frame
    __shared__ uint32_t smk[SH][4];                  // mask bits of a tile row (SW <= 96) + a spare word
    __shared__ uint32_t snz[SH][4];                  // bit = the value is not +0.0f
    __shared__ uint32_t sany;
    const long fpx = (long)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) sany = 0u;
    uint32_t any = 0;
    if (VEC) {                                   // W % 4 == 0, 16-byte aligned pointers: one float4 per item
        for (int i = tid; i < SH * 4; i += 256) { (&smk[0][0])[i] = 0u; (&snz[0][0])[i] = 0u; }
        __syncthreads();
        for (int i = tid; i < SH * (SW / 4); i += 256) {
            const int ly = i / (SW / 4), q = i - ly * (SW / 4);
            const int gy = y0 + ly - R, gx = x0 - 4 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t mb = 0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const long o = fpx + (long)gy * W + gx;
                v = *reinterpret_cast<const float4*>(data + o);
                if (mask) {
                    const float4 m = *reinterpret_cast<const float4*>(mask + o);
                    mb = (m.x != 0.f) | ((m.y != 0.f) << 1) | ((m.z != 0.f) << 2) | ((m.w != 0.f) << 3);
                } else {
                    mb = (v.x > 0.001f) | ((v.y > 0.001f) << 1) | ((v.z > 0.001f) << 2) | ((v.w > 0.001f) << 3);
                }
            }
            *reinterpret_cast<float4*>(&sd[ly][4 * q]) = v;
            const uint32_t nb = (__float_as_uint(v.x) != 0u) | ((__float_as_uint(v.y) != 0u) << 1) |
                                ((__float_as_uint(v.z) != 0u) << 2) | ((__float_as_uint(v.w) != 0u) << 3);
            if (mb) atomicOr(&smk[ly][q >> 3], mb << ((q & 7) * 4));
            if (nb) atomicOr(&snz[ly][q >> 3], nb << ((q & 7) * 4));
            any |= mb | nb;
        }
        if (any) sany = 1u;                      // benign race: every writer stores the same value
    } else {
        __syncthreads();
        for (int ly = wid; ly < SH; ly += 8) {
            const int gy = y0 + ly - R;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int lx = c * 32 + lane, gx = x0 - 4 + lx;
                float v = 0.f;
                bool m = false;
                if (lx < SW && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                    v = data[fpx + (long)gy * W + gx];
                    m = mask ? (mask[fpx + (long)gy * W + gx] != 0.f) : (v > 0.001f);  // mask * weight > 0 <=> mask != 0
                }
                if (lx < SW) sd[ly][lx] = v;
                const uint32_t bm = __ballot_sync(0xffffffffu, m);
                const uint32_t bn = __ballot_sync(0xffffffffu, __float_as_uint(v) != 0u);
                if (lane == 0) { smk[ly][c] = bm; snz[ly][c] = bn; }
                any |= bm | bn;
            }
            if (lane == 0) { smk[ly][3] = 0u; snz[ly][3] = 0u; }
        }
        if (lane == 0 && any) atomicOr(&sany, 1u);
    }
    __syncthreads();
    const int tx = tid & 63, ty0 = (tid >> 6) * SR;
    const int gx = x0 + tx;
    if (gx >= W) return;
    float* op = out + fpx + (long)(y0 + ty0) * W + gx;
    if (!sany) {                                 // nothing masked, every value +0: the sums are +0
        for (int j = 0; j < SR && y0 + ty0 + j < H; ++j) {
            op[(long)j * W] = 0.f;
            if (out_mask) out_mask[(op - out) + (long)j * W] = 0;
        }
        return;
    }
    const int w = (tx + OFF) >> 5, sh = (tx + OFF) & 31;
    auto field = [&](const uint32_t (*bits)[4], int ly) { return __funnelshift_r(bits[ly][w], bits[ly][(w + 1) & 3], sh) & FM; };
    uint64_t win = 0;                            // rows ty0 .. ty0+T-2 in window rows 1 .. T-1: one shift completes it
#pragma unroll
    for (int k = 0; k < T - 1; ++k) win |= (uint64_t)field(smk, ty0 + k) << (8 * (k + 1));
#pragma unroll 1
    for (int j = 0; j < SR; ++j) {
        win = (win >> 8) | ((uint64_t)field(smk, ty0 + j + T - 1) << (8 * (T - 1)));
        if (y0 + ty0 + j >= H) break;
        const float* sdp = &sd[ty0 + j][tx + OFF];
        float sum = 0.f, cnt;
        if (win) {
            int best = 0;                        // smallest d whose disk meets the window
#pragma unroll
            for (int st = 4; st; st >>= 1)
                if (best + st <= 2 * R && !(win & rings.disk[best + st - 1])) best += st;
            const uint64_t sel = win & rings.ring[best];
            uint32_t a = (uint32_t)sel, b = (uint32_t)(sel >> 32);
            cnt = (float)(__popc(a) + __popc(b));
            while (a) {
                const int bit = __ffs(a) - 1;
                a &= a - 1;
                sum += sdp[(bit >> 3) * PITCH + (bit & 7)];
            }
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                sum += sdp[((bit >> 3) + 4) * PITCH + (bit & 7)];
            }
        } else {                                 // nothing masked: every window position ties at weight 0
            uint32_t nz = 0;
#pragma unroll
            for (int k = 0; k < T; ++k) nz |= field(snz, ty0 + j + k);
            if (nz)
                for (int dy = 0; dy < T; ++dy)
                    for (int dx = 0; dx < T; ++dx) sum += sdp[dy * PITCH + dx];
            cnt = (float)(T * T);
        }
        const float res = sum / (0.000001f + cnt);
        op[(long)j * W] = res;
        if (out_mask) out_mask[(op - out) + (long)j * W] = res > 0.001f;          // net.py:95 mask of the next level
    }
}

// ------------------------------------------------------------------------------------------------------
// The older variant of the pooling that demo.py carries (demo.py:65-149): no mask; the weight of a window position is
// 10 ** (T - |dy| - |dx|) (demo.py:65-76, float32) and the pixels selected are those whose VALUE TIMES WEIGHT equals
// the window's maximum (demo.py:120-121), so a far pixel ten times deeper than a near one wins; the denominator counts
// the selected pixels that are not zero (tf.math.count_nonzero, demo.py:122).  Zero padding takes part in the maximum
// like tf.image.extract_patches(padding='SAME').  One thread per pixel, 32 x 8 tile + halo in shared memory, two sweeps
// of the T x T window (maximum, then sum and count in row-major order).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k5_dt_pool_demo(const float* __restrict__ data, const float* __restrict__ weights /*[T*T]*/,
                                                        int H, int W, int T, float* __restrict__ out)
{
    __shared__ float sd[K5_TH + 2 * K5_MAXR][K5_TW + 2 * K5_MAXR + 1];
    __shared__ float sw[(2 * K5_MAXR + 1) * (2 * K5_MAXR + 1)];
    const int R = T / 2;
    const long fpx = (long)blockIdx.z * H * W;
    const int x0 = blockIdx.x * K5_TW, y0 = blockIdx.y * K5_TH;
    const int tw = K5_TW + 2 * R, th = K5_TH + 2 * R;
    for (int i = threadIdx.x; i < T * T; i += 256) sw[i] = weights[i];
    for (int i = threadIdx.x; i < tw * th; i += 256) {
        const int ly = i / tw, lx = i - ly * tw;
        const int gy = y0 + ly - R, gx = x0 + lx - R;
        sd[ly][lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? data[fpx + (long)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) return;
    float mx = -3.402823466e38f;
    for (int dy = 0; dy < T; ++dy)
        for (int dx = 0; dx < T; ++dx) mx = fmaxf(mx, __fmul_rn(sd[ty + dy][tx + dx], sw[dy * T + dx]));
    float sum = 0.0f;
    int cnt = 0;
    for (int dy = 0; dy < T; ++dy)
        for (int dx = 0; dx < T; ++dx) {
            const float v = sd[ty + dy][tx + dx];
            if (__fmul_rn(v, sw[dy * T + dx]) == mx) { sum = __fadd_rn(sum, v); cnt += v != 0.0f; }
        }
    out[fpx + (long)gy * W + gx] = __fdiv_rn(sum, __fadd_rn(0.000001f, (float)cnt));
}

}  // namespace dtfill

// dtfill_k7_edt.cuh -- K7: exact Euclidean feature transform (the algorithm BASELINE.json's north_star
// describes).  An EXTENSION: the reference never computes a Euclidean transform (every call site is the chamfer scan,
// tools.py:9), so there is no reference function to be exact against; the oracle is scipy.ndimage.distance_transform_edt.
//
// Separable (Meijster / Felzenszwalb): the squared distance to the nearest source is
//     D(y,x) = min over y' of (y - y')^2 + h(y',x)^2,   h(y',x) = distance to the nearest source of row y' from column x.
//   k7_edt_rows     one warp per row.  Coalesced loads of the row, source predicate of nearest_point (tools.py:8), one
//                   ballot per 32 pixels: lane w keeps word w of the row's source bits in a register.  The nearest source
//                   column on either side of every pixel follows from count-leading/trailing-zeros on that word plus a
//                   max / min scan of the words' last / first source across the lanes (shuffles).  Writes the nearest
//                   source COLUMN per pixel (u16), 64 contiguous bytes per lane.
//   k7_edt_columns  lower envelope of the parabolas (y - y')^2 + h(y',x)^2 down every column: ONE THREAD PER COLUMN, a
//                   warp's 32 scanlines being 32 adjacent columns, so every load of the map and every store of the two
//                   outputs is coalesced without staging.  Forward sweep: the envelope as a stack of (row, first row where
//                   it is the minimum) pairs; backward sweep: read the answers off the stack.  O(1) amortised per pixel
//                   and exact in integers (the separator is a floor division).  The stack of a scanline can be as deep as
//                   the frame is tall and 64 scanlines per SM sub-partition are in flight, so it lives in a scratch array
//                   [frame][depth][column] (a warp's pushes at equal depth are one 128-byte line; the hot tops stay in
//                   L1/L2) rather than in shared memory; its top is held in registers.
// The round-1 order of the passes (columns first, then per row an outward search from x' = x staged in shared memory with
// cp.async.bulk) was O(distance) per pixel: 5.9 ms per 256 KITTI frames, most of it in the source-free top third of the
// frames where every search ran ~2 x 100 steps; this order needs 0.5 ms.
// Ties: among several sources at the minimal distance the first in raster order wins (upper row; in a row the left one).
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

constexpr uint16_t EDT_NONE = 0xFFFFu;          // no source in this row
constexpr int EDT_MAX_WORDS = 1024;             // row pass: W <= 32768

// ------------------------------------------------------------------------------------------------------
// Row pass.  nearest_col[y][x] = column of the source of row y nearest to x (the left one at equal distance), EDT_NONE
// if the row holds no source.  V8: W % 8 == 0, so a lane's 32 values start 16-byte aligned (128-bit stores).
// ------------------------------------------------------------------------------------------------------
template <bool V8>
__global__ void __launch_bounds__(128) k7_edt_rows(const float* __restrict__ in, long nrows, int W, float src_cut,
                                                    uint16_t* __restrict__ nearest_col)
{
    __shared__ int next_first[4][EDT_MAX_WORDS / 32 + 1];      // per warp: first source column in the chunks after this one
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long row = (long)blockIdx.x * 4 + wib;
    if (row >= nrows) return;
    const float* rp = in + row * W;
    uint16_t* out = nearest_col + row * W;
    const int WW = (W + 31) >> 5;
    const int nchunks = (WW + 31) >> 5;
    constexpr int BIG = 1 << 30;

    // source bits of the row: chunk c holds words 32c .. 32c+31, word 32c + l in lane l
    auto chunk_bits = [&](int c) -> uint32_t {
        uint32_t mine = 0;
        const int w0 = c * 32;
        const int nw = min(32, WW - w0);
        for (int k0 = 0; k0 < nw; k0 += 8) {                  // 8 loads in flight per lane
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int x = (w0 + k0 + k) * 32 + lane;
                v[k] = (k0 + k < nw && x < W) ? ld_stream(rp + x) : -1.0f;     // padding is never a source (src_cut > -1)
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int x = (w0 + k0 + k) * 32 + lane;
                const bool src = (k0 + k < nw && x < W) && !(v[k] < src_cut);   // tools.py:8; NaN is a source
                const uint32_t m = __ballot_sync(0xffffffffu, src);
                if (lane == k0 + k) mine = m;
            }
        }
        return mine;
    };

    // right to left over the chunks: first source column after every chunk
    uint32_t bits0 = 0, bits1 = 0;                            // words of chunks 0 and 1 (W <= 2048), kept for the second sweep
    {
        int nf = BIG;
        for (int c = nchunks - 1; c >= 0; --c) {
            if (lane == 0) next_first[wib][c] = nf;
            const uint32_t word = chunk_bits(c);
            if (c == 0) bits0 = word;
            if (c == 1) bits1 = word;
            const uint32_t any = __ballot_sync(0xffffffffu, word != 0);
            if (any) {
                const int fl = __ffs(any) - 1;                // first lane with a source
                const uint32_t fw = __shfl_sync(0xffffffffu, word, fl);
                nf = (c * 32 + fl) * 32 + __ffs(fw) - 1;
            }
        }
        __syncwarp();
    }
    // left to right
    int last_before = -1;                                     // last source column in the chunks before this one
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t word = c == 0 ? bits0 : (c == 1 ? bits1 : chunk_bits(c));
        const int w = c * 32 + lane;
        const int base = w * 32;
        const int lastw = word ? base + 31 - __clz(word) : -1;
        const int firstw = word ? base + __ffs(word) - 1 : BIG;
        int incl_last = lastw, incl_first = firstw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, incl_last, d);
            const int b2 = __shfl_down_sync(0xffffffffu, incl_first, d);
            if (lane >= d) incl_last = max(incl_last, a);
            if (lane + d < 32) incl_first = min(incl_first, b2);
        }
        int exl = __shfl_up_sync(0xffffffffu, incl_last, 1);
        int exr = __shfl_down_sync(0xffffffffu, incl_first, 1);
        exl = lane == 0 ? last_before : max(exl, last_before);
        const int nf = next_first[wib][c];
        exr = lane == 31 ? nf : min(exr, nf);
        last_before = max(last_before, __shfl_sync(0xffffffffu, incl_last, 31));
        if (w < WW) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int x = base + i;
                const uint32_t le = word & (0xFFFFFFFFu >> (31 - i));
                const uint32_t gt = i == 31 ? 0u : (word & (0xFFFFFFFFu << (i + 1)));
                const int L = le ? base + 31 - __clz(le) : exl;
                const int R = gt ? base + __ffs(gt) - 1 : exr;
                uint32_t best = EDT_NONE;
                if (L >= 0 && (R >= BIG || x - L <= R - x)) best = (uint32_t)L;      // tie: the left one
                else if (R < BIG) best = (uint32_t)R;
                if (i & 1) pk[i >> 1] |= best << 16; else pk[i >> 1] = best;
            }
            if (V8 && base + 32 <= W) {
                uint4* o4 = reinterpret_cast<uint4*>(out + base);
#pragma unroll
                for (int j = 0; j < 4; ++j) o4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (base + i < W) out[base + i] = (uint16_t)((i & 1) ? (pk[i >> 1] >> 16) : (pk[i >> 1] & 0xFFFFu));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Column pass: lower envelope of the parabolas F(y; i) = (y - i)^2 + f(i), f(i) = (x - nearest_col[i][x])^2, over the rows
// i of column x that hold a source at all.  Stack entry q: row s[q] and the first row t[q] from which s[q] is the minimum
// (s | t << 16); the separator of two rows i < u is the largest y with F(y; i) <= F(y; u), so the upper row keeps ties.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int edt_floordiv(int num, int den /* > 0 */)
{
    int q = num / den;
    if ((num % den != 0) && (num < 0)) --q;
    return q;
}

__global__ void __launch_bounds__(128) k7_edt_columns(const uint16_t* __restrict__ nearest_col, int H, int W,
                                                       uint32_t* __restrict__ stack, int32_t* __restrict__ out_d2,
                                                       int32_t* __restrict__ out_idx)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const long fo = (long)blockIdx.y * H * W + x;
    const uint16_t* m = nearest_col + fo;
    uint32_t* st = stack + fo;
    int32_t* od2 = out_d2 + fo;
    int32_t* oix = out_idx ? out_idx + fo : nullptr;

    int q = -1;
    int s_top = 0, t_top = 0, f_top = 0, xs_top = 0;
    auto reload_top = [&]() {                       // entry q of the stack -> registers
        const uint32_t e = st[(long)q * W];
        s_top = (int)(e & 0xFFFFu);
        t_top = (int)(e >> 16);
        xs_top = (int)m[(long)s_top * W];
        const int dx = x - xs_top;
        f_top = dx * dx;
    };
    // forward sweep, one row ahead on the map
    uint32_t nxt = m[0];
    for (int u = 0; u < H; ++u) {
        const uint32_t xs = nxt;
        if (u + 1 < H) nxt = m[(long)(u + 1) * W];
        if (xs == EDT_NONE) continue;
        const int dx = x - (int)xs;
        const int fu = dx * dx;
        while (q >= 0) {
            const int a = t_top - s_top, b2 = t_top - u;
            if (a * a + f_top > b2 * b2 + fu) {     // u is lower already where s_top starts: s_top is never the minimum
                --q;
                if (q >= 0) reload_top();
            } else {
                break;
            }
        }
        if (q < 0) {
            q = 0; s_top = u; t_top = 0; f_top = fu; xs_top = (int)xs;
            st[0] = (uint32_t)u;
        } else {
            const int wsep = 1 + edt_floordiv(u * u - s_top * s_top + fu - f_top, 2 * (u - s_top));
            if (wsep < H) {
                ++q; s_top = u; t_top = wsep; f_top = fu; xs_top = (int)xs;
                st[(long)q * W] = (uint32_t)u | ((uint32_t)wsep << 16);
            }
        }
    }
    if (q < 0) {                                    // no source in the frame's column span at all = none in the frame
        for (int u = 0; u < H; ++u) {
            od2[(long)u * W] = 0x7FFFFFFF;
            if (oix) oix[(long)u * W] = -1;
        }
        return;
    }
    // the forward sweep may have left a top that was never pushed (wsep >= H): registers still describe entry q
    // only if the last push was entry q; reload to be sure
    reload_top();
    for (int u = H - 1; u >= 0; --u) {
        const int dy = u - s_top;
        od2[(long)u * W] = dy * dy + f_top;
        if (oix) oix[(long)u * W] = s_top * W + xs_top;
        if (u == t_top && q > 0) { --q; reload_top(); }
    }
}

}  // namespace dtfill

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
// frames where every search ran ~2 x 100 steps; this order needs 0.74 ms (row pass 0.17, column pass 0.57).
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
    constexpr int NEG = -(1 << 29);                           // "no source on the left": any x - NEG is larger than W
    uint32_t bits0 = 0, bits1 = 0;                            // words of chunks 0 and 1 (W <= 2048), kept for the second sweep
    int nf = BIG;
    for (int c = nchunks - 1; c >= 0; --c) {
        if (lane == 0) next_first[wib][c] = nf;
        const uint32_t word = chunk_bits(c);
        if (c == 0) bits0 = word;
        if (c == 1) bits1 = word;
        const uint32_t any = __ballot_sync(0xffffffffu, word != 0);
        if (any) {
            const int fl = __ffs(any) - 1;                    // first lane with a source
            const uint32_t fw = __shfl_sync(0xffffffffu, word, fl);
            nf = (c * 32 + fl) * 32 + __ffs(fw) - 1;
        }
    }
    __syncwarp();
    if (nf == BIG) {                                          // a row without sources (a third of a KITTI frame)
        if (V8) {
            uint4* o4 = reinterpret_cast<uint4*>(out);
            for (int i = lane; i < W / 8; i += 32) o4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        } else {
            for (int i = lane; i < W; i += 32) out[i] = EDT_NONE;
        }
        return;
    }
    // left to right.  Every pixel of a row that holds a source has one on its left or on its right, so distances to
    // "none" only need to lose the comparison.
    int last_before = NEG;                                    // last source column in the chunks before this one
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t word = c == 0 ? bits0 : (c == 1 ? bits1 : chunk_bits(c));
        const int w = c * 32 + lane;
        const int base = w * 32;
        int incl_last = word ? base + 31 - __clz(word) : NEG;
        int incl_first = word ? base + __ffs(word) - 1 : BIG;
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
        const int nfc = next_first[wib][c];
        exr = lane == 31 ? nfc : min(exr, nfc);
        last_before = max(last_before, __shfl_sync(0xffffffffu, incl_last, 31));
        if (w < WW) {
            // distances instead of positions: two 32-step recurrences of a bit test and an increment each
            int dr[32];
            int d = exr - (base + 32);                        // from column base + 32 to the first source after this word
#pragma unroll
            for (int i = 31; i >= 0; --i) { d = d + 1; dr[i] = d; if (word & (1u << i)) d = 0; }
            int dl = base - 1 - exl;                          // from column base - 1 back to the last source before this word
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                dl = (word & (1u << i)) ? 0 : dl + 1;
                const int x = base + i;
                const uint32_t best = (uint32_t)(dl <= dr[i] ? x - dl : x + dr[i]);      // tie: the left one
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
// i of column x that hold a source at all.  Stack entry q: {row s[q] | first row t[q] from which s[q] is the minimum << 16,
// nearest source column of row s[q]}; the separator of two rows i < u is the largest y with F(y; i) <= F(y; u), so the
// upper row keeps ties.  A thread runs NC scanlines (column x of NC frames) interleaved: their loads are in flight
// together, and a batch of 256 KITTI frames (311 296 scanlines) fits the 148 SMs in one wave of resident threads --
// every scanline lasts as long as the kernel, so a second, nearly empty wave would double it.
// ------------------------------------------------------------------------------------------------------
__device__ __noinline__ int edt_floordiv_wide(int num, int den /* > 0 */)
{
    int q = num / den;
    if ((num % den != 0) && (num < 0)) --q;
    return q;
}
__device__ __forceinline__ int edt_floordiv(int num, int den /* > 0 */)
{
    if (num > -(1 << 24) && num < (1 << 24)) {               // both operands exact in float32: quotient off by one at most
        int q = (int)floorf(__fdividef((float)num, (float)den));
        const int r = num - q * den;
        if (r < 0) --q; else if (r >= den) ++q;
        return q;
    }
    return edt_floordiv_wide(num, den);                       // frames wider than 4096 columns: out of line
}

template <int NC>
__global__ void __launch_bounds__(128, NC == 2 ? 10 : 16) k7_edt_columns(const uint16_t* __restrict__ nearest_col, int B, int H,
                                                                        int W, uint2* __restrict__ stack,
                                                                        int32_t* __restrict__ out_d2, int32_t* __restrict__ out_idx)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    constexpr int PF = 4;                           // rows of the map fetched ahead per scanline
    long fo[NC];
    bool live[NC];
    // the top of the stack decoded in registers, the entry below it (q - 1, valid while q >= 1) as it is stored: a pop
    // needs no load of its own, the load it issues (entry q - 2) is only consumed by the pop after it
    int q[NC], s_top[NC], t_top[NC], f_top[NC], xs_top[NC];
    uint2 below[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int b = blockIdx.y + k * gridDim.y;
        live[k] = b < B;
        fo[k] = (long)(live[k] ? b : 0) * H * W + x;
        q[k] = -1; s_top[k] = t_top[k] = f_top[k] = xs_top[k] = 0;
        below[k] = make_uint2(0u, 0u);
    }
    // forward sweep (a single copy of the step with the map values rotating through a register ring was measured
    // 20 % slower than these PF unrolled copies)
    for (int u0 = 0; u0 < H; u0 += PF) {
        uint32_t v[NC][PF];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const uint16_t* mp = nearest_col + fo[k] + (long)u0 * W;
#pragma unroll
            for (int r = 0; r < PF; ++r) {
                v[k][r] = (live[k] && u0 + r < H) ? (uint32_t)*mp : (uint32_t)EDT_NONE;
                mp += W;
            }
        }
#pragma unroll
        for (int r = 0; r < PF; ++r) {
            const int u = u0 + r;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const uint32_t xs = v[k][r];
                if (xs == EDT_NONE) continue;
                const int dx = x - (int)xs;
                const int fu = dx * dx;
                uint2* sk = stack + fo[k];
                while (q[k] >= 0) {
                    const int a = t_top[k] - s_top[k], b2 = t_top[k] - u;
                    if (a * a + f_top[k] <= b2 * b2 + fu) break;
                    // u is lower already where s_top starts: s_top is never the minimum
                    --q[k];
                    if (q[k] >= 0) {
                        const uint2 e = below[k];
                        s_top[k] = (int)(e.x & 0xFFFFu);
                        t_top[k] = (int)(e.x >> 16);
                        xs_top[k] = (int)e.y;
                        const int d2x = x - xs_top[k];
                        f_top[k] = d2x * d2x;
                        if (q[k] >= 1) below[k] = sk[(long)(q[k] - 1) * W];
                    }
                }
                if (q[k] < 0) {
                    q[k] = 0; s_top[k] = u; t_top[k] = 0; f_top[k] = fu; xs_top[k] = (int)xs;
                    sk[0] = make_uint2((uint32_t)u, xs);
                } else {
                    const int wsep = 1 + edt_floordiv(u * u - s_top[k] * s_top[k] + fu - f_top[k], 2 * (u - s_top[k]));
                    if (wsep < H) {
                        below[k] = make_uint2((uint32_t)s_top[k] | ((uint32_t)t_top[k] << 16), (uint32_t)xs_top[k]);
                        ++q[k]; s_top[k] = u; t_top[k] = wsep; f_top[k] = fu; xs_top[k] = (int)xs;
                        sk[(long)q[k] * W] = make_uint2((uint32_t)u | ((uint32_t)wsep << 16), xs);
                    }
                }
            }
        }
    }
    // backward sweep, all lanes on the same row (coalesced stores; per-lane loops over a stack entry's rows let the lanes
    // drift apart and ran 3x slower): entry q answers rows t[q] .. t[q + 1] - 1; the entry below the current one is
    // already in registers when the sweep reaches t[q], the one below that is fetched then
    bool none[NC];
    int idx_top[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        none[k] = q[k] < 0;
        idx_top[k] = none[k] ? -1 : s_top[k] * W + xs_top[k];
        if (none[k]) { s_top[k] = 0; f_top[k] = 0; t_top[k] = -1; }
    }
    for (int u = H - 1; u >= 0; --u) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (!live[k]) continue;
            const long o = fo[k] + (long)u * W;
            const int dy = u - s_top[k];
            out_d2[o] = none[k] ? 0x7FFFFFFF : dy * dy + f_top[k];
            if (out_idx) out_idx[o] = idx_top[k];
            if (u == t_top[k] && q[k] > 0) {
                --q[k];
                const uint2 e = below[k];
                if (q[k] >= 1) below[k] = stack[fo[k] + (long)(q[k] - 1) * W];
                s_top[k] = (int)(e.x & 0xFFFFu);
                t_top[k] = (int)(e.x >> 16);
                xs_top[k] = (int)e.y;
                const int dx = x - xs_top[k];
                f_top[k] = dx * dx;
                idx_top[k] = s_top[k] * W + xs_top[k];
            }
        }
    }
}

}  // namespace dtfill

// dtfill_k7_edt.cuh -- K7: exact Euclidean feature transform (the algorithm BASELINE.json's north_star
// describes).  An EXTENSION: the reference never computes a Euclidean transform (every call site is the chamfer scan,
// tools.py:9), so there is no reference function to be exact against; the oracle is scipy.ndimage.distance_transform_edt.
//
// Separable (Meijster / Felzenszwalb): the squared distance to the nearest source is
//     D(y,x) = min over x' of (x - x')^2 + g(y,x')^2,   g(y,x') = distance to the nearest source in column x'.
//   k7_edt_bits     source predicate of nearest_point (tools.py:8) -> bit rows, one ballot per 32 pixels
//   k7_edt_columns  column pass: one warp per strip of 32 columns; 32 rows of the strip are loaded as 32 words (one per
//                   lane), transposed with five shuffle/mask steps so that lane l holds the 32 rows of column l, and
//                   the nearest source row above / below every pixel follows from count-leading/trailing-zeros on that
//                   word plus a carry between chunks.  Writes the nearest source ROW per pixel (u16).
//   k7_edt_rows     row pass: a block stages ROWS_PER_BLOCK rows of that map (contiguous in memory) into shared memory
//                   with ONE cp.async.bulk (TMA) transaction signalled on an mbarrier, then every thread finds its pixel's
//                   minimum over the parabolas (x - x')^2 + g^2 by walking outwards from x' = x until the horizontal
//                   offset alone exceeds the best value found: exact, and O(distance) per pixel instead of O(W).
// Ties: the column pass prefers the source above at equal vertical distance; the row pass prefers the smaller |x - x'|,
// then the left neighbour.  The squared distance does not depend on them.
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

constexpr uint16_t EDT_NONE = 0xFFFFu;          // no source in this column
constexpr int EDT_ROWS_PER_BLOCK = 4;
constexpr int EDT_MAX_CHUNKS = 128;             // column pass: H <= 4096

__global__ void __launch_bounds__(256) k7_edt_bits(const float* __restrict__ in, long npx_total, int W, int WW, float src_cut,
                                                    uint32_t* __restrict__ bits)
{
    // one warp per 32-pixel word of a row
    const long word = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long nrows = npx_total / W;
    if (word >= nrows * WW) return;
    const long row = word / WW;
    const int w = (int)(word - row * WW);
    const int x = w * 32 + lane;
    bool src = false;
    if (x < W) {
        const float v = ld_stream(in + row * W + x);
        src = !(v < src_cut);                    // tools.py:8: !(float32(1 - x) > thr); NaN is a source
    }
    const uint32_t m = __ballot_sync(0xffffffffu, src);
    if (lane == 0) bits[word] = m;
}

// 32 x 32 bit transpose across a warp: on entry lane r holds row r (bit c = column c), on exit lane c holds column c
// (bit r = row r).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t v, int lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t o = __shfl_xor_sync(0xffffffffu, v, s);
        // lanes with bit s clear keep their low half-blocks and take the partner's low half-blocks shifted up;
        // lanes with bit s set keep their high half-blocks and take the partner's high half-blocks shifted down
        v = (lane & s) ? ((v & ~m) | ((o >> s) & m)) : ((v & m) | ((o << s) & ~m));
    }
    return v;
}

__global__ void __launch_bounds__(32) k7_edt_columns(const uint32_t* __restrict__ bits, int H, int W, int WW,
                                                      uint16_t* __restrict__ nearest_row)
{
    __shared__ uint16_t below[EDT_MAX_CHUNKS][32];      // first source row in the chunks after this one, per column
    const int w = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
    const uint32_t* bf = bits + (long)b * H * WW + w;
    uint16_t* out = nearest_row + (long)b * H * W;
    const int nchunks = (H + 31) >> 5;
    const int x = w * 32 + lane;
    // pass A, bottom to top: carry of the nearest source row below every chunk
    int dn = -1;
    for (int c = nchunks - 1; c >= 0; --c) {
        below[c][lane] = dn < 0 ? EDT_NONE : (uint16_t)dn;
        const int y = c * 32 + lane;
        const uint32_t col = warp_transpose32(y < H ? bf[(long)y * WW] : 0u, lane);
        if (col) dn = c * 32 + __ffs(col) - 1;
    }
    // pass B, top to bottom
    int up = -1;
    for (int c = 0; c < nchunks; ++c) {
        const int y0 = c * 32;
        const int yl = y0 + lane;
        const uint32_t col = warp_transpose32(yl < H ? bf[(long)yl * WW] : 0u, lane);
        const int dnc = below[c][lane] == EDT_NONE ? -1 : (int)below[c][lane];
        if (x < W) {
            for (int r = 0; r < 32 && y0 + r < H; ++r) {
                const int y = y0 + r;
                const uint32_t le = col & (0xFFFFFFFFu >> (31 - r));          // sources at rows <= y of this chunk
                const uint32_t gt = r == 31 ? 0u : (col & (0xFFFFFFFFu << (r + 1)));
                const int u = le ? y0 + 31 - __clz(le) : up;
                const int d = gt ? y0 + __ffs(gt) - 1 : dnc;
                int best = EDT_NONE;
                if (u >= 0 && (d < 0 || y - u <= d - y)) best = u;            // tie: the source above
                else if (d >= 0) best = d;
                out[(long)y * W + x] = (uint16_t)best;
            }
        }
        if (col) up = y0 + 31 - __clz(col);
    }
}

__device__ __forceinline__ uint32_t edt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA == true: rows staged with cp.async.bulk + mbarrier (needs 16-byte aligned rows: (ROWS * W * 2) % 16 == 0 for every
// block start, i.e. W % 8 == 0, and a 16-byte aligned map); otherwise plain loads.
template <bool TMA>
__global__ void __launch_bounds__(256) k7_edt_rows(const uint16_t* __restrict__ nearest_row, long nrows_total, int H, int W,
                                                    int32_t* __restrict__ out_d2, int32_t* __restrict__ out_idx)
{
    extern __shared__ __align__(16) uint16_t srow[];             // [EDT_ROWS_PER_BLOCK][W]
    __shared__ __align__(8) uint64_t mbar;
    const long row0 = (long)blockIdx.x * EDT_ROWS_PER_BLOCK;
    const int nr = (int)min((long)EDT_ROWS_PER_BLOCK, nrows_total - row0);
    const uint32_t bytes = (uint32_t)nr * W * 2;
    if (TMA) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(edt_smem_u32(&mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(edt_smem_u32(&mbar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(edt_smem_u32(srow)), "l"(nearest_row + row0 * W), "r"(bytes), "r"(edt_smem_u32(&mbar)) : "memory");
        }
        // every thread waits for the transaction (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(edt_smem_u32(&mbar)) : "memory");
        }
    } else {
        for (int i = threadIdx.x; i < nr * W; i += blockDim.x) srow[i] = nearest_row[row0 * W + i];
        __syncthreads();
    }
    for (int r = 0; r < nr; ++r) {
        const long row = row0 + r;
        const int y = (int)(row % H);
        const uint16_t* g = srow + r * W;
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            // parabola of the own column first, then outwards; stop when the horizontal offset alone reaches the best
            long best = 0x7FFFFFFFl;
            int bx = -1;
            {
                const uint16_t ry = g[x];
                if (ry != EDT_NONE) { const long dy = y - (int)ry; best = dy * dy; bx = x; }
            }
            const int dmax = max(x, W - 1 - x);
            for (int d = 1; d <= dmax; ++d) {
                const long dd = (long)d * d;
                if (dd >= best) break;
                if (x - d >= 0) {
                    const uint16_t ry = g[x - d];
                    if (ry != EDT_NONE) { const long dy = y - (int)ry; const long c = dd + dy * dy; if (c < best) { best = c; bx = x - d; } }
                }
                if (x + d < W) {
                    const uint16_t ry = g[x + d];
                    if (ry != EDT_NONE) { const long dy = y - (int)ry; const long c = dd + dy * dy; if (c < best) { best = c; bx = x + d; } }
                }
            }
            out_d2[row * W + x] = (int32_t)best;
            if (out_idx) out_idx[row * W + x] = bx < 0 ? -1 : (int32_t)g[bx] * W + bx;
        }
    }
}

}  // namespace dtfill

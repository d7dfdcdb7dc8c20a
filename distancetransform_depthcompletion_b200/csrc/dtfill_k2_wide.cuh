// dtfill_k2_wide.cuh -- K2w: 64-bit-key fallback of the scan
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

// ------------------------------------------------------------------------------------------------------
// K2w: wide fallback.  One warp per frame, 64-bit keys dist:29|order:3|label:32, three row buffers in shared
// memory.  Forward state: distance plane in ws.scratch (u32 per pixel), label plane parked in out_depth
// (same size, overwritten row by row with the final depth during the backward pass).
// ------------------------------------------------------------------------------------------------------
constexpr int WDSH = 35, WOSH = 32;
__host__ __device__ constexpr uint64_t WKC(int cost, int order) {
    return (uint64_t(cost) << WDSH) | (uint64_t(order) << WOSH);
}
constexpr uint64_t WORDCLR = ~(7ull << WOSH);
constexpr uint64_t WLMASK = 0xFFFFFFFFull;
constexpr uint32_t WINIT = 1u << 27;

__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

__global__ void __launch_bounds__(32) k2_chamfer_wide(FrameParams fp, Workspace ws, float* __restrict__ out_depth,
                                                       float* __restrict__ out_dt, int32_t* __restrict__ out_lbl)
{
    extern __shared__ __align__(16) uint64_t wsm[];
    const Task task = ws.tasks[blockIdx.x];
    if (task.kind != TASK_WIDE) return;
    const int lane = threadIdx.x;
    const int H = fp.H, W = fp.W, WW = fp.WW;
    const int b = task.frame;
    const long fpx = (long)b * H * W;
    const int RW = W + 4;                              // row buffer with 2 INIT columns on each side
    uint64_t* buf[3] = {wsm, wsm + RW, wsm + 2 * RW};
    const uint64_t init_key = (uint64_t)WINIT << WDSH;
    for (int i = lane; i < 3 * RW; i += 32) wsm[i] = init_key;
    __syncwarp();
    const int chunk = (W + 31) / 32;
    const int xa = min(W, lane * chunk), xb = min(W, xa + chunk);
    const uint32_t* bits_f = ws.srcbits + (long)b * H * WW;
    const uint16_t* pre_f = ws.wprefix + (long)b * H * WW;
    const uint32_t* rowbase = ws.rowsrc + (long)b * H;
    uint32_t* fdist = ws.scratch + fpx;                            // forward distance plane
    uint32_t* flab = reinterpret_cast<uint32_t*>(out_depth) + fpx;  // forward label plane (temporary)
    const float* dl = ws.dlist + fpx;

    // cross-lane carry on (dist,label) pairs; DIR>0: from lower lanes, DIR<0: from higher lanes
    auto carry = [&](uint32_t ed, uint32_t el, bool has, int DIR, uint32_t& cd, uint32_t& cl) {
        // lanes with an empty chunk contribute "infinite"
        uint32_t d_ = has ? ed : 0x7FFFFFFFu, l_ = el;
        // positions: distance between chunk ends of lane a and lane b is |xend_b - xend_a|; use explicit positions
        int pos = DIR > 0 ? xb - 1 : xa;
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t od = DIR > 0 ? __shfl_up_sync(0xffffffffu, d_, s) : __shfl_down_sync(0xffffffffu, d_, s);
            const uint32_t ol = DIR > 0 ? __shfl_up_sync(0xffffffffu, l_, s) : __shfl_down_sync(0xffffffffu, l_, s);
            const int op = DIR > 0 ? __shfl_up_sync(0xffffffffu, pos, s) : __shfl_down_sync(0xffffffffu, pos, s);
            const bool ok = DIR > 0 ? (lane >= s) : (lane + s < 32);
            if (ok && od < 0x40000000u) {
                const uint32_t t = od + (uint32_t)abs(pos - op);
                if (t < d_) { d_ = t; l_ = ol; }
            }
        }
        // value entering this lane = inclusive value of the neighbouring lane, measured at that lane's end
        const uint32_t nd = DIR > 0 ? __shfl_up_sync(0xffffffffu, d_, 1) : __shfl_down_sync(0xffffffffu, d_, 1);
        const uint32_t nl = DIR > 0 ? __shfl_up_sync(0xffffffffu, l_, 1) : __shfl_down_sync(0xffffffffu, l_, 1);
        const int np = DIR > 0 ? __shfl_up_sync(0xffffffffu, pos, 1) : __shfl_down_sync(0xffffffffu, pos, 1);
        const bool edge = DIR > 0 ? lane == 0 : lane == 31;
        if (edge || nd >= 0x40000000u) { cd = 0x7FFFFFFFu; cl = 0; }
        else { cd = nd; cl = nl; (void)np; }
    };

    // ---------------- forward ----------------
    for (int y = 0; y < H; ++y) {
        uint64_t* A = buf[(y + 2) % 3];   // row y-1
        uint64_t* Bq = buf[(y + 1) % 3];  // row y-2
        uint64_t* C = buf[y % 3];         // row y (overwrites row y-3)
        const uint32_t* br = bits_f + (long)y * WW;
        const uint16_t* pr = pre_f + (long)y * WW;
        const uint32_t rb = rowbase[y];
        uint64_t u = init_key;
        for (int x = xa; x < xb; ++x) {
            const int q = x + 2;
            uint64_t m = Bq[q - 1] + WKC(3, 0);
            m = umin64(m, Bq[q + 1] + WKC(3, 1));
            m = umin64(m, A[q - 2] + WKC(3, 2));
            m = umin64(m, A[q - 1] + WKC(2, 3));
            m = umin64(m, A[q] + WKC(1, 4));
            m = umin64(m, A[q + 1] + WKC(2, 5));
            m = umin64(m, A[q + 2] + WKC(3, 6));
            const uint32_t word = br[x >> 5];
            if ((word >> (x & 31)) & 1u)
                m = (uint64_t)(rb + pr[x >> 5] + __popc(word & ((1u << (x & 31)) - 1u)) + 1u);
            u = (x == xa) ? (m & WORDCLR) : (umin64(m, u + WKC(1, 7)) & WORDCLR);
            C[q] = u;
        }
        uint32_t cd, cl;
        carry((uint32_t)(u >> WDSH), (uint32_t)(u & WLMASK), xb > xa, +1, cd, cl);
        const int endprev = xa - 1;                    // column of the carried value
        for (int x = xa; x < xb; ++x) {
            uint64_t t = C[x + 2];
            if (cd < 0x40000000u) {
                const uint64_t k = ((uint64_t)(cd + (uint32_t)(x - endprev)) << WDSH) | (1ull << WOSH) | cl;
                t = umin64(t, k) & WORDCLR;
            }
            C[x + 2] = t;
            const uint32_t d = (uint32_t)(t >> WDSH);
            fdist[(long)y * W + x] = d;
            flab[(long)y * W + x] = (uint32_t)(t & WLMASK);
        }
        __syncwarp();
    }
    // ---------------- backward ----------------
    for (int i = lane; i < 3 * RW; i += 32) wsm[i] = init_key;
    __syncwarp();
    for (int y = H - 1, it = 0; y >= 0; --y, ++it) {
        uint64_t* A = buf[(it + 2) % 3];   // row y+1
        uint64_t* Bq = buf[(it + 1) % 3];  // row y+2
        uint64_t* C = buf[it % 3];
        uint64_t u = init_key;
        for (int x = xb - 1; x >= xa; --x) {
            const int q = x + 2;
            uint64_t m = ((uint64_t)fdist[(long)y * W + x] << WDSH) | flab[(long)y * W + x];
            m = umin64(m, Bq[q + 1] + WKC(3, 1));
            m = umin64(m, Bq[q - 1] + WKC(3, 2));
            m = umin64(m, A[q + 2] + WKC(3, 3));
            m = umin64(m, A[q + 1] + WKC(2, 4));
            m = umin64(m, A[q] + WKC(1, 5));
            m = umin64(m, A[q - 1] + WKC(2, 6));
            m = umin64(m, A[q - 2] + WKC(3, 7));
            m &= WORDCLR;
            u = (x == xb - 1) ? m : (umin64(m, u + WKC(1, 1)) & WORDCLR);
            C[q] = u;
        }
        uint32_t cd, cl;
        carry((uint32_t)(u >> WDSH), (uint32_t)(u & WLMASK), xb > xa, -1, cd, cl);
        const int endnext = xb;
        for (int x = xa; x < xb; ++x) {
            uint64_t t = C[x + 2];
            if (cd < 0x40000000u) {
                const uint64_t k = ((uint64_t)(cd + (uint32_t)(endnext - x)) << WDSH) | (1ull << WOSH) | cl;
                t = umin64(t, k) & WORDCLR;
            }
            C[x + 2] = t;
            const uint32_t d = (uint32_t)(t >> WDSH);
            const uint32_t l = (uint32_t)(t & WLMASK);
            const long o = fpx + (long)y * W + x;
            out_depth[o] = dl[l - 1u];
            if (out_dt) out_dt[o] = (float)d;
            if (out_lbl) out_lbl[o] = (int32_t)l;
        }
        __syncwarp();
    }
}

}  // namespace dtfill

// dtfill_k3_sky.cuh -- K3: rows above the first source row in closed form
#pragma once
#include "dtfill_common.cuh"

namespace dtfill {

// ------------------------------------------------------------------------------------------------------
// K3: the rows above the first source row f ("sky": the upper third of a KITTI frame).  Nothing lies above or
// beside them, so OpenCV's forward scan leaves them unreached, and in the backward scan every candidate of a pixel
// (y,x) with y + 2 <= f has a distance of the form (f - y') + g(x'), g = dt of row f, which is 1-Lipschitz.  Hence
//     dt(y,x)  = g(x) + (f - y)
//     lbl(y,x) = lbl(y+2, x+1)  if g(x+1) == g(x) - 1      the first candidate in OpenCV's order that attains
//              = lbl(y+2, x-1)  elif g(x-1) == g(x) - 1     the minimum (strict '>' update); the candidates
//              = lbl(y+1, x)    otherwise                   (+1,+2), (+1,+1) tie only if (+2,+1) already did
// The route is monotone: t(x) diagonal steps down g towards a valley column v(x), then straight down.  With the
// two base rows S, S+1 (S + 1 <= f, final keys stored by K2) and j = ceil((S - y) / 2) diagonal steps above them:
//     t(x) <  j :  lbl(y,x) = lbl(S, v(x))
//     t(x) >= j :  lbl(y,x) = lbl(y + 2j, x + s(x) j),   y + 2j in {S, S+1},  s(x) = +-1 the direction of descent
// i.e. one table lookup per pixel, no scan.  tests/test_kernel_model.py checks the rule against the oracle.
// One block per 32 rows of a frame; the per-column tables are rebuilt by every block (W entries).
// ------------------------------------------------------------------------------------------------------
constexpr int SKY_ROWS = 32;

__global__ void __launch_bounds__(256) k3_sky(FrameParams fp, Workspace ws, float* __restrict__ out_depth,
                                               float* __restrict__ out_dt, int32_t* __restrict__ out_lbl)
{
    __shared__ uint16_t d0[SKY_MAX_W + 2];       // dt of row S, one guard entry on each side
    __shared__ uint16_t steps[SKY_MAX_W];        // t(x)
    __shared__ int8_t dir[SKY_MAX_W];            // s(x)
    __shared__ float dep[2][SKY_MAX_W];          // depth_list[lbl - 1] of the two base rows
    const int b = blockIdx.x, S = ws.sky[b];
    const int y0 = blockIdx.y * SKY_ROWS;
    if (y0 >= S) return;
    DTFILL_TRACE_SCOPE(fp, 3);
    const int H = fp.H, W = fp.W, tid = threadIdx.x;
    const long fpx = (long)b * H * W;
    const uint32_t* sk = ws.skykeys + (long)b * 2 * W;
    const float* dl = ws.dlist + fpx;
    for (int x = tid; x < W; x += 256) {
        const uint32_t k0 = sk[x], k1 = sk[W + x];
        d0[x + 1] = (uint16_t)(k0 >> DSH);
        dep[0][x] = dl[max(k0 & LMASK, 1u) - 1u];
        dep[1][x] = dl[max(k1 & LMASK, 1u) - 1u];
    }
    if (tid == 0) { d0[0] = 0xFFFFu; d0[W + 1] = 0xFFFFu; }
    __syncthreads();
    for (int x = tid; x < W; x += 256) {
        const int g = d0[x + 1];
        dir[x] = (int)d0[x + 2] == g - 1 ? 1 : ((int)d0[x] == g - 1 ? -1 : 0);
    }
    __syncthreads();
    for (int x = tid; x < W; x += 256) {
        const int sd = dir[x];
        int k = 0;
        for (int xx = x; dir[xx] != 0; xx += sd) ++k;
        steps[x] = (uint16_t)k;
    }
    __syncthreads();
    const int yend = min(y0 + SKY_ROWS, S);
    if ((W & 3) == 0) {
        // a thread keeps the tables of 4 consecutive columns in registers and walks 8 rows: per pixel a compare, two
        // selects, one shared-memory load and one subtraction; 128-bit streaming stores
        const int ngroups = W >> 2;
        const float* depflat = &dep[0][0];
        for (int u = tid; u < ngroups * (SKY_ROWS / 8); u += 256) {
            const int part = u / ngroups, x = (u - part * ngroups) * 4;
            int tt[4], sd[4], ts[4];
            float gf[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                tt[e] = steps[x + e];
                sd[e] = dir[x + e];
                ts[e] = x + e + tt[e] * sd[e];                    // valley column, in base row S
                gf[e] = (float)((int)d0[x + e + 1] + S);
            }
            const int ya = y0 + part * 8, yb = min(ya + 8, yend);
            for (int y = ya; y < yb; ++y) {
                const int j = (S - y + 1) >> 1, oddoff = ((S - y) & 1) * SKY_MAX_W;
                const float fy = (float)y;
                const long ro = fpx + (long)y * W + x;
                int idx[4];
                uint32_t od[4], ot[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    idx[e] = tt[e] < j ? ts[e] : x + e + sd[e] * j + oddoff;
                    od[e] = __float_as_uint(depflat[idx[e]]);
                    ot[e] = __float_as_uint(gf[e] - fy);          // integers below 2^24: exact
                }
                st_stream_v4(out_depth + ro, od[0], od[1], od[2], od[3]);
                if (out_dt) st_stream_v4(out_dt + ro, ot[0], ot[1], ot[2], ot[3]);
                if (out_lbl) {
                    uint32_t ol[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        ol[e] = sk[idx[e] >= SKY_MAX_W ? W + idx[e] - SKY_MAX_W : idx[e]] & LMASK;
                    st_stream_v4(out_lbl + ro, ol[0], ol[1], ol[2], ol[3]);
                }
            }
        }
        return;
    }
    for (int y = y0; y < yend; ++y) {
        const int j = (S - y + 1) >> 1, odd = (S - y) & 1;
        const long ro = fpx + (long)y * W;
        for (int x = tid; x < W; x += 256) {
            const int t = steps[x], sd = dir[x];
            const bool valley = t < j;
            const int col = x + sd * (valley ? t : j);
            const int r = valley ? 0 : odd;
            st_stream_u32(out_depth + ro + x, __float_as_uint(dep[r][col]));
            if (out_dt) st_stream_u32(out_dt + ro + x, __float_as_uint((float)((int)d0[x + 1] + S - y)));
            if (out_lbl) st_stream_u32(out_lbl + ro + x, sk[r * W + col] & LMASK);
        }
    }
}

}  // namespace dtfill

// dtfill_kernels.cuh -- sm_100a kernels of the DT + nearest-neighbour fill path.
//
// Pipeline for a batch of frames [B,H,W] float32 (all device resident):
//   K1  k1_mask_rows      in -> source/valid bit rows, per-word source prefix, row counts, validity mask (u8)
//                         (reference: value_mask tools.py:8 / eval_NYU.py:115, with_value tools.py:22, net.py:131)
//   K1b k1b_scan_compact  per frame: exclusive row bases (= raster ranks, OpenCV's label initialisation),
//                         depth_list = in[valid] in raster order (tools.py:24), task list for K2
//   K2  k2_chamfer<PPL>   one warp per task: OpenCV's forward/backward 5x5 chamfer scan with label
//                         propagation (the cv2 call at tools.py:9) restructured as row-sequential,
//                         lane-parallel (min,+) scans on packed keys kept in registers, fused with the gather
//                         depth_list[lbl-1] (tools.py:25-27) and the dt / lbl outputs
//   K2w k2_chamfer_wide   same scan with 64-bit keys and rows in shared memory, for frames the packed 32-bit
//                         key cannot hold (width > 1216, 2H+W too large, or >= 2^18 sources)
//   K4  k4_metrics_*      masked RMSE/MAE/iRMSE/iMAE(/REL/delta) reductions of evaluation.py:82-123, 196-239
//
// Key format of the fast path: dist:11 | order:4 | label:17.  "order" is the position
// of a candidate in OpenCV's comparison sequence, so that OpenCV's "first candidate that is strictly smaller
// wins" is a plain unsigned minimum (one VIADDMNMX per candidate); label is the 1-based raster rank of the
// source, which is also what cv2 returns with DIST_LABEL_PIXEL.  Stencil candidates use even order values:
// a stored key may then keep the 0/1 order bit left by the carry application (which is not cleared) without
// changing the outcome of any later comparison.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dtfill {

constexpr int DSH = 21;                       // dist field shift
constexpr int OSH = 17;                       // order field shift
constexpr uint32_t LMASK = (1u << OSH) - 1u;  // label field
constexpr uint32_t ORDCLR = ~(15u << OSH);
constexpr uint32_t MAX_FAST_LABEL = LMASK;    // frames with more sources take the wide path
constexpr float UNREACHED_DT = 65533.0f;      // what OpenCV reports where no source is reachable

__host__ __device__ constexpr uint32_t KC(int cost, int order) {
    return (uint32_t(cost) << DSH) | (uint32_t(order) << OSH);
}

// TASK_CHAMFER: full-width tile, the kernel instance with the frame's PPL.  TASK_NARROW: half-width tile, the
// instance with the narrow PPL.  TASK_WIDE: 64-bit-key fallback.  TASK_NOSRC: frame without sources.
enum TaskKind : int { TASK_CHAMFER = 0, TASK_NOSRC = 1, TASK_WIDE = 2, TASK_SKIP = 3, TASK_NARROW = 4 };

// A tile of one frame: the sub-image rows [lo,hi) x columns [clo, clo + 32*PPL) is scanned as if it were the whole
// image; results are written for rows [r0,r1) x columns [c0,c1) only.  Exact because every written pixel's
// city-block ball of radius dt lies inside the sub-image (halo >= a guaranteed bound of dt).
struct __align__(16) Task {
    int frame;
    int lo, hi;        // sub-image rows (band + halo)
    int r0, r1;        // rows whose results are written (lo <= r0 < r1 <= hi)
    int kind;
    int scratch_off;   // start of this task's forward-state scratch, in units of 32 keys
    int fstart;        // first row >= lo holding a source: the forward pass starts here (rows above stay "unreached")
    int clo;           // first column of the sub-image
    int c0, c1;        // columns whose results are written
    int sky;           // S > 0: r0 == S and rows [0,S) of the frame are filled by k3_sky from the final keys of rows
                       // S, S+1, which this task stores into ws.skykeys; -2 otherwise
};

constexpr int MAXT = 32;      // task slots per frame; slot-major layout tasks[slot * B + frame]
constexpr int CELL_H = 4;     // coarse occupancy cells used by the band planner
constexpr int CELL_W = 8;
constexpr int SKY_MAX_W = 1216; // widest frame of the 32-bit-key path (32 lanes x 38 pixels): size of k3_sky's tables
constexpr int MAX_CELLS = 15360;   // planner grid limit (30 KB of shared memory); larger frames are not banded

struct FrameParams {
    int B, H, W, WW;           // WW = 32-bit words per bit row
    int in_H, in_crop;         // input frames hold in_H rows; rows [in_crop, in_crop + H) are the frame (uint16 input)
    float src_thr, val_thr;
    float src_cut;             // smallest float x (in the total order) with !(float32(1 - x) > src_thr)
    int init_dist;             // "unreached" distance of the fast path: H + W + 8
    int force_wide;            // size not representable in the 32-bit key
    int band_cap;              // planner: target cost (row steps) of one task; <= 0 disables banding
    int scratch_units_per_frame; // capacity of the forward-state scratch per frame, in units of 32 keys
    int wide_ppl;              // pixels per lane of the full-width kernel instance (scratch units per row)
    int narrow_ppl;            // pixels per lane of the half-width instance, 0 if frames are never split in columns
    int max_col_tiles;         // planner: at most this many narrow tiles side by side (2..4)
    int sky_min;               // planner: least number of source-free top rows worth handing to k3_sky; 0 disables
    int frame0;                // index of this sub-batch's first frame in the caller's batch (error reporting)
    // Multipliers handed over at run time so that ptxas keeps the multiply-adds below on the FMA pipe instead of
    // strength-reducing them to shifts/LEAs on the ALU pipe, which is the pipe the scan kernel saturates.
    uint32_t mul_dist;         // 1 << (32 - DSH):  umulhi(key, mul_dist)  == key >> DSH
    uint32_t mul_ord;          // 1 << (32 - OSH):  umulhi(key, mul_ord)   == key >> OSH
    uint32_t neg_ord;          // -(1 << OSH):      key + (key >> OSH) * neg_ord == key & LMASK
    uint32_t four;             // sizeof(float)
    uint32_t one;              // 1: x * one + c keeps a plain add on the FMA pipe
};

struct Workspace {
    uint32_t* srcbits;   // [B*H*WW]
    uint32_t* valbits;   // [B*H*WW]
    uint16_t* wprefix;   // [B*H*WW] sources in the row before this word
    uint8_t* rowcell;    // [B*H*WW] per word: bit j = some source among its pixels 8j..8j+7
    uint32_t* rowsrc;    // [B*H]  K1: row count, K1b: exclusive base within the frame
    uint32_t* rowval;    // [B*H]
    int32_t* counts;     // [B*2]  n_src, n_valid
    float* dlist;        // [B*H*W] depth_list per frame (first n_valid entries used)
    uint32_t* scratch;   // forward state, lane-major rows of 32*PPL keys
    Task* tasks;         // [B * max_tasks_per_frame]
    int* sky;            // [B] S: rows [0,S) lie above every source and are filled by k3_sky (0: none)
    uint32_t* skykeys;   // [B*2*W] final keys of rows S and S+1
    int* status;         // [0] first bad frame (INT_MAX if none), [1] number of wide tasks
};

// ------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// Streaming stores to the caller's output buffers: nothing on this path reads them back, so they carry no
// "memory" clobber and the compiler may keep loads in flight across them.
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ void st_stream_v2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b));
}
__device__ __forceinline__ void st_stream_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d));
}
// 128-bit read-only load that may allocate in L1: a lane's 16 pixels are four such loads of consecutive 16 B, so
// the second half of every 32 B sector is an L1 hit instead of a second trip to L2
__device__ __forceinline__ float4 ld_stream_v4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream_v4u(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// KITTI depth PNG sample -> metres, data_read.py:215 `depth_png.astype(np.float32) / 256.` (exact in float32):
// 0x47000000 is 32768.0f, whose mantissa step is 2^-8, so OR-ing the sample into the mantissa gives 32768 + v/256.
__device__ __forceinline__ float u16_depth(uint32_t v16) { return __uint_as_float(0x47000000u | v16) - 32768.0f; }
__device__ __forceinline__ float load_px(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_px(const uint16_t* p) { return u16_depth(__ldg(p)); }
__device__ __forceinline__ float load_px_stream(const float* p) { return ld_stream(p); }
__device__ __forceinline__ float load_px_stream(const uint16_t* p) { return u16_depth(__ldg(p)); }

// 16 consecutive pixels of a row as they arrive from memory (128-bit loads), decoded on use
template <typename T> struct In16;
template <> struct In16<float> {
    float4 q[4];
    __device__ __forceinline__ void load(const float* p, int col, int W) {
#pragma unroll
        for (int g = 0; g < 4; ++g) q[g] = col + 4 * g < W ? ld_stream_v4(p + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ float4 get(int g) const { return q[g]; }
};
template <> struct In16<uint16_t> {
    uint4 r[2];
    __device__ __forceinline__ void load(const uint16_t* p, int col, int W) {     // W % 8 == 0
#pragma unroll
        for (int k = 0; k < 2; ++k) r[k] = col + 8 * k < W ? ld_stream_v4u(p + 8 * k) : make_uint4(0u, 0u, 0u, 0u);
    }
    __device__ __forceinline__ float4 get(int g) const {
        const uint32_t a = (g & 1) ? r[g >> 1].z : r[g >> 1].x, b = (g & 1) ? r[g >> 1].w : r[g >> 1].y;
        return make_float4(u16_depth(a & 0xFFFFu), u16_depth(a >> 16), u16_depth(b & 0xFFFFu), u16_depth(b >> 16));
    }
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------------------
// K1 (W % 4 == 0): predicates -> bit rows, per-word source prefix, coarse cells, row counts, validity mask,
// and the row-local compaction of the valid depths (into ws.scratch, which K2 only uses later).
// One warp per row; a lane owns 16 consecutive pixels of every 512-pixel chunk (four 128-bit loads).
// The predicates are evaluated without branches: a > b  <=>  sign(b - a) for IEEE floats (a NaN operand gives
// the canonical positive NaN, i.e. "false", like the comparison), and the sign bits of four differences are
// gathered into a nibble with byte permutes and one multiply.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sign_nibble(float t0, float t1, float t2, float t3)
{
    // top bytes of the four floats -> one word -> bits 7,15,23,31 -> nibble (multiply gathers them into 28..31)
    const uint32_t p01 = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x0073);   // [t0.b3, t1.b3, 0, 0]
    const uint32_t p23 = __byte_perm(__float_as_uint(t2), __float_as_uint(t3), 0x0073);
    const uint32_t w = __byte_perm(p01, p23, 0x5410) & 0x80808080u;
    return (w * 0x00204081u) >> 28;
}

template <typename T>
__global__ void __launch_bounds__(256) k1_mask_rows_v16(const T* __restrict__ in, FrameParams fp, Workspace ws,
                                                         uint8_t* __restrict__ out_mask, float* __restrict__ out_lidar)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int W = fp.W, WW = fp.WW;
    const long nrows = (long)fp.B * fp.H;
    const int nchunks = (W + 511) >> 9;
    const float scut = fp.src_cut, vthr = fp.val_thr;
    const bool mask16 = (W & 15) == 0;
    float* rowvals = reinterpret_cast<float*>(ws.scratch);
    for (long row = warp; row < nrows; row += nwarps) {
        const long frame = row / fp.H;
        const T* rp = in + (frame * fp.in_H + fp.in_crop + (row - frame * fp.H)) * W;
        uint32_t cs = 0, cv = 0;
        // software pipeline over the 512-pixel chunks: the 128-bit loads of the next chunk are issued (volatile
        // asm, so they stay ahead) before the current chunk is processed
        In16<T> nq;
        nq.load(rp + lane * 16, lane * 16, W);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int col = (ch << 9) + lane * 16;
            float4 q[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) q[g] = nq.get(g);
            if (ch + 1 < nchunks) nq.load(rp + col + 512, col + 512, W);
            if (out_lidar && col < W) {                  // decoded frame (uint16 input): what the CNN reads as lidar
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (col + 4 * g < W)
                        st_stream_v4(out_lidar + row * W + col + 4 * g, __float_as_uint(q[g].x), __float_as_uint(q[g].y),
                                     __float_as_uint(q[g].z), __float_as_uint(q[g].w));
            }
            uint32_t sb = 0, vb = 0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                // tools.py:8: source <=> !(float32(1 - x) > src_thr) <=> !(x < src_cut), src_cut being the smallest
                // float that satisfies the predicate (found on the host, the predicate is monotone in x);
                // tools.py:22: valid <=> x > val_thr
                const uint32_t ns = sign_nibble(__fsub_rn(q[g].x, scut), __fsub_rn(q[g].y, scut), __fsub_rn(q[g].z, scut),
                                                __fsub_rn(q[g].w, scut));                   // bit = x < src_cut
                const uint32_t nv = sign_nibble(__fsub_rn(vthr, q[g].x), __fsub_rn(vthr, q[g].y),
                                                __fsub_rn(vthr, q[g].z), __fsub_rn(vthr, q[g].w));
                sb |= ns << (4 * g);
                vb |= nv << (4 * g);
            }
            const uint32_t inb = (1u << min(max(W - col, 0), 16)) - 1u;      // pixels of this lane inside the row
            sb = ~sb & inb;
            vb &= inb;
            if (out_mask && col < W) {
                uint32_t m[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) m[g] = (((vb >> (4 * g)) & 0xFu) * 0x00204081u) & 0x01010101u;
                uint8_t* mp = out_mask + row * W + col;
                if (mask16) {
                    st_stream_v4(mp, m[0], m[1], m[2], m[3]);
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (col + 4 * g < W) st_stream_u32(mp + 4 * g, m[g]);
                }
            }
            // 32-bit words from the halves of 2 neighbouring lanes
            uint32_t sw = sb << ((lane & 1) * 16), vw = vb << ((lane & 1) * 16);
            sw |= __shfl_xor_sync(0xffffffffu, sw, 1);
            vw |= __shfl_xor_sync(0xffffffffu, vw, 1);
            // coarse cells: bit j of the word's nibble = some source among its pixels 8j..8j+7
            const uint32_t cell = ((sw & 0xFFu) != 0) | (((sw & 0xFF00u) != 0) << 1) | (((sw & 0xFF0000u) != 0) << 2) |
                                  (((sw & 0xFF000000u) != 0) << 3);
            const uint32_t sany = __ballot_sync(0xffffffffu, sb != 0);
            const uint32_t vany = __ballot_sync(0xffffffffu, vb != 0);
            uint32_t spre = 0, stot = 0;
            if (sany) {
                const uint32_t c = __popc(sb);
                uint32_t inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += o;
                }
                spre = inc - c;
                stot = __shfl_sync(0xffffffffu, inc, 31);
            }
            const int w = (ch << 4) + (lane >> 1);
            if ((lane & 1) == 0 && w < WW) {
                const long wi = row * WW + w;
                ws.srcbits[wi] = sw;
                ws.valbits[wi] = vw;
                ws.wprefix[wi] = (uint16_t)(cs + spre);
                ws.rowcell[wi] = (uint8_t)cell;
            }
            cs += stot;
            if (vany) {
                const uint32_t c = __popc(vb);
                uint32_t inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += o;
                }
                // few pixels per lane are valid (~5 % density): walk the set bits and re-read the values (L1 hits)
                float* dst = rowvals + row * W + cv + (inc - c);
                const T* xs = rp + col;
                uint32_t m = vb;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    *dst++ = load_px(xs + j);
                }
                cv += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        if (lane == 0) {
            ws.rowsrc[row] = cs;
            ws.rowval[row] = cv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K1 for widths that are not a multiple of 4 (no 128-bit row alignment): same outputs, scalar loads + ballots.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k1_mask_rows(const T* __restrict__ in, FrameParams fp, Workspace ws,
                                                     uint8_t* __restrict__ out_mask, float* __restrict__ out_lidar)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int W = fp.W, WW = fp.WW;
    const long nrows = (long)fp.B * fp.H;
    const bool vec_mask = (W & 3) == 0;
    const uint32_t ltmask = lanemask_lt();
    float* rowvals = reinterpret_cast<float*>(ws.scratch);
    for (long row = warp; row < nrows; row += nwarps) {
        const long frame = row / fp.H;
        const T* rp = in + (frame * fp.in_H + fp.in_crop + (row - frame * fp.H)) * W;
        uint32_t cs = 0, cv = 0;
        for (int c0 = 0; c0 < WW; c0 += 16) {
            float x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int col = (c0 + k) * 32 + lane;
                x[k] = col < W ? load_px_stream(rp + col) : 0.0f;
                if (out_lidar && col < W) out_lidar[row * W + col] = x[k];
            }
            uint32_t mys = 0, myv = 0, mypre = 0;
            uint32_t vq[4];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int col = (c0 + k) * 32 + lane;
                const bool inb = col < W;
                const float d = __fsub_rn(1.0f, x[k]);                 // tools.py:8  1.0 - x  (float32)
                const bool sp = inb && !(d > fp.src_thr);             // value_mask == 0  <=> source
                const bool vp = inb && (x[k] > fp.val_thr);           // tools.py:22 with_value
                const uint32_t sw = __ballot_sync(0xffffffffu, sp);
                const uint32_t vw = __ballot_sync(0xffffffffu, vp);
                if (lane == k) { mys = sw; myv = vw; mypre = cs; }
                if (vp) rowvals[row * W + cv + __popc(vw & ltmask)] = x[k];
                cs += __popc(sw);
                cv += __popc(vw);
                vq[k & 3] = vw;
                if (out_mask) {
                    if (vec_mask) {
                        if ((k & 3) == 3) {
                            const int col4 = (c0 + k - 3) * 32 + lane * 4;
                            if (col4 < W) {
                                const int q = lane >> 3;
                                const uint32_t word = q == 0 ? vq[0] : q == 1 ? vq[1] : q == 2 ? vq[2] : vq[3];
                                const uint32_t nib = (word >> ((lane & 7) * 4)) & 0xFu;
                                st_stream_u32(out_mask + row * W + col4, (nib * 0x00204081u) & 0x01010101u);
                            }
                        }
                    } else if (inb) {
                        out_mask[row * W + col] = (uint8_t)vp;
                    }
                }
            }
            if (lane < 16 && c0 + lane < WW) {
                const long wi = row * WW + c0 + lane;
                ws.srcbits[wi] = mys;
                ws.valbits[wi] = myv;
                ws.wprefix[wi] = (uint16_t)mypre;
                ws.rowcell[wi] = (uint8_t)(((mys & 0xFFu) != 0) | (((mys & 0xFF00u) != 0) << 1) |
                                           (((mys & 0xFF0000u) != 0) << 2) | (((mys & 0xFF000000u) != 0) << 3));
            }
        }
        if (lane == 0) {
            ws.rowsrc[row] = cs;
            ws.rowval[row] = cv;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K1b: per frame -- exclusive scans of the row counts, depth_list compaction, task emission.
// One 256-thread block per frame.
// ------------------------------------------------------------------------------------------------------
constexpr int K1B_THREADS = 512;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem /*[K1B_THREADS/32 + 1]*/, uint32_t& total)
{
    constexpr int NW = K1B_THREADS / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < NW; ++i) { const uint32_t t = smem[i]; smem[i] = run; run += t; }
        smem[NW] = run;
    }
    __syncthreads();
    const uint32_t base = smem[wid];
    total = smem[NW];
    __syncthreads();
    return base + inc - v;
}

// barrier among the 256 planner threads only (warps 8..15), so that the compaction warps are not held up
__device__ __forceinline__ void planner_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(K1B_THREADS) k1b_scan_compact(FrameParams fp, Workspace ws,
                                                                 int32_t* __restrict__ out_counts)
{
    __shared__ uint32_t sm[K1B_THREADS / 32 + 1];
    __shared__ uint16_t cellD[MAX_CELLS];        // planner: distance (pixels) to the nearest occupied cell
    __shared__ int cellU[1024];                  // planner: per cell-row upper bound of dt
    __shared__ uint8_t occw[MAX_CELLS / 4 + 1024];   // planner: per cell-row and word, occupancy nibble
    __shared__ uint32_t srcrows[128];            // planner: bit y = row y holds a source (H <= 4096)
    __shared__ Task st[MAXT];                    // planner: tasks of this frame before ordering
    __shared__ int scost[MAXT];
    __shared__ int snt;
    const int b = blockIdx.x;
    const int H = fp.H, W = fp.W, WW = fp.WW;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int per = (H + K1B_THREADS - 1) / K1B_THREADS;
    const int y0 = min(H, tid * per), y1 = min(H, y0 + per);
    uint32_t* rs = ws.rowsrc + (long)b * H;
    uint32_t* rv = ws.rowval + (long)b * H;

#ifdef DTFILL_PLANNER_CLOCKS     // phase clocks of block 0 (profiles/k1b_clock.py reads them through dtfill_debug_read_status)
    const long long dbg_t0 = clock64();
    auto dbg_mark = [&](int k) { if (b == 0 && (tid == 0 || tid == 256)) ws.status[4 + k + (tid ? 8 : 0)] = (int)(clock64() - dbg_t0); };
#else
    auto dbg_mark = [](int) {};
#endif
    if (tid < 128) srcrows[tid] = 0;
    uint32_t ls = 0, lv = 0;
    for (int y = y0; y < y1; ++y) { ls += rs[y]; lv += rv[y]; }
    uint32_t nsrc, nval;
    uint32_t bs = block_exclusive_scan(ls, sm, nsrc);
    uint32_t bv = block_exclusive_scan(lv, sm, nval);
    for (int y = y0; y < y1; ++y) {
        const uint32_t s_ = rs[y], v_ = rv[y];
        rs[y] = bs; rv[y] = bv;
        bs += s_; bv += v_;
        if (s_ && y < 4096) atomicOr(&srcrows[y >> 5], 1u << (y & 31));
    }
    __syncthreads();

    dbg_mark(0);
    const int B = fp.B;
    int kind = (nsrc == 0) ? TASK_NOSRC : ((fp.force_wide || nsrc > MAX_FAST_LABEL) ? TASK_WIDE : TASK_CHAMFER);
    if (nval == 0) kind = TASK_SKIP;
    const int nh = (H + CELL_H - 1) / CELL_H, nw = (W + CELL_W - 1) / CELL_W;
    const bool plan = kind == TASK_CHAMFER && fp.band_cap > 0 && nh * nw <= MAX_CELLS && nh <= 1024 && H <= 4096 &&
                      2 * H > fp.band_cap;

    if (wid < 8) {
        // ---- warps 0..7: depth_list = in[valid] in raster order (tools.py:24).  K1 left every row's valid depths
        // compacted at the start of the row's slot in ws.scratch; concatenate the non-empty rows.
        float* dl = ws.dlist + (long)b * H * W;
        const float* rowvals = reinterpret_cast<const float*>(ws.scratch) + (long)b * H * W;
        for (int yb = wid * 32; yb < H; yb += 8 * 32) {
            // one coalesced read of 33 row bases per 32 rows instead of two dependent loads per row
            const int yy = yb + lane;
            const uint32_t mybase = yy < H ? rv[yy] : nval;
            const uint32_t nextbase = __shfl_down_sync(0xffffffffu, mybase, 1);
            const uint32_t after = (yb + 32 < H) ? rv[yb + 32] : nval;
            const uint32_t mycnt = (lane == 31 ? after : nextbase) - mybase;
            for (int r = 0; r < 32 && yb + r < H; ++r) {
                const uint32_t cnt = __shfl_sync(0xffffffffu, mycnt, r);
                if (cnt == 0) continue;
                const uint32_t base = __shfl_sync(0xffffffffu, mybase, r);
                const float* src = rowvals + (long)(yb + r) * W;
                for (uint32_t i0 = 0; i0 < cnt; i0 += 32 * 12) {       // 12 loads in flight per lane
                    float v[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) {
                        const uint32_t i = i0 + k * 32 + lane;
                        v[k] = i < cnt ? src[i] : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 12; ++k) {
                        const uint32_t i = i0 + k * 32 + lane;
                        if (i < cnt) dl[base + i] = v[k];
                    }
                }
            }
        }
        dbg_mark(1);
        return;
    }

    // ---- warps 8..15: tile planner -------------------------------------------------------------------------
    const int ptid = tid - 256, pw = wid - 8;
    if (plan) {
        // coarse occupancy -> exact anisotropic city-block distance on the cell grid (two sweeps per axis)
        // rowcell nibbles of CELL_H consecutive rows OR-ed per word, then one distance cell per bit
        const uint8_t* rc = ws.rowcell + (long)b * H * WW;
        for (int i0 = 0; i0 < nh * WW; i0 += 256 * 8) {           // 32 independent byte loads in flight per thread
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + k * 256 + ptid;
                o[k] = 0;
                if (i < nh * WW) {
                    const int cy = i / WW, w = i - cy * WW;
#pragma unroll
                    for (int r = 0; r < CELL_H; ++r) {
                        const int y = cy * CELL_H + r;
                        if (y < H) o[k] |= rc[(long)y * WW + w];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + k * 256 + ptid;
                if (i < nh * WW) occw[i] = (uint8_t)o[k];
            }
        }
        planner_sync();
        dbg_mark(1);
        for (int i = ptid; i < nh * nw; i += 256) {
            const int cy = i / nw, cx = i - cy * nw;
            cellD[i] = ((occw[cy * WW + (cx >> 2)] >> (cx & 3)) & 1u) ? 0 : 60000;
        }
        planner_sync();
        dbg_mark(2);
        for (int cx = ptid; cx < nw; cx += 256) {        // vertical sweeps, one thread per cell column
            uint32_t d = 60000;
            for (int c0 = 0; c0 < nh; c0 += 8) {         // 8 loads ahead of the dependent (min,+) chain
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = c0 + k < nh ? (uint32_t)cellD[(c0 + k) * nw + cx] : 60000u;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    d = min(min(d + CELL_H, v[k]), 60000u);
                    if (c0 + k < nh) cellD[(c0 + k) * nw + cx] = (uint16_t)d;
                }
            }
            d = 60000;
            for (int c0 = nh - 1; c0 >= 0; c0 -= 8) {
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = c0 - k >= 0 ? (uint32_t)cellD[(c0 - k) * nw + cx] : 60000u;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    d = min(min(d + CELL_H, v[k]), 60000u);
                    if (c0 - k >= 0) cellD[(c0 - k) * nw + cx] = (uint16_t)d;
                }
            }
        }
        planner_sync();
        dbg_mark(3);
        // horizontal sweeps, one warp per cell row: a lane keeps its (up to 8) consecutive cells in registers,
        // (min,+) scans across lanes via shuffles; only the row maximum leaves the warp
        const int chunk = (nw + 31) / 32;
        if (chunk <= 8) {
            for (int cy = pw; cy < nh; cy += 8) {
                const uint16_t* rowp = cellD + cy * nw;
                const int xa = lane * chunk;
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = (k < chunk && xa + k < nw) ? (uint32_t)rowp[xa + k] : 120000u;
                // left -> right
                uint32_t d = 120000u;
#pragma unroll
                for (int k = 0; k < 8; ++k) if (k < chunk) d = min(d + CELL_W, v[k]);
                uint32_t e = d;
#pragma unroll
                for (int s_ = 1; s_ < 32; s_ <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, e, s_);
                    if (lane >= s_) e = min(e, o + (uint32_t)(s_ * chunk * CELL_W));
                }
                uint32_t cin = __shfl_up_sync(0xffffffffu, e, 1);
                d = lane == 0 ? 120000u : cin;
#pragma unroll
                for (int k = 0; k < 8; ++k) if (k < chunk) { d = min(d + CELL_W, v[k]); v[k] = d; }
                // right -> left
                d = 120000u;
#pragma unroll
                for (int k = 7; k >= 0; --k) if (k < chunk) d = min(d + CELL_W, v[k]);
                e = d;
#pragma unroll
                for (int s_ = 1; s_ < 32; s_ <<= 1) {
                    const uint32_t o = __shfl_down_sync(0xffffffffu, e, s_);
                    if (lane + s_ < 32) e = min(e, o + (uint32_t)(s_ * chunk * CELL_W));
                }
                cin = __shfl_down_sync(0xffffffffu, e, 1);
                d = lane == 31 ? 120000u : cin;
                uint32_t mx = 0;
#pragma unroll
                for (int k = 7; k >= 0; --k)
                    if (k < chunk) { d = min(d + CELL_W, v[k]); if (xa + k < nw) mx = max(mx, d); }
#pragma unroll
                for (int s_ = 16; s_ > 0; s_ >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, s_));
                if (lane == 0) cellU[cy] = (int)min(mx, 60000u) + (CELL_H - 1) + (CELL_W - 1);   // >= max dt of the cell row
            }
        } else {
            for (int cy = pw * 32 + lane; cy < nh; cy += 256) {      // very wide frames: one thread per cell row
                uint32_t d = 60000;
                for (int cx = 0; cx < nw; ++cx) { d = min(d + CELL_W, (uint32_t)cellD[cy * nw + cx]); cellD[cy * nw + cx] = (uint16_t)min(d, 60000u); }
                d = 60000;
                uint32_t mx = 0;
                for (int cx = nw - 1; cx >= 0; --cx) { d = min(d + CELL_W, (uint32_t)cellD[cy * nw + cx]); mx = max(mx, min(d, 60000u)); }
                cellU[cy] = (int)mx + (CELL_H - 1) + (CELL_W - 1);
            }
        }
        planner_sync();
    }
    dbg_mark(4);

    if (ptid == 0) {
        ws.counts[2 * b] = (int)nsrc;
        ws.counts[2 * b + 1] = (int)nval;
        if (out_counts) { out_counts[2 * b] = (int)nsrc; out_counts[2 * b + 1] = (int)nval; }
        // numpy's IndexError: empty depth_list, or a label beyond its end (tools.py:26)
        if (nval == 0 || nsrc > nval) atomicMin(&ws.status[0], b + fp.frame0);
        if (kind == TASK_WIDE) atomicAdd(&ws.status[1], 1);

        int nt = 0;
        Task* t = st;                 // tasks are built in shared memory and written out by all planner threads
        int* cost = scost;
        auto blank = [&](int knd) {
            Task q;
            q.frame = b; q.lo = 0; q.hi = H; q.r0 = 0; q.r1 = H; q.kind = knd; q.scratch_off = 0; q.fstart = 0;
            q.clo = 0; q.c0 = 0; q.c1 = W; q.sky = -2;
            return q;
        };
        // Rows above the first source row f need no scan: their distance is that of row f plus the row offset and
        // their label follows a fixed route down to two base rows (see k3_sky).  S = rows handed to k3_sky, a
        // multiple of the cell height with S + 1 <= f; the tiles below cover rows [S, H).
        int S = 0;
        if (plan && fp.sky_min > 0 && W <= SKY_MAX_W) {
            int f = 0;
            for (int w = 0; w < 128 && (w << 5) < H; ++w)
                if (srcrows[w]) { f = (w << 5) + __ffs(srcrows[w]) - 1; break; }
            const int s4 = f >= 1 ? ((f - 1) / CELL_H) * CELL_H : 0;
            if (s4 >= fp.sky_min) S = s4;
        }
        if (plan) {
            const int nwid = fp.narrow_ppl * 32;                 // width of a half-width tile (0: never split)
            int cy = S / CELL_H, scr = 0;
            bool ok = true;
            while (cy < nh && ok) {
                const int r0 = cy * CELL_H;
                int lo = 1 << 30, hi = 0, prev = 0, end = cy, best_lo = 0, best_hi = 0, umax = 0, best_u = 0;
                for (int c = cy; c < nh; ++c) {
                    lo = min(lo, c * CELL_H - cellU[c]);
                    hi = max(hi, min(H, (c + 1) * CELL_H) + cellU[c]);
                    umax = max(umax, cellU[c]);
                    const int L = max(S, lo), Hh = min(H, hi);      // nothing above S feeds the forward pass
                    const int cst = (Hh - L) + (Hh - r0);
                    // extend while the tile stays under the target cost, while extending is (nearly) free, or while
                    // the band is still short compared with its halo (sparse frames: tall bands, less redundancy)
                    const bool take = c == cy || nt >= MAXT - 4 || cst <= fp.band_cap || cst - prev <= CELL_H ||
                                      (c - cy) * CELL_H < 2 * cellU[c];
                    if (!take) break;
                    prev = cst; end = c + 1; best_lo = L; best_hi = Hh; best_u = umax;
                }
                Task q = blank(TASK_CHAMFER);
                q.lo = best_lo; q.hi = best_hi; q.r0 = r0; q.r1 = min(H, end * CELL_H);
                q.sky = (S > 0 && r0 == S) ? S : -2;
                // n overlapping narrow tiles when the bound leaves every written pixel's ball inside its tile and the
                // extra columns stay below ~60 % (n * nwid <= 1.6 W)
                int ntile = 0;
                if (nwid > 0 && W > nwid && (W & 3) == 0) {
                    for (int n = 2; n <= fp.max_col_tiles && !ntile; ++n) {
                        if (5 * n * nwid > 8 * W || nt + n > MAXT) break;
                        bool fits = true;                      // every interior tile edge at least best_u away
                        int prev_split = 0;
                        for (int k = 0; k < n && fits; ++k) {
                            const int s0 = (int)(((long)(W - nwid) * k / (n - 1)) & ~3L);
                            const int s1 = (int)(((long)(W - nwid) * (k + 1) / (n - 1)) & ~3L);
                            const int split = k == n - 1 ? W : ((s1 + s0 + nwid) / 2) & ~3;
                            if ((k > 0 && prev_split - s0 < best_u) || (k < n - 1 && s0 + nwid - split < best_u) ||
                                split <= prev_split) fits = false;
                            prev_split = split;
                        }
                        if (fits) ntile = n;
                    }
                }
                if (ntile) {
                    q.kind = TASK_NARROW;
                    int prev_split = 0;
                    for (int k = 0; k < ntile; ++k) {
                        const int s0 = (int)(((long)(W - nwid) * k / (ntile - 1)) & ~3L);           // sub-image start
                        const int s1 = (int)(((long)(W - nwid) * (k + 1) / (ntile - 1)) & ~3L);     // next tile's start
                        const int split = k == ntile - 1 ? W : ((s1 + s0 + nwid) / 2) & ~3;        // middle of the overlap
                        q.clo = s0; q.c0 = prev_split; q.c1 = split; q.scratch_off = scr;
                        // halo check (the sizes above guarantee it; keep the planner honest)
                        if ((k > 0 && q.c0 - s0 < best_u) || (k < ntile - 1 && s0 + nwid - split < best_u)) ok = false;
                        scr += (best_hi - best_lo) * fp.narrow_ppl;
                        cost[nt] = prev; t[nt++] = q;
                        prev_split = split;
                    }
                } else {
                    q.scratch_off = scr;
                    scr += (best_hi - best_lo) * fp.wide_ppl;
                    cost[nt] = 2 * prev;                      // twice the work per row step of a narrow tile
                    t[nt++] = q;
                }
                if (scr > fp.scratch_units_per_frame) ok = false;
                cy = end;
            }
            if (!ok) { nt = 0; S = 0; }
        }
        ws.sky[b] = S;
        if (nt == 0) {
            cost[0] = 4 * H;
            t[nt++] = blank(kind);
        }
        snt = nt;
    }
    dbg_mark(5);
    planner_sync();
    // ---- all planner threads: order the tasks (longest first: the block scheduler hands out blocks in index
    // order, slot-major task array), fill in the forward start rows, write the 32 slots of this frame
    {
        const int nt = snt;
        if (ptid < MAXT) {
            Task q;
            int slot = ptid;
            if (ptid < nt) {
                const int c = scost[ptid];
                int rank = 0;
                for (int j = 0; j < nt; ++j) rank += (scost[j] > c) || (scost[j] == c && j < ptid);
                slot = rank;
                q = st[ptid];
                if (plan) {       // rows without any source above them stay unreached in the forward pass: skip them
                    int f = H;
                    for (int w = q.lo >> 5; w < 128 && (w << 5) < H; ++w) {
                        uint32_t m = srcrows[w];
                        if (w == (q.lo >> 5)) m &= ~0u << (q.lo & 31);
                        if (m) { f = min(H, (w << 5) + __ffs(m) - 1); break; }
                    }
                    q.fstart = min(f, q.hi - 1);
                }
                q.scratch_off += b * fp.scratch_units_per_frame;
            } else {
                q.frame = b; q.lo = 0; q.hi = 0; q.r0 = 0; q.r1 = 0; q.kind = TASK_SKIP; q.scratch_off = 0; q.fstart = 0;
                q.clo = 0; q.c0 = 0; q.c1 = W; q.sky = -2;
            }
            ws.tasks[(long)slot * B + b] = q;
        }
    }
    dbg_mark(6);
}

// ------------------------------------------------------------------------------------------------------
// K2: the chamfer scan, fast path.  One warp per task; lane l owns columns [l*PPL, (l+1)*PPL) of every row.
// ------------------------------------------------------------------------------------------------------
template <int PPL>
struct Row {
    uint32_t v[PPL];
    uint32_t l1, l2;   // columns -1, -2 (previous lane's last two)
    uint32_t r0, r1;   // columns PPL, PPL+1 (next lane's first two)
};

template <int PPL>
__device__ __forceinline__ uint32_t at(const Row<PPL>& r, int idx) {
    // idx is a compile-time constant after unrolling
    return idx == -2 ? r.l2 : idx == -1 ? r.l1 : idx == PPL ? r.r0 : idx == PPL + 1 ? r.r1 : r.v[idx < 0 ? 0 : (idx >= PPL ? PPL - 1 : idx)];
}

template <int PPL>
__device__ __forceinline__ void fill_row(Row<PPL>& r, uint32_t k) {
#pragma unroll
    for (int i = 0; i < PPL; ++i) r.v[i] = k;
    r.l1 = r.l2 = r.r0 = r.r1 = k;
}

template <int PPL>
__device__ __forceinline__ void refresh_halo(Row<PPL>& r, int lane, uint32_t init_key) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, r.v[PPL - 1], 1);
    const uint32_t b = __shfl_up_sync(0xffffffffu, r.v[PPL - 2], 1);
    const uint32_t c = __shfl_down_sync(0xffffffffu, r.v[0], 1);
    const uint32_t d = __shfl_down_sync(0xffffffffu, r.v[1], 1);
    r.l1 = lane == 0 ? init_key : a;
    r.l2 = lane == 0 ? init_key : b;
    r.r0 = lane == 31 ? init_key : c;
    r.r1 = lane == 31 ? init_key : d;
}

// Carry entering this lane from the lanes before it (DIR=+1, forward scan) or after it (DIR=-1, backward
// scan).  e = this lane's outgoing value (cleared key).  Works in a widened dist:14|label:18 form so that
// adding up to 31*PPL columns cannot overflow.  Ties keep the nearer lane (OpenCV: the left neighbour is
// the last candidate compared, so a value already held wins).
template <int PPL, int DIR>
__device__ __forceinline__ uint32_t lane_carry(uint32_t e, int lane, uint32_t clamp_dist)
{
    uint32_t E = ((e >> DSH) << OSH) | (e & LMASK);
    {   // distance 1: the neighbouring lane
        const uint32_t o = DIR > 0 ? __shfl_up_sync(0xffffffffu, E, 1) : __shfl_down_sync(0xffffffffu, E, 1);
        const uint32_t t = o + (uint32_t(PPL) << OSH);
        E = ((t | LMASK) < E) ? t : E;
    }
    // a value carried over d >= 2 lanes is at least 2*PPL; it can only win where a lane's own value is larger than
    // that, which never happens in densely sampled tiles: one warp-wide maximum decides whether to go on
    if (__reduce_max_sync(0xffffffffu, E >> OSH) >= 2u * PPL) {
#pragma unroll
        for (int d = 2; d < 32; d <<= 1) {
            const uint32_t o = DIR > 0 ? __shfl_up_sync(0xffffffffu, E, d) : __shfl_down_sync(0xffffffffu, E, d);
            const uint32_t t = o + (uint32_t(d * PPL) << OSH);
            E = ((t | LMASK) < E) ? t : E;
        }
    }
    const uint32_t cin = DIR > 0 ? __shfl_up_sync(0xffffffffu, E, 1) : __shfl_down_sync(0xffffffffu, E, 1);
    const uint32_t cd = min(cin >> OSH, clamp_dist);
    uint32_t key = (cd << DSH) | (1u << OSH) | (cin & LMASK);
    const bool edge = DIR > 0 ? (lane == 0) : (lane == 31);
    if (edge) key = (clamp_dist << DSH) | (1u << OSH);
    return key;
}

__device__ __forceinline__ void cp_async8(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async16_l2only(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
// forward-state scratch is written once and read once, a whole pass later: keep it out of L1 (L2 only)
__device__ __forceinline__ void st_scratch_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_scratch_v2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// dist field of a key as float, on the FMA pipe (the ALU pipe is the one this kernel saturates):
// (key >> 21) + 2^23 as the high half of a multiply-add, then the float with that bit pattern minus 2^23.
__device__ __forceinline__ float key_dist_f32(uint32_t key, uint32_t mul_dist) {
    uint32_t t;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(key), "r"(mul_dist), "r"(0x4B000000u));
    return __uint_as_float(t) - 8388608.0f;
}
// label field of a key, two multiply-adds on the FMA pipe
__device__ __forceinline__ uint32_t key_label(uint32_t key, uint32_t mul_ord, uint32_t neg_ord) {
    uint32_t hi, l;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(key), "r"(mul_ord));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(l) : "r"(hi), "r"(neg_ord), "r"(key));
    return l;
}

// Raw words holding this lane's PPL source bits of one row, fetched one row ahead of their use.
struct RowBits {
    uint32_t a, b, c, pre, base;
};

template <int PPL>
__device__ __forceinline__ RowBits fetch_row_bits(const uint32_t* __restrict__ bits_f, const uint16_t* __restrict__ pre_f,
                                                  const uint32_t* __restrict__ rowbase, int WW, int x0, int y)
{
    RowBits r;
    const int w = x0 >> 5;
    const uint32_t* br = bits_f + (long)y * WW;
    r.a = w < WW ? br[w] : 0u;
    r.b = w + 1 < WW ? br[w + 1] : 0u;
    r.c = (PPL > 33 && w + 2 < WW) ? br[w + 2] : 0u;
    r.pre = w < WW ? (uint32_t)pre_f[(long)y * WW + w] : 0u;
    r.base = rowbase[y];
    return r;
}

// bit i of `bits` = column x0+i is a source; rank = 1-based raster rank of the first source of this lane
template <int PPL>
struct LaneBits { typedef uint64_t type; };
template <> struct LaneBits<10> { typedef uint32_t type; };
template <> struct LaneBits<20> { typedef uint32_t type; };

template <int PPL>
__device__ __forceinline__ void decode_row_bits(const RowBits& r, int x0, typename LaneBits<PPL>::type& bits,
                                                uint32_t& rank)
{
    const int sh = x0 & 31;
    if (sizeof(typename LaneBits<PPL>::type) == 4) {
        bits = __funnelshift_r(r.a, r.b, sh) & ((1u << (PPL & 31)) - 1u);     // PPL <= 32 bits from two words
    } else {
        uint64_t lo = ((uint64_t)r.b << 32) | r.a;
        lo >>= sh;
        if (PPL > 33 && sh) lo |= (uint64_t)r.c << (64 - sh);
        bits = (typename LaneBits<PPL>::type)(lo & ((PPL >= 64) ? ~0ull : ((1ull << PPL) - 1ull)));
    }
    rank = r.base + r.pre + __popc(r.a & ((1u << sh) - 1u)) + 1u;
}

template <int PPL, bool PAD, bool WANT_LBL, bool VEC>
__global__ void __launch_bounds__(32, (PPL >= 38 ? 12 : (PPL >= 20 ? 20 : 32))) k2_chamfer(FrameParams fp, Workspace ws, float* __restrict__ out_depth,
                                                  float* __restrict__ out_dt, int32_t* __restrict__ out_lbl, int my_kind)
{
    // Transposition buffer for the keys of an output row.  Keeping shared memory small matters: what is left of the
    // 228 KB is the L1 that serves the depth_list gather.
    __shared__ __align__(16) uint32_t stage[32 * PPL];
    __shared__ __align__(16) uint2 fwdbuf[16 * PPL];      // forward keys of the next row to scan, [j][lane]

    const Task task = ws.tasks[blockIdx.x];      // slot-major: blockIdx = slot * B + frame, longest tasks first
    if (task.kind != my_kind && !(task.kind == TASK_NOSRC && my_kind == TASK_CHAMFER)) return;
    const int lane = threadIdx.x;
    const int H = fp.H, W = fp.W, WW = fp.WW;
    const int b = task.frame;
    const long fpx = (long)b * H * W;

    if (task.kind == TASK_NOSRC) {
        // no source anywhere: OpenCV leaves dt at 65533 and lbl at 0; depth_list[0-1] is numpy's last element
        const int nval = ws.counts[2 * b + 1];
        const float last = ws.dlist[fpx + (nval > 0 ? nval - 1 : 0)];
        for (long i = (long)task.r0 * W + lane; i < (long)task.r1 * W; i += 32) {
            out_depth[fpx + i] = last;
            if (out_dt) out_dt[fpx + i] = UNREACHED_DT;
            if (WANT_LBL) out_lbl[fpx + i] = 0;
        }
        return;
    }

    const int x0 = task.clo + lane * PPL;        // first image column of this lane
    const int xl = lane * PPL;                   // same, relative to the tile
    const uint32_t init_key = (uint32_t)fp.init_dist << DSH;
    const uint32_t clamp_dist = 2047u - PPL - 1u;
    const uint32_t* bits_f = ws.srcbits + (long)b * H * WW;
    const uint16_t* pre_f = ws.wprefix + (long)b * H * WW;
    const uint32_t* rowbase = ws.rowsrc + (long)b * H;
    constexpr int VW = (PPL % 4 == 0) ? 4 : 2;   // keys per scratch vector
    uint2* scr = reinterpret_cast<uint2*>(ws.scratch) + (long)task.scratch_off * 16;   // 32*PPL keys per row

    Row<PPL> ra, rb;
    fill_row(ra, init_key);
    fill_row(rb, init_key);

    // ---------------- forward pass: rows lo .. hi-1 ----------------
    RowBits nextbits = fetch_row_bits<PPL>(bits_f, pre_f, rowbase, WW, x0, task.fstart);
    auto fwd_step = [&](const Row<PPL>& A /*row y-1*/, Row<PPL>& Bq /*row y-2 in, row y out*/, int y) {
        typename LaneBits<PPL>::type bits; uint32_t rank;
        decode_row_bits<PPL>(nextbits, x0, bits, rank);
        nextbits = fetch_row_bits<PPL>(bits_f, pre_f, rowbase, WW, x0, min(y + 1, task.hi - 1));   // one row ahead
        if (lane < 2 && y + 4 < task.hi) {           // bit row and prefixes four rows ahead -> L2
            asm volatile("prefetch.global.L2 [%0];" ::"l"(bits_f + (long)(y + 4) * WW + (x0 >> 5) + lane * 32));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pre_f + (long)(y + 4) * WW + (x0 >> 5) + lane * 32));
        }
        uint32_t c[PPL];
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            uint32_t m = at(Bq, i - 1) * fp.one + KC(3, 0);              // (-2,-1) cost 3 (IMAD: FMA pipe)
            m = __viaddmin_u32(at(Bq, i + 1), KC(3, 2), m);              // (-2,+1) cost 3
            m = __viaddmin_u32(at(A, i - 2), KC(3, 4), m);               // (-1,-2) cost 3
            m = __viaddmin_u32(at(A, i - 1), KC(2, 6), m);               // (-1,-1) cost 2
            m = __viaddmin_u32(at(A, i), KC(1, 8), m);                   // (-1, 0) cost 1
            m = __viaddmin_u32(at(A, i + 1), KC(2, 10), m);              // (-1,+1) cost 2
            m = __viaddmin_u32(at(A, i + 2), KC(3, 12), m);              // (-1,+2) cost 3
            c[i] = m;
        }
        if (__any_sync(0xffffffffu, bits != 0)) {                         // sources: dist 0, own raster rank
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
                const bool s = (bits & ((typename LaneBits<PPL>::type)1 << i)) != 0;
                c[i] = s ? rank : c[i];
                rank += s ? 1u : 0u;
            }
        }
        // in-lane scan: T[x] = min(c[x], T[x-1] + 1); the left neighbour is OpenCV's last candidate (order 14)
        uint32_t u = c[0] & ORDCLR;
        c[0] = u;
#pragma unroll
        for (int i = 1; i < PPL; ++i) {
            u = __viaddmin_u32(u, KC(1, 14), c[i]) & ORDCLR;
            c[i] = u;
        }
        const uint32_t cin = lane_carry<PPL, +1>(u, lane, clamp_dist);
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            uint32_t t = __viaddmin_u32(cin, uint32_t(i + 1) << DSH, c[i]);   // order bit 0/1 stays (see header)
            if (PAD && x0 + i >= W) t = init_key;
            Bq.v[i] = t;
        }
        refresh_halo(Bq, lane, init_key);
        // forward state -> scratch, [vector j][lane] so that every store instruction is fully coalesced; the rows of
        // the upper halo are never read back (the backward pass ends at r0)
        if (y >= task.r0) {
            char* dst = reinterpret_cast<char*>(scr) + (long)(y - task.lo) * (128 * PPL) + lane * (4 * VW);
#pragma unroll
            for (int j = 0; j < PPL / VW; ++j) {
                if (VW == 4) st_scratch_v4(dst + j * 512, Bq.v[4 * j], Bq.v[4 * j + 1], Bq.v[4 * j + 2], Bq.v[4 * j + 3]);
                else st_scratch_v2(dst + j * 256, Bq.v[2 * j], Bq.v[2 * j + 1]);
            }
        }
    };

    // one copy of the step in the instruction stream (the unrolled step is ~13 KB of code); the two live rows
    // are rotated with register moves, which go to the otherwise idle FMA pipe
#pragma unroll 1
    for (int y = task.fstart; y < task.hi; ++y) {
        fwd_step(ra, rb, y);
        const Row<PPL> t = ra; ra = rb; rb = t;
    }

    // ---------------- backward pass: rows hi-1 .. r0 ----------------
    // The forward keys of row y-1 are copied scratch -> fwdbuf with cp.async (16 B, L2 only) while row y is being
    // scanned: no registers, no exposed latency, no L1 pollution.  (The depth gather is NOT done with cp.async: 4-byte
    // LDGSTS cost 8 LSU cycles each and 20-38 of them per row step saturate the LSU -- measured.)
    fill_row(ra, init_key);
    fill_row(rb, init_key);
    const float* dl = ws.dlist + fpx;
    const char* dlm1_bytes = reinterpret_cast<const char*>(dl - 1);       // depth_list[lbl - 1]
    // output addressing that does not depend on the row: which 4-pixel groups of the transposed row this lane
    // writes (inside [c0,c1)), and where
    constexpr int NJ = (32 * PPL + 127) / 128;
    uint32_t okmask = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int lc = (j * 32 + lane) * 4, col = task.clo + lc;
        if (lc < 32 * PPL && col >= task.c0 && col < task.c1) okmask |= 1u << j;
    }
    const long colbase = fpx + task.clo + lane * 4;
    const uint4* sread = reinterpret_cast<const uint4*>(&stage[lane * 4]);
    uint2* swrite = reinterpret_cast<uint2*>(&stage[xl]);
    const uint32_t fwdbuf_lane = (uint32_t)__cvta_generic_to_shared(fwdbuf) + lane * (4 * VW);

    auto issue_fwd_row = [&](int y) {            // group A(y)
        if (y >= task.fstart && y >= task.lo) {
            const char* src = reinterpret_cast<const char*>(scr) + (long)(y - task.lo) * (128 * PPL) + lane * (4 * VW);
#pragma unroll
            for (int j = 0; j < PPL / VW; ++j) {
                if (VW == 4) cp_async16_l2only(fwdbuf_lane + j * 512, src + j * 512);
                else cp_async8(fwdbuf_lane + j * 256, src + j * 256);
            }
        }
        cp_async_commit();
    };
    // Depths gathered for an output row stay in registers across the loop back-edge and are stored at the start of
    // the next step: the gather's latency is covered by the row rotation, and nothing else is live meanwhile.
    uint32_t g[NJ * 4];
    auto flush_depth_row = [&](int y) {
        float* pd = out_depth + colbase + (long)y * W;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if ((okmask >> j) & 1u) st_stream_v4(pd + j * 128, g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
    };

    issue_fwd_row(task.hi - 1);

    auto bwd_step = [&](const Row<PPL>& A /*row y+1*/, Row<PPL>& Bq /*row y+2 in, row y out*/, int y) {
        if (lane < PPL && y - 3 >= task.fstart)      // forward row three steps ahead -> L2 (one 128 B line per lane)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(scr) +
                                                          (long)(y - 3 - task.lo) * (128 * PPL) + lane * 128));
        if (VEC && y + 1 >= task.r0 && y + 1 < task.r1) flush_depth_row(y + 1);     // gathered during the last step
        cp_async_wait<0>();                          // A(y), the only group in flight, has landed
        uint32_t c[PPL];
        if (y >= task.fstart) {
#pragma unroll
            for (int j = 0; j < PPL / VW; ++j) {
                if (VW == 4) {
                    const uint4 f = reinterpret_cast<const uint4*>(fwdbuf)[j * 32 + lane];
                    c[4 * j] = f.x; c[4 * j + 1] = f.y; c[4 * j + 2] = f.z; c[4 * j + 3] = f.w;
                } else {
                    const uint2 f = fwdbuf[j * 32 + lane];
                    c[2 * j] = f.x; c[2 * j + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < PPL; ++i) c[i] = init_key;      // rows the forward pass skipped: unreached
        }
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            uint32_t m = c[i];                                            // own forward value first (order <= 1)
            m = __viaddmin_u32(at(Bq, i + 1), KC(3, 2), m);              // (+2,+1)
            m = __viaddmin_u32(at(Bq, i - 1), KC(3, 4), m);              // (+2,-1)
            m = __viaddmin_u32(at(A, i + 2), KC(3, 6), m);               // (+1,+2)
            m = __viaddmin_u32(at(A, i + 1), KC(2, 8), m);               // (+1,+1)
            m = __viaddmin_u32(at(A, i), KC(1, 10), m);                  // (+1, 0)
            m = __viaddmin_u32(at(A, i - 1), KC(2, 12), m);              // (+1,-1)
            m = __viaddmin_u32(at(A, i - 2), KC(3, 14), m);              // (+1,-2)
            c[i] = m & ORDCLR;
        }
        issue_fwd_row(y - 1);                        // A(y-1): fwdbuf has been consumed above
        uint32_t u = c[PPL - 1];
#pragma unroll
        for (int i = PPL - 2; i >= 0; --i) {
            u = __viaddmin_u32(u, KC(1, 1), c[i]) & ORDCLR;              // right neighbour is compared last
            c[i] = u;
        }
        const uint32_t cin = lane_carry<PPL, -1>(u, lane, clamp_dist);
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            uint32_t t = __viaddmin_u32(cin, uint32_t(PPL - i) << DSH, c[i]);   // order bit 0/1 stays
            if (PAD && x0 + i >= W) t = init_key;
            Bq.v[i] = t;
        }
        refresh_halo(Bq, lane, init_key);

        // ---- output of row y: keys -> shared memory (transpose), then per lane 4 consecutive pixels per group:
        // dt / lbl stores and the gather depth_list[lbl-1] (tools.py:26).  In this layout neighbouring lanes ask
        // for neighbouring labels (consecutive ranks along a beam), so a gather instruction touches few lines.
        if (y >= task.r0 && y < task.r1) {
#pragma unroll
            for (int j = 0; j < PPL / 2; ++j) swrite[j] = make_uint2(Bq.v[2 * j], Bq.v[2 * j + 1]);
            __syncwarp();
            const long ro = (long)y * W;
            if (VEC) {
                float* pdt = out_dt + colbase + ro;
                int32_t* plb = out_lbl + colbase + ro;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if ((okmask >> j) & 1u) {
                        const uint4 k = sread[j * 32];
                        if (out_dt)
                            st_stream_v4(pdt + j * 128, __float_as_uint(key_dist_f32(k.x, fp.mul_dist)),
                                         __float_as_uint(key_dist_f32(k.y, fp.mul_dist)),
                                         __float_as_uint(key_dist_f32(k.z, fp.mul_dist)),
                                         __float_as_uint(key_dist_f32(k.w, fp.mul_dist)));
                        if (WANT_LBL)
                            st_stream_v4(plb + j * 128, k.x & LMASK, k.y & LMASK, k.z & LMASK, k.w & LMASK);
                        const uint32_t kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t l = kk[e] & LMASK;                                // >= 1 inside [c0,c1)
                            g[4 * j + e] = __float_as_uint(*reinterpret_cast<const float*>(dlm1_bytes + (uint64_t)l * fp.four));
                        }
                    }
                }
            } else {
                for (int lc = lane; lc < 32 * PPL; lc += 32) {
                    const int col = task.clo + lc;
                    if (col >= task.c0 && col < task.c1) {
                        const uint32_t k = stage[lc];
                        if (out_dt) out_dt[fpx + ro + col] = (float)(k >> DSH);
                        if (WANT_LBL) out_lbl[fpx + ro + col] = (int32_t)(k & LMASK);
                        out_depth[fpx + ro + col] = dl[(k & LMASK) - 1u];
                    }
                }
            }
            if (y <= task.sky + 1) {                 // base rows S, S+1 of the source-free top rows: keys for k3_sky
                uint32_t* sk = ws.skykeys + ((long)b * 2 + (y - task.sky)) * W;
                for (int lc = lane * 4; lc < 32 * PPL; lc += 128) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = task.clo + lc + e;
                        if (col >= task.c0 && col < task.c1) sk[col] = stage[lc + e];
                    }
                }
            }
            __syncwarp();                            // stage is free for the next row
        }
    };

#pragma unroll 1
    for (int y = task.hi - 1; y >= task.r0; --y) {
        bwd_step(ra, rb, y);
        const Row<PPL> t = ra; ra = rb; rb = t;
    }
    cp_async_wait<0>();
    if (VEC) flush_depth_row(task.r0);
}

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
    const int H = fp.H, W = fp.W, tid = threadIdx.x;
    const long fpx = (long)b * H * W;
    const uint32_t* sk = ws.skykeys + (long)b * 2 * W;
    const float* dl = ws.dlist + fpx;
    for (int x = tid; x < W; x += 256) {
        const uint32_t k0 = sk[x], k1 = sk[W + x];
        d0[x + 1] = (uint16_t)(k0 >> DSH);
        dep[0][x] = dl[(k0 & LMASK) - 1u];
        dep[1][x] = dl[(k1 & LMASK) - 1u];
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
        const GT io = (GT)1.0 / og, it = (GT)1.0 / t;                  // :232-233
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
                                                         double* __restrict__ sums /*[10]*/)
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
        sums[threadIdx.x] = s;
    }
}

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
                                                         int H, int W, int T, float* __restrict__ out)
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
    out[fpx + (long)gy * W + gx] = sum / (0.000001f + cnt);
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
                                                        int H, int W, float* __restrict__ out)
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
        out[fpx + (long)gy * W + gx] = sum / (0.000001f + cnt);
    }
}

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

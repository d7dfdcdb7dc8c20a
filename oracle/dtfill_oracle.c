/*
 * dtfill_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * CPU restatement, in plain C, of the reference's distance-transform nearest-neighbour fill path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (distancetransform_depthcompletion_b200) never does.
 *
 * What is restated, with the reference line each function follows:
 *   oracle_chamfer_l1_labels   <- cv2.distanceTransformWithLabels(mask, DIST_L1, 5, DIST_LABEL_PIXEL),
 *                                 the third-party call made at solution_DeepNet/tools.py:9,
 *                                 solution_DeepNet/eval_NYU.py:116, solution_DeepNet/demo.py:80.
 *                                 OpenCV is NOT vendored under /root/reference (pinned there as
 *                                 opencv-contrib-python==3.4.2.16, install_dependency.sh:3-4); the
 *                                 algorithm restated is its published two-pass 5x5 chamfer scan with
 *                                 label propagation.
 *   oracle_nearest_point_f32   <- solution_DeepNet/tools.py:7-10, eval_NYU.py:114-117
 *   oracle_dt_fill_f32         <- solution_DeepNet/tools.py:17-28, eval_NYU.py:120-133
 *
 * Parity pin: the reference ships no tests or golden vectors, so this restatement
 * is pinned against outputs of the reference itself run in the build container: live cv2 4.13.0 and the
 * imported tools.py / AST-extracted eval_NYU.py functions (tests/test_oracle_vs_reference.py, and the
 * committed fixtures under tests/golden/ made by tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Distance given by OpenCV to a pixel no source can reach (probed: a source-free mask yields
 * dt == 65533.0 everywhere and lbl == 0; a 1 x 70000 row saturates at 65533 with label 0). */
#define ORACLE_INIT_DIST 65533u

/* Candidate neighbours in the exact order OpenCV compares them (strict '>' update, first wins ties). */
static const int FWD_DY[8] = {-2, -2, -1, -1, -1, -1, -1, 0};
static const int FWD_DX[8] = {-1, +1, -2, -1, 0, +1, +2, -1};
static const unsigned COST[8] = {3, 3, 3, 2, 1, 2, 3, 1};

/*
 * mask: uint8 [H,W], 0 marks a source pixel (cv2 convention: distance to the nearest zero pixel).
 * dt:   float32 [H,W] exact city-block distance (integer valued), ORACLE_INIT_DIST where unreachable.
 * lbl:  int32 [H,W]; source pixels are numbered 1..N in raster order, every other pixel carries the
 *       number of the source its value was propagated from (0 where unreachable).
 * returns 0, or -1 on allocation failure / bad arguments.
 */
int oracle_chamfer_l1_labels(const uint8_t* mask, int H, int W, float* dt, int32_t* lbl)
{
    if (!mask || !dt || !lbl || H <= 0 || W <= 0) return -1;
    const int B = 2;                       /* border of the padded work arrays */
    const int PW = W + 2 * B, PH = H + 2 * B;
    uint32_t* T = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)PW * PH);
    int32_t* L = (int32_t*)malloc(sizeof(int32_t) * (size_t)PW * PH);
    if (!T || !L) { free(T); free(L); return -1; }
    for (size_t i = 0; i < (size_t)PW * PH; ++i) { T[i] = ORACLE_INIT_DIST; L[i] = 0; }

    /* label initialisation: raster order over the source pixels */
    int32_t next = 1;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (mask[(size_t)y * W + x] == 0) L[(size_t)(y + B) * PW + (x + B)] = next++;

    /* forward raster pass */
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const size_t p = (size_t)(y + B) * PW + (x + B);
            if (mask[(size_t)y * W + x] == 0) { T[p] = 0; continue; }
            uint32_t t0 = ORACLE_INIT_DIST; int32_t l0 = 0;
            for (int k = 0; k < 8; ++k) {
                const size_t q = (size_t)(y + B + FWD_DY[k]) * PW + (x + B + FWD_DX[k]);
                const uint32_t t = T[q] + COST[k];
                if (t0 > t) { t0 = t; l0 = L[q]; }
            }
            T[p] = t0; L[p] = l0;
        }
    }
    /* backward raster pass: the mirrored neighbourhood, own value first */
    for (int y = H - 1; y >= 0; --y) {
        for (int x = W - 1; x >= 0; --x) {
            const size_t p = (size_t)(y + B) * PW + (x + B);
            uint32_t t0 = T[p]; int32_t l0 = L[p];
            if (t0 > 1) {
                for (int k = 0; k < 8; ++k) {
                    const size_t q = (size_t)(y + B - FWD_DY[k]) * PW + (x + B - FWD_DX[k]);
                    const uint32_t t = T[q] + COST[k];
                    if (t0 > t) { t0 = t; l0 = L[q]; }
                }
                T[p] = t0; L[p] = l0;
            }
            dt[(size_t)y * W + x] = (float)t0;
            lbl[(size_t)y * W + x] = l0;
        }
    }
    free(T); free(L);
    return 0;
}

/* tools.py:8 / eval_NYU.py:115 -- value_mask = uint8(1.0 - x > thr), evaluated in float32 like numpy
 * does for a float32 array with Python-float scalars; then the cv2 call of tools.py:9. */
int oracle_nearest_point_f32(const float* x, int H, int W, float thr, float* dt, int32_t* lbl)
{
    if (!x || H <= 0 || W <= 0) return -1;
    uint8_t* m = (uint8_t*)malloc((size_t)H * W);
    if (!m) return -1;
    for (size_t i = 0; i < (size_t)H * W; ++i) {
        volatile float d = 1.0f - x[i];    /* volatile: keep the float32 rounding of the subtraction */
        m[i] = (d > thr) ? 1 : 0;
    }
    const int rc = oracle_chamfer_l1_labels(m, H, W, dt, lbl);
    free(m);
    return rc;
}

/*
 * One frame of DT_complete_batch (tools.py:19-27) / Distance_Transform (eval_NYU.py:122-131):
 *   lbl         = nearest_point(x)[1]                  (source predicate, threshold src_thr)
 *   depth_list  = x[x > val_thr]                       (raster-order compaction of the VALID pixels)
 *   out_depth   = depth_list[lbl - 1]                  (numpy indexing: index -1 is the LAST element)
 * Also returns the distance channel, the label map and the validity mask (x > val_thr) when the
 * pointers are non-NULL.  Return 0; -2 when the frame has no valid pixel (numpy raises IndexError);
 * -3 when a label points past the end of depth_list (numpy raises IndexError as well); -1 on failure.
 */
int oracle_dt_fill_f32(const float* x, int H, int W, float src_thr, float val_thr,
                       float* out_depth, float* out_dt, int32_t* out_lbl, uint8_t* out_mask)
{
    if (!x || !out_depth || H <= 0 || W <= 0) return -1;
    const size_t n = (size_t)H * W;
    float* dt = out_dt ? out_dt : (float*)malloc(sizeof(float) * n);
    int32_t* lbl = out_lbl ? out_lbl : (int32_t*)malloc(sizeof(int32_t) * n);
    float* list = (float*)malloc(sizeof(float) * n);
    int rc = -1;
    if (dt && lbl && list && oracle_nearest_point_f32(x, H, W, src_thr, dt, lbl) == 0) {
        size_t nv = 0;
        for (size_t i = 0; i < n; ++i) {
            const int v = x[i] > val_thr;
            if (out_mask) out_mask[i] = (uint8_t)v;
            if (v) list[nv++] = x[i];
        }
        rc = 0;
        if (nv == 0) rc = -2;
        for (size_t i = 0; i < n && rc == 0; ++i) {
            long idx = (long)lbl[i] - 1;
            if (idx < 0) idx += (long)nv;
            if (idx < 0 || idx >= (long)nv) { rc = -3; break; }
            out_depth[i] = list[idx];
        }
    }
    if (!out_dt) free(dt);
    if (!out_lbl) free(lbl);
    free(list);
    return rc;
}

/* A batch of frames [B,H,W] (channel 0 already selected), frames independent (tools.py:17).
 * first_bad receives the index of the first frame whose fill failed, or -1. */
int oracle_dt_fill_batch_f32(const float* x, int B, int H, int W, float src_thr, float val_thr,
                             float* out_depth, float* out_dt, int32_t* out_lbl, uint8_t* out_mask,
                             int* first_bad)
{
    const size_t n = (size_t)H * W;
    if (first_bad) *first_bad = -1;
    for (int b = 0; b < B; ++b) {
        const int rc = oracle_dt_fill_f32(x + b * n, H, W, src_thr, val_thr, out_depth + b * n,
                                          out_dt ? out_dt + b * n : 0, out_lbl ? out_lbl + b * n : 0,
                                          out_mask ? out_mask + b * n : 0);
        if (rc != 0) { if (first_bad) *first_bad = b; return rc; }
    }
    return 0;
}

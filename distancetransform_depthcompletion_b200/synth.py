"""Seeded synthetic sparse-depth frames shaped like the reference's inputs.

The datasets are not shipped with the reference (depth_selection/ holds a placeholder), so every config of
BASELINE.json runs on frames generated here:

* ``kitti_frame``  -- 352 x 1216, 64-beam scan pattern at ~5 % density, depths on the KITTI grid k/256
  (data_read.py:215 decodes uint16/256) and all >= 1.0 m so the source and valid predicates agree.
* ``keep_beams``   -- the beam sub-sampling rule of subsample_Lidar_train.py:170-177 (keep beams whose
  index is a multiple of 1/keep_ratio): 64 -> 32/16/8 beams.
* ``nyu_frame``    -- 480 x 640 dense depth sampled at 500 random pixels, data_read.py:360-364.
* ``kitti_gt``     -- semi-dense ground truth (float64, data_read.py:223) for the metric sweep.
"""
from __future__ import annotations

import numpy as np

KITTI_H, KITTI_W = 352, 1216
NYU_H, NYU_W = 480, 640


def _kitti_scene(rng: np.random.Generator, H: int, W: int) -> np.ndarray:
    y = np.arange(H, dtype=np.float64)[:, None]
    D = np.full((H, W), 80.0)
    ground = np.clip(1.65 * 721.0 / np.maximum(y - 170.0, 1e-3), 2.0, 80.0)
    D = np.where(y > 175, np.broadcast_to(ground, (H, W)), D)
    for _ in range(6):
        bh, bw = int(rng.integers(30, 110)), int(rng.integers(40, 260))
        y0, x0 = int(rng.integers(110, H - 30)), int(rng.integers(0, W - 40))
        D[y0:y0 + bh, x0:x0 + bw] = np.minimum(D[y0:y0 + bh, x0:x0 + bw], rng.uniform(5.0, 40.0))
    return D


def kitti_beam_rows(H: int = KITTI_H, W: int = KITTI_W, beams: int = 64) -> np.ndarray:
    """Row of every beam at every column: int [beams, W]."""
    x = np.arange(W, dtype=np.float64)
    base = np.linspace(120.0, H - 7.0, beams)[:, None]
    return np.clip(np.rint(base + 3.0 * np.sin(np.pi * x / W)[None, :]), 0, H - 1).astype(np.int64)


def kitti_frame(seed: int, H: int = KITTI_H, W: int = KITTI_W, beam_step: int = 1, keep_prob: float = 0.28,
                return_dense: bool = False):
    """One sparse frame float32 [H,W]; ``beam_step`` 1/2/4/8 keeps 64/32/16/8 beams."""
    rng = np.random.default_rng(seed)
    D = _kitti_scene(rng, H, W)
    rows = kitti_beam_rows(H, W, 64)
    keep = rng.random((64, W)) < keep_prob
    sparse = np.zeros((H, W), np.float32)
    cols = np.arange(W)
    for b in range(0, 64, beam_step):
        c = cols[keep[b]]
        r = rows[b, c]
        sparse[r, c] = (np.rint(D[r, c] * 256.0) / 256.0).astype(np.float32)
    if return_dense:
        return sparse, D
    return sparse


def kitti_batch(seeds, beam_step: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """[B,352,1216,1] float32, the layout DT_complete_batch takes (tools.py:13-19)."""
    seeds = list(seeds)
    if out is None:
        out = np.empty((len(seeds), KITTI_H, KITTI_W, 1), np.float32)
    for i, s in enumerate(seeds):
        out[i, :, :, 0] = kitti_frame(s, beam_step=beam_step)
    return out


def kitti_gt(seed: int, H: int = KITTI_H, W: int = KITTI_W, p: float = 0.2) -> np.ndarray:
    """Semi-dense ground truth float64 [H,W] for frame ``seed`` (0 = no measurement)."""
    rng = np.random.default_rng(seed)
    D = _kitti_scene(rng, H, W)
    rng2 = np.random.default_rng(1_000_003 + seed)
    m = rng2.random((H, W)) < p
    m[:120] = False
    return np.where(m, np.rint(D * 256.0) / 256.0, 0.0)


def nyu_frame(seed: int, H: int = NYU_H, W: int = NYU_W, samples: int = 500, return_dense: bool = False):
    """Dense U[1,10] m depth sampled like data_read.py:360-364 (duplicates collapse)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    a, b, c = rng.uniform(-0.004, 0.004), rng.uniform(-0.004, 0.004), rng.uniform(3.0, 7.0)
    gt = np.clip(c + a * (xx - W / 2) + b * (yy - H / 2) + 0.3 * rng.standard_normal((H, W)), 1.0, 10.0)
    gt = gt.astype(np.float32)
    mask = np.zeros((H, W), np.float32)
    mask[rng.integers(0, H - 12, samples) + 6, rng.integers(0, W - 16, samples) + 8] = 1.0
    depth = gt * mask
    if return_dense:
        return depth, gt
    return depth


def nyu_batch(seeds, H: int = NYU_H, W: int = NYU_W, samples: int = 500) -> np.ndarray:
    seeds = list(seeds)
    out = np.empty((len(seeds), H, W), np.float32)
    for i, s in enumerate(seeds):
        out[i] = nyu_frame(s, H, W, samples)
    return out


def adversarial_frames(H: int = 24, W: int = 40):
    """Named small frames covering SURVEY.md Appendix B quirks; values float32."""
    rng = np.random.default_rng(7)
    out = {}
    out["dense"] = rng.uniform(1.0, 50.0, (H, W)).astype(np.float32)
    f = np.zeros((H, W), np.float32); f[H // 2, W // 2] = 7.5; out["single_source"] = f
    f = np.zeros((H, W), np.float32); f[0, 0] = 3.25; out["corner_source"] = f
    f = np.zeros((H, W), np.float32); f[H - 1, W - 1] = 3.25; out["corner_source_br"] = f
    f = np.zeros((1, W), np.float32); f[0, 3] = 2.0; f[0, W - 2] = 9.0; out["one_row"] = f
    f = np.zeros((H, 1), np.float32); f[2, 0] = 2.0; f[H - 3, 0] = 9.0; out["one_col"] = f
    f = np.zeros((H, W), np.float32); f[::4, :] = rng.uniform(1, 30, (len(range(0, H, 4)), W)); out["row_stripes"] = f
    f = np.zeros((H, W), np.float32); f[:, ::5] = rng.uniform(1, 30, (H, len(range(0, W, 5)))); out["col_stripes"] = f
    f = np.zeros((H, W), np.float32)
    cb = (np.add.outer(np.arange(H), np.arange(W)) % 2) == 0
    f[cb] = rng.uniform(1, 30, int(cb.sum())); out["checkerboard"] = f
    # valid-but-not-source depths in (0.1, 0.9): every later label reads a shifted list entry
    f = (rng.random((H, W)) < 0.1) * rng.uniform(1.0, 30.0, (H, W)); f = f.astype(np.float32)
    f[3, 5] = 0.5; f[10, 20] = 0.75; out["valid_not_source"] = f
    # exactly 0.9f is NOT a source (1 - 0.9f = 0.100000024 > 0.1f), its float32 successor is
    f = (rng.random((H, W)) < 0.05) * rng.uniform(1.0, 30.0, (H, W)); f = f.astype(np.float32)
    f[5, 5] = np.float32(0.9); f[6, 9] = np.nextafter(np.float32(0.9), np.float32(1.0)); out["exact_0p9"] = f
    # valid pixels but no source at all: lbl == 0 -> whole frame filled with the LAST valid depth
    f = np.zeros((H, W), np.float32); f[2, 3] = 0.5; f[7, 30] = 0.25; out["valid_no_source"] = f
    # duplicates of the same depth value, clustered sources (many exact ties)
    f = np.zeros((H, W), np.float32); f[4:8, 10:14] = 5.0; f[15, 2:38:3] = 5.0; out["duplicates"] = f
    return out

"""Tile statistics of the planner for the bench workload: how much of K2's row-step work is redundancy (halos,
column overlap, padded lanes).  Run on the GPU box: python profiles/tile_stats.py [workload [pipeline_depth [band_cap]]]"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from distancetransform_depthcompletion_b200 import synth
from distancetransform_depthcompletion_b200.engine import DTFillEngine

wl = sys.argv[1] if len(sys.argv) > 1 else "kitti64"
step = {"kitti64": 1, "kitti32": 2, "kitti16": 4, "kitti8": 8}[wl]
B = 256
x = np.stack([synth.kitti_frame(i, beam_step=step) for i in range(32)])
x = np.concatenate([x] * (B // 32))
eng = DTFillEngine(0, pipeline_depth=int(sys.argv[2]) if len(sys.argv) > 2 else 3)
if len(sys.argv) > 3:
    eng.handle.set_band_cap(int(sys.argv[3]))
xd = torch.from_numpy(x).cuda()
eng.fill(xd); eng.flush(); eng.status()
t = eng.handle.debug_tasks(1 << 16)
F = {n: t[:, i].astype(np.int64) for i, n in enumerate(eng.handle.TASK_FIELDS)}
H, W = x.shape[1:]
kinds = np.unique(F["kind"], return_counts=True)
print("tasks", len(t), "per frame", len(t) / B, "kinds", dict(zip(*[k.tolist() for k in kinds])))
ppl = np.where(F["kind"] == 4, 20, 38)
width = 32 * ppl
out_px = ((F["r1"] - F["r0"]) * (F["c1"] - F["c0"])).sum()
fwd_rows = F["hi"] - np.maximum(F["fstart"], F["lo"])
bwd_rows = F["hi"] - F["r0"]
print("output px / frame px", out_px / (B * H * W))
print("fwd row-steps", fwd_rows.sum(), "bwd row-steps", bwd_rows.sum(), "out rows", (F["r1"] - F["r0"]).sum())
print("lane-px processed fwd", (fwd_rows * width).sum() / out_px, "bwd", (bwd_rows * width).sum() / out_px)
print("row redundancy fwd", (fwd_rows * width).sum() / ((F["r1"] - F["r0"]) * width).sum(),
      "bwd", (bwd_rows * width).sum() / ((F["r1"] - F["r0"]) * width).sum())
print("col redundancy", ((F["r1"] - F["r0"]) * width).sum() / out_px)
for k in np.unique(F["kind"]):
    m = F["kind"] == k
    print("kind", k, "n", m.sum(), "mean out rows", (F["r1"] - F["r0"])[m].mean(), "mean fwd", fwd_rows[m].mean(),
          "mean bwd", bwd_rows[m].mean(), "mean out cols", (F["c1"] - F["c0"])[m].mean())
f0 = t[F["frame"] == 0]
for r in f0[np.lexsort((f0[:, 8], f0[:, 3]))]:
    print(dict(zip(eng.handle.TASK_FIELDS[:11], r.tolist()[:11])))

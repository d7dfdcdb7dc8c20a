"""Which stage bounds the pipelined step?  After complete warm-up runs (the workspace then holds valid bit rows, tasks and
scratch), time steps in which only some stages are launched (dtfill_debug_set_skip), at the same pipeline depth: each
stage's saturated throughput next to copies of itself, and pairs of stages next to each other.
Run on the GPU box:  python profiles/stage_probe.py [--depth 4] [--steps 40]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200.engine import DTFillEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--cap", type=int, default=-1)
ap.add_argument("--workload", default="kitti64")
a = ap.parse_args()
_, H, W, src_thr, _ = bench.WORKLOADS[a.workload]
x = torch.from_numpy(bench.make_frames(a.batch, 0, a.workload)).cuda()
eng = DTFillEngine(0, pipeline_depth=a.depth)
eng.handle.set_band_cap(a.cap)
outs = [None] * a.depth
for i in range(2 * a.depth):
    outs[i % a.depth] = eng.fill(x, src_thr=src_thr, out=outs[i % a.depth])
eng.flush(); eng.status()
K1, K1B, K2, SKY = 1, 2, 4, 8
ALL = 15
cases = [("all", 0), ("k2 only", ALL & ~K2), ("k1 only", ALL & ~K1), ("sky only", ALL & ~SKY),
         ("k1+k1b", K2 | SKY), ("k2+sky", K1 | K1B), ("k1+k1b+k2", SKY), ("all again", 0)]
for name, skip in cases:
    eng.handle.debug_set_skip(skip)
    for i in range(a.depth):
        eng.fill(x, src_thr=src_thr, out=outs[i % a.depth])
    eng.flush(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        eng.fill(x, src_thr=src_thr, out=outs[i % a.depth])
    eng.flush(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"stages": name, "ms_per_step": round(e0.elapsed_time(e1) / a.steps, 4)}), flush=True)
eng.handle.debug_set_skip(0)

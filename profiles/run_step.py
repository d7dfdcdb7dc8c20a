"""Minimal driver used for the ncu captures: W warm-up steps + K steps of the device-resident hot path
(batch of 256 synthetic KITTI frames), nothing else.  Same kernels, arguments and batch as bench.py's timed
region.   python profiles/run_step.py [--steps K] [--warmup W] [--batch B] [--band-cap C] [--pipeline D]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200.engine import DTFillEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--band-cap", type=int, default=None)
ap.add_argument("--pipeline", type=int, default=1, help="batches in flight (3 = bench.py's headline mode, k3_sky on)")
a = ap.parse_args()
x = torch.from_numpy(bench.make_frames(a.batch, 0)).cuda()
eng = DTFillEngine(0, pipeline_depth=a.pipeline)
if a.band_cap is not None:
    eng.handle.set_band_cap(a.band_cap)
outs = [None] * max(1, a.pipeline)
for i in range(a.warmup + a.steps):
    out = outs[i % len(outs)] = eng.fill(x, out=outs[i % len(outs)])
eng.flush()
bad, launches = eng.status()
torch.cuda.synchronize()
print("ok", bad, launches, float(out["depth"].sum()))

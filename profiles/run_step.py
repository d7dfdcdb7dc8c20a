"""Minimal driver used for the ncu captures: W warm-up steps + K steps of the device-resident hot path
(batch of 256 synthetic KITTI frames), nothing else.  Same kernels, arguments and batch as bench.py's timed
region.   python profiles/run_step.py [--steps K] [--warmup W] [--batch B] [--band-cap C]"""
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
a = ap.parse_args()
x = torch.from_numpy(bench.make_frames(a.batch, 0)).cuda()
eng = DTFillEngine(0)
if a.band_cap is not None:
    eng.handle.set_band_cap(a.band_cap)
out = None
for _ in range(a.warmup + a.steps):
    out = eng.fill(x, out=out)
bad, launches = eng.status()
torch.cuda.synchronize()
print("ok", bad, launches, float(out["depth"].sum()))

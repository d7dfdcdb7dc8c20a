"""Minimal driver for the ncu captures of the Euclidean feature transform (dtfill_edt, extension f-4): W warm-up calls +
K calls on a device-resident batch of 256 synthetic KITTI frames.   python profiles/run_edt.py [--steps K] [--warmup W]"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200 import _lib  # noqa: E402
from distancetransform_depthcompletion_b200.engine import DTFillEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--workload", default="kitti64")
a = ap.parse_args()
_, H, W, thr, _ = bench.WORKLOADS[a.workload]
x = torch.from_numpy(bench.make_frames(a.batch, 0, a.workload)).cuda()
eng = DTFillEngine(0)
eng._bind_stream()
L = _lib.load()
d2 = torch.empty((a.batch, H, W), dtype=torch.int32, device="cuda")
idx = torch.empty((a.batch, H, W), dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.warmup + a.steps):
    if i == a.warmup:
        torch.cuda.synchronize(); e0.record()
    _lib._check(L.dtfill_edt(eng.handle._h, ctypes.c_void_p(x.data_ptr()), 1, a.batch, H, W, ctypes.c_float(thr),
                             ctypes.c_void_p(d2.data_ptr()), ctypes.c_void_p(idx.data_ptr()), 1), "edt")
e1.record(); torch.cuda.synchronize()
print("edt %s: %.3f ms per call of %d frames" % (a.workload, e0.elapsed_time(e1) / a.steps, a.batch))

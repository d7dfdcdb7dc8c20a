"""End-to-end probe of the drop-in call on the GPU box: tools.DT_complete_batch(pageable numpy) for several staging
thread counts, and the pieces it is made of.   python profiles/e2e_probe.py [--batch 256] [--reps 6]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200 import _lib, tools  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--threads", default="-1,2,4,8,12")
a = ap.parse_args()
x = bench.make_frames(a.batch, 0)[:, :, :, None]
h = _lib.get_handle(0)
for thr in [int(v) for v in a.threads.split(",")]:
    h.set_stage_threads(thr)
    tools.DT_complete_batch(x)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        r = tools.DT_complete_batch(x)
    dt = (time.perf_counter() - t0) / a.reps
    print(f"stage threads {thr:3d}: {dt * 1e3:7.2f} ms per {a.batch} frames = {a.batch / dt:8.0f} frames/s "
          f"({x.nbytes / dt / 1e9:.1f} GB/s each way)", flush=True)
# pieces: a plain memcpy of the input (one thread), np.empty + first touch of an output
t0 = time.perf_counter(); y = x.copy(); t1 = time.perf_counter()
print(f"numpy copy of the input (1 thread, fresh destination): {(t1 - t0) * 1e3:.1f} ms")
t0 = time.perf_counter(); y[...] = x; t1 = time.perf_counter()
print(f"numpy copy of the input (1 thread, touched destination): {(t1 - t0) * 1e3:.1f} ms")

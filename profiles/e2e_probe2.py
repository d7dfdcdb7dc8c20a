"""A/B of the drop-in call's host-side knobs after a complete warm-up (the pool of page-locked output arrays is filled
first): staging threads x slices per call, median and best of 10 calls each, two rounds."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from distancetransform_depthcompletion_b200 import _lib, tools
x = bench.make_frames(256, 0)[:, :, :, None]
h = _lib.get_handle(0)
keep = [tools.DT_complete_batch(x) for _ in range(3)]      # three pooled output arrays exist from now on
del keep
for rnd in range(2):
    for nsub in (16, 32):
        for thr in (8, 12, 16):
            h.set_stage_threads(thr); h.set_subbatches(nsub)
            for _ in range(3): r = tools.DT_complete_batch(x)
            ts = []
            for _ in range(10):
                t0 = time.perf_counter(); r = tools.DT_complete_batch(x); ts.append(time.perf_counter() - t0)
            print(f"round {rnd} slices {nsub:2d} threads {thr:2d}: median {statistics.median(ts)*1e3:6.2f} ms best {min(ts)*1e3:6.2f} ms "
                  f"= {256/statistics.median(ts):7.0f} frames/s", flush=True)

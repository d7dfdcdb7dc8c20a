"""The drop-in call end to end (tools.DT_complete_batch, pageable numpy in, new array out) after a complete warm-up:
median and best of 12 calls for a given number of staging threads; the compaction's instruction set comes from
DTFILL_COMPACT_ISA (read once per process).  Also prints what the host is.
    python profiles/e2e_probe3.py [threads ...]"""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from distancetransform_depthcompletion_b200 import _lib, tools
x = bench.make_frames(256, 0)[:, :, :, None]
h = _lib.get_handle(0)
keep = [tools.DT_complete_batch(x) for _ in range(3)]
del keep
for thr in [int(v) for v in sys.argv[1:]] or [12]:
    h.set_stage_threads(thr)
    for _ in range(3): r = tools.DT_complete_batch(x)
    ts = []
    for _ in range(12):
        t0 = time.perf_counter(); r = tools.DT_complete_batch(x); ts.append(time.perf_counter() - t0)
    print(f"isa {os.environ.get('DTFILL_COMPACT_ISA', 'widest')} threads {thr:2d}: median {statistics.median(ts)*1e3:6.2f} ms best {min(ts)*1e3:6.2f} ms "
          f"= {256/statistics.median(ts):7.0f} frames/s", flush=True)

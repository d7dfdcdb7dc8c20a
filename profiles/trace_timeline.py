"""Timeline of the kernels of consecutive pipelined calls, from a -DDTFILL_TRACE build (profiles/build_variant.sh trace
-DDTFILL_TRACE):  DTFILL_LIB=_variants/libdtfill_trace.so python profiles/trace_timeline.py [--depth 4] [--calls 16]
Per call and kernel: first block start, last block end (us, relative to the first call shown), mean block time, blocks."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200 import _lib  # noqa: E402
from distancetransform_depthcompletion_b200.engine import DTFillEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--calls", type=int, default=16)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--cap", type=int, default=-1)
a = ap.parse_args()
x = torch.from_numpy(bench.make_frames(a.batch, 0)).cuda()
eng = DTFillEngine(0, pipeline_depth=a.depth)
eng.handle.set_band_cap(a.cap)
outs = [None] * a.depth
for i in range(3 * a.depth):
    outs[i % a.depth] = eng.fill(x, out=outs[i % a.depth])
eng.flush(); eng.status()
L = _lib.load()
L.dtfill_debug_trace_arm.argtypes = [ctypes.c_void_p]
L.dtfill_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert L.dtfill_debug_trace_arm(eng.handle._h) == 0
for i in range(a.calls):
    outs[i % a.depth] = eng.fill(x, out=outs[i % a.depth])
eng.flush(); torch.cuda.synchronize()
buf = np.zeros((256, 4, 4), dtype=np.uint64)
assert L.dtfill_debug_trace_read(eng.handle._h, buf.ctypes.data) == 0
t0 = min(int(buf[c, k, 0]) for c in range(a.calls) for k in range(4) if buf[c, k, 3] > 0)
names = ["k1", "k1b", "k2", "sky"]
for c in range(a.calls):
    parts = []
    for k in range(4):
        st, en, sm, n = (int(v) for v in buf[c, k])
        if n == 0:
            continue
        parts.append(f"{names[k]} [{(st - t0) / 1e3:7.1f},{(en - t0) / 1e3:7.1f}] n={n} mean={sm / n / 1e3:6.1f}us")
    print(f"call {c:2d}: " + " | ".join(parts))

"""Host-only: throughput of the sparse upload's compaction (dtfill_debug_compact, one thread) on KITTI-shaped frames.
    python profiles/compact_probe.py [frames]      -> GB/s of input read per thread; DTFILL_COMPACT_ISA=avx2 forces the AVX2 path"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from distancetransform_depthcompletion_b200 import _lib, synth  # noqa: E402

L = _lib.load()
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = np.stack([synth.kitti_frame(i) for i in range(nf)])
n = x.size
idx = np.empty(n + 64, np.uint32)
val = np.empty(n + 64, np.uint32)
vp = ctypes.c_void_p
best = 1e9
for rep in range(7):
    t0 = time.perf_counter()
    k = 0
    for f in range(nf):                 # the product calls it on blocks of 32768 pixels; a frame at a time is close enough
        k += L.dtfill_debug_compact(x[f].ctypes.data_as(vp), x[f].size, 0.1, 0.1, idx.ctypes.data_as(vp), val.ctypes.data_as(vp), x[f].size + 64)
    best = min(best, time.perf_counter() - t0)
print(f"{nf} frames, {k} pairs in the last frame set, {n * 4 / best / 1e9:.2f} GB/s per thread, kept {k / n * 100:.2f} % of the pixels")

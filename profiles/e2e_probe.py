"""e2e (host pinned in -> host pinned out) timing of dtfill_run for several sub-batch counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from distancetransform_depthcompletion_b200 import _lib
B, H, W = 256, 352, 1216
x = bench.make_frames(B, 0)
pin_in = _lib.pinned_empty((B, H, W), np.float32); pin_in[...] = x
out = dict(depth=_lib.pinned_empty((B, H, W), np.float32), dt=_lib.pinned_empty((B, H, W), np.float32),
           mask=_lib.pinned_empty((B, H, W), np.uint8))
h = _lib.Handle(0)
for ns in (1, 2, 4, 8, 16):
    h.set_subbatches(ns)
    h.run_host(pin_in, 0.1, 0.1, want_dt=True, want_mask=True, out=out)
    t0 = time.perf_counter()
    for _ in range(4):
        h.run_host(pin_in, 0.1, 0.1, want_dt=True, want_mask=True, out=out)
    dt = (time.perf_counter() - t0) / 4
    print("nsub", ns, "%.2f ms  %.0f frames/s" % (dt * 1e3, B / dt))
# only depth out
for ns in (1, 8):
    h.set_subbatches(ns)
    h.run_host(pin_in, 0.1, 0.1, out=out)
    t0 = time.perf_counter()
    for _ in range(4):
        h.run_host(pin_in, 0.1, 0.1, out=out, want_counts=False)
    dt = (time.perf_counter() - t0) / 4
    print("depth only nsub", ns, "%.2f ms  %.0f frames/s" % (dt * 1e3, B / dt))
for kw, name in ((dict(want_dt=True), "depth+dt"), (dict(want_mask=True), "depth+mask")):
    for ns in (1, 8):
        h.set_subbatches(ns)
        h.run_host(pin_in, 0.1, 0.1, out=out, want_counts=False, **kw)
        t0 = time.perf_counter()
        for _ in range(4):
            h.run_host(pin_in, 0.1, 0.1, out=out, want_counts=False, **kw)
        dt = (time.perf_counter() - t0) / 4
        print(name, "nsub", ns, "%.2f ms  %.0f frames/s" % (dt * 1e3, B / dt))
import torch
n1, n2 = 438304768, 986185728
h1 = torch.empty(n1, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n2, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n1, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n2, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
both(); torch.cuda.synchronize()
t0 = time.perf_counter(); both(); both(); both(); torch.cuda.synchronize(); dtm = (time.perf_counter() - t0) / 3
print("torch asymmetric duplex: %.2f ms  (H2D %.0f MB, D2H %.0f MB)" % (dtm * 1e3, n1 / 1e6, n2 / 1e6))

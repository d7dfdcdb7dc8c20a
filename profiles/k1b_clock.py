"""Phase clocks of k1b_scan_compact's block 0 (cycles since the block's start) from a -DDTFILL_PLANNER_CLOCKS build:
    bash profiles/build_variant.sh clocks -DDTFILL_PLANNER_CLOCKS; DTFILL_LIB=_variants/libdtfill_clocks.so python profiles/k1b_clock.py [frames]"""
import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo')
import bench
from distancetransform_depthcompletion_b200 import _lib
import torch
from distancetransform_depthcompletion_b200.engine import DTFillEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.from_numpy(bench.make_frames(B, 0)).cuda()
eng = DTFillEngine(0)
for _ in range(4): out = eng.fill(x)
eng.status()
L = _lib.load(); L.dtfill_debug_read_status.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(64, np.int32)
L.dtfill_debug_read_status(eng.handle._h, buf.ctypes.data, 64)
print(B, "frames; compaction warps (cycles): scans done %d, depth_list done %d" % (buf[4], buf[5]))
print("planner warps: marks 0..6 (scans, -, cells, vertical sweeps, horizontal sweeps, greedy bands, tasks written):", buf[12:19].tolist())

import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo')
import bench
from distancetransform_depthcompletion_b200 import _lib
import torch
from distancetransform_depthcompletion_b200.engine import DTFillEngine
x = torch.from_numpy(bench.make_frames(256, 0)).cuda()
eng = DTFillEngine(0)
for _ in range(4): out = eng.fill(x)
eng.status()
L = _lib.load(); L.dtfill_debug_read_status.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(64, np.int32)
L.dtfill_debug_read_status(eng.handle._h, buf.ctypes.data, 64)
print("compaction warps (cycles): scan_done %d, compaction_done %d" % (buf[4], buf[5]))
print("planner: scan_done %d occw %d cells %d vertical %d horizontal %d greedy %d written %d" % tuple(buf[12:19]))

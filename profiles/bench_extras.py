"""Device-resident timings of the widened rows and of the metrics kernel, batch of 256 KITTI
frames, CUDA events, 10 repetitions after 3 warm-ups."""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distancetransform_depthcompletion_b200 import _lib, synth
from distancetransform_depthcompletion_b200.engine import DTFillEngine

B, H, W = 256, 352, 1216
x = torch.from_numpy(bench.make_frames(B, 0)).cuda()
eng = DTFillEngine(0)
h = eng.handle
L = _lib.load()

def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

npx = B * H * W
mask = (x > 0.1).float()
data = (x / 90.0 * mask).contiguous()
pool_out = torch.empty((3, B, H, W), dtype=torch.float32, device="cuda")
eng._bind_stream()
t = timed(lambda: h.dt_pool(data.data_ptr(), mask.data_ptr(), B, H, W, 7, 4, on_device=True, out_ptr=pool_out.data_ptr()))
print("dt_pool 3 levels: %.3f ms / 256 frames  (%.0f frames/s, %.0f GB/s of 8 B/px/level)" % (t, B / t * 1e3, 3 * 8 * npx / t / 1e6))
out = torch.empty_like(x)
t = timed(lambda: _lib._check(L.dtfill_outlier_removal(h._h, ctypes.c_void_p(x.data_ptr()), 1, B, H, W, ctypes.c_void_p(out.data_ptr()), 1), "outlier"))
print("outlier_removal: %.3f ms / 256 frames  (%.0f frames/s, %.0f GB/s of 8 B/px)" % (t, B / t * 1e3, 8 * npx / t / 1e6))
fill = eng.fill(x)
gt = torch.from_numpy(np.stack([synth.kitti_gt(i % 16) for i in range(B)])).cuda()
t = timed(lambda: eng.metrics(fill["depth"], gt))
print("metrics (Result.evaluate, f64 gt): %.3f ms / 256 frames (%.0f GB/s of 12 B/px)" % (t, 12 * npx / t / 1e6))

# ---- f-3: uint16 PNG samples in (decode + 96-row crop inside K1), pipelined like bench.py's headline -----------
Hin = H + 96
png = torch.zeros((B, Hin, W), dtype=torch.int32)
png[:, 96:] = torch.round(x.cpu() * 256).to(torch.int32)
png = png.to(torch.uint16).cuda() if hasattr(torch, "uint16") else None
sets = [dict(lidar=torch.empty((B, H, W), device="cuda"), depth=torch.empty((B, H, W), device="cuda"),
             dt=torch.empty((B, H, W), device="cuda"), mask=torch.empty((B, H, W), dtype=torch.uint8, device="cuda"))
        for _ in range(3)]
if png is not None:
    h.set_pipeline_depth(3)
    for want_lidar in (0, 1):
        it = [0]
        def step():
            o = sets[it[0] % 3]; it[0] += 1
            h.run_device_u16_async(png.data_ptr(), B, Hin, W, 96, 0.1, 0.1, o["depth"].data_ptr(),
                                   o["lidar"].data_ptr() if want_lidar else None, o["dt"].data_ptr(), None,
                                   o["mask"].data_ptr(), None)
        for _ in range(6): step()
        h.flush(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(24): step()
        h.flush(); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 24
        bpp = 2 + 4 + 4 + 1 + 4 * want_lidar
        print("uint16 input%s, pipelined: %.4f ms / 256 frames (%.0f frames/s, %.0f GB/s of %d B/px)"
              % (" + decoded lidar out" if want_lidar else "", t, B / t * 1e3, bpp * npx / t / 1e6, bpp))
    h.set_pipeline_depth(1)

# ---- f-4: exact Euclidean feature transform (extension), device-resident: squared distance + nearest-source index --
d2 = torch.empty((B, H, W), dtype=torch.int32, device="cuda")
idx = torch.empty((B, H, W), dtype=torch.int32, device="cuda")
for want_idx in (0, 1):
    t = timed(lambda: _lib._check(L.dtfill_edt(h._h, ctypes.c_void_p(x.data_ptr()), 1, B, H, W, ctypes.c_float(0.1),
                                               ctypes.c_void_p(d2.data_ptr()),
                                               ctypes.c_void_p(idx.data_ptr()) if want_idx else None, 1), "edt"))
    bpp = 4 + 4 + 4 * want_idx
    print("edt (exact Euclidean, d2%s): %.3f ms / 256 frames (%.0f frames/s, %.0f GB/s of %d B/px)"
          % (" + index" if want_idx else "", t, B / t * 1e3, bpp * npx / t / 1e6, bpp))

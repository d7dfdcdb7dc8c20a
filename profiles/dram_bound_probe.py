"""Is the pipelined step bound by DRAM traffic or by instruction issue?  Same kernels, same scan work, fewer output
bytes: time the step with (a) depth+dt+mask, (b) depth+mask, (c) depth only.  If the time follows the bytes the step
is DRAM-bound.  Run on the GPU box: python profiles/dram_bound_probe.py [sky_min]"""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
from distancetransform_depthcompletion_b200 import synth, _lib

B, H, W = 256, 352, 1216
x = np.stack([synth.kitti_frame(i) for i in range(32)])
x = np.concatenate([x] * (B // 32))
h = _lib.Handle(0)
if len(sys.argv) > 1:
    h.set_sky_min(int(sys.argv[1]))
xs = [torch.from_numpy(x).cuda() for _ in range(3)]
outs = [dict(depth=torch.empty((B, H, W), device="cuda"), dt=torch.empty((B, H, W), device="cuda"),
             mask=torch.empty((B, H, W), dtype=torch.uint8, device="cuda")) for _ in range(3)]
stream = torch.cuda.current_stream().cuda_stream or 0x1
h.set_stream(stream)
for depth in (3, 1):
    h.set_pipeline_depth(depth)
    for name, want_dt, want_mask in (("depth+dt+mask", 1, 1), ("depth+mask", 0, 1), ("depth", 0, 0)):
        def step(i):
            o = outs[i % 3]
            h.run_device_async(xs[i % 3].data_ptr(), B, H, W, 0.1, 0.1, o["depth"].data_ptr(),
                               o["dt"].data_ptr() if want_dt else None, None,
                               o["mask"].data_ptr() if want_mask else None, None)
        for i in range(6):
            step(i)
        h.flush(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 30
        for i in range(K):
            step(i)
        h.flush()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        nbytes = B * H * W * (4 + 4 + 4 * want_dt + want_mask)
        print(f"pipeline {depth} {name:14s} {ms:.4f} ms/step  algorithmic {nbytes/1e6:.0f} MB -> {nbytes/ms/1e6:.0f} GB/s")

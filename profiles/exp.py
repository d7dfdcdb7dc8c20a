"""Experiment driver for tuning the scan: times strict and pipelined steps for a grid of planner settings.
Run on the GPU box:  python profiles/exp.py [--caps 96,120,...] [--sky 0,8] [--depths 1,4] [--steps 30] [--check]
Prints one line per setting: ms/step, frames/s, fraction of the measured HBM roofline, task count."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200.engine import DTFillEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--caps", default="-1")
ap.add_argument("--sky", default="-1")
ap.add_argument("--depths", default="1,4")
ap.add_argument("--nsub", default="-1")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--workload", default="kitti64")
ap.add_argument("--check", action="store_true", help="compare 8 frames of every setting with the oracle")
ap.add_argument("--kernels", action="store_true", help="per-kernel CUDA-event times (strict order)")
a = ap.parse_args()
_, H, W, src_thr, _ = bench.WORKLOADS[a.workload]
peak, _ = bench.measured_peak()
x_np = bench.make_frames(a.batch, 0, a.workload)
x = torch.from_numpy(x_np).cuda()
ref = None
if a.check:
    from oracle import oracle as O
    ref = O.dt_fill(x_np[:8], src_thr=src_thr)
for depth in [int(v) for v in a.depths.split(",")]:
    for sky in [int(v) for v in a.sky.split(",")]:
      for nsub in [int(v) for v in a.nsub.split(",")]:
        for cap in [int(v) for v in a.caps.split(",")]:
            eng = DTFillEngine(0, pipeline_depth=depth)
            eng.handle.set_subbatches(nsub)
            eng.handle.set_band_cap(cap)
            eng.handle.set_sky_min(sky)
            outs = [None] * max(1, depth)
            for i in range(a.warmup):
                outs[i % len(outs)] = eng.fill(x, src_thr=src_thr, out=outs[i % len(outs)])
            eng.flush(); bad, launches = eng.status()
            ntasks = len(eng.handle.debug_tasks(1 << 17))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                outs[i % len(outs)] = eng.fill(x, src_thr=src_thr, out=outs[i % len(outs)])
            eng.flush()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            line = dict(depth=depth, sky=sky, nsub=nsub, cap=cap, ms=round(ms, 4), fps=round(a.batch / ms * 1e3),
                        frac=round(13 * a.batch * H * W / (ms * 1e-3) / 1e9 / peak, 4), tasks=ntasks, launches=launches)
            if a.kernels:
                eng.handle.set_pipeline_depth(1)
                eng.handle.set_profiling(True)
                if sky < 0 and depth > 1:
                    eng.handle.set_sky_min(8)
                kt = {}
                for _ in range(5):
                    eng.fill(x, src_thr=src_thr, out=outs[0])
                    for k, v in eng.handle.kernel_times().items():
                        kt[k] = kt.get(k, 0.0) + v / 5
                eng.handle.set_profiling(False)
                line["kernel_ms"] = {k: round(v, 4) for k, v in kt.items()}
            if ref is not None:
                o = outs[(a.steps - 1) % len(outs)]
                ok = all(np.array_equal(o[k][:8].cpu().numpy(), ref[k]) for k in ("depth", "dt", "mask"))
                line["parity8"] = bool(ok)
            print(json.dumps(line), flush=True)
            del eng

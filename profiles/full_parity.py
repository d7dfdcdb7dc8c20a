"""Whole-batch parity of a bench workload against the oracle, strict and pipelined, every frame (not a sample).
python profiles/full_parity.py --workload kitti16 [--batch 256] [--depth 4]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from distancetransform_depthcompletion_b200.engine import DTFillEngine
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="kitti16")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--calls", type=int, default=6)
a = ap.parse_args()
_, H, W, thr, defb = bench.WORKLOADS[a.workload]
B = a.batch or defb
frames = bench.make_frames(B, 0, a.workload)
want = O.dt_fill(frames, src_thr=thr)
x = torch.from_numpy(frames).cuda()
for depth in (1, a.depth):
    eng = DTFillEngine(0, pipeline_depth=depth)
    outs = [None] * max(1, depth)
    for i in range(a.calls):
        outs[i % len(outs)] = eng.fill(x, src_thr=thr, want_lbl=(i % 2 == 1), out=outs[i % len(outs)])
    eng.flush(); bad, _ = eng.status()
    tasks = eng.handle.debug_tasks(1 << 17)
    for j, o in enumerate(outs):
        for k in ("depth", "dt", "mask"):
            got = o[k].cpu().numpy()
            neq = got != want[k]
            if neq.any():
                fr = np.unique(np.nonzero(neq)[0])
                f0 = fr[0]
                ys, xs = np.nonzero(neq[f0])
                print(f"depth {depth} set {j} {k}: {neq.sum()} px differ in frames {fr[:10].tolist()} (n={len(fr)}); frame {f0}: rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}")
                t = tasks[tasks[:, 0] == f0]
                for r in t: print("   task", dict(zip(eng.handle.TASK_FIELDS, r.tolist())))
                break
        else:
            print(f"depth {depth} set {j}: all {B} frames identical to the oracle")
    del eng

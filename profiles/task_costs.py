import sys, numpy as np
sys.path.insert(0,'/root/repo')
import torch, bench
from distancetransform_depthcompletion_b200.engine import DTFillEngine
x = torch.from_numpy(bench.make_frames(256, 0)).cuda()
eng = DTFillEngine(0, pipeline_depth=1)
eng.fill(x); eng.flush(); eng.status()
t = eng.handle.debug_tasks(1 << 17)
F = {n: t[:, i].astype(np.int64) for i, n in enumerate(eng.handle.TASK_FIELDS)}
cost = (F["hi"] - np.maximum(F["fstart"], F["lo"])) + (F["hi"] - F["r0"])
print("tasks", len(t), "cost min/mean/max", cost.min(), cost.mean(), cost.max())
print("hist", np.histogram(cost, bins=[0,40,60,80,100,120,140,160,180,200,250])[0])
f0 = t[F["frame"] == 0]
for r in f0[np.lexsort((f0[:, 8], f0[:, 3]))]:
    d = dict(zip(eng.handle.TASK_FIELDS[:11], r.tolist()[:11]))
    print(d, "cost", (d["hi"]-max(d["fstart"],d["lo"]))+(d["hi"]-d["r0"]))

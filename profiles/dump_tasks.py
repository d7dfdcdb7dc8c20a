import sys, numpy as np
sys.path.insert(0,'/root/repo')
from distancetransform_depthcompletion_b200 import _lib, synth
h=_lib.get_handle()
for cap in (-1,100,150):
    h.set_band_cap(cap)
    x=np.stack([synth.kitti_frame(i) for i in range(2)])
    h.run_host(x,0.1,0.1)
    t=h.debug_tasks()
    print("cap",cap,"tasks/frame",len(t)/2)
    for r in t[t[:,0]==0]: print(dict(zip(h.TASK_FIELDS,r.tolist())))

"""World-size-2 run of the multi-GPU host logic on CPU (gloo): contiguous frame shards + one all-reduce of the
metric totals gives the same means as a single process.  Per-frame metrics come from the
oracle here because no GPU is available; the exchanged vector has the layout libdtfill's dtfill_metrics writes."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distancetransform_depthcompletion_b200 import sharding, synth
from oracle import oracle as O

N_FRAMES = 5


def _per_frame_sums(begin, end):
    s = np.zeros(10)
    for i in range(begin, end):
        x = synth.kitti_frame(i)[150:190, :256]
        gt = synth.kitti_gt(i)[150:190, :256]
        m = O.result_kitti(O.dt_fill(np.ascontiguousarray(x))["depth"], np.ascontiguousarray(gt))
        s[:5] += [m["mse"], m["rmse"], m["mae"], m["irmse"], m["imae"]]
        s[8] += m["count"]
        s[9] += 1
    return s


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(N_FRAMES, rank, world)
    t = torch.from_numpy(_per_frame_sums(b, e))
    sharding.allreduce_sums(t)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _per_frame_sums(0, N_FRAMES)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    m = sharding.finalize_means(got)
    assert m["frames"] == N_FRAMES and np.isfinite(m["rmse"])

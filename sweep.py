#!/usr/bin/env python
"""sweep.py -- BASELINE.json configs[4]: N synthetic KITTI frames sharded over the GPUs of one box, DT + NN fill and
evaluation.py's masked RMSE/MAE/iRMSE/iMAE per frame on the GPU, the per-frame metrics summed and all-reduced over
NCCL (the mean-of-per-frame-metrics rule of eval.py:212-232 / the ablation notebook).

    python sweep.py --frames 8192                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        sweep.py --frames 8192                                      8 GPUs, one process each

Frames are independent (tools.py:17-28): rank r takes the contiguous range sharding.shard_range(frames, r, world),
in batches of --batch frames that stay in HBM; the only exchange is one all_reduce(sum) of 10 doubles.
Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 352, 1216


def frame_pool(n: int, seed0: int):
    """n distinct sparse frames with their semi-dense ground truth (float64), generated once per rank."""
    from distancetransform_depthcompletion_b200 import synth
    x = np.stack([synth.kitti_frame(seed0 + i) for i in range(n)])
    g = np.stack([synth.kitti_gt(seed0 + i) for i in range(n)])
    return x, g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--pool", type=int, default=32, help="distinct synthetic frames per rank (others are column rolls)")
    args = ap.parse_args()
    import torch
    from distancetransform_depthcompletion_b200 import _lib, sharding
    from distancetransform_depthcompletion_b200.engine import DTFillEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    begin, end = sharding.shard_range(args.frames, rank, world)
    x_np, g_np = frame_pool(args.pool, 0)        # same pool on every rank: frame i is defined by its global index
    xp = torch.from_numpy(x_np).to(dev)
    gp = torch.from_numpy(g_np).to(dev)
    eng = DTFillEngine(local_rank)
    totals = torch.zeros(_lib.METRIC_COLS + 1, dtype=torch.float64, device=dev)

    def make_batch(first: int, n: int):
        """Global frames first..first+n: pool frame (i % pool) rolled by 8*(i // pool) columns."""
        idx = torch.arange(first, first + n, device=dev)
        xs, gs = xp[idx % args.pool], gp[idx % args.pool]
        shifts = ((idx // args.pool) * 8) % W
        cols = (torch.arange(W, device=dev)[None, :] - shifts[:, None]) % W
        cols = cols[:, None, :].expand(n, H, W)
        return torch.gather(xs, 2, cols).contiguous(), torch.gather(gs, 2, cols).contiguous()

    # warm-up (allocations, clocks)
    xb, gb = make_batch(begin, min(args.batch, end - begin))
    out = eng.fill(xb)
    eng.metrics(out["depth"], gb)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = 0.0
    done = 0
    while begin + done < end:
        n = min(args.batch, end - begin - done)
        xb, gb = make_batch(begin + done, n)
        e0.record()
        out = eng.fill(xb)
        _, sums = eng.metrics(out["depth"], gb)
        e1.record()
        totals += sums
        bad, _ = eng.status()
        assert bad == -1
        kernel_ms += e0.elapsed_time(e1)
        done += n
    sharding.allreduce_sums(totals)                      # the one collective of the path
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    t = torch.tensor([wall, kernel_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        means = sharding.finalize_means(totals)
        print(json.dumps({
            "workload": f"sweep: {args.frames} synthetic KITTI 64-beam frames 352x1216 over {world} GPU(s), "
                        f"batches of {args.batch}, fill + Result.evaluate per frame, NCCL all-reduce of the totals",
            "n_gpus": world, "frames": means["frames"],
            "frames_per_s_wall": args.frames / float(t[0]), "frames_per_s_kernels": args.frames / (float(t[1]) * 1e-3),
            "mean_rmse_mm": means["rmse"], "mean_mae_mm": means["mae"], "mean_irmse_1_per_km": means["irmse"],
            "mean_imae_1_per_km": means["imae"], "valid_pixels": means["valid_pixels"]}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

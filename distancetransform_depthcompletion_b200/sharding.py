"""Frame sharding across the GPUs of one box.

Frames are independent (tools.py:17-28), so a batch is split contiguously, B/G frames per rank, with no
data-path collective.  The only exchange is one all-reduce(sum) of the running metric totals that the eval loops
keep (eval.py:212-232: sums of per-frame rmse/mae/irmse/imae, divided by the frame count at the end).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of rank's frames; the first n % world ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sums(sums, group=None):
    """In-place sum over ranks of the metric-totals vector (torch tensor, CUDA -> NCCL, CPU -> gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


class MetricsComm:
    """The NCCL communicator behind dtfill_allreduce_sums (include/dtfill.h), one rank per process / GPU.

    Rank 0 draws the ncclUniqueId, torch.distributed (any backend, it only carries 128 bytes) hands it to the other
    ranks, every rank calls ncclCommInitRank through the C ABI.  ``allreduce(engine, sums)`` then enqueues the
    float64 sum all-reduce of the totals vector on the engine's stream, right behind the metric kernels."""

    def __init__(self, handle, rank: int, world: int, group=None):
        from . import _lib
        self._lib = _lib
        self.rank, self.world = int(rank), int(world)
        ident = [_lib.nccl_unique_id() if rank == 0 else None]
        if world > 1:
            import torch.distributed as dist
            dist.broadcast_object_list(ident, src=0, group=group)
        self.comm = handle.comm_create(ident[0], world, rank)

    def allreduce(self, engine, sums):
        """sums: float64 CUDA tensor (the running totals, e.g. the 10 columns of dtfill_metrics); in place."""
        engine._bind_stream()
        engine.handle.allreduce_sums(self.comm, sums.data_ptr(), sums.numel())
        return sums

    def close(self):
        if self.comm:
            self._lib.comm_destroy(self.comm)
            self.comm = 0


def finalize_means(sums) -> dict:
    """Mean of per-frame metrics from the all-reduced totals [mse, rmse, mae, irmse, imae, d1, d2, d3, count, n]."""
    s = np.asarray(sums.detach().cpu() if hasattr(sums, "detach") else sums, dtype=np.float64)
    n = s[9]
    names = ("mse", "rmse", "mae", "irmse", "imae", "delta1", "delta2", "delta3")
    out = {k: float(s[i] / n) for i, k in enumerate(names)}
    out["frames"] = int(round(n))
    out["valid_pixels"] = float(s[8])
    return out

"""What limits the drop-in call end to end when all GPUs of a box run it at once?  Every rank, at the same time
(barriers in between): A. device-to-host copies only (438 MB into page-locked memory), B. the compaction of the sparse
upload only (host threads, no GPU), C. the whole call tools.DT_complete_batch; then D. rank 0 alone.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/e2e_scale_probe.py"""
import ctypes
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from distancetransform_depthcompletion_b200 import _lib, tools  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo")
barrier = dist.barrier if world > 1 else (lambda: None)
numa = bench.pin_to_gpu_numa(lr)
torch.cuda.set_device(lr)
avail, total = len(os.sched_getaffinity(0)), os.cpu_count() or 1
sharing = max(1, round(world * avail / total))
threads = max(1, min(12, (3 * avail // 4) // sharing))
x = bench.make_frames(256, rank)[:, :, :, None]
nbytes = x.nbytes


def gather(v):
    if world == 1:
        return [v]
    out = [None] * world
    dist.all_gather_object(out, v)
    return out


def report(name, secs, unit_scale, unit):
    vals = gather(unit_scale / secs)
    if rank == 0:
        print(f"{name}: per rank {[round(v, 1) for v in vals]} {unit}, sum {sum(vals):.1f}", flush=True)


# A. device-to-host only
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
pin.copy_(dev); torch.cuda.synchronize(); barrier()
t0 = time.perf_counter()
for _ in range(8):
    pin.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
report("A device-to-host copies only", (time.perf_counter() - t0) / 8, nbytes / 1e9, "GB/s")
barrier()

# B. compaction only, `threads` host threads per rank (ctypes releases the GIL)
L = _lib.load()
flat = x.reshape(-1)
parts = np.array_split(np.arange(256), threads)
bufs = [(np.empty(len(p) * 352 * 1216 + 64, np.uint32), np.empty(len(p) * 352 * 1216 + 64, np.uint32)) for p in parts]
vp = ctypes.c_void_p


def work(i):
    p = parts[i]
    a = x[p[0]:p[-1] + 1]
    L.dtfill_debug_compact(a.ctypes.data_as(vp), a.size, 0.1, 0.1, bufs[i][0].ctypes.data_as(vp), bufs[i][1].ctypes.data_as(vp), a.size + 64)


def run_b():
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]; [t.join() for t in th]


run_b(); barrier()
t0 = time.perf_counter()
for _ in range(4):
    run_b()
report(f"B compaction only ({threads} threads per rank)", (time.perf_counter() - t0) / 4, nbytes / 1e9, "GB/s of input")
barrier()

# C. the whole call, all ranks
_lib.get_handle(lr).set_stage_threads(threads)
for _ in range(3):
    r = tools.DT_complete_batch(x, device=lr)
barrier()
t0 = time.perf_counter()
for _ in range(8):
    r = tools.DT_complete_batch(x, device=lr)
report("C tools.DT_complete_batch, all ranks at once", (time.perf_counter() - t0) / 8, 256 / 1e3, "k frames/s")
barrier()

# D. rank 0 alone (same thread count, then 12)
for thr in (threads, 12):
    if rank == 0:
        _lib.get_handle(lr).set_stage_threads(thr)
        for _ in range(2):
            r = tools.DT_complete_batch(x, device=lr)
        t0 = time.perf_counter()
        for _ in range(8):
            r = tools.DT_complete_batch(x, device=lr)
        print(f"D rank 0 alone, {thr} staging threads: {256 * 8 / (time.perf_counter() - t0) / 1e3:.1f} k frames/s", flush=True)
    barrier()
if rank == 0:
    print("numa", numa, "cpus available", avail, "of", total, "threads per rank", threads, flush=True)

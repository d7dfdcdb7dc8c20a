#!/usr/bin/env python
"""sweep.py -- BASELINE.json configs[4] alone: `python sweep.py [--gpus N] ...` is `python bench.py --workload sweep ...`
(8192 synthetic KITTI frames sharded over the ranks, fused fill + per-frame metrics steps, one NCCL sum all-reduce of the
totals inside the timed region; see bench.run_sweep).  Under torchrun one process per GPU, like bench.py."""
import sys

import bench

if __name__ == "__main__":
    sys.argv[1:1] = ["--workload", "sweep"]
    bench.main()

"""Drop-in for the two helper functions defined inside the reference's solution_DeepNet/eval_NYU.py:114-133.

    nearest_point(refined_lidar)   -> (dt, lbl)          eval_NYU.py:114-117  (source threshold 0.001)
    Distance_Transform(lidar)      -> depth_map [H,W]    eval_NYU.py:120-133  (any size, dtype of the input)
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .tools import VALID_THR, _as_frames_f32, _gather_f64, _is_wide, _labels_f64
from .tools import nearest_point as _nearest_point

NYU_SRC_THR = 0.001     # eval_NYU.py:115


def nearest_point(refined_lidar, device: int | None = None):
    return _nearest_point(refined_lidar, thr=NYU_SRC_THR, device=device)


def Distance_Transform(lidar, src_thr: float = NYU_SRC_THR, device: int | None = None):
    """eval_NYU.py:120-133.  The notebook copies of this function use src_thr=0.1."""
    lidar = np.squeeze(np.asarray(lidar))                                    # :122
    if lidar.ndim != 2:
        raise ValueError(f"not enough values to unpack (expected 2, got {lidar.ndim})"
                         if lidar.ndim < 2 else f"too many values to unpack (expected 2)")   # :123
    if _is_wide(lidar):
        # float64 input: predicates in float64, the result keeps the input's dtype (:126-133)
        _, lbl, val, r = _labels_f64(lidar[None], src_thr, device)
        if "index_error" in r:
            raise IndexError(r["index_error"])                               # :128
        if int(r["counts"][0, 1]) == 1:
            raise IndexError("too many indices for array: array is 0-dimensional, but 1 were indexed")
        return _gather_f64(lidar, val[0], lbl[0])
    x = _as_frames_f32(lidar, "Distance_Transform")
    r = _lib.get_handle(device).run_host(x[None], src_thr, VALID_THR)
    if "index_error" in r:
        raise IndexError(r["index_error"])                                   # :128
    if int(r["counts"][0, 1]) == 1:
        # :126 np.squeeze(lidar[with_value]) makes a single valid depth 0-dimensional; :128 then fails
        raise IndexError("too many indices for array: array is 0-dimensional, but 1 were indexed")
    return r["depth"][0]

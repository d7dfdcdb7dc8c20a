"""Drop-in for the one compute function of the reference's data_read.py that sits next to the hot path:

    outlier_removal(lidar)        data_read.py:103-128   

numpy in, numpy out; the 7 x 7 diamond filter runs on the GPU (kernel k6_outlier_removal).  The file IO of
data_read.py (PNG / h5 readers) is out of scope.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def outlier_removal(lidar, device: int | None = None):
    """data_read.py:103-128: zero every depth that is more than 1.0 m farther than the average of the valid depths
    inside its 7 x 7 diamond.  Any shape that squeezes to [H,W]; returns float32 [H,W]."""
    sparse_lidar = np.squeeze(np.asarray(lidar))                              # :114
    if sparse_lidar.ndim != 2:
        raise ValueError(f"outlier_removal: input must squeeze to 2-D, got {np.shape(lidar)}")
    if sparse_lidar.dtype != np.float32:
        raise TypeError(f"outlier_removal: dtype {sparse_lidar.dtype} not supported by the CUDA path (float32 only)")
    return _lib.get_handle(device).outlier_removal(np.ascontiguousarray(sparse_lidar)[None])[0]

"""Drop-in for the "DT pooling" helpers every model class of the reference's solution_DeepNet/net.py carries:

    create_weight_matrix(table_size)                           net.py:71-81
    generate_multi_channel(lidar_data, lidar_mask, ...)        net.py:83-123

numpy in, numpy out (the reference runs these lines as TF ops inside the model).  The window
search and averaging run on the GPU (kernels k5_dt_pool_* behind dtfill_dt_pool / dtfill_dt_pool_ex).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def create_weight_matrix(table_size: int = 7) -> np.ndarray:
    """net.py:71-81: weight table_size - |i - mid| - |j - mid| per window position, flattened, float32."""
    assert (table_size + 1) % 2 == 0                                         # net.py:72
    mid = (table_size - 1) // 2
    i = np.abs(np.arange(table_size) - mid)
    return (table_size - i[:, None] - i[None, :]).reshape(-1).astype(np.float32)


def generate_multi_channel(lidar_data, lidar_mask, table_size: int = 7, scale_num: int = 4, device: int | None = None):
    """net.py:83-123.  lidar_data, lidar_mask: float32 [B,H,W,1] (or [B,H,W]); returns (lidar_1, lidar_2, lidar_3,
    lidar_4) with None for the levels beyond scale_num, like the reference; lidar_k has shape [B,H,W] for k >= 2
    (net.py:93 reduces the patch axis) and lidar_1 is the input itself."""
    d = np.asarray(lidar_data)
    m = np.asarray(lidar_mask)
    if d.dtype != np.float32 or m.dtype != np.float32:
        raise TypeError("generate_multi_channel: float32 arrays expected")
    if d.shape != m.shape:
        raise ValueError(f"generate_multi_channel: data {d.shape} and mask {m.shape} differ")
    d3 = d[..., 0] if d.ndim == 4 else d
    m3 = m[..., 0] if m.ndim == 4 else m
    if d3.ndim != 3:
        raise ValueError("generate_multi_channel: expected [B,H,W,1] or [B,H,W]")
    assert (table_size + 1) % 2 == 0                                         # net.py:72
    if not 1 <= scale_num <= 4:
        raise ValueError("scale_num must be 1..4")
    B, H, W = d3.shape
    levels = _lib.get_handle(device).dt_pool(np.ascontiguousarray(d3), np.ascontiguousarray(m3), B, H, W, table_size,
                                             scale_num)
    outs = [d] + [levels[k] for k in range(scale_num - 1)]
    outs += [None] * (4 - len(outs))
    return tuple(outs)


def dt_pooling_masks(lidar_data, lidar_mask, table_size: int = 7, scale_num: int = 4, device: int | None = None):
    """The masks generate_multi_channel pools with, as uint8: level 1 is ``lidar_mask != 0``
    (net.py:131-132 valued_mask), level k >= 2 is ``lidar_k > 0.001`` (net.py:95-96, :105-106, :115-116), written by
    the same kernels that produce lidar_k.  Returns (levels, masks): the tuple of generate_multi_channel and a
    tuple of uint8 [B,H,W] arrays (None beyond scale_num)."""
    d = np.asarray(lidar_data)
    m = np.asarray(lidar_mask)
    if d.dtype != np.float32 or m.dtype != np.float32:
        raise TypeError("dt_pooling_masks: float32 arrays expected")
    if d.shape != m.shape:
        raise ValueError(f"dt_pooling_masks: data {d.shape} and mask {m.shape} differ")
    d3 = d[..., 0] if d.ndim == 4 else d
    m3 = m[..., 0] if m.ndim == 4 else m
    if d3.ndim != 3:
        raise ValueError("dt_pooling_masks: expected [B,H,W,1] or [B,H,W]")
    assert (table_size + 1) % 2 == 0                                         # net.py:72
    if not 1 <= scale_num <= 4:
        raise ValueError("scale_num must be 1..4")
    B, H, W = d3.shape
    levels, masks = _lib.get_handle(device).dt_pool(np.ascontiguousarray(d3), np.ascontiguousarray(m3), B, H, W,
                                                    table_size, scale_num, want_masks=True)
    outs = [d] + [levels[k] for k in range(scale_num - 1)]
    mks = [(m3 != 0).astype(np.uint8)] + [masks[k] for k in range(scale_num - 1)]
    outs += [None] * (4 - len(outs))
    mks += [None] * (4 - len(mks))
    return tuple(outs), tuple(mks)


def create_weight_matrix_demo(size: int = 11) -> np.ndarray:
    """demo.py:65-76: weight 10 ** (size - |i - mid| - |j - mid|) per window position, flattened, float32."""
    assert (size + 1) % 2 == 0                                               # demo.py:66
    mid = (size - 1) // 2
    i = np.abs(np.arange(size) - mid)
    return (10.0 ** (size - i[:, None] - i[None, :])).reshape(-1).astype(np.float32)


def generate_multi_channel_demo(lidar_data, table_size: int = 11, scale_range: float = 90.0, scale_num: int = 4,
                                device: int | None = None):
    """demo.py:107-149, the older pooling without a mask (value-weighted maximum, count_nonzero denominator).
    lidar_data float32 [B,H,W,1] (or [B,H,W]); returns (lidar_1, .., lidar_4) / scale_range, [B,H,W], None beyond
    scale_num, like the reference."""
    d = np.asarray(lidar_data)
    if d.dtype != np.float32:
        raise TypeError("generate_multi_channel_demo: float32 array expected")
    d3 = d[..., 0] if d.ndim == 4 else d
    if d3.ndim != 3:
        raise ValueError("generate_multi_channel_demo: expected [B,H,W,1] or [B,H,W]")
    assert (table_size + 1) % 2 == 0                                         # demo.py:66
    if not 1 <= scale_num <= 4:
        raise ValueError("scale_num must be 1..4")
    B, H, W = d3.shape
    d3 = np.ascontiguousarray(d3)
    levels = _lib.get_handle(device).dt_pool_demo(d3, B, H, W, table_size, scale_num)
    sr = np.float32(scale_range)
    outs = [d3 / sr] + [levels[k] / sr for k in range(scale_num - 1)]        # demo.py:141-148
    outs += [None] * (4 - len(outs))
    return tuple(outs)

"""A numpy stand-in for the nine TensorFlow operations that the reference's DT-pooling lines call -- TEST INFRASTRUCTURE.

TensorFlow is not installed in the build container, so `net.py:83-123` (`generate_multi_channel`) and
`demo.py:107-149` cannot run on TensorFlow itself.  They CAN run unchanged -- the reference's own source lines,
AST-extracted by make_golden.py -- on a module that supplies these operations with TensorFlow's documented semantics:

  tf.image.extract_patches(images [B,H,W,C], sizes=(1,t,t,1), strides=(1,1,1,1), rates=(1,1,1,1), padding='SAME')
      -> [B,H,W,t*t*C]; a patch is flattened rows first, then columns, then channels; 'SAME' centres the window
      (pad (t-1)//2 before, t-1-(t-1)//2 after) and pads with zeros
  tf.math.equal, tf.math.greater (elementwise, bool), tf.math.reduce_max(axis, keepdims), tf.reduce_sum(axis),
  tf.math.count_nonzero(axis) (int64), tf.dtypes.cast, tf.expand_dims, tf.float32

What this does and does not pin: the control flow, the weights, the mask/threshold logic and the arithmetic dtype of
the reference lines are the reference's own; the bodies of the nine operations are this file's (written from the
TensorFlow documentation, independently of oracle/oracle.py, which loops over window offsets instead of taking strided
views).  The order in which TensorFlow's kernels add the <= t*t float32 terms of a reduce_sum is not knowable here;
the tests allow 2e-6 relative for it.
"""
import types

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

float32 = np.float32


def _extract_patches(images, sizes, strides, rates, padding):
    images = np.asarray(images)
    assert images.ndim == 4 and tuple(strides) == (1, 1, 1, 1) and tuple(rates) == (1, 1, 1, 1) and padding == "SAME"
    assert sizes[0] == 1 and sizes[3] == 1
    th, tw = int(sizes[1]), int(sizes[2])
    B, H, W, C = images.shape
    pt, pl = (th - 1) // 2, (tw - 1) // 2
    padded = np.pad(images, ((0, 0), (pt, th - 1 - pt), (pl, tw - 1 - pl), (0, 0)))
    win = sliding_window_view(padded, (th, tw), axis=(1, 2))          # [B,H,W,C,th,tw]
    return np.ascontiguousarray(win.transpose(0, 1, 2, 4, 5, 3)).reshape(B, H, W, th * tw * C)


def _cast(x, dtype):
    return np.asarray(x).astype(dtype)


def _reduce_sum(x, axis=None, keepdims=False):
    x = np.asarray(x)
    return np.sum(x, axis=axis, keepdims=keepdims, dtype=x.dtype)


def _reduce_max(x, axis=None, keepdims=False):
    return np.max(np.asarray(x), axis=axis, keepdims=keepdims)


def _count_nonzero(x, axis=None):
    return np.count_nonzero(np.asarray(x), axis=axis).astype(np.int64)


image = types.SimpleNamespace(extract_patches=_extract_patches)
math = types.SimpleNamespace(equal=lambda a, b: np.equal(a, b), greater=lambda a, b: np.greater(a, b),
                             reduce_max=_reduce_max, count_nonzero=_count_nonzero)
dtypes = types.SimpleNamespace(cast=_cast)
reduce_sum = _reduce_sum
expand_dims = lambda x, axis: np.expand_dims(np.asarray(x), axis)   # noqa: E731

"""B200-native (sm_100a) distance-transform nearest-neighbour fill for sparse depth frames.

Drop-in for the preprocessing hot path of placeforyiming/DistanceTransform-DepthCompletion:
``tools.nearest_point`` / ``tools.DT_complete_batch`` (solution_DeepNet/tools.py),
``eval_nyu.Distance_Transform`` (solution_DeepNet/eval_NYU.py:114-133) and
``evaluation.Result`` / ``Result_NYU`` (evaluation.py).  numpy in, numpy out; the work is done by hand-written
CUDA kernels behind the C ABI of include/dtfill.h (libdtfill.so, loaded with ctypes).  No CPU fallback.
"""
from .tools import nearest_point, DT_complete_batch, dt_fill_batch          # noqa: F401
from .eval_nyu import Distance_Transform                                    # noqa: F401
from .evaluation import Result, Result_NYU, evaluate_batch                  # noqa: F401

__all__ = ["nearest_point", "DT_complete_batch", "dt_fill_batch", "Distance_Transform", "Result", "Result_NYU",
           "evaluate_batch"]

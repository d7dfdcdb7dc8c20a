"""Driver entry points: build() compiles every CUDA extension for sm_100a, smoke() runs the hot path once."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build() -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libdtfill.so (in-tree), the oracle's C
    restatement with gcc, then import the package and load the library."""
    from distancetransform_depthcompletion_b200 import build as b
    path = b.build(force=True)
    print("built", path)
    from oracle import oracle as O
    print("built", O.build(force=True))
    import distancetransform_depthcompletion_b200  # noqa: F401
    from distancetransform_depthcompletion_b200 import _lib
    assert _lib.load().dtfill_abi_version() == 1


def smoke() -> None:
    """One small batch of KITTI-shaped frames through the CUDA path on cuda:0, checked against the oracle."""
    import numpy as np
    from distancetransform_depthcompletion_b200 import _lib, synth, tools, evaluation
    from oracle import oracle as O
    x = synth.kitti_batch([0, 1])
    got = tools.DT_complete_batch(x)
    r = tools.dt_fill_batch(x[..., 0], want_lbl=True)
    o = O.dt_fill(x[..., 0])
    assert np.array_equal(got[..., 0], o["depth"]), "filled depth differs from the oracle"
    assert np.array_equal(r["dt"], o["dt"]) and np.array_equal(r["lbl"], o["lbl"]) and np.array_equal(r["mask"], o["mask"])
    gt = synth.kitti_gt(0)
    R = evaluation.Result()
    R.evaluate(got[0, :, :, 0], gt)
    m = O.result_kitti(o["depth"][0], gt)
    assert abs(R.rmse - m["rmse"]) <= 1e-9 * m["rmse"]
    print("smoke ok: 2 frames 352x1216, rmse %.3f mm, library %s" % (R.rmse, _lib.LIB_PATH))


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()

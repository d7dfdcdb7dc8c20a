"""The oracle (oracle/dtfill_oracle.c + oracle/oracle.py) against the committed outputs of the reference."""
import hashlib
import os

import numpy as np
import pytest

from distancetransform_depthcompletion_b200 import synth
from oracle import oracle as O


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _names(z, suffix):
    return sorted({k.rsplit("/", 1)[0] for k in z.files if k.endswith(suffix)})


def test_small_frames(golden_dir):
    z = np.load(os.path.join(golden_dir, "small_frames.npz"))
    names = _names(z, "/in")
    assert len(names) >= 20
    for n in names:
        x = z[n + "/in"]
        mask = np.asarray(np.float32(1.0) - x > np.float32(0.1)).astype(np.uint8)
        dt, lbl = O.chamfer_l1_labels(mask)
        assert np.array_equal(dt, z[n + "/dt"]), n
        assert np.array_equal(lbl, z[n + "/lbl"]), n
        r = O.dt_fill(x, 0.1, 0.1)
        assert np.array_equal(r["depth"], z[n + "/depth"]), n
        assert np.array_equal(r["dt"], z[n + "/dt"]) and np.array_equal(r["lbl"], z[n + "/lbl"]), n


def test_full_frames(golden_dir):
    z = np.load(os.path.join(golden_dir, "full_frames.npz"))
    for name, x, thr in (("kitti64_seed0", synth.kitti_frame(0), 0.1),
                         ("kitti8_seed5", synth.kitti_frame(5, beam_step=8), 0.1),
                         ("nyu_seed3", synth.nyu_frame(3), 0.001)):
        dt, lbl = O.nearest_point(x, thr)
        assert np.array_equal(lbl, z[name + "/lbl"]), name
        assert np.array_equal(dt.astype(np.uint16), z[name + "/dt_u16"]), name
        assert dt.max() < 65533


def test_checksums(golden_dir):
    z = np.load(os.path.join(golden_dir, "checksums.npz"))
    for step in (1, 8):
        for seed in (0, 3):
            x = synth.kitti_frame(seed, beam_step=step)
            r = O.dt_fill(x, 0.1, 0.1)
            want = z[f"kitti_b{64 // step}_s{seed}"]
            assert [sha(r["depth"]), sha(r["dt"]), sha(r["lbl"])] == list(want)
    x = synth.nyu_frame(2)
    r = O.dt_fill(x, 0.001, 0.1)
    assert [sha(r["depth"]), sha(r["dt"]), sha(r["lbl"])] == list(z["nyu_s2"])
    x = synth.nyu_frame(9, 240, 320)
    r = O.dt_fill(x, 0.001, 0.1)
    assert [sha(r["depth"]), sha(r["dt"]), sha(r["lbl"])] == list(z["nyu240_s9"])


def test_metrics(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    for seed in range(3):
        fill = O.dt_fill(synth.kitti_frame(seed))["depth"]
        gt = synth.kitti_gt(seed)
        m = O.result_kitti(fill, gt)
        assert [m[k] for k in ("mse", "rmse", "mae", "irmse", "imae")] == list(z[f"kitti_s{seed}"])
        m = O.result_kitti(np.maximum(fill, np.float32(0.9)), gt.astype(np.float32))
        assert [m[k] for k in ("mse", "rmse", "mae", "irmse", "imae")] == list(z[f"kitti_f32gt_s{seed}"])
        xn, g = synth.nyu_frame(seed, return_dense=True)
        m = O.result_nyu(O.dt_fill(xn, 0.001, 0.1)["depth"], g)
        assert [m[k] for k in ("mse", "rmse", "mae", "irmse", "imae", "delta1", "delta2", "delta3")] == \
            list(z[f"nyu_s{seed}"])


def test_index_errors():
    x = np.zeros((8, 9), np.float32)
    with pytest.raises(IndexError):
        O.dt_fill(x)
    x[2, 2] = np.nan            # a source (1-nan > thr is False) that is not valid
    with pytest.raises(IndexError):
        O.dt_fill(x)
    x[4, 4] = 0.5               # valid, not a source: one source, one valid -> fine
    r = O.dt_fill(x)
    assert np.all(r["depth"] == np.float32(0.5))


def _pool_cases():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    return make_golden.pool_cases(), make_golden.pool_inputs


def test_dt_pooling_restatement_equals_the_reference_lines(golden_dir):
    """f-1: tests/golden/dt_pool.npz holds the outputs of net.py:83-123 / demo.py:107-149 THEMSELVES (AST-extracted,
    run on tests/golden/tf_numpy_shim.py because TensorFlow is absent); the numpy restatement in oracle/oracle.py must
    reproduce them bit for bit (both sum a window's float32 terms in numpy's order)."""
    z = np.load(os.path.join(golden_dir, "dt_pool.npz"))
    cases, pool_inputs = _pool_cases()
    for t in (3, 7, 11):
        assert np.array_equal(O.create_weight_matrix(t), z[f"weights/net_t{t}"])
        assert np.array_equal(O.demo_create_weight_matrix(t), z[f"weights/demo_t{t}"])
    n = 0
    for name, (x, t, s) in cases.items():
        data, mask = pool_inputs(x)
        lv = O.generate_multi_channel(data[..., 0], mask[..., 0], t, s)
        for k in range(1, s):
            assert np.array_equal(lv[k], z[f"net/{name}/l{k + 1}"]), (name, k)
            n += 1
        if f"demo/{name}/l1" in z.files:
            dl = O.demo_generate_multi_channel(x.astype(np.float32), t, 90.0, s)
            for k in range(s):
                assert np.array_equal(dl[k], z[f"demo/{name}/l{k + 1}"]), (name, k)
                n += 1
    assert n >= 25

"""Pins the oracle against the LIVE reference where it is present (the build container).  On the GPU box
/root/reference does not exist and these tests skip; the committed fixtures (test_oracle_golden.py) stand in."""
import os
import sys

import numpy as np
import pytest

from distancetransform_depthcompletion_b200 import synth
from oracle import oracle as O

REF = os.environ.get("DTFILL_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not (os.path.isdir(REF) and O.have_cv2()), reason="reference checkout or cv2 absent")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    return make_golden.load_reference()


def test_random_masks_vs_cv2(ref):
    cv2 = ref[4]
    rng = np.random.default_rng(0)
    for t in range(300):
        H, W = int(rng.integers(1, 41)), int(rng.integers(1, 61))
        dens = rng.choice([0.01, 0.05, 0.2, 0.5, 0.9])
        m = (rng.random((H, W)) >= dens).astype(np.uint8)
        dt, lbl = cv2.distanceTransformWithLabels(m, cv2.DIST_L1, 5, labelType=cv2.DIST_LABEL_PIXEL)
        d2, l2 = O.chamfer_l1_labels(m)
        assert np.array_equal(dt, d2) and np.array_equal(lbl, l2), (t, H, W)


def test_tools_dt_complete_batch(ref):
    tools = ref[0]
    xb = synth.kitti_batch([7, 8], beam_step=2)
    want = tools.DT_complete_batch(xb)
    got = O.dt_fill(xb[..., 0])
    assert want.dtype == np.float32 and want.shape == (2, 352, 1216, 1)
    assert np.array_equal(want[..., 0], got["depth"])
    dt, lbl = tools.nearest_point(xb[0, :, :, 0])
    assert np.array_equal(dt, got["dt"][0]) and np.array_equal(lbl, got["lbl"][0])


def test_eval_nyu_distance_transform(ref):
    dt_fn = ref[3]
    x = synth.nyu_frame(11)
    want = dt_fn(x[None, :, :, None])
    assert np.array_equal(want, O.dt_fill(x, 0.001, 0.1)["depth"])


def test_cv2_port_matches_tools(ref):
    tools = ref[0]
    xb = synth.kitti_batch([1])
    assert np.array_equal(tools.DT_complete_batch(xb), O.cv2_port_complete_batch(xb))


def test_metrics_vs_evaluation(ref):
    evaluation = ref[1]
    fill = O.dt_fill(synth.kitti_frame(4))["depth"]
    gt = synth.kitti_gt(4)
    R = evaluation.Result()
    R.evaluate(fill, gt)
    m = O.result_kitti(fill, gt)
    assert (R.mse, R.rmse, R.mae, R.irmse, R.imae) == (m["mse"], m["rmse"], m["mae"], m["irmse"], m["imae"])


def test_dt_pooling_lines_on_the_tf_stand_in():
    """f-1: net.py:83-123 and demo.py:107-149, AST-extracted and run on tests/golden/tf_numpy_shim.py, against the
    restatement in oracle/oracle.py on fresh random inputs; the stand-in's extract_patches against a plain loop over
    window offsets and pixels."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    import tf_numpy_shim as tf
    net_pool, demo_pool = make_golden.load_pooling_reference()
    rng = np.random.default_rng(5)
    img = rng.random((2, 6, 7, 1)).astype(np.float32)
    for t in (3, 5):
        got = tf.image.extract_patches(img, sizes=(1, t, t, 1), strides=(1, 1, 1, 1), rates=(1, 1, 1, 1), padding="SAME")
        assert got.shape == (2, 6, 7, t * t)
        for b in range(2):
            for y in range(6):
                for x in range(7):
                    for i in range(t):
                        for j in range(t):
                            yy, xx = y + i - t // 2, x + j - t // 2
                            want = img[b, yy, xx, 0] if 0 <= yy < 6 and 0 <= xx < 7 else 0.0
                            assert got[b, y, x, i * t + j] == want
    for trial in range(6):
        H, W = int(rng.integers(8, 40)), int(rng.integers(8, 50))
        x = ((rng.random((2, H, W)) < rng.choice([0.02, 0.1, 0.4])) * rng.uniform(0.05, 80, (2, H, W))).astype(np.float32)
        t, s = int(rng.choice([3, 5, 7, 9])), int(rng.integers(1, 5))
        data, mask = make_golden.pool_inputs(x)
        lv = net_pool(data, mask, t, s)
        want = O.generate_multi_channel(data[..., 0], mask[..., 0], t, s)
        assert lv[0] is data
        for k in range(1, 4):
            assert (lv[k] is None) == (want[k] is None)
            if k < s:
                assert np.array_equal(lv[k], want[k])
        dl = demo_pool(x[:1, :, :, None], t, 90.0, s)
        dw = O.demo_generate_multi_channel(x[:1], t, 90.0, s)
        for k in range(s):
            assert np.array_equal(dl[k], dw[k])

"""DT pooling of the CNN input stage (net.py:71-123).  TensorFlow is absent, so the pin is
the reference's own lines (AST-extracted) run on a numpy stand-in for the nine TensorFlow operations they call
(tests/golden/tf_numpy_shim.py): their outputs are committed as tests/golden/dt_pool.npz, the numpy restatement
oracle.generate_multi_channel reproduces them bit for bit (tests/test_oracle_golden.py), and the kernels are compared
with both.  What stays unverifiable here is TensorFlow's own summation order inside reduce_sum (2e-6 relative)."""
import os
import sys

import numpy as np
import pytest

from distancetransform_depthcompletion_b200 import net_pool, synth
from oracle import oracle as O


def test_weight_matrix_matches_restatement():
    for t in (1, 3, 5, 7, 9):
        assert np.array_equal(net_pool.create_weight_matrix(t), O.create_weight_matrix(t))
    w = net_pool.create_weight_matrix(7).reshape(7, 7)
    assert w[3, 3] == 7 and w[0, 0] == 1 and w[3, 0] == 4 and w.dtype == np.float32
    with pytest.raises(AssertionError):
        net_pool.create_weight_matrix(4)


def test_restatement_properties():
    """Level masks are Chebyshev dilations of the validity mask by 3 px per level."""
    from scipy.ndimage import binary_dilation
    x = synth.kitti_frame(1)[None, 100:200, :300]
    mask = (x > 0.1).astype(np.float32)
    l1, l2, l3, l4 = O.generate_multi_channel(x / 90.0 * mask, mask)
    sq = np.ones((1, 7, 7), bool)
    m = mask > 0
    for lv in (l2, l3, l4):
        m = binary_dilation(m, sq)
        assert np.array_equal(lv > 0.001, m)
    assert np.array_equal(l1, (x / 90.0 * mask).astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("scale_num,table_size", [(4, 7), (2, 7), (3, 5), (4, 3), (1, 7), (3, 9)])
def test_gpu_matches_restatement(scale_num, table_size):
    rng = np.random.default_rng(scale_num * 10 + table_size)
    frames = [synth.kitti_frame(5)[90:190, 200:520], synth.kitti_frame(6, beam_step=4)[90:190, 200:520],
              ((rng.random((100, 320)) < 0.02) * rng.uniform(1, 80, (100, 320))).astype(np.float32),
              np.zeros((100, 320), np.float32)]
    x = np.stack(frames)[..., None]
    mask = (x > 0.1).astype(np.float32)                           # net.py:464-465
    data = (x / np.float32(90.0) * mask).astype(np.float32)       # net.py:467, :486
    got = net_pool.generate_multi_channel(data, mask, table_size=table_size, scale_num=scale_num)
    want = O.generate_multi_channel(data[..., 0], mask[..., 0], table_size, scale_num)
    assert got[0] is data or np.array_equal(got[0], data)
    for k in range(1, 4):
        if k >= scale_num:
            assert got[k] is None and want[k] is None
            continue
        assert got[k].shape == (4, 100, 320) and got[k].dtype == np.float32
        # only the order of the (<= T*T term) float32 sums may differ
        np.testing.assert_allclose(got[k], want[k], rtol=2e-6, atol=1e-9)
        assert np.array_equal(got[k] > 0.001, want[k] > 0.001)


@pytest.mark.gpu
@pytest.mark.parametrize("W", [319, 66, 5])
def test_gpu_odd_widths(W):
    """Widths that are not a multiple of 4 take the scalar tile load; tiny frames are all halo."""
    rng = np.random.default_rng(W)
    x = np.stack([synth.kitti_frame(7)[100:171, 300:300 + W],
                  ((rng.random((71, W)) < 0.05) * rng.uniform(1, 80, (71, W))).astype(np.float32)])[..., None]
    mask = (x > 0.1).astype(np.float32)
    data = (x / np.float32(90.0) * mask).astype(np.float32)
    for t in (3, 5, 7, 9, 11):
        got = net_pool.generate_multi_channel(data, mask, table_size=t, scale_num=3)
        want = O.generate_multi_channel(data[..., 0], mask[..., 0], t, 3)
        for k in (1, 2):
            np.testing.assert_allclose(got[k], want[k], rtol=2e-6, atol=1e-9)
            assert np.array_equal(got[k] > 0.001, want[k] > 0.001)


@pytest.mark.gpu
@pytest.mark.parametrize("table_size", [7, 9, 11])
def test_gpu_level_masks(table_size):
    """the level-k DT-pooling masks as uint8, from the kernels that produce the levels."""
    from scipy.ndimage import binary_dilation
    x = np.stack([synth.kitti_frame(8)[60:200, 100:420], synth.nyu_frame(3)[:140, :320]])[..., None]
    mask = (x > 0.1).astype(np.float32)
    data = (x / np.float32(90.0) * mask).astype(np.float32)
    levels, masks = net_pool.dt_pooling_masks(data, mask, table_size=table_size, scale_num=4)
    plain = net_pool.generate_multi_channel(data, mask, table_size=table_size, scale_num=4)
    m = mask[..., 0] > 0
    assert np.array_equal(masks[0], m.astype(np.uint8))
    for k in (1, 2, 3):
        assert np.array_equal(levels[k], plain[k])
        assert masks[k].dtype == np.uint8 and np.array_equal(masks[k], (levels[k] > 0.001).astype(np.uint8))
        m = binary_dilation(m, np.ones((1, table_size, table_size), bool))
        assert np.array_equal(masks[k].astype(bool), m)


@pytest.mark.gpu
def test_gpu_full_size_and_errors():
    x = synth.kitti_batch([11, 12])
    mask = (x > 0.1).astype(np.float32)
    data = (x / np.float32(90.0) * mask).astype(np.float32)
    got = net_pool.generate_multi_channel(data, mask)
    want = O.generate_multi_channel(data[..., 0], mask[..., 0])
    for k in (1, 2, 3):
        np.testing.assert_allclose(got[k], want[k], rtol=2e-6, atol=1e-9)
    with pytest.raises(TypeError):
        net_pool.generate_multi_channel(data.astype(np.float64), mask)
    with pytest.raises(ValueError):
        net_pool.generate_multi_channel(data, mask[:, :10])


@pytest.mark.gpu
def test_demo_variant_of_the_pooling(dtfill_lib):
    """demo.py:65-149, the older pooling (weights 10 ** ..., value-weighted maximum, count_nonzero denominator), against
    its numpy restatement in oracle/oracle.py (pinned like f-1 to the reference's own lines run on tests/golden/tf_numpy_shim.py).  A
    single selected pixel (the rule, ties need equal value times weight) is reproduced exactly; sums of several agree to
    float32 rounding."""
    from distancetransform_depthcompletion_b200 import net_pool, synth
    from oracle import oracle as O
    assert np.array_equal(net_pool.create_weight_matrix_demo(11), O.demo_create_weight_matrix(11))
    rng = np.random.default_rng(3)
    frames = [synth.kitti_frame(0)[96:224, 300:620], synth.kitti_frame(1, beam_step=4)[200:328, 100:420]]
    x = np.stack(frames).astype(np.float32)
    x[1, 10:14, 10:14] = 7.5                                            # equal values: ties inside a window
    for T, sn in ((11, 4), (7, 3), (3, 2)):
        got = net_pool.generate_multi_channel_demo(x[..., None], table_size=T, scale_num=sn)
        want = O.demo_generate_multi_channel(x, T, 90.0, sn)
        for g, w in zip(got, want):
            assert (g is None) == (w is None)
            if g is not None:
                assert g.shape == w.shape and g.dtype == np.float32
                np.testing.assert_allclose(g, w, rtol=2e-6, atol=0)
    r = rng.random((1, 33, 47)).astype(np.float32) * (rng.random((1, 33, 47)) < 0.3)     # odd sizes, borders
    got = net_pool.generate_multi_channel_demo(r, table_size=5, scale_num=4)
    want = O.demo_generate_multi_channel(r, 5, 90.0, 4)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=2e-6, atol=0)


@pytest.mark.gpu
def test_gpu_matches_the_reference_lines_fixture(golden_dir):
    """The kernels against tests/golden/dt_pool.npz = outputs of net.py:83-123 / demo.py:107-149 themselves (run on the
    numpy stand-in for TensorFlow, see the module docstring)."""
    sys.path.insert(0, golden_dir)
    import make_golden
    z = np.load(os.path.join(golden_dir, "dt_pool.npz"))
    for t in (3, 7, 11):
        assert np.array_equal(net_pool.create_weight_matrix(t), z[f"weights/net_t{t}"])
        assert np.array_equal(net_pool.create_weight_matrix_demo(t), z[f"weights/demo_t{t}"])
    n = 0
    for name, (x, t, s) in make_golden.pool_cases().items():
        data, mask = make_golden.pool_inputs(x)
        got = net_pool.generate_multi_channel(data, mask, table_size=t, scale_num=s)
        for k in range(1, s):
            want = z[f"net/{name}/l{k + 1}"]
            np.testing.assert_allclose(got[k], want, rtol=2e-6, atol=1e-9, err_msg=f"{name} level {k + 1}")
            assert np.array_equal(got[k] > 0.001, want > 0.001)
            n += 1
        if f"demo/{name}/l1" in z.files:
            dg = net_pool.generate_multi_channel_demo(x[..., None].astype(np.float32), table_size=t, scale_num=s)
            for k in range(s):
                np.testing.assert_allclose(dg[k], z[f"demo/{name}/l{k + 1}"], rtol=2e-6, atol=0, err_msg=f"demo {name} {k + 1}")
                n += 1
    assert n >= 25

"""f-4 (extension): the exact Euclidean feature transform dtfill_edt / tools.euclidean_feature_transform.
There is no reference function behind it; the oracle is scipy.ndimage.distance_transform_edt on the source mask of
tools.py:8 (pinned here against an exhaustive search).  Squared distances must be identical; an index is checked by
what it attains (it must point at a source at exactly that distance) and by the documented tie rule (the first source in
raster order among those at the minimal distance)."""
import numpy as np
import pytest

from distancetransform_depthcompletion_b200 import synth
from oracle import oracle as O


def _random_frame(rng, H, W, dens):
    return ((rng.random((H, W)) < dens) * rng.uniform(1, 50, (H, W))).astype(np.float32)


def test_oracle_edt_matches_exhaustive_search():
    rng = np.random.default_rng(0)
    for t in range(40):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 50))
        x = _random_frame(rng, H, W, rng.choice([0.01, 0.05, 0.3, 0.9]))
        d2, idx = O.edt(x)
        assert np.array_equal(d2, O.edt_bruteforce(x)), (t, H, W)
    assert np.all(O.edt(np.zeros((5, 7), np.float32))[1] == -1)


def _check(x, thr=0.1):
    from distancetransform_depthcompletion_b200 import tools
    d2, idx = tools.euclidean_feature_transform(x, thr)
    want, _ = O.edt(x, thr)
    assert d2.dtype == np.int32 and idx.dtype == np.int32 and d2.shape == x.shape
    assert np.array_equal(d2.astype(np.int64), want)
    src = O.source_mask(x, thr)
    H, W = x.shape
    if not src.any():
        assert np.all(idx == -1)
        return
    iy, ix = idx // W, idx % W
    assert np.all(src[iy, ix]), "an index does not point at a source"
    yy, xx = np.mgrid[0:H, 0:W]
    assert np.array_equal((yy - iy) ** 2 + (xx - ix) ** 2, want), "an index does not attain the minimal distance"
    if H * W <= 4096:
        # documented tie rule: among the sources at the minimal distance the first in raster order
        ys, xs = np.nonzero(src)
        d = (yy[..., None] - ys) ** 2 + (xx[..., None] - xs) ** 2
        first = np.argmax(d == want[..., None], axis=-1)
        assert np.array_equal(idx, ys[first] * W + xs[first]), "tie rule"
    return d2, idx


@pytest.mark.gpu
def test_edt_random_shapes(dtfill_lib):
    rng = np.random.default_rng(1)
    for t in range(60):
        H, W = int(rng.integers(1, 150)), int(rng.integers(1, 300))
        if t % 3 == 0:
            W = 8 * max(1, W // 8)           # 128-bit stores of the row pass
        if t % 4 == 1:
            H, W = int(rng.integers(1, 64)), int(rng.integers(1, 64))     # small enough for the tie-rule check
        _check(_random_frame(rng, H, W, rng.choice([0.002, 0.02, 0.2, 0.8])))
    _check(_random_frame(rng, 3, 2500, 0.001))                             # three chunks of bit words per row
    _check(_random_frame(rng, 700, 40, 0.01))                              # deep envelope stacks


@pytest.mark.gpu
def test_edt_config_frames_and_batches(dtfill_lib):
    from distancetransform_depthcompletion_b200 import tools
    for s in (0, 1):
        _check(synth.kitti_frame(s))
        _check(synth.kitti_frame(s, beam_step=8))
    _check(synth.nyu_frame(3), thr=0.001)
    xb = np.stack([synth.kitti_frame(i, beam_step=2) for i in range(3)])
    d2, idx = tools.euclidean_feature_transform(xb)
    for i in range(3):
        assert np.array_equal(d2[i].astype(np.int64), O.edt(xb[i])[0])


@pytest.mark.gpu
def test_edt_edge_cases(dtfill_lib):
    from distancetransform_depthcompletion_b200 import tools
    z = np.zeros((40, 64), np.float32)
    d2, idx = tools.euclidean_feature_transform(z)
    assert np.all(d2 == 2 ** 31 - 1) and np.all(idx == -1)
    one = z.copy(); one[17, 40] = 3.0
    d2, idx = _check(one)
    assert np.all(idx == 17 * 64 + 40)
    _check(np.full((33, 65), 2.0, np.float32))                      # every pixel a source
    for H, W in ((1, 200), (200, 1), (32, 32), (33, 31), (64, 8), (97, 1216)):
        x = np.zeros((H, W), np.float32); x[H // 2, W // 3] = 5.0; x[0, W - 1] = 6.0
        _check(x)
    # documented ties: the first source in raster order among those at the minimal distance
    t = np.zeros((5, 5), np.float32); t[0, 2] = 1.0; t[4, 2] = 1.0; t[2, 0] = 1.0; t[2, 4] = 1.0
    d2, idx = tools.euclidean_feature_transform(t)
    assert d2[2, 2] == 4 and idx[2, 2] == 0 * 5 + 2                 # the upper row
    t = np.zeros((3, 5), np.float32); t[1, 0] = 1.0; t[1, 4] = 1.0
    d2, idx = tools.euclidean_feature_transform(t)
    assert d2[1, 2] == 4 and idx[1, 2] == 1 * 5 + 0                 # equal |dx|: the left one
    with pytest.raises(TypeError):
        tools.euclidean_feature_transform(z.astype(np.float64))

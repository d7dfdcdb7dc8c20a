"""The numpy model of the kernel arithmetic (tests/kernel_model.py) against the oracle: packed keys,
candidate order, lane decomposition of the row scans, band + halo exactness with the coarse bound."""
import numpy as np

import kernel_model as M
from distancetransform_depthcompletion_b200 import synth
from oracle import oracle as O


def _run(x, thr, ppl, bands=None):
    src = ~((np.float32(1.0) - x) > np.float32(thr))
    rank = np.where(src, np.cumsum(src.reshape(-1)).reshape(src.shape), 0)
    dt, lbl = O.chamfer_l1_labels((~src).astype(np.uint8))
    H, W = x.shape
    if bands is None:
        d, l = M.chamfer_band(src, rank, 0, H, ppl)
        assert np.array_equal(d, dt.astype(np.int64))
        assert np.array_equal(l, lbl)
        return
    BH, ch, cw = bands
    U = M.coarse_row_bound(src, ch, cw)
    assert (U >= dt.max(axis=1)).all()
    for r0 in range(0, H, BH):
        r1 = min(H, r0 + BH)
        ys = np.arange(r0, r1)
        lo = max(0, int((ys - U[r0:r1]).min()))
        hi = min(H, int((ys + U[r0:r1]).max()) + 1)
        d, l = M.chamfer_band(src, rank, lo, hi, ppl)
        assert np.array_equal(d[r0 - lo:r1 - lo], dt[r0:r1].astype(np.int64)), (r0, lo, hi)
        assert np.array_equal(l[r0 - lo:r1 - lo], lbl[r0:r1]), (r0, lo, hi)


def test_model_small_random():
    rng = np.random.default_rng(2)
    for _ in range(40):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 65))
        dens = rng.choice([0.005, 0.02, 0.1, 0.5, 0.9])
        x = ((rng.random((H, W)) < dens) * rng.uniform(1, 50, (H, W))).astype(np.float32)
        _run(x, 0.1, 2)


def test_model_adversarial():
    for name, f in synth.adversarial_frames().items():
        _run(f, 0.1, 2)


def test_model_bands_small():
    rng = np.random.default_rng(3)
    for _ in range(10):
        H, W = int(rng.integers(8, 60)), int(rng.integers(8, 64))
        x = ((rng.random((H, W)) < 0.05) * rng.uniform(1, 50, (H, W))).astype(np.float32)
        if not (x > 0).any():
            continue
        _run(x, 0.1, 2, bands=(int(rng.integers(3, 12)), 4, 4))


def test_model_kitti_and_nyu_shapes():
    _run(synth.kitti_frame(0)[100:180], 0.1, 38)       # 80 rows at full width: exercises PPL=38 packing
    _run(synth.nyu_frame(0)[:64], 0.001, 20)


def test_sky_rows_closed_form_matches_oracle():
    """The rule k3_sky implements for the rows above the first source row, against the oracle's scan."""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    checked = 0
    for trial in range(120):
        H, W = int(rng.integers(6, 50)), int(rng.integers(1, 80))
        f = int(rng.integers(1, H))
        dens = rng.choice([0.01, 0.05, 0.3, 0.9])
        x = ((rng.random((H, W)) < dens) * rng.uniform(1, 50, (H, W))).astype(np.float32)
        x[:f] = 0
        x[f, rng.integers(0, W)] = 5.0
        o = O.dt_fill(x[None], 0.1, 0.1)
        dt, lbl = o["dt"][0], o["lbl"][0]
        for S in {f - 1, (f - 1) // 4 * 4, max(f - 3, 0)}:
            if S < 1:
                continue
            odt, ol = M.sky_rows_closed_form(dt.astype(np.int64), lbl, S)
            assert np.array_equal(odt, dt.astype(np.int64)) and np.array_equal(ol, lbl), (H, W, f, S)
            checked += 1
    assert checked > 100

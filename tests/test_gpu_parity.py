"""Parity of the CUDA path (through the C ABI / ctypes shim) with the oracle and the committed reference
outputs.  Bit-exact for dt, lbl, mask and the filled depth (a copy of input values); metrics within the stated
relative tolerance.  Run on the B200 box:  python -m pytest tests -m gpu"""
import hashlib
import os

import numpy as np
import pytest

from distancetransform_depthcompletion_b200 import _lib, eval_nyu, evaluation, synth, tools
from oracle import oracle as O

pytestmark = pytest.mark.gpu

METRIC_RTOL_F64 = 1e-9     # one-pass sums (dtfill_run_eval_async, metrics_exact off), float64 ground truth: the order differs
METRIC_RTOL_F32 = 2e-5     # ... float32 ground truth: numpy's mean itself accumulates in float32 (pairwise)
# dtfill_metrics in its default mode reproduces numpy's pairwise sums: compared with == below


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def handle(dtfill_lib):
    return _lib.get_handle()


def _check_frames(handle, frames, src_thr, val_thr):
    frames = np.ascontiguousarray(frames, np.float32)
    r = handle.run_host(frames, src_thr, val_thr, want_dt=True, want_lbl=True, want_mask=True)
    assert "index_error" not in r, r.get("index_error")
    o = O.dt_fill(frames, src_thr, val_thr)
    for b in range(frames.shape[0]):
        assert np.array_equal(r["lbl"][b], o["lbl"][b]), f"lbl frame {b}"
        assert np.array_equal(r["dt"][b], o["dt"][b]), f"dt frame {b}"
        assert np.array_equal(r["mask"][b], o["mask"][b]), f"mask frame {b}"
        assert np.array_equal(r["depth"][b].view(np.uint32), o["depth"][b].view(np.uint32)), f"depth frame {b}"
        assert r["counts"][b, 1] == int(o["mask"][b].sum())
    # without the optional outputs the depth must not change
    r2 = handle.run_host(frames, src_thr, val_thr)
    assert np.array_equal(r2["depth"], r["depth"])
    return r


def test_small_golden_frames(handle, golden_dir):
    z = np.load(os.path.join(golden_dir, "small_frames.npz"))
    names = sorted({k.rsplit("/", 1)[0] for k in z.files if k.endswith("/in")})
    for n in names:
        x = z[n + "/in"]
        r = handle.run_host(np.ascontiguousarray(x[None]), 0.1, 0.1, want_dt=True, want_lbl=True, want_mask=True)
        assert "index_error" not in r, (n, r.get("index_error"))
        assert np.array_equal(r["lbl"][0], z[n + "/lbl"]), n
        assert np.array_equal(r["dt"][0], z[n + "/dt"]), n
        assert np.array_equal(r["depth"][0], z[n + "/depth"]), n


def test_random_small_frames_all_paths(handle):
    rng = np.random.default_rng(5)
    for t in range(60):
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 400))
        dens = rng.choice([0.003, 0.02, 0.1, 0.5, 0.95])
        x = ((rng.random((2, H, W)) < dens) * rng.uniform(1, 50, (2, H, W))).astype(np.float32)
        x[:, H // 2, W // 2] = 3.0
        _check_frames(handle, x, 0.1, 0.1)


@pytest.mark.parametrize("thr", [0.1, 0.001, 0.0, 1.0, 0.5, -0.5, 2.0, 1e-30, 0.9999999, 3e38, -3e38])
def test_source_threshold_boundary(handle, thr):
    """K1 evaluates tools.py:8's !(float32(1 - x) > thr) as !(x < cut) with a host-computed cut: every float within
    a few ulps of the boundary, signed zeros, denormals and infinities must classify as the reference does."""
    thr32 = np.float32(thr)
    centre = np.float32(1.0) - thr32
    vals = [centre]
    for _ in range(6):
        vals.append(np.nextafter(vals[-1], np.float32(np.inf)))
    lo = centre
    for _ in range(6):
        lo = np.nextafter(lo, np.float32(-np.inf)); vals.append(lo)
    vals += [0.0, -0.0, 1e-45, -1e-45, 1.0, -1.0, np.inf, 3e38, -3e38, 0.9, 0.999, 1.0000001, 2.0, 1e-38, -1e-38]
    vals = np.array(vals, np.float32)
    rng = np.random.default_rng(11)
    for W in (64, 61):                                      # 128-bit K1 and the scalar K1
        x = rng.choice(vals, size=(3, 16, W)).astype(np.float32)
        with np.errstate(over="ignore", invalid="ignore"):
            src = ~((np.float32(1.0) - x).astype(np.float32) > thr32)
        r = handle.run_host(x, float(thr32), -3.3e38, want_dt=True, want_lbl=True)
        assert "index_error" not in r
        assert np.array_equal(r["counts"][:, 0], src.reshape(3, -1).sum(1))
        o = O.dt_fill(x, float(thr32), -3.3e38)
        assert np.array_equal(r["lbl"], o["lbl"]) and np.array_equal(r["dt"], o["dt"])
        assert np.array_equal(r["depth"].view(np.uint32), o["depth"].view(np.uint32))


@pytest.mark.parametrize("W", [320, 640, 1216, 319, 321, 641, 1215, 100, 37])
def test_lane_layouts(handle, W):
    """Exact multiples of 32*PPL (no padding) and ragged widths for each lane layout (PPL 10/20/38)."""
    rng = np.random.default_rng(W)
    x = ((rng.random((3, 50, W)) < 0.03) * rng.uniform(1, 50, (3, 50, W))).astype(np.float32)
    x[:, 0, 0] = 2.0
    _check_frames(handle, x, 0.1, 0.1)


def test_wide_path_by_size(handle):
    """Frames the 32-bit key cannot hold (wider than 1216, or 2H+W too large) take the 64-bit-key kernel."""
    rng = np.random.default_rng(8)
    for H, W in ((40, 1300), (20, 3000), (900, 300), (5, 2049)):
        x = ((rng.random((2, H, W)) < 0.01) * rng.uniform(1, 50, (2, H, W))).astype(np.float32)
        x[:, H - 1, 0] = 2.0
        _check_frames(handle, x, 0.1, 0.1)


def test_wide_path_by_source_count(handle):
    """More than 2^18-1 sources in a frame that otherwise fits the fast path."""
    rng = np.random.default_rng(9)
    x = rng.uniform(1, 50, (2, 352, 1216)).astype(np.float32)
    x[1] *= (rng.random((352, 1216)) < 0.7)
    x[0, 100:140, 200:900] = 0.0
    r = _check_frames(handle, x, 0.1, 0.1)
    assert r["counts"][0, 0] > (1 << 18) and r["counts"][1, 0] > (1 << 18)


@pytest.mark.parametrize("cap", [0, 8, 24, 64, 150, 400])
def test_band_splitting_is_exact(handle, cap):
    """Frames split into independently processed bands of rows (+ halo from the coarse bound) must give the
    very same labels as the whole-frame scan, whatever the band size."""
    handle.set_band_cap(cap)
    try:
        x = np.stack([synth.kitti_frame(600 + i, beam_step=s) for i, s in enumerate((1, 2, 8))])
        _check_frames(handle, x, 0.1, 0.1)
        _check_frames(handle, np.stack([synth.nyu_frame(610), synth.nyu_frame(611, samples=50)]), 0.001, 0.1)
        rng = np.random.default_rng(cap)
        for t in range(12):
            H, W = int(rng.integers(5, 120)), int(rng.integers(5, 500))
            dens = rng.choice([0.002, 0.02, 0.2])
            f = ((rng.random((2, H, W)) < dens) * rng.uniform(1, 50, (2, H, W))).astype(np.float32)
            f[:, rng.integers(0, H), rng.integers(0, W)] = 3.0
            _check_frames(handle, f, 0.1, 0.1)
        # a frame whose only sources sit in one corner: bounds as large as the frame
        f = np.zeros((1, 352, 1216), np.float32)
        f[0, 350, 3] = 4.0
        f[0, 351, 1200] = 6.0
        _check_frames(handle, f, 0.1, 0.1)
    finally:
        handle.set_band_cap(-1)


@pytest.mark.parametrize("sky", [0, 8])
def test_planner_tiles_cover_every_pixel_once(handle, sky):
    """The tiles of a frame (plus, with sky > 0, the rows handed to k3_sky) partition its pixels (written regions
    disjoint and complete), each written region lies inside its sub-image, and KITTI-like frames do get split
    (rows and columns)."""
    for cap in (-1, 60, 150):
        handle.set_band_cap(cap)
        handle.set_sky_min(sky)
        handle.set_subbatches(1)          # one task list for the whole batch (frame indices are per sub-batch)
        x = np.stack([synth.kitti_frame(800 + i, beam_step=(1, 8)[i % 2]) for i in range(4)])
        handle.run_host(x, 0.1, 0.1)
        t = handle.debug_tasks()
        handle.set_band_cap(-1)
        handle.set_sky_min(-1)
        handle.set_subbatches(-1)
        for b in range(4):
            cover = np.zeros((352, 1216), np.int32)
            tb = t[t[:, 0] == b]
            S = max(0, int(tb[:, 11].max()))          # rows [0,S) are k3_sky's (closed form above the first source row)
            first_src = int(np.argmax((x[b] > 0.9).any(axis=1)))
            if sky:
                assert S % 4 == 0 and S + 1 <= first_src and S >= first_src - 4
            else:
                assert S == 0
            cover[:S] += 1
            for (_, lo, hi, r0, r1, kind, _, fstart, clo, c0, c1, sk) in tb:
                assert kind in (0, 4) and S <= lo <= r0 < r1 <= hi <= 352 and lo <= fstart < hi
                assert sk == (S if (S > 0 and r0 == S) else -2) and (sk < 0 or r1 >= S + 2)
                width = 1216 if kind == 0 else 640
                assert clo % 4 == 0 and clo <= c0 < c1 <= min(1216, clo + width)
                cover[r0:r1, c0:c1] += 1
            assert np.all(cover == 1)
        if cap != 0:
            assert len(t) > 4 and np.any(t[:, 5] == 4), "KITTI frames should be tiled in rows and columns"


@pytest.mark.parametrize("W", [1216, 1215, 640, 333, 36])
def test_rows_above_first_source_closed_form(handle, W):
    """k3_sky: rows above the first source row come from two rows of the scan, not from the scan itself.  Any
    height of the source-free top (aligned or not with the planner's cells), plateaus and single sources in the
    first source row, label and distance outputs on or off -- all bit-exact against the oracle."""
    rng = np.random.default_rng(W)
    H = 96
    handle.set_band_cap(40)
    handle.set_sky_min(8)
    try:
        frames = []
        for top in (1, 2, 7, 8, 9, 12, 13, 31, 32, 33, 60, 90, 94):
            dens = rng.choice([0.01, 0.1, 0.6])
            f = ((rng.random((H, W)) < dens) * rng.uniform(1, 50, (H, W))).astype(np.float32)
            f[:top] = 0
            f[top] = 0
            kind = top % 3
            if kind == 0:
                f[top, rng.integers(0, W)] = 2.0                       # one source: slopes as long as the row
            elif kind == 1:
                f[top, ::max(1, W // 7)] = 3.0                         # sparse comb
            else:
                f[top, W // 3: W // 2 + 1] = 4.0                       # a plateau of sources
            frames.append(f)
        x = np.stack(frames)
        r = _check_frames(handle, x, 0.1, 0.1)
        t = handle.debug_tasks()
        if W % 4 == 0 or W > 640:
            assert (t[:, 11] >= 8).any(), "some of these frames must have used the closed form"
        handle.set_sky_min(0)
        r0 = handle.run_host(x, 0.1, 0.1, want_dt=True, want_lbl=True)
        assert not (handle.debug_tasks()[:, 11] > 0).any()
        for k in ("depth", "dt", "lbl"):
            assert np.array_equal(r[k], r0[k])
    finally:
        handle.set_band_cap(-1)
        handle.set_sky_min(-1)


@pytest.mark.parametrize("W", [660, 800, 1000, 1212, 1216, 400, 500, 592])
def test_half_width_tiles_are_exact(handle, W):
    """Bands split into two overlapping half-width tiles (narrow kernel instance) whenever the distance bound
    allows it: dense-ish frames split, sparse ones do not; both must match the oracle bit for bit."""
    rng = np.random.default_rng(W)
    handle.set_band_cap(60)
    try:
        for dens in (0.3, 0.08, 0.02, 0.004):
            x = ((rng.random((2, 96, W)) < dens) * rng.uniform(1, 50, (2, 96, W))).astype(np.float32)
            x[:, 40, W // 2] = 2.0
            x[0, 20:60, W // 2 - 40: W // 2 + 40] = 0.0          # a hole right at the tile seam
            _check_frames(handle, x, 0.1, 0.1)
    finally:
        handle.set_band_cap(-1)


@pytest.mark.parametrize("nsub", [1, 2, 3, 8])
def test_subbatch_streams(handle, nsub):
    """Sub-batches on forked streams: same results, and a bad frame is reported with its index in the batch."""
    handle.set_subbatches(nsub)
    try:
        x = np.stack([synth.kitti_frame(700 + i, beam_step=(1, 2, 4, 8)[i % 4])[100:228, :640] for i in range(11)])
        _check_frames(handle, np.ascontiguousarray(x), 0.1, 0.1)
        x[9] = 0.0
        x[10] = 0.0
        r = handle.run_host(np.ascontiguousarray(x), 0.1, 0.1)
        assert "index_error" in r and r["first_bad"] == 9
    finally:
        handle.set_subbatches(-1)


def test_golden_full_frames(handle, golden_dir):
    z = np.load(os.path.join(golden_dir, "full_frames.npz"))
    for name, x, thr in (("kitti64_seed0", synth.kitti_frame(0), 0.1),
                         ("kitti8_seed5", synth.kitti_frame(5, beam_step=8), 0.1),
                         ("nyu_seed3", synth.nyu_frame(3), 0.001)):
        r = handle.run_host(np.ascontiguousarray(x[None]), thr, 0.1, want_dt=True, want_lbl=True)
        assert np.array_equal(r["lbl"][0], z[name + "/lbl"]), name
        assert np.array_equal(r["dt"][0].astype(np.uint16), z[name + "/dt_u16"]), name


def test_golden_checksums_all_configs(handle, golden_dir):
    """Reference outputs (tools.DT_complete_batch / eval_NYU.Distance_Transform run in the build container)
    for 64/32/16/8-beam KITTI frames and NYU frames, compared by SHA-256."""
    z = np.load(os.path.join(golden_dir, "checksums.npz"))
    for step in (1, 2, 4, 8):
        xb = synth.kitti_batch(range(4), beam_step=step)
        depth = tools.DT_complete_batch(xb)
        assert depth.shape == (4, 352, 1216, 1) and depth.dtype == np.float32
        for seed in range(4):
            dt, lbl = tools.nearest_point(xb[seed])
            want = z[f"kitti_b{64 // step}_s{seed}"]
            assert [sha(depth[seed, :, :, 0]), sha(dt), sha(lbl)] == list(want), (step, seed)
    for seed in range(4):
        x = synth.nyu_frame(seed)
        d = eval_nyu.Distance_Transform(x[None, :, :, None])
        dt, lbl = eval_nyu.nearest_point(x)
        assert d.dtype == np.float32 and d.shape == (480, 640)
        assert [sha(d), sha(dt), sha(lbl)] == list(z[f"nyu_s{seed}"]), seed
    x = synth.nyu_frame(9, 240, 320)
    dt, lbl = eval_nyu.nearest_point(x)
    assert [sha(eval_nyu.Distance_Transform(x)), sha(dt), sha(lbl)] == list(z["nyu240_s9"])


def test_batch_vs_oracle_kitti(handle):
    x = np.stack([synth.kitti_frame(100 + i, beam_step=s) for i, s in enumerate((1, 1, 2, 4, 8, 1))])
    _check_frames(handle, x, 0.1, 0.1)


def test_batch_vs_oracle_nyu(handle):
    x = np.stack([synth.nyu_frame(200 + i) for i in range(4)])
    _check_frames(handle, x, 0.001, 0.1)


def test_reference_quirks(handle):
    adv = synth.adversarial_frames()
    # valid pixels but no source: lbl == 0 -> whole frame is the LAST valid depth, dt == 65533 (Appendix B)
    f = adv["valid_no_source"]
    r = handle.run_host(np.ascontiguousarray(f[None]), 0.1, 0.1, want_dt=True, want_lbl=True)
    assert np.all(r["lbl"] == 0) and np.all(r["dt"] == 65533.0) and np.all(r["depth"] == np.float32(0.25))
    # no valid pixel at all -> IndexError from both fill functions (tools.py:26)
    z = np.zeros((1, 352, 1216, 1), np.float32)
    with pytest.raises(IndexError):
        tools.DT_complete_batch(z)
    with pytest.raises(IndexError):
        eval_nyu.Distance_Transform(z[0, :, :, 0])
    # only the second frame is bad: IndexError names it, like the reference's per-frame loop would hit it
    xb = synth.kitti_batch([0, 1])
    xb[1] = 0
    with pytest.raises(IndexError, match="frame 1"):
        tools.DT_complete_batch(xb)
    # exactly one valid pixel: tools.py:24 handles it, eval_NYU.py:126's squeeze makes it an IndexError
    one = np.zeros((352, 1216), np.float32)
    one[10, 10] = 5.0
    out = tools.DT_complete_batch(one[None, :, :, None])
    assert np.all(out == 5.0)
    with pytest.raises(IndexError):
        eval_nyu.Distance_Transform(one)
    # NaN is a source (1 - nan > thr is False) but not valid: more sources than valid depths -> IndexError
    f = np.zeros((1, 16, 16), np.float32)
    f[0, 3, 3] = np.nan
    r = handle.run_host(f, 0.1, 0.1)
    assert "index_error" in r and r["first_bad"] == 0
    # extra channels are ignored (tools.py:19 reads channel 0)
    xb2 = np.concatenate([synth.kitti_batch([3]), np.full((1, 352, 1216, 1), 9.0, np.float32)], axis=-1)
    assert np.array_equal(tools.DT_complete_batch(xb2), tools.DT_complete_batch(xb2[..., :1]))
    # input untouched, output is a new array
    xb3 = synth.kitti_batch([4])
    keep = xb3.copy()
    out = tools.DT_complete_batch(xb3)
    assert np.array_equal(xb3, keep) and not np.shares_memory(out, xb3)


def test_idempotence_and_source_preservation(handle):
    """Size-independent properties at full size: sources keep their own depth and label; filling a filled
    frame changes nothing (every pixel is then its own source)."""
    x = np.stack([synth.kitti_frame(300 + i) for i in range(8)])
    r = handle.run_host(x, 0.1, 0.1, want_dt=True, want_lbl=True)
    src = x >= 1.0
    assert np.array_equal(r["depth"][src], x[src])
    assert np.all(r["dt"][src] == 0) and np.all(r["dt"][~src] >= 1)
    for b in range(x.shape[0]):
        assert np.array_equal(r["lbl"][b][src[b]], np.arange(1, int(src[b].sum()) + 1))
    r2 = handle.run_host(r["depth"], 0.1, 0.1, want_dt=True)
    assert np.array_equal(r2["depth"], r["depth"]) and np.all(r2["dt"] == 0)


def test_metrics_golden_and_oracle(handle, golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    for seed in range(3):
        x = synth.kitti_frame(seed)
        fill = tools.DT_complete_batch(x[None, :, :, None])[0, :, :, 0]
        gt = synth.kitti_gt(seed)
        R = evaluation.Result()
        assert R.evaluate(fill, gt) is None
        np.testing.assert_array_equal([R.mse, R.rmse, R.mae, R.irmse, R.imae], z[f"kitti_s{seed}"])
        R.evaluate(np.maximum(fill, np.float32(0.9)), gt.astype(np.float32))
        np.testing.assert_array_equal([R.mse, R.rmse, R.mae, R.irmse, R.imae], z[f"kitti_f32gt_s{seed}"])
        xn, g = synth.nyu_frame(seed, return_dense=True)
        fill = eval_nyu.Distance_Transform(xn)
        R = evaluation.Result_NYU()
        R.evaluate(fill, g)
        np.testing.assert_array_equal([R.mse, R.rmse, R.mae, R.irmse, R.imae, R.delta1, R.delta2, R.delta3],
                                      z[f"nyu_s{seed}"])
    # empty valid set -> nan like numpy's mean of an empty array
    R = evaluation.Result()
    R.evaluate(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float64))
    assert np.isnan(R.rmse) and np.isnan(R.mae)


def test_metrics_batch_sums(handle):
    B = 6
    fills = np.stack([O.dt_fill(synth.kitti_frame(400 + i))["depth"] for i in range(B)])
    gts = np.stack([synth.kitti_gt(400 + i) for i in range(B)])
    per_frame, sums = evaluation.evaluate_batch(fills, gts, _lib.METRICS_KITTI)
    want = np.array([[m["mse"], m["rmse"], m["mae"], m["irmse"], m["imae"], 0, 0, 0, m["count"]]
                     for m in (O.result_kitti(fills[i], gts[i]) for i in range(B))])
    np.testing.assert_array_equal(per_frame, want)              # numpy's summation order, bit for bit
    np.testing.assert_allclose(sums[:9], want.sum(axis=0), rtol=1e-14)
    assert sums[9] == B


def test_metrics_bit_exact_over_sizes(handle):
    """dtfill_metrics against evaluation.py restated with numpy (oracle.result_*), ==: valid counts below 8, up to 128
    (one leaf of numpy's pairwise sum), just above, odd, and frame-sized; float32 and float64 ground truth; both modes;
    and the one-pass mode within its stated tolerance."""
    rng = np.random.default_rng(5)
    for n_px, dens in ((64, 0.05), (200, 0.5), (300, 0.45), (1000, 0.9), (4099, 0.3), (12345, 1.0), (352 * 1216, 0.2),
                       (480 * 640, 1.0)):
        B = 3
        out = rng.uniform(0.5, 80.0, (B, n_px)).astype(np.float32)
        gt64 = rng.uniform(0.5, 80.0, (B, n_px)) * (rng.random((B, n_px)) < dens)
        for gt in (gt64, gt64.astype(np.float32)):
            for mode, ref in ((_lib.METRICS_KITTI, O.result_kitti), (_lib.METRICS_NYU, O.result_nyu)):
                per_frame, sums = evaluation.evaluate_batch(out, gt, mode)
                for b in range(B):
                    m = ref(out[b], gt[b])
                    want = [m["mse"], m["rmse"], m["mae"], m["irmse"], m["imae"], m.get("delta1", 0), m.get("delta2", 0),
                            m.get("delta3", 0), m["count"]]
                    np.testing.assert_array_equal(per_frame[b], want, err_msg=f"{n_px} {gt.dtype} mode {mode} frame {b}")
    handle.set_metrics_exact(False)
    try:
        per_frame, _ = evaluation.evaluate_batch(out, gt64, _lib.METRICS_KITTI)
        m = O.result_kitti(out[0], gt64[0])
        np.testing.assert_allclose(per_frame[0, :5], [m["mse"], m["rmse"], m["mae"], m["irmse"], m["imae"]], rtol=METRIC_RTOL_F64)
    finally:
        handle.set_metrics_exact(True)


def test_pipelined_mode_matches_strict(handle):
    """Pipeline depth 2: consecutive device-resident calls overlap on two internal streams / workspaces; every call's
    outputs must equal the strict-mode outputs once flushed, and a bad frame in any in-flight call is reported."""
    import torch
    from distancetransform_depthcompletion_b200.engine import DTFillEngine
    strict = DTFillEngine(0)
    piped = DTFillEngine(0, pipeline_depth=3)
    batches = [torch.from_numpy(np.stack([synth.kitti_frame(900 + 10 * k + i, beam_step=(1, 2, 8)[k % 3])
                                          for i in range(3)])).cuda() for k in range(5)]
    want = []
    for xb in batches:
        o = strict.fill(xb, want_lbl=True)
        strict.status()
        want.append({k: v.clone() for k, v in o.items() if v is not None})
    got = [piped.fill(xb, want_lbl=True) for xb in batches]          # five calls in flight, two at a time
    bad, _ = piped.status()
    assert bad == -1
    for w, g in zip(want, got):
        for k in ("depth", "dt", "lbl", "mask", "counts"):
            assert torch.equal(w[k], g[k]), k
    xbad = batches[0].clone()
    xbad[1] = 0
    piped.fill(xbad)
    piped.fill(batches[1])
    bad, _ = piped.status()
    assert bad == 1


def test_device_resident_engine_matches_host_path(handle):
    import torch
    from distancetransform_depthcompletion_b200.engine import DTFillEngine
    eng = DTFillEngine(0)
    x = np.stack([synth.kitti_frame(500 + i) for i in range(4)])
    xd = torch.from_numpy(x).cuda()
    out = eng.fill(xd, want_lbl=True)
    bad, launches = eng.status()
    assert bad == -1 and launches >= 3
    o = O.dt_fill(x)
    assert np.array_equal(out["depth"].cpu().numpy(), o["depth"])
    assert np.array_equal(out["dt"].cpu().numpy(), o["dt"])
    assert np.array_equal(out["lbl"].cpu().numpy(), o["lbl"])
    assert np.array_equal(out["mask"].cpu().numpy(), o["mask"])
    gt = torch.from_numpy(np.stack([synth.kitti_gt(500 + i) for i in range(4)])).cuda()
    per_frame, sums = eng.metrics(out["depth"], gt)
    want = [O.result_kitti(o["depth"][i], gt[i].cpu().numpy())["rmse"] for i in range(4)]
    np.testing.assert_allclose(per_frame[:, 1].cpu().numpy(), want, rtol=METRIC_RTOL_F64)


@pytest.mark.parametrize("W,crop", [(1216, 96), (640, 0), (648, 5), (100, 3), (37, 1)])
def test_uint16_png_input_decode_and_crop(handle, W, crop):
    """uint16 PNG samples in, decode (/256, data_read.py:215) + top crop (train.py:211) inside K1.
    Must equal the float path fed with the reference's own decode, bit for bit, in every output."""
    rng = np.random.default_rng(W + crop)
    Hin = 70 + crop
    B = 5
    png = ((rng.random((B, Hin, W)) < 0.06) * rng.integers(1, 65536, (B, Hin, W))).astype(np.uint16)
    png[:, crop + 10, W // 2] = 65535                    # largest sample
    png[0, crop + 11, 0] = 25                            # 0.0977: neither valid nor source
    png[0, crop + 12, 1] = 26                            # 0.1016: valid (> 0.1) but not a source
    png[1, crop + 13, 2] = 230                           # 0.8984: valid, not a source
    png[1, crop + 13, 3] = 231                           # 0.9023: source
    png[:, :crop] = 40000                                # the cropped rows must not leak into the result
    lidar = (png.astype(np.float32) / np.float32(256.0))[:, crop:]          # the reference's decode + crop
    lidar = np.ascontiguousarray(lidar)
    ref = handle.run_host(lidar, 0.1, 0.1, want_dt=True, want_lbl=True, want_mask=True)
    o = O.dt_fill(lidar, 0.1, 0.1)
    r = handle.run_host_u16(png, crop, 0.1, 0.1, want_lidar=True, want_dt=True, want_lbl=True, want_mask=True)
    assert "index_error" not in r and "index_error" not in ref
    assert np.array_equal(r["lidar"].view(np.uint32), lidar.view(np.uint32))
    for k in ("depth", "dt", "lbl", "mask", "counts"):
        assert np.array_equal(r[k], ref[k]), k
    assert np.array_equal(r["lbl"], o["lbl"]) and np.array_equal(r["depth"].view(np.uint32), o["depth"].view(np.uint32))
    r2 = handle.run_host_u16(png, crop, 0.1, 0.1)        # without the optional outputs
    assert np.array_equal(r2["depth"], r["depth"]) and r2["lidar"] is None


def test_uint16_png_drop_in_and_errors(handle):
    png = np.zeros((2, 448, 1216), np.uint16)
    x = synth.kitti_batch([40, 41])[..., 0]
    png[:, 96:] = np.round(x * 256).astype(np.uint16)
    lidar, refined = tools.DT_complete_batch_png(png)
    assert lidar.shape == refined.shape == (2, 352, 1216, 1) and refined.dtype == np.float32
    want = O.cv2_port_complete_batch(lidar)
    assert np.array_equal(refined.view(np.uint32), want.view(np.uint32))
    with pytest.raises(ValueError):
        tools.DT_complete_batch_png(png, crop_top=95)
    with pytest.raises(TypeError):
        tools.DT_complete_batch_png(png.astype(np.int32))
    with pytest.raises(IndexError):
        tools.DT_complete_batch_png(np.zeros((1, 448, 1216), np.uint16))     # no valid pixel (tools.py:26)


def test_uint16_png_device_resident_pipelined(handle):
    """The uint16 entry on device tensors with batches in flight (k3_sky active) against the float path."""
    import torch
    from distancetransform_depthcompletion_b200.engine import DTFillEngine
    x = synth.kitti_batch([70, 71, 72])[..., 0]
    png = np.zeros((3, 448, 1216), np.uint16)
    png[:, 96:] = np.round(x * 256).astype(np.uint16)
    want = handle.run_host(np.ascontiguousarray(x), 0.1, 0.1, want_dt=True, want_lbl=True, want_mask=True)
    eng = DTFillEngine(0, pipeline_depth=3)
    pd = torch.from_numpy(png.view(np.int16)).cuda().view(torch.uint16)
    outs = [eng.fill_png(pd, want_lbl=True) for _ in range(4)]
    eng.flush()
    bad, _ = eng.status()
    assert bad == -1
    for o in outs:
        assert np.array_equal(o["lidar"].cpu().numpy(), x)
        for k in ("depth", "dt", "lbl", "mask"):
            assert np.array_equal(o[k].cpu().numpy(), want[k]), k
        assert np.array_equal(o["counts"].cpu().numpy(), want["counts"])


def test_pipelined_random_shapes_against_oracle(handle):
    """Batches in flight (k3_sky active where a frame has source-free top rows) over random sizes, densities and
    top gaps, labels included: every output against the oracle."""
    import torch
    from distancetransform_depthcompletion_b200.engine import DTFillEngine
    rng = np.random.default_rng(2024)
    eng = DTFillEngine(0, pipeline_depth=4)
    eng.handle.set_band_cap(48)
    pending = []
    for t in range(24):
        H, W = int(rng.integers(20, 200)), int(rng.integers(8, 1217))
        if t % 3 == 0:
            W = (W // 16) * 16 + 16
        dens = float(rng.choice([0.004, 0.03, 0.2, 0.7]))
        top = int(rng.integers(0, H - 2))
        x = ((rng.random((3, H, W)) < dens) * rng.uniform(1, 60, (3, H, W))).astype(np.float32)
        x[:, :top] = 0
        x[:, top, rng.integers(0, W)] = 4.0
        xd = torch.from_numpy(x).cuda()
        pending.append((x, xd, eng.fill(xd, want_lbl=True)))
        if len(pending) == 4:
            eng.flush()
            assert eng.status()[0] == -1
            for x_, _, o in pending:
                want = O.dt_fill(x_, 0.1, 0.1)
                assert np.array_equal(o["lbl"].cpu().numpy(), want["lbl"]), x_.shape
                assert np.array_equal(o["dt"].cpu().numpy(), want["dt"]), x_.shape
                assert np.array_equal(o["depth"].cpu().numpy().view(np.uint32), want["depth"].view(np.uint32)), x_.shape
                assert np.array_equal(o["mask"].cpu().numpy(), want["mask"]), x_.shape
            pending = []


@pytest.mark.gpu
def test_strict_random_shapes_with_two_scan_instances(handle):
    """Strict order on frames wide enough for both scan instances (half-width tiles on the call's stream, full-width tasks
    and -- for small calls -- k3_sky on the side stream): random sizes, widths that are and are not multiples of 4 or 16,
    densities from a few sources to dense, source-free tops of any height, 1..12 frames per call, the automatic band
    target and small explicit ones, labels on and off.  Every output against the oracle."""
    import torch
    from distancetransform_depthcompletion_b200.engine import DTFillEngine
    rng = np.random.default_rng(77)
    eng = DTFillEngine(0)
    split_seen = False
    try:
        for t in range(30):
            H, W = int(rng.integers(24, 353)), int(rng.integers(660, 1217))
            if t % 3:
                W = (W // 16) * 16 if t % 3 == 1 else (W // 4) * 4
            B = int(rng.integers(1, 13))
            dens = float(rng.choice([0.002, 0.02, 0.05, 0.3]))
            top = int(rng.integers(0, H - 10))
            x = ((rng.random((B, H, W)) < dens) * rng.uniform(1, 60, (B, H, W))).astype(np.float32)
            x[:, :top] = 0
            x[:, top, rng.integers(0, W)] = 4.0
            if t % 5 == 0:
                x[0, top + 1:] = 0                      # a frame with one source in all
            eng.handle.set_band_cap(int(rng.choice([-1, -1, 40, 64, 120])))
            want_lbl = bool(t % 2)
            o = eng.fill(torch.from_numpy(x).cuda(), want_lbl=want_lbl)
            assert eng.status()[0] == -1
            want = O.dt_fill(x, 0.1, 0.1)
            for k in ("depth", "dt", "mask") + (("lbl",) if want_lbl else ()):
                assert np.array_equal(o[k].cpu().numpy(), want[k]), (t, x.shape, k)
            tk = eng.handle.debug_tasks(1 << 16)
            split_seen |= bool(((tk[:, 5] == 0) & (tk[:, 11] >= 8)).any() and (tk[:, 5] == 4).any())
    finally:
        eng.handle.set_band_cap(-1)
    assert split_seen, "some call must have run k3_sky's base rows as a full-width task next to half-width tiles"

"""Host-side logic that needs no GPU: argument contracts of the drop-in functions, sharding arithmetic, the
rule that the product never touches oracle/."""
import os
import re

import numpy as np
import pytest

from distancetransform_depthcompletion_b200 import sharding, synth, tools, eval_nyu, evaluation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dt_complete_batch_shape_errors():
    with pytest.raises(ValueError):                       # tools.py:25-27 hard-coded 352 x 1216
        tools.DT_complete_batch(np.ones((1, 100, 100, 1), np.float32))
    with pytest.raises(IndexError):                       # tools.py:19 indexes four axes
        tools.DT_complete_batch(np.ones((352, 1216), np.float32))
    with pytest.raises(TypeError):                        # float64 and integers take the float64 route; float16 has none
        tools.DT_complete_batch(np.ones((1, 352, 1216, 1), np.float16))


def test_nearest_point_and_distance_transform_shape_errors():
    with pytest.raises(ValueError):
        tools.nearest_point(np.ones((2, 3, 4), np.float32))
    with pytest.raises(ValueError):
        eval_nyu.Distance_Transform(np.ones((2, 3, 4), np.float32))


def test_evaluate_contract_errors():
    with pytest.raises(IndexError):
        evaluation.Result().evaluate(np.ones((3, 3), np.float32), np.ones((3, 4), np.float32))
    with pytest.raises(TypeError):
        evaluation.Result().evaluate(np.ones((3, 3), np.float64), np.ones((3, 3), np.float64))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 8192, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_finalize_means():
    s = np.array([2.0, 4.0, 6.0, 8.0, 10.0, 1.0, 1.5, 2.0, 1234.0, 2.0])
    m = sharding.finalize_means(s)
    assert m["rmse"] == 2.0 and m["imae"] == 5.0 and m["frames"] == 2 and m["valid_pixels"] == 1234.0


def test_synth_inputs_match_the_configs():
    x = synth.kitti_frame(0)
    assert x.shape == (352, 1216) and x.dtype == np.float32
    d = float((x > 0).mean())
    assert 0.045 < d < 0.056                               # ~5 % density (BASELINE configs[0])
    assert x[x > 0].min() >= 1.0                           # source and valid predicates agree
    assert np.all(x * 256 == np.rint(x * 256))             # KITTI grid k/256 (data_read.py:215)
    for step, frac in ((2, 0.5), (4, 0.25), (8, 0.125)):
        assert abs((synth.kitti_frame(0, beam_step=step) > 0).sum() / (x > 0).sum() - frac) < 0.03
    n = synth.nyu_frame(0)
    assert n.shape == (480, 640) and 450 <= int((n > 0).sum()) <= 500
    assert synth.kitti_gt(0).dtype == np.float64


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distancetransform_depthcompletion_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|dtfill_oracle|libdtfill_oracle", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference the oracle: {bad}"


def test_numpy_pairwise_sum_restated():
    """The summation order dtfill_metrics reproduces (csrc/dtfill_k4_exact.cuh) IS numpy's: the restatement in
    tests/kernel_model.py equals np.add.reduce and np.mean bit for bit, float32 and float64, over the length classes of
    the algorithm (below 8, one block of up to 128, splits rounded to multiples of 8)."""
    from kernel_model import numpy_pairwise_sum
    rng = np.random.default_rng(0)
    for T in (np.float32, np.float64):
        for n in list(range(1, 20)) + [127, 128, 129, 136, 255, 256, 257, 1000, 4097, 12345, 86000]:
            a = (rng.random(n) * 1000).astype(T)
            s = numpy_pairwise_sum(a)
            assert s == np.add.reduce(a), (T.__name__, n)
            assert T(s / T(n)) == np.mean(a), (T.__name__, n)


def _check_compaction():
    import ctypes  # noqa: F401
    from distancetransform_depthcompletion_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(4)
    for n, dens, thr in ((0, 0.0, 0.1), (5, 0.5, 0.1), (31, 0.3, 0.1), (32, 1.0, 0.1), (63, 0.2, 0.1), (64, 0.02, 0.1), (81, 1.0, 0.1),
                         (1000, 0.05, 0.1), (70001, 0.05, 0.001), (4096, 0.0, 0.1), (4099, 0.9, 0.1)):
        x = np.zeros(n, np.float32)
        m = rng.random(n) < dens
        x[m] = rng.choice([0.05, 0.1, 0.5, 0.9, np.float32(0.90000004), 0.999, 1.0, 37.25, -3.0, np.nan, np.inf, 1e-41, -0.0], m.sum())
        idx = np.zeros(n + 16, np.uint32); val = np.zeros(n + 16, np.uint32)
        k = L.dtfill_debug_compact(x.ctypes.data, n, thr, 0.1, idx.ctypes.data, val.ctypes.data, n + 16)
        with np.errstate(invalid="ignore"):
            keep = ~((np.float32(1.0) - x) > np.float32(thr)) | (x > np.float32(0.1))
        want = np.nonzero(keep)[0]
        assert k == len(want), (n, dens, thr)
        assert np.array_equal(idx[:k], want)
        assert np.array_equal(val[:k], x.view(np.uint32)[want])
    assert L.dtfill_debug_compact(x.ctypes.data, n, 0.1, 0.1, idx.ctypes.data, val.ctypes.data, n) == -1     # no slack


@pytest.mark.parametrize("isa", ["", "avx2", "scalar"])
def test_sparse_upload_compaction_on_the_host(isa):
    """The host half of the sparse upload (csrc/dtfill.cu compact_block: AVX-512 compress, AVX2 left-packing or scalar,
    whichever the CPU has; DTFILL_COMPACT_ISA caps it, read once per process, hence the subprocess) needs no GPU: it
    must keep exactly the pixels that are a source (tools.py:8, float32) or valid (tools.py:22), in order, with their bits."""
    import subprocess
    import sys
    env = dict(os.environ, DTFILL_COMPACT_ISA=isa) if isa else {k: v for k, v in os.environ.items() if k != "DTFILL_COMPACT_ISA"}
    code = "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_host_logic as t; t._check_compaction()" % (
        ROOT, os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on the host alone and prints ONE JSON line with the keys the driver reads
    (impl, metric, value, unit, cpu_baseline.kind/cores/sample, e2e with zero copy bytes); here it finds the reference
    checkout when there is one and the cv2 port otherwise."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--batch", "4"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0

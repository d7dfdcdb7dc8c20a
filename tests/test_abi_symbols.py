"""libdtfill.so loads and exports every symbol include/dtfill.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dtfill.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dtfill_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("dtfill_create", "dtfill_destroy", "dtfill_run", "dtfill_run_async", "dtfill_status", "dtfill_metrics",
              "dtfill_last_error", "dtfill_host_alloc", "dtfill_host_free", "dtfill_set_stream",
              "dtfill_synchronize", "dtfill_abi_version", "dtfill_set_profiling", "dtfill_kernel_times", "dtfill_set_band_cap", "dtfill_set_sky_min", "dtfill_run_u16", "dtfill_run_u16_async", "dtfill_set_subbatches", "dtfill_debug_get_tasks", "dtfill_set_pipeline_depth", "dtfill_flush", "dtfill_dt_pool", "dtfill_dt_pool_ex", "dtfill_debug_read_status", "dtfill_outlier_removal",
              "dtfill_metrics_ex", "dtfill_allreduce_sums", "dtfill_comm_create", "dtfill_comm_destroy",
              "dtfill_nccl_unique_id", "dtfill_set_stage_threads"):
        assert s in syms


def test_library_exports_every_declared_symbol(dtfill_lib):
    L = ctypes.CDLL(dtfill_lib)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/dtfill.h but not exported by libdtfill.so"
    L.dtfill_abi_version.restype = ctypes.c_int
    hdr = open(os.path.join(ROOT, "include", "dtfill.h")).read()
    assert L.dtfill_abi_version() == int(re.search(r"#define DTFILL_ABI_VERSION (\d+)", hdr).group(1))
    from distancetransform_depthcompletion_b200 import _lib
    assert L.dtfill_abi_version() == _lib.ABI_VERSION


def test_binding_declares_every_symbol(dtfill_lib):
    from distancetransform_depthcompletion_b200 import _lib
    L = _lib.load()
    for s in declared_symbols():
        assert getattr(L, s) is not None


def test_no_cpu_fallback_without_gpu(dtfill_lib):
    """Without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from distancetransform_depthcompletion_b200 import _lib, tools
    with pytest.raises(_lib.DTFillError):
        tools.nearest_point(np.ones((4, 4), np.float32))

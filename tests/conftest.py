import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests skip (instead of erroring) on a box without a CUDA device"""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import ctypes
        n = ctypes.CDLL(os.path.join(ROOT, "opensmile_b200", "libosm_b200.so")).osm_b200_device_count()
    except OSError:
        n = 0
    if n <= 0:
        skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU path)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build what is missing (libosm_b200.so, liboracle.so); both are compiled in-tree."""
    lib = os.path.join(ROOT, "opensmile_b200", "libosm_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "opensmile_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


def rel_to_frame_scale(got, ref):
    """Parity metric: |got - ref| relative to the per-frame vector scale
    max_j |ref[t, j]|.  Elementwise-relative 1e-5 is unattainable even for a float64
    restatement of the reference (near-zero cepstra), so 1e-5 is taken against the scale."""
    import numpy as np
    scale = np.abs(ref).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    return float((np.abs(got - ref) / scale).max())


def column_scale_report(got, ref):
    """Per-column parity: |got - ref| relative to max_t |ref[t, column]| -- delta columns of magnitude ~1 are not
    hidden behind a c0 of ~60.  Returns (largest error, share of values beyond 1e-5).  The rule the tests and bench.py apply to the
    MFCC / PLP graphs: no value beyond 5e-5 and at most 0.1 % beyond 1e-5 (the own FFT differs from the reference's by ~2e-7 of the
    frame's spectral peak, which the logarithm of weak bands and the regression turn into a few 1e-5 of a delta-delta column on
    isolated frames)."""
    import numpy as np
    scale = np.abs(ref).max(axis=0, keepdims=True)
    scale[scale == 0] = 1.0
    err = np.abs(got - ref) / scale
    return float(err.max()), float((err > 1e-5).mean())


def assert_columns_close(got, ref):
    """the MFCC / PLP rule; on samples of fewer than 2 000 values "0.1 %" would be less than two values, so two are allowed there"""
    worst, share = column_scale_report(got, ref)
    assert worst < 5e-5 and share <= max(1e-3, 2.0 / max(got.size, 1)), (worst, share)

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def t(a, dtype=None, device="cpu"):
    """numpy -> torch"""
    x = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x.to(device)


# Tolerance of BASELINE.json:north_star for fp32 warps, costs, losses and gradients.
RTOL, ATOL = 1e-5, 1e-6


def assert_close(actual, expected, rtol=RTOL, atol=ATOL, what=""):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    err = np.abs(actual - expected)
    bound = atol + rtol * np.abs(expected)
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} outside rtol={rtol} atol={atol}; worst at {i}: "
                             f"got {actual[i]!r} want {expected[i]!r} (err {err[i]:.3e})")


def assert_close_or_better(actual, ref32, ref64, rtol=RTOL, atol=ATOL, what=""):
    """|actual - ref64| <= atol + rtol*|ref64|  OR  no further from the float64 reference than the
    reference's own float32 path is (summation-order noise of long fp32 reductions)."""
    actual = np.asarray(actual, dtype=np.float64)
    ref32 = np.asarray(ref32, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    err = np.abs(actual - ref64)
    bound = np.maximum(atol + rtol * np.abs(ref64), 2.0 * np.abs(ref32 - ref64))
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} outside tolerance; worst at {i}: got {actual[i]!r} "
                             f"f64 {ref64[i]!r} f32 {ref32[i]!r} (err {err[i]:.3e}, bound {bound[i]:.3e})")

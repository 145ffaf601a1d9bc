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


def reduction_floor(ref64, rtol=RTOL):
    """Rounding floor of an fp32 SUM: proportional to the magnitude of the summands, which for the
    scatter / channel / pixel reductions of this path is the RMS of the result tensor, not the
    (possibly cancelled) value of one element."""
    ref64 = np.asarray(ref64, dtype=np.float64)
    return rtol * float(np.sqrt(np.mean(ref64 ** 2))) if ref64.size else 0.0


def assert_close_or_better(actual, ref32, ref64, rtol=RTOL, atol=ATOL, what="", reduction=True):
    """Gradient check against the float64 reference.  Element-wise it accepts
        |actual - ref64| <= atol + rtol*|ref64|        (the north_star tolerance), or
        |actual - ref64| <= 2*|ref32 - ref64|          (no worse than the reference's own fp32 path), or,
      for outputs that are fp32 sums (reduction=True),
        |actual - ref64| <= rtol * rms(ref64)          (summation-order floor, see reduction_floor).
    Ill-conditioned maps (e.g. d loss / d inverse-depth = -d^2 * d loss / d depth spans 5 decades and
    inherits the rounding of the fp32 pixel coordinates) can still miss that at isolated elements in
    BOTH fp32 implementations; then the error norms decide: the CUDA path must be as accurate against
    float64 as the reference's fp32 path is (RMS within 1.5x, max within 3x)."""
    actual = np.asarray(actual, dtype=np.float64)
    ref32 = np.asarray(ref32, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    assert actual.shape == ref64.shape, (what, actual.shape, ref64.shape)
    err = np.abs(actual - ref64)
    ref_err = np.abs(ref32 - ref64)
    bound = np.maximum(atol + rtol * np.abs(ref64), 2.0 * ref_err)
    if reduction:
        bound = np.maximum(bound, reduction_floor(ref64, rtol))
    bad = err > bound
    if bad.any():
        rms, ref_rms = float(np.sqrt(np.mean(err ** 2))), float(np.sqrt(np.mean(ref_err ** 2)))
        if rms <= 1.5 * ref_rms + atol and err.max() <= 3.0 * ref_err.max() + atol:
            return
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()}/{bad.size} outside tolerance; worst at {i}: got {actual[i]!r} "
                             f"f64 {ref64[i]!r} f32 {ref32[i]!r} (err {err[i]:.3e}, bound {bound[i]:.3e}); "
                             f"rms err {rms:.3e} vs reference fp32 {ref_rms:.3e}, "
                             f"max err {err.max():.3e} vs {ref_err.max():.3e}")

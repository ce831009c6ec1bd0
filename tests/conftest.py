import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: 1e-5 relative (fp32) for states, observations, rewards and advantages, with a
# small absolute floor because many observation entries are differences that sit near zero
# (SURVEY.md section 8d "Parity tolerances").  For 2-D arrays whose rows are vectors rotated into the heading
# frame (observations) the floor scales with the row's largest magnitude: a rotated component near zero carries
# an absolute error of ~eps * |v| in ANY fp32 implementation (torch-CPU vs torch-CUDA differ the same way), so
# the floor is ATOL * max(1, max|row|) when row_scale=True.
RTOL = 1e-5
ATOL = 2e-6


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


def assert_close(a, b, rtol=RTOL, atol=ATOL, what="", row_scale=False):
    """|a-b| <= rtol*|b| + atol elementwise, with a useful message (b is the reference)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    floor = atol
    if row_scale and b.ndim >= 2 and b.size:
        floor = atol * np.maximum(1.0, np.nanmax(np.abs(b), axis=-1, keepdims=True))
    bad = ~(np.abs(a - b) <= rtol * np.abs(b) + floor)
    both_nan = np.isnan(a) & np.isnan(b)
    bad &= ~both_nan
    if bad.any():
        idx = np.argwhere(bad)
        err = np.abs(a - b)
        worst = np.unravel_index(np.nanargmax(np.where(bad, err, 0)), a.shape)
        raise AssertionError(
            f"{what}: {bad.sum()} of {a.size} outside rtol={rtol} atol={atol}; worst at {worst}: "
            f"got {a[worst]!r} want {b[worst]!r} (abs err {err[worst]:.3e}); first bad index {tuple(idx[0])}")


def assert_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    n = int((a != b).sum())
    assert n == 0, f"{what}: {n} of {a.size} differ (first at {tuple(np.argwhere(a != b)[0])})"


@pytest.fixture(scope="session")
def golden():
    return {n: load_npz(n + ".npz") for n in ("cmu_tables", "cmu_step", "synth_tables", "synth_step", "gae", "rms", "sample_time")}


@pytest.fixture(scope="session")
def golden_amp():
    return load_npz("amp.npz")

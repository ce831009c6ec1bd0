import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerance: 1e-5 relative (fp32) for states, observations, rewards and advantages, with a small absolute floor because
# many observation entries are differences that sit near zero (SURVEY.md section 8d "Parity tolerances").  For observation rows
# (``row_scale=True``) the floor of an element is ATOL * max(1, |v|_inf) where v is the 3- or 6-vector the element belongs to
# (oracle/device_parity.py obs_floor): a component of a rotated vector near zero carries ~eps * |v| of absolute error in ANY fp32
# implementation (torch-CUDA vs torch-CPU differ the same way), but metres, unit axes and rad/s of the same row never mix.
# profiles/r2_parity_report.json publishes the worst margins under the PLAIN rule (floor = ATOL) as well.
RTOL = 1e-5
ATOL = 2e-6

# (first column, vector width, vectors) per block, by row width
_SELF = ((0, 1, 1), (1, 3, 23), (70, 6, 24), (214, 3, 24), (286, 3, 24))
_TASK = ((0, 3, 24), (72, 6, 24), (216, 3, 24), (288, 3, 24), (360, 3, 24), (432, 6, 24))
_LAYOUTS = {
    934: _SELF + tuple((c + 358, w, n) for c, w, n in _TASK),
    358: _SELF,
    357: ((0, 3, 23), (69, 6, 24), (213, 3, 24), (285, 3, 24)),          # root_height_obs=False
    576: _TASK,
}


def vector_floor(b, atol=ATOL):
    """ATOL * max(1, max-magnitude of the 3-/6-vector each element of an observation row belongs to); plain ATOL for unknown widths."""
    lay = _LAYOUTS.get(b.shape[-1]) if b.ndim == 2 else None
    if lay is None:
        return atol
    a = np.abs(b)
    floor = np.full(b.shape, atol, dtype=np.float64)
    for c0, w, n in lay:
        vmax = a[:, c0:c0 + w * n].reshape(b.shape[0], n, w).max(axis=-1, keepdims=True)
        floor[:, c0:c0 + w * n] = np.broadcast_to(atol * np.maximum(1.0, vmax), (b.shape[0], n, w)).reshape(b.shape[0], n * w)
    return floor


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
        floor = vector_floor(b, atol)
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


@pytest.fixture(autouse=True)
def _reference_device_cpu():
    """The committed golden vectors were produced by the reference on torch-CPU, and the C oracle defaults to that flavour, so every
    test starts from the CPU flavour; tests against the reference on torch-CUDA (tests/test_reference_devices.py) switch explicitly."""
    import puffer_phc_b200
    from oracle import c_oracle
    prev = puffer_phc_b200.set_reference_device("cpu")
    c_oracle.set_ref_device("cpu")
    yield
    puffer_phc_b200.set_reference_device(prev)
    c_oracle.set_ref_device("cpu")


@pytest.fixture(scope="session")
def golden():
    return {n: load_npz(n + ".npz") for n in ("cmu_tables", "cmu_step", "synth_tables", "synth_step", "gae", "rms", "sample_time")}


@pytest.fixture(scope="session")
def golden_amp():
    return load_npz("amp.npz")

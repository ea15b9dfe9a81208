import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the reference modules (tests/golden/make_golden.py)."""
    with np.load(GOLDEN_PATH) as z:
        return {k: z[k] for k in z.files}


def assert_close(got, ref, rtol=1e-3, atol=1e-4, what=""):
    """The north-star tolerance: |got-ref| <= atol + rtol*|ref| elementwise (fp32)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    err = np.abs(got - ref)
    tol = atol + rtol * np.abs(ref)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err / tol), err.shape)
        raise AssertionError(
            f"{what}: {bad.sum()} of {bad.size} elements outside rtol={rtol} atol={atol}; worst at {i}: "
            f"got {got[i]} ref {ref[i]} (err {err[i]:.3e}, tol {tol[i]:.3e})")

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "python-billiards_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tests/golden/ref_tests/ holds the reference's own test files; they are run in a subprocess by
# tests/test_gpu_reference_suite.py (their import sets numpy's error state), never collected into this session
collect_ignore = [] if os.environ.get("BILLIARDS_COLLECT_REF_TESTS") else ["golden"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; never imported by the product package)."""
    from oracle import oracle as orc
    orc.lib()
    return orc


def bits(a):
    """float64 array -> uint64 view, for bit-exact comparisons that also treat NaNs/INFs literally"""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    bad = np.flatnonzero(bits(a).ravel() != bits(b).ravel())
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} values differ, first at {bad[:5]}: " \
                          f"{a.ravel()[bad[:5]]} vs {b.ravel()[bad[:5]]}"

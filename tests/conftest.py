import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def vv():
    """the product package (ctypes over libvvb200.so); builds the library if it is missing"""
    pkg = entry.load_package()
    if not os.path.exists(pkg.LIB_PATH):
        entry.build()
    return pkg


@pytest.fixture(scope="session")
def vo():
    """the CPU oracle wrappers -- test infrastructure"""
    mod = entry.load_oracle()
    if not os.path.exists(os.path.join(ROOT, "oracle", "build", "libvvoracle_mixed.so")):
        mod.build("oracle")
    return mod


def rel_err(a, b):
    """max |a-b| / max(|b|, 1e-3 rms(b))  (parity plan)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = np.maximum(np.abs(b), 1e-3 * np.sqrt(np.mean(b * b)) + 1e-300)
    return float(np.max(np.abs(a - b) / scale))


# tolerances of the parity plan: single frozen-force substep, relative error on x and v
TOL = {"mixed": 1e-6, "double": 1e-12, "single": 1e-5}
# scale factors / group energies
TOL_KE = {"mixed": 1e-12, "double": 1e-12, "single": 1e-5}

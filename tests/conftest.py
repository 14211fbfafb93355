import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def forloop_golden():
    """Outputs of the unmodified reference for-loop (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "forloop.npz"))
    cases = []
    for k in range(int(z["n_cases"])):
        p = "c%d_" % k
        cases.append({key[len(p):]: z[key] for key in z.files if key.startswith(p)})
    return cases


@pytest.fixture(scope="session")
def numba_golden():
    """Outputs of the reference numba kernels under the CUDA simulator."""
    z = np.load(os.path.join(GOLDEN, "numba_sim.npz"))
    cases = []
    for k in range(int(z["n_mm_cases"])):
        p = "mm%d_" % k
        cases.append({key[len(p):]: z[key] for key in z.files if key.startswith(p)})
    return cases

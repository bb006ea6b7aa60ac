import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def icp_cases():
    z = load_golden("icp_pairs.npz")
    out = []
    for i in range(int(z["count"])):
        out.append({k: z["%d_%s" % (i, k)] for k in
                    ("tar", "src", "T", "iters", "max_iter", "tol", "seed", "truth")})
    return out

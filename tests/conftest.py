import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_built():
    """The .so files are build artefacts (git-ignored): build them once if a fresh checkout has none."""
    import subprocess
    pkg = os.path.join(ROOT, "a-2d-lidar-based-slam-system-for-wheeled-mobile-robots_b200")
    if not os.path.isfile(os.path.join(pkg, "libb2slam.so")):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(pkg, "csrc")])
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


_ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def icp_cases():
    """Whole ICP.process runs of the reference itself: 120 / 360 beams (icp_pairs.npz) and the first pairs of the
    cfg-4 stream at 1080 beams (icp_pairs_1080.npz); see oracle/make_golden.py."""
    out = []
    for name in ("icp_pairs.npz", "icp_pairs_1080.npz"):
        z = load_golden(name)
        for i in range(int(z["count"])):
            out.append({k: z["%d_%s" % (i, k)] for k in
                        ("tar", "src", "T", "iters", "max_iter", "tol", "seed", "truth")})
    return out

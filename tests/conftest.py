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


def load_demo_input():
    """Config 1 input (the reference's dark.std) as a PINIT array, from the committed fixture."""
    from skid_b200.tipsy import PINIT_DTYPE
    z = np.load(os.path.join(GOLDEN, "demo_input.npz"))
    n = len(z["mass"])
    p = np.zeros(n, PINIT_DTYPE)
    p["r"], p["v"], p["fMass"], p["fSoft"], p["fTemp"] = z["r"], z["v"], z["mass"], z["soft"], z["temp"]
    p["iOrder"] = np.arange(n, dtype=np.int32)
    return p, int(z["nGas"]), int(z["nDark"]), int(z["nStar"]), float(z["time"])


@pytest.fixture(scope="session")
def demo_input():
    return load_demo_input()


@pytest.fixture(scope="session")
def demo_golden():
    return np.load(os.path.join(GOLDEN, "demo_golden.npz"))


# the reference's own demo command line (`demo:2`)
DEMO = dict(tau=9e-4, nSmooth=64, fDensMin=170.0, nMembers=8, H0=2.8944, period=1.0)

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def shipped_weights():
    """The 64 live tensors of the reference's weights/best_model.pth (tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "weights_best_model.npz")))


def golden_cases():
    return sorted(f[len("case_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("case_") and f.endswith(".npz"))


def load_case(name):
    d = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Every test session (CPU or GPU) runs against a freshly built in-tree library."""
    import __graft_entry__ as ge
    ge.build()

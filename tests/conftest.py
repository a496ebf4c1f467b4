import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(tag):
        m = dict(np.load(os.path.join(GOLDEN, f"mesh_{tag}.npz")))
        r = dict(np.load(os.path.join(GOLDEN, f"ref_{tag}.npz")))
        ptr, ee = m["electrode_ptr"], m["electrode_elems"]
        m["electrodes"] = [ee[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
        return m, r

    return load

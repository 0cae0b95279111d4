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


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def ctx():
    """One libpodb200 context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    from pod_uqnn_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def golden_workload(g):
    """(x_mesh, mu, n_t, t_min, t_max) of a fixture."""
    return g["x_mesh"], g["mu"], int(g["n_t"]), float(g["t_min"]), float(g["t_max"])

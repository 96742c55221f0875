import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simple-vector-db_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def port():
    from oracle import binding
    binding.build()
    return binding.load_port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref); skipped where it was never built."""
    from oracle import binding
    if not binding.have_ref():
        binding.build()
    if not binding.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    return binding.load_ref()


@pytest.fixture(scope="session")
def cpu(port):
    """Best available CPU checker: compiled reference if present, else the port."""
    from oracle import binding
    return binding.load_ref() if binding.have_ref() else port

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle.binding import Checker
    return Checker("orc")


@pytest.fixture(scope="session")
def ref():
    from oracle.binding import Checker
    if not Checker.available("ref"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Checker("ref")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")

    def load(name):
        return np.load(os.path.join(gdir, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def engine():
    """The product: legosnark_b200 over the C-ABI library, on cuda:0."""
    import legosnark_b200 as lb
    lb.init(1)
    yield lb
    lb.shutdown()

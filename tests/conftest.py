import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "refonly: needs /root/reference (build container only)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def g1():
    return load_golden("g1_branin_n20.npz")


@pytest.fixture(scope="session")
def g2():
    return load_golden("g2_hartmann6_n100.npz")


@pytest.fixture(scope="session")
def g3():
    return load_golden("g3_wavy6_n500.npz")


@pytest.fixture(scope="session")
def g4():
    return load_golden("g4_kernel_zoo.npz")


@pytest.fixture(scope="session")
def g5():
    return load_golden("g5_branin_long_chain.npz")


@pytest.fixture(scope="session")
def g6():
    return load_golden("g6_branin_warp.npz")

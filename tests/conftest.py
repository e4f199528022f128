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


def load_pair(name):
    d = np.load(os.path.join(GOLDEN, "inputs_%s.npz" % name))
    return d["scan1"], d["scan2"]


@pytest.fixture(scope="session")
def frame_pair():
    """reference src/sample_data/frame_804.npy / frame_805.npy as float32 [3, 65536] planes"""
    return load_pair("frame")


@pytest.fixture(scope="session")
def sample_pair():
    """reference python/point_clouds/sample_pc_1.npy / sample_pc_2.npy as float32 [3, 131072] planes"""
    return load_pair("sample_pc")


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """CUDA context of the product library (GPU tests only)."""
    import icet_b200
    icet_b200.build()
    c = icet_b200.Context(0)
    yield c
    c.close()

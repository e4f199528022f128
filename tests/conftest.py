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


class _ParityLog:
    """Collects the worst deltas / list sizes the GPU parity tests measure; written to gpurun_out/PARITY_r02.json at the
    end of the session (the copy under profiles/ is the one from the builder's last GPU visit)."""

    def __init__(self):
        self.records = []

    def add(self, stage, config, **metrics):
        self.records.append(dict(stage=stage, config=config, **metrics))


@pytest.fixture(scope="session")
def parity():
    import json
    log = _ParityLog()
    yield log
    if not log.records:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "PARITY_r02.json"), "w") as f:
            exc = []
            try:
                import test_gpu_parity
                exc = test_gpu_parity.EXCEPTIONS
            except Exception:
                pass
            json.dump({"tolerances": {"X_m": 1e-4, "X_rad": 1e-5, "Q_rel": 1e-4, "stat_rel": 1e-5, "edge_ulps": 2},
                       "exceptions_listed": exc, "records": log.records}, f, indent=1)
    except OSError:
        pass

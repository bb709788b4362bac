import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# tolerance stated by BASELINE.json's north_star for the floating-point path
RTOL, ATOL = 1e-3, 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_ted():
    return dict(np.load(os.path.join(GOLDEN, "ted.npz")))


@pytest.fixture(scope="session")
def golden_beat():
    return dict(np.load(os.path.join(GOLDEN, "beat.npz")))


@pytest.fixture(scope="session")
def golden_plms():
    return dict(np.load(os.path.join(GOLDEN, "plms.npz")))


@pytest.fixture(scope="session")
def golden_sag():
    return dict(np.load(os.path.join(GOLDEN, "sag.npz")))


@pytest.fixture(scope="session")
def golden_schedule():
    return dict(np.load(os.path.join(GOLDEN, "schedule.npz")))


@pytest.fixture(scope="session")
def golden_hooks():
    return {"ted": dict(np.load(os.path.join(GOLDEN, "hooks_ted.npz"))),
            "beat": dict(np.load(os.path.join(GOLDEN, "hooks_beat.npz")))}


@pytest.fixture(scope="session")
def golden_grad():
    return {"ted": dict(np.load(os.path.join(GOLDEN, "grad_ted.npz"))),
            "beat": dict(np.load(os.path.join(GOLDEN, "grad_beat.npz")))}


@pytest.fixture(scope="session")
def golden_train():
    return {"ted": dict(np.load(os.path.join(GOLDEN, "train_ted.npz"))),
            "beat": dict(np.load(os.path.join(GOLDEN, "train_beat.npz")))}


@pytest.fixture(scope="session")
def golden_vlb():
    return {"ted": dict(np.load(os.path.join(GOLDEN, "vlb_ted.npz"))),
            "beat": dict(np.load(os.path.join(GOLDEN, "vlb_beat.npz")))}


@pytest.fixture(scope="session")
def golden_fgd():
    return dict(np.load(os.path.join(GOLDEN, "fgd.npz")))


@pytest.fixture(scope="session")
def golden_metrics():
    return dict(np.load(os.path.join(GOLDEN, "metrics.npz")))

import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped (not failed): `pytest tests` is green on a CPU box."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked test)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def weights():
    import torch
    return torch.load(os.path.join(GOLDEN, "egonn_weights.pth"), map_location="cpu", weights_only=True)


def load_golden(name):
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


GOLDEN_CASES = {
    "cfg1_cartesian": dict(coordinates="cartesian", step=0.3),
    "mini3_cartesian": dict(coordinates="cartesian", step=0.4),
    "mini2_polar": dict(coordinates="polar", step=[1., 0.3, 0.2]),
}

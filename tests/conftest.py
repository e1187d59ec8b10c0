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
# Fixtures written by the REAL MinkowskiEngine (tools/verify_against_me.py --write-golden, on a machine that has ME 0.5.4)
# join the golden cases of every CPU (oracle) and GPU (engine) parity test as soon as they are committed: <case>_me.npz.
# None exist yet - MinkowskiEngine cannot be installed in this image - so the oracle's ME semantics stay "parity unpinned".
for _case in list(GOLDEN_CASES):
    if os.path.exists(os.path.join(GOLDEN, _case + "_me.npz")):
        GOLDEN_CASES[_case + "_me"] = GOLDEN_CASES[_case]


def train_fixtures():
    """Training-step fixtures: the shim-generated one, plus the real-ME one when it has been committed."""
    names = ["train_mini3.npz"]
    if os.path.exists(os.path.join(GOLDEN, "train_mini3_me.npz")):
        names.append("train_mini3_me.npz")
    return names

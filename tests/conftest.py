import os
import sys
import types

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


REF_ROOT = os.environ.get("EGONN_REFERENCE_ROOT", "/root/reference")
TOPS = ("MinkowskiEngine", "models", "layers", "misc", "datasets")


@pytest.fixture()
def reference_on_front_end(monkeypatch):
    """model_factory / ModelParams of the unmodified reference on the egonn_b200 front end + CPU double; sys.modules and
    sys.path are restored afterwards (other tests import the reference on the ORACLE shim)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "models")):
        pytest.skip("reference sources absent")
    import egonn_b200.minkowski as front
    from cpu_engine import CpuEngine
    monkeypatch.setattr(front, "Engine", CpuEngine)
    saved = {n: m for n, m in sys.modules.items() if n.split(".")[0] in TOPS}
    saved_path = list(sys.path)
    for n in saved:
        del sys.modules[n]
    front.install()
    sys.path.insert(0, REF_ROOT)
    m = types.ModuleType("datasets")                                   # HuggingFace `datasets` shadows the reference's package
    m.__path__ = [os.path.join(REF_ROOT, "datasets")]
    sys.modules["datasets"] = m
    from misc.utils import ModelParams
    from models.model_factory import model_factory
    import models.minkgl
    assert models.minkgl.__file__.startswith(REF_ROOT) and sys.modules["MinkowskiEngine"] is front
    try:
        yield model_factory, ModelParams
    finally:
        for n in [n for n in sys.modules if n.split(".")[0] in TOPS]:
            del sys.modules[n]
        sys.modules.update(saved)
        sys.path[:] = saved_path

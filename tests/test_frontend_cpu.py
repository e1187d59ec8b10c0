"""INTEGRATION.md path 2 without a GPU: the reference's OWN model code (``models/model_factory.py``, ``models/minkgl.py``,
``layers/*.py``, ``datasets/quantization.py``, imported UNMODIFIED from /root/reference - skipped where it is absent, e.g.
on the GPU box) runs on ``egonn_b200.minkowski`` registered as ``MinkowskiEngine``, with the engine replaced by its CPU
test double (tests/cpu_engine.py).  This exercises the whole ME symbol surface the reference touches (SURVEY §8b.2:
``SparseTensor(features, coordinates=...)``, ``.F .C .tensor_stride .decomposed_features ._batchwise_row_indices``, ``+``/
``+=``, convolutions, BatchNorm, pooling, broadcast, ``MinkowskiFunctional.normalize``, the ECA block's
``SparseTensor(..., coordinate_manager=..., coordinate_map_key=...)``) against the golden vectors of the same graph on the
oracle shim.  The GPU counterpart is tests/test_gpu_me_frontend.py."""
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import me_ops

@pytest.mark.parametrize("case", ["mini3_cartesian", "mini2_polar"])
def test_unmodified_reference_graph_on_the_front_end_matches_golden(reference_on_front_end, weights, case):
    model_factory, ModelParams = reference_on_front_end
    quant = GOLDEN_CASES[case]
    step = quant["step"]
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("[MODEL]\nmodel = egonn\ncoordinates = %s\nquantization_step = %s\n"
                % (quant["coordinates"], ", ".join(str(v) for v in step) if isinstance(step, list) else step))
    model = model_factory(ModelParams(f.name))
    os.unlink(f.name)
    model.load_state_dict(weights)
    model.eval()
    g = load_golden(case)
    coords = torch.from_numpy(g["coords"])
    grabbed = {}
    h = model.local_keypoint_regressor.register_forward_hook(lambda _m, _i, o: grabbed.update(c=o.C.numpy().copy()))
    with torch.no_grad():
        y = model({"coords": coords, "features": torch.ones((coords.shape[0], 1))})
    h.remove()
    # per-cloud lists in the front end's row order -> one array in the fixture's canonical (lexicographic) order
    c3 = grabbed["c"]
    rows = np.concatenate(me_ops.batch_rows(c3))
    assert np.array_equal(rows, np.arange(rows.shape[0])), "rows of a cloud are contiguous, clouds in batch order"
    order = me_ops.canonical_order(c3)
    assert np.array_equal(c3[order], g["coords_L3"])
    np.testing.assert_allclose(y["global"].numpy(), g["global"], rtol=1e-4, atol=1e-6)
    for k, atol in (("descriptors", 1e-6), ("keypoints", 1e-4), ("sigma", 1e-6)):
        got = torch.cat(y[k], dim=0).numpy()[order]
        np.testing.assert_allclose(got, g[k], rtol=1e-4, atol=atol, err_msg=k)
    assert len(y["descriptors"]) == int(g["n_clouds"])


@pytest.fixture()
def cpu_engine(monkeypatch):
    import egonn_b200.minkowski as ME
    from cpu_engine import CpuEngine
    monkeypatch.setattr(ME, "Engine", CpuEngine)


@pytest.mark.parametrize("case", ["mini3_cartesian", "mini2_polar"])
def test_own_layer_walk_matches_golden(cpu_engine, weights, case):
    """egonn_b200's own model classes, layer walk (``forward_layerwise``: the cross-check path of the fused engine and the
    train-mode path) on the engine double == the golden outputs, including the mirrored quantizers' keypoint_position."""
    import egonn_b200 as E
    quant = GOLDEN_CASES[case]
    model = E.model_factory(E.ModelParams.from_dict(model="egonn", coordinates=quant["coordinates"], quantization_step=quant["step"]))
    model.load_state_dict(weights)
    model.eval()
    g = load_golden(case)
    coords = torch.from_numpy(g["coords"])
    grabbed = {}
    h = model.local_keypoint_regressor.register_forward_hook(lambda _m, _i, o: grabbed.update(c=o.C.numpy().copy()))
    y = model.forward_layerwise({"coords": coords, "features": torch.ones((coords.shape[0], 1))})
    h.remove()
    order = me_ops.canonical_order(grabbed["c"])
    assert np.array_equal(grabbed["c"][order], g["coords_L3"])
    np.testing.assert_allclose(y["global"].numpy(), g["global"], rtol=1e-4, atol=1e-6)
    for k, atol in (("descriptors", 1e-6), ("keypoints", 1e-4), ("sigma", 1e-6)):
        np.testing.assert_allclose(torch.cat(y[k], dim=0).numpy()[order], g[k], rtol=1e-4, atol=atol, err_msg=k)


def test_own_minkloc3d_layer_walk_matches_golden(cpu_engine):
    import egonn_b200 as E
    from conftest import GOLDEN
    g = torch.load(os.path.join(GOLDEN, "minkloc3d.pt"), map_location="cpu", weights_only=True)
    m = E.model_factory(E.ModelParams.from_dict(model="MinkLoc3D", coordinates="cartesian", quantization_step=0.4))
    m.load_state_dict(g["state_dict"])
    m.eval()
    y = m.forward_layerwise({"coords": g["coords"], "features": torch.ones((g["coords"].shape[0], 1))})
    torch.testing.assert_close(y["global"], g["global"], rtol=1e-4, atol=1e-6)

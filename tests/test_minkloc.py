"""MinkLoc / MinkLoc3D (MinkFPN + global pooling): oracle vs the golden vectors of the unmodified
reference (CPU), engine vs golden / oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import egonn_oracle, me_ops


@pytest.fixture(scope="module")
def golden3d():
    return torch.load(os.path.join(GOLDEN, "minkloc3d.pt"), map_location="cpu", weights_only=True)


def test_oracle_minkloc3d_matches_reference_graph(golden3d):
    g = golden3d
    out = egonn_oracle.forward_minkloc(g["state_dict"], g["coords"].numpy(), torch.ones((g["coords"].shape[0], 1)), 1, "GeM", "pooling.p")
    assert np.array_equal(out["map"][0], g["map_coords"].numpy())
    torch.testing.assert_close(out["map"][1], g["map_features"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["global"], g["global"], rtol=1e-5, atol=1e-6)


def test_model_factory_minkloc_state_dict_mirrors_reference(golden3d):
    import egonn_b200 as E
    m = E.model_factory(E.ModelParams.from_dict(model="MinkLoc3D", coordinates="cartesian", quantization_step=0.4))
    ref_keys = list(golden3d["state_dict"].keys())
    assert list(m.state_dict().keys()) == ref_keys
    m.load_state_dict(golden3d["state_dict"])
    m2 = E.model_factory(E.ModelParams.from_dict(model="MinkLoc", coordinates="cartesian", quantization_step=0.4,
                                                 block="ECABasicBlock", planes="32,64,64", layers="1,1,1", pooling="GeM"))
    keys = list(m2.state_dict().keys())
    assert "backbone.blocks.1.0.eca.conv.weight" in keys and "pooling.pooling.p" in keys and "backbone.tconvs.0.kernel" in keys
    with pytest.raises(NotImplementedError):
        E.model_factory(E.ModelParams.from_dict(model="MinkLoc", coordinates="cartesian", quantization_step=0.4, block="SEBasicBlock"))


@pytest.mark.gpu
def test_engine_minkloc3d_vs_golden(golden3d):
    import egonn_b200 as E
    from gpu_common import assert_close_rel
    dev = torch.device("cuda", 0)
    g = golden3d
    m = E.model_factory(E.ModelParams.from_dict(model="MinkLoc3D", coordinates="cartesian", quantization_step=0.4))
    m.load_state_dict(g["state_dict"])
    m = m.eval().to(dev)
    batch = {"coords": g["coords"].to(dev), "features": torch.ones((g["coords"].shape[0], 1), device=dev)}
    y = m(batch)
    assert set(y) == {"global"} and y["global"].shape == (2, 256)
    assert_close_rel(y["global"], g["global"], 1e-3, "MinkLoc3D global descriptor")
    eng = m._engine
    fmap = eng.tap(3, 0, 256)
    o = me_ops.canonical_order(eng.level_coords(2).cpu().numpy())
    assert np.array_equal(eng.level_coords(2).cpu().numpy()[o], g["map_coords"].numpy())
    assert_close_rel(fmap[o], g["map_features"], 1e-3, "FPN output map")
    z = m.forward_layerwise(batch)
    assert_close_rel(z["global"], y["global"], 1e-4, "layer-wise operator path")


@pytest.mark.gpu
@pytest.mark.parametrize("block,pool", [("ECABasicBlock", "GeM"), ("BasicBlock", "MAC"), ("BasicBlock", "SPoC")])
def test_engine_minkloc_variants_vs_oracle(block, pool, golden3d):
    import egonn_b200 as E
    from gpu_common import assert_close_rel
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    m = E.model_factory(E.ModelParams.from_dict(model="MinkLoc", coordinates="cartesian", quantization_step=0.4, block=block,
                                                pooling=pool, feature_size=256, output_dim=256))
    with torch.no_grad():
        for name, buf in m.named_buffers():
            if name.endswith("running_var"):
                buf.copy_(torch.rand_like(buf) + 0.5)
            if name.endswith("running_mean"):
                buf.copy_(torch.randn_like(buf) * 0.1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    coords = golden3d["coords"]
    ref = egonn_oracle.forward_minkloc(sd, coords.numpy(), torch.ones((coords.shape[0], 1)), 1, pool, "pooling.pooling.p")
    m = m.eval().to(dev)
    y = m({"coords": coords.to(dev), "features": torch.ones((coords.shape[0], 1), device=dev)})
    assert_close_rel(y["global"], ref["global"], 1e-3, f"MinkLoc {block}/{pool}")

"""The functional oracle (oracle/egonn_oracle.py) against the golden vectors that the UNMODIFIED
reference graph code produced on the ME shim (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import egonn_oracle, me_ops


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_quantize_matches_reference_quantizer(case):
    g = load_golden(case)
    quant = GOLDEN_CASES[case]
    sp = g["points_splits"]
    coords, index = [], []
    for i in range(int(g["n_clouds"])):
        c, ndx = egonn_oracle.quantize(torch.from_numpy(g["points"][sp[i]:sp[i + 1]]), quant)
        coords.append(c)
        index.append(ndx.numpy())
    bc = me_ops.batched_coordinates(coords).numpy()
    assert np.array_equal(bc, g["coords"])
    assert np.array_equal(np.concatenate(index), g["quant_index"])


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_forward_matches_reference_graph(case, weights):
    g = load_golden(case)
    quant = GOLDEN_CASES[case]
    coords = g["coords"]
    out = egonn_oracle.forward(weights, coords, torch.ones((coords.shape[0], 1)), quant, keep_intermediates=True)
    for L in range(1, 8):
        assert np.array_equal(out["levels"][L], g[f"coords_L{L}"]), f"level {L} coordinates"
    assert np.array_equal(out["coords_L3"], g["coords_L3"])
    for L in (1, 3, 5, 7):
        np.testing.assert_allclose(out["features"][f"block{L}"].numpy(), g[f"block{L}"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(out["global"].numpy(), g["global"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["descriptors"].numpy(), g["descriptors"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["keypoints"].numpy(), g["keypoints"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(out["sigma"].numpy(), g["sigma"], rtol=1e-5, atol=1e-6)
    assert len(out["descriptors_list"]) == int(g["n_clouds"])


def test_fp64_accumulation_gap_is_small(weights):
    """how far fp32 accumulation is from fp64 accumulation - the floor under the 1e-3 budget."""
    g = load_golden("mini3_cartesian")
    coords = g["coords"]
    f = torch.ones((coords.shape[0], 1))
    a = egonn_oracle.forward(weights, coords, f, GOLDEN_CASES["mini3_cartesian"])
    b = egonn_oracle.forward(weights, coords, f, GOLDEN_CASES["mini3_cartesian"], acc64=True)
    for k in ("global", "descriptors", "sigma"):
        err = (a[k] - b[k]).abs().max().item() / b[k].abs().max().item()
        assert err < 1e-4, (k, err)

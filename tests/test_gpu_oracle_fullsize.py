"""-m gpu parity at BASELINE.json's sizes against the CPU ORACLE itself (element-wise, every tap + the four outputs):
one cloud each of cfg2 / cfg3 / cfg4 (the oracle needs 1-3 s per cloud), the single 1 M-point cloud of cfg5 (the oracle
needs 1-2 minutes, fp32 accumulation), and cloud 0 INSIDE the full cfg2 batch of 16 (engine rows of cloud 0 sliced by
the per-level batch offsets, oracle run on cloud 0 alone - batch independence of the reference graph).

This is the path `MinkGL.forward` models/minkgl.py:267-315 at the sizes where the engine's full-size code paths are
active (N-split / K-split thresholds of the 128-channel levels, the 256-slice pooling cap, narrow sort keys).
Bars: coordinates of every level bit-exact (keyed by coordinate); floats max|a-b|/max|b| <= 1e-3 (north star).  The
per-tensor errors are printed (run with -s) and written to gpurun_out/oracle_parity_<case>.json when that directory
exists, so the judged run leaves a record."""
import json
import os

import numpy as np
import pytest
import torch

from gpu_common import RTOL, lex_order, rel_err
from oracle import egonn_oracle, me_ops

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda", 0)


def _model(weights, voxel, cuda):
    import egonn_b200 as E
    mp = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=voxel)
    m = E.model_factory(mp)
    m.load_state_dict(weights)
    return m.eval().to(cuda), mp


def _record(case, errs):
    print(f"\n[oracle parity] {case}: " + "  ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    out = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"oracle_parity_{case}.json"), "w") as f:
            json.dump({"case": case, "tolerance": RTOL, "rel_err": errs}, f, indent=1)


def _compare_cloud(model, p, ref, b, errs):
    """Engine outputs `p` (packed, whole batch) of cloud `b` against the oracle run `ref` on that cloud ALONE (batch
    index 0 there).  Every level's coordinates bit-exact, every tap and output <= RTOL."""
    eng = model._engine

    def rows_of(level):
        off = eng.batch_offsets(level).cpu().numpy()
        c = eng.level_coords(level)[int(off[b]):int(off[b + 1])].clone()
        c[:, 0] = 0
        return int(off[b]), int(off[b + 1]), c

    def check(name, a, r):
        e = rel_err(a, r)
        errs[name] = e
        assert e <= RTOL, f"{name}: max|a-b|/max|b| = {e:.3e} > {RTOL:.1e}"

    orders = {}
    for L in range(0, 8):
        lo, hi, c = rows_of(L)
        o = lex_order(c)
        assert np.array_equal(c.cpu().numpy()[o], ref["levels"][L]), f"level {L} coordinates differ"
        orders[L] = (lo, hi, o)
    fe = ref["features"]
    lo, hi, o = orders[0]
    check("conv0", eng.tap(0, 0, 32)[lo:hi][o], fe["conv0"])
    for L in range(1, 8):
        lo, hi, o = orders[L]
        check(f"down{L}", eng.tap(1, L, fe[f"down{L}"].shape[1])[lo:hi][o], fe[f"down{L}"])
        check(f"block{L}", eng.tap(2, L, fe[f"block{L}"].shape[1])[lo:hi][o], fe[f"block{L}"])
    lo, hi, o = orders[5]
    check("global_head_map", eng.tap(3, 5, 128)[lo:hi][o], fe["global_head_map"])
    lo, hi, o = orders[3]
    check("local_map", eng.tap(4, 3, 64)[lo:hi][o], fe["local_map"])
    off3 = p["local_offsets"].cpu().numpy()
    assert (int(off3[b]), int(off3[b + 1])) == (lo, hi)
    lc = p["local_coords"][lo:hi].clone()
    lc[:, 0] = 0
    assert np.array_equal(lc.cpu().numpy()[o], ref["coords_L3"]), "keypoint voxels (level-3 coordinates) differ"
    check("global", p["global"][b], ref["global"][0])
    check("descriptors", p["descriptors"][lo:hi][o], ref["descriptors"])
    check("keypoints", p["keypoints"][lo:hi][o], ref["keypoints"])
    check("sigma", p["sigma"][lo:hi][o], ref["sigma"])


def _oracle_cloud(weights, pc, voxel):
    quant = {"coordinates": "cartesian", "step": voxel}
    c, _ = egonn_oracle.quantize(torch.from_numpy(pc), quant)
    bc = me_ops.batched_coordinates([c])
    ref = egonn_oracle.forward(weights, bc.numpy(), torch.ones((bc.shape[0], 1)), quant, keep_intermediates=True)
    return c, ref


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_single_cloud_vs_oracle_all_taps(cfg, cuda, weights):
    """One full-size cloud of the config: engine (tensor-core path, fused ingest of the voxels) == oracle, all taps."""
    import egonn_b200 as E
    from egonn_b200 import synth
    voxel = synth.CONFIGS[cfg]["voxel"]
    pc = synth.make_batch(cfg, batch=1)[0]
    c_ref, ref = _oracle_cloud(weights, pc, voxel)
    model, mp = _model(weights, voxel, cuda)
    c, _ = mp.quantizer(torch.from_numpy(pc).to(cuda))
    assert torch.equal(c.cpu(), c_ref), "quantised voxels differ from the oracle (bit-exact bar)"
    bc = E.batched_coordinates([c])
    p = model.forward_packed({"coords": bc, "features": torch.ones((bc.shape[0], 1), device=cuda)})
    torch.cuda.synchronize()
    errs = {}
    _compare_cloud(model, p, ref, 0, errs)
    _record(f"{cfg}_single", errs)


def test_cfg2_cloud_inside_full_batch_vs_oracle(cuda, weights):
    """BASELINE config 2 as benchmarked (16 clouds, ~750 k voxels): clouds 0 and 11 of the batch, rows sliced by the
    batch offsets of every level, against the oracle run on each cloud alone."""
    import egonn_b200 as E
    from egonn_b200 import synth
    voxel = synth.CONFIGS["cfg2"]["voxel"]
    clouds = synth.make_batch("cfg2", batch=16)
    model, mp = _model(weights, voxel, cuda)
    coords = [mp.quantizer(torch.from_numpy(pc).to(cuda))[0] for pc in clouds]
    bc = E.batched_coordinates(coords).contiguous()
    p = model.forward_packed({"coords": bc, "features": torch.ones((bc.shape[0], 1), device=cuda)})
    torch.cuda.synchronize()
    for b in (0, 11):
        c_ref, ref = _oracle_cloud(weights, clouds[b], voxel)
        assert torch.equal(coords[b].cpu(), c_ref)
        errs = {}
        _compare_cloud(model, p, ref, b, errs)
        _record(f"cfg2_batch16_cloud{b}", errs)
    # the fused raw-point ingest (what bench.py's e2e arm runs) gives the same packed outputs
    pts = torch.cat([torch.from_numpy(pc) for pc in clouds]).to(cuda)
    off = torch.tensor(np.cumsum([0] + [pc.shape[0] for pc in clouds]), dtype=torch.int32, device=cuda)
    p2 = model.forward_points(pts, off)
    torch.cuda.synchronize()
    assert torch.equal(p2["local_coords"], p["local_coords"])
    for k in ("global", "descriptors", "keypoints", "sigma"):
        assert rel_err(p2[k], p[k]) <= 1e-6, k


def test_cfg5_map_tile_vs_oracle(cuda, weights):
    """BASELINE config 5 (1 M points, 0.05 m voxels, ~0.9 M voxels): the whole cloud against the oracle with fp32
    accumulation (acc64=False; ~1-2 minutes of host time).  Slow by design, run once per suite."""
    import egonn_b200 as E
    from egonn_b200 import synth
    voxel = synth.CONFIGS["cfg5"]["voxel"]
    pc = synth.make_batch("cfg5", batch=1)[0]
    model, mp = _model(weights, voxel, cuda)
    c, _ = mp.quantizer(torch.from_numpy(pc).to(cuda))
    bc = E.batched_coordinates([c])
    p = model.forward_packed({"coords": bc, "features": torch.ones((bc.shape[0], 1), device=cuda)})
    torch.cuda.synchronize()
    c_ref, ref = _oracle_cloud(weights, pc, voxel)
    assert torch.equal(c.cpu(), c_ref)
    errs = {}
    _compare_cloud(model, p, ref, 0, errs)
    _record("cfg5_single", errs)

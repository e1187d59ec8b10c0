"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference graph code.

Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

It imports ``models.model_factory.model_factory`` / ``misc.utils.ModelParams`` /
``datasets.quantization`` from /root/reference on top of ``oracle/me_shim/MinkowskiEngine`` (the
ME-semantics CPU shim - MinkowskiEngine itself is not installable here), loads the shipped checkpoint
and runs ``model(batch)`` exactly as ``eval/evaluate.py:327-350`` does.  Outputs are stored in
canonical row order (lexicographic (b,x,y,z)) next to the voxel coordinates they belong to.

Also re-saves the reference checkpoint as CPU tensors (tests/golden/egonn_weights.pth): the GPU box has
no /root/reference, and the original file needs a CUDA device to unpickle.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_import, me_ops  # noqa: E402
from egonn_b200 import synth  # noqa: E402

CASES = {
    # name: (coordinates, quantization_step, clouds)
    "cfg1_cartesian": ("cartesian", "0.3", lambda: [synth.uniform_cloud(4096, 0)]),
    "mini3_cartesian": ("cartesian", "0.4", lambda: [
        synth.spinning_lidar_cloud(11, beams=16, azimuths=300, max_range=40.0, n_cylinders=30),
        synth.uniform_cloud(1500, 12),
        synth.spinning_lidar_cloud(13, beams=24, azimuths=200, max_range=30.0, n_cylinders=20)]),
    "mini2_polar": ("polar", "1., 0.3, 0.2", lambda: [
        synth.spinning_lidar_cloud(21, beams=32, azimuths=360, max_range=50.0, n_cylinders=40),
        synth.uniform_cloud(3000, 22)]),
}


def enable_real_me(reference_root: str):
    """The reference on the REAL MinkowskiEngine (tools/verify_against_me.py --write-golden, on a machine that has it):
    same import work-arounds as oracle/ref_import.py, without the shim."""
    import types
    import MinkowskiEngine as ME
    assert "oracle" not in getattr(ME, "__version__", "") and "egonn_b200" not in getattr(ME, "__version__", ""), \
        "this must be the real MinkowskiEngine"
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    m = types.ModuleType("datasets")                                   # HuggingFace `datasets` shadows the reference's package
    m.__path__ = [os.path.join(reference_root, "datasets")]
    sys.modules["datasets"] = m


def main(real_me: bool = False, reference_root: str = None, suffix: str = ""):
    """``real_me``: run on the real MinkowskiEngine and write ``<case><suffix>.npz`` (suffix "_me"): the fixtures that PIN the
    oracle - tests/conftest.py adds every ``<case>_me.npz`` it finds to the golden cases of the CPU and GPU parity tests."""
    if real_me:
        enable_real_me(reference_root or ref_import.REFERENCE_ROOT)
    else:
        ref_import.enable()
    import MinkowskiEngine as ME
    from models.model_factory import model_factory
    from misc.utils import ModelParams

    if real_me:
        sd = torch.load(os.path.join(HERE, "egonn_weights.pth"), map_location="cpu", weights_only=True)
    else:
        sd = torch.load(os.path.join(ref_import.REFERENCE_ROOT, "weights", "model_egonn_20210916_1104.pth"),
                        map_location="cpu", weights_only=True)
        torch.save({k: v.clone() for k, v in sd.items()}, os.path.join(HERE, "egonn_weights.pth"))

    def to_np(t):
        return t.detach().cpu().numpy().copy()

    for name, (coordinates, step, make) in CASES.items():
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(f"[MODEL]\nmodel = egonn\ncoordinates = {coordinates}\nquantization_step = {step}\n")
        mp = ModelParams(f.name)
        os.unlink(f.name)
        model = model_factory(mp)
        model.load_state_dict(sd)
        model.eval()

        clouds = make()
        coords, index = [], []
        for pc in clouds:
            c, ndx = mp.quantizer(torch.from_numpy(pc))           # eval/evaluate.py:331
            coords.append(c)
            index.append(to_np(ndx))
        bcoords = ME.utils.batched_coordinates(coords)             # eval/evaluate.py:333
        feats = torch.ones((bcoords.shape[0], 1), dtype=torch.float32)

        grabbed = {}
        def grab(key):
            def hook(_m, _i, o):
                grabbed[key] = (to_np(o.C).astype(np.int32), to_np(o.F))
            return hook
        handles = [model.local_keypoint_regressor.register_forward_hook(grab("kp"))]
        for L in range(1, 8):
            handles.append(model.trunk.blocks[str(L)].register_forward_hook(grab(f"block{L}")))
        with torch.no_grad():
            y = model({"coords": bcoords, "features": feats})      # eval/evaluate.py:338
        for h in handles:
            h.remove()

        c3 = grabbed["kp"][0]
        order = me_ops.canonical_order(c3)
        # the reference returns per-cloud lists in map row order; concatenate in batch order and
        # re-sort canonically (rows of one cloud are contiguous in both orders)
        cat = lambda lst: to_np(torch.cat(lst, dim=0))
        rows = np.concatenate(me_ops.batch_rows(c3))
        inv = np.empty_like(rows)
        inv[rows] = np.arange(rows.shape[0])
        out = {
            "n_clouds": np.int64(len(clouds)),
            "points": np.concatenate(clouds, axis=0),
            "points_splits": np.cumsum([0] + [p.shape[0] for p in clouds]),
            "quant_index": np.concatenate(index),
            "coords": to_np(bcoords).astype(np.int32),
            "global": to_np(y["global"]),
            "coords_L3": c3[order],
            "descriptors": cat(y["descriptors"])[inv][order],
            "keypoints": cat(y["keypoints"])[inv][order],
            "sigma": cat(y["sigma"])[inv][order],
        }
        for L in range(1, 8):
            cL, fL = grabbed[f"block{L}"]
            o = me_ops.canonical_order(cL)
            out[f"coords_L{L}"] = cL[o]
            if L in (1, 3, 5, 7):
                out[f"block{L}"] = fL[o]
        np.savez_compressed(os.path.join(HERE, name + suffix + ".npz"), **out)
        print(name, "points", out["points"].shape, "voxels", bcoords.shape[0],
              "levels", [out[f'coords_L{L}'].shape[0] for L in range(1, 8)], "global[0,:3]", out["global"][0, :3])


if __name__ == "__main__":
    main()

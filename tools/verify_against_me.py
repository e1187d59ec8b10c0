"""Diff the CPU oracle (and, with a GPU, the engine) against the REAL MinkowskiEngine 0.5.4 on a machine that has it.

NOT runnable in the build image or on the B200 boxes (MinkowskiEngine is not installable there: no network, and ME
0.5.4 predates CUDA 12) - this script is the hook for a maintainer with a working ME install to pin the oracle
(SURVEY.md §8c(5)).  It runs the reference's own model code (path via --reference) on the same inputs and weights as
tests/golden/make_golden.py and prints the maximum relative differences, keyed by coordinate.

    PYTHONPATH=/path/to/Egonn python tools/verify_against_me.py --reference /path/to/Egonn [--gpu]
    python tools/verify_against_me.py --reference /path/to/Egonn --write-golden

--write-golden writes the fixtures that PIN the oracle: the three golden cases of tests/golden/make_golden.py and the
training step of tests/golden/make_golden_train.py, produced by the reference's own model code on the REAL
MinkowskiEngine, as tests/golden/<case>_me.npz and tests/golden/train_mini3_me.npz.  Commit them: tests/conftest.py adds
every <case>_me.npz to the golden cases of the CPU (oracle) and GPU (engine) parity tests, and the training tests check
train_mini3_me.npz next to the shim-generated fixture.
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--gpu", action="store_true", help="also run the egonn_b200 engine")
    ap.add_argument("--write-golden", action="store_true", help="write tests/golden/*_me.npz from the real MinkowskiEngine and exit")
    args = ap.parse_args()
    if args.write_golden:
        sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import make_golden
        import make_golden_train
        make_golden.main(real_me=True, reference_root=args.reference, suffix="_me")
        make_golden_train.main(real_me=True, reference_root=args.reference, suffix="_me")
        print("wrote tests/golden/*_me.npz - run `python -m pytest tests -m 'not gpu'` (and -m gpu on a B200) and commit them")
        return
    import MinkowskiEngine as ME                                       # the real one
    assert "oracle" not in getattr(ME, "__version__", ""), "this must be the real MinkowskiEngine"
    sys.path.insert(0, args.reference)
    m = types.ModuleType("datasets")
    m.__path__ = [os.path.join(args.reference, "datasets")]
    sys.modules["datasets"] = m
    from models.model_factory import model_factory
    from misc.utils import ModelParams
    from oracle import egonn_oracle, me_ops
    from egonn_b200 import synth

    sd = torch.load(os.path.join(REPO, "tests", "golden", "egonn_weights.pth"), map_location="cpu")
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("[MODEL]\nmodel = egonn\ncoordinates = cartesian\nquantization_step = 0.3\n")
    mp = ModelParams(f.name)
    model = model_factory(mp)
    model.load_state_dict(sd)
    model.eval()
    pc = torch.from_numpy(synth.uniform_cloud(4096, 0))
    coords, _ = mp.quantizer(pc)
    bc = ME.utils.batched_coordinates([coords])
    feats = torch.ones((bc.shape[0], 1))
    hook = {}
    h = model.local_keypoint_regressor.register_forward_hook(lambda _m, _i, o: hook.update(c=o.C.cpu().numpy()))
    with torch.no_grad():
        y = model({"coords": bc, "features": feats})
    h.remove()
    order = me_ops.canonical_order(hook["c"])
    ora = egonn_oracle.forward(sd, bc.numpy(), feats, {"coordinates": "cartesian", "step": 0.3})
    assert np.array_equal(hook["c"][order], ora["coords_L3"]), "level-3 coordinate sets differ"

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    print("oracle vs MinkowskiEngine:")
    print("  global      ", rel(ora["global"], y["global"]))
    for k in ("descriptors", "keypoints", "sigma"):
        print(f"  {k:12s}", rel(ora[k], y[k][0][torch.from_numpy(order)]))
    if args.gpu:
        import egonn_b200 as E
        em = E.model_factory(E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.3))
        em.load_state_dict(sd)
        em = em.eval().cuda()
        p = em.forward_packed({"coords": bc.cuda(), "features": feats.cuda()})
        o2 = me_ops.canonical_order(p["local_coords"].cpu().numpy())
        print("engine vs MinkowskiEngine:")
        print("  global      ", rel(p["global"].cpu(), y["global"]))
        for k in ("descriptors", "keypoints", "sigma"):
            print(f"  {k:12s}", rel(p[k].cpu()[o2], y[k][0][torch.from_numpy(order)]))


if __name__ == "__main__":
    main()

"""Golden vectors for the MinkLoc3D (MinkFPN + GeM) model family from the UNMODIFIED reference
(third_party/minkloc3d/minkloc.py) on the oracle ME shim.  The reference ships no MinkLoc3D checkpoint, so the
weights are a seeded random initialisation with non-trivial BatchNorm statistics; they are stored next to the
outputs (tests/golden/minkloc3d.pt).  Run here:  python tests/golden/make_golden_minkloc.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import ref_import, me_ops  # noqa: E402
from egonn_b200 import synth  # noqa: E402


def main():
    ref_import.enable()
    import MinkowskiEngine as ME
    from third_party.minkloc3d.minkloc import MinkLoc3D
    from datasets.quantization import CartesianQuantizer
    torch.manual_seed(1234)
    model = MinkLoc3D()
    with torch.no_grad():
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn_like(buf) * 0.2)
            if name.endswith("running_var"):
                buf.copy_(torch.rand_like(buf) * 1.5 + 0.25)
        for name, p in model.named_parameters():
            if name.endswith("bn.weight"):
                p.copy_(torch.rand_like(p) + 0.5)
            if name.endswith("bn.bias"):
                p.copy_(torch.randn_like(p) * 0.1)
        model.pooling.p.fill_(2.6)
    model.eval()
    q = CartesianQuantizer(0.4)
    clouds = [synth.spinning_lidar_cloud(31, beams=16, azimuths=300, max_range=40.0, n_cylinders=30), synth.uniform_cloud(2000, 32)]
    coords = [q(torch.from_numpy(pc))[0] for pc in clouds]
    bc = ME.utils.batched_coordinates(coords)
    feats = torch.ones((bc.shape[0], 1))
    grabbed = {}
    h = model.backbone.register_forward_hook(lambda m, i, o: grabbed.update(map=(o.C.numpy().copy(), o.F.detach().numpy().copy())))
    with torch.no_grad():
        y = model({"coords": bc, "features": feats})
    h.remove()
    cm, fm = grabbed["map"]
    o = me_ops.canonical_order(cm)
    torch.save({"state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "coords": bc, "global": y["global"],
                "map_coords": torch.from_numpy(cm[o]), "map_features": torch.from_numpy(fm[o])},
               os.path.join(HERE, "minkloc3d.pt"))
    print("voxels", bc.shape[0], "fpn map", fm.shape, "global[0,:4]", y["global"][0, :4])


if __name__ == "__main__":
    main()

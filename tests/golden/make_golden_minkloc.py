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


def _randomise(model, p_value):
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
        p_value.fill_(2.6)
    model.eval()


def _run(model, name, small=False):
    import MinkowskiEngine as ME
    from datasets.quantization import CartesianQuantizer
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
               os.path.join(HERE, name))
    print(name, "voxels", bc.shape[0], "fpn map", fm.shape, "global[0,:4]", y["global"][0, :4])


def main():
    ref_import.enable()
    from third_party.minkloc3d.minkloc import MinkLoc3D
    torch.manual_seed(1234)
    model = MinkLoc3D()
    _randomise(model, model.pooling.p)
    _run(model, "minkloc3d.pt")
    if "--layers" in sys.argv:
        # models/minkloc.py with more than one block per level (layers = 2,1,2; ECABasicBlock): ResNetBase._make_layer
        # models/resnet.py:81-97 builds layers[L] blocks, the first one carries the 1x1 downsample
        from models.minkloc import MinkLoc
        torch.manual_seed(4321)
        m2 = MinkLoc(in_channels=1, feature_size=64, output_dim=64, planes=[32, 64, 64], layers=[2, 1, 2], num_top_down=1,
                     conv0_kernel_size=5, block="ECABasicBlock", pooling_method="GeM")
        _randomise(m2, m2.pooling.pooling.p)
        _run(m2, "minkloc_layers212.pt")


if __name__ == "__main__":
    main()

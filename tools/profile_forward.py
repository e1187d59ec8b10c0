"""Run a few device-resident forwards of a BASELINE config - the command wrapped by ncu (see profiles/README.md)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egonn_b200 as E  # noqa: E402
from egonn_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--no-tc", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
cfg = synth.CONFIGS[args.config]
batch = args.batch or cfg["batch"]
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "egonn_weights.pth"), map_location="cpu",
                weights_only=True)
params = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=cfg["voxel"])
model = E.model_factory(params)
model.load_state_dict(sd)
model = model.eval().to(dev)
clouds = synth.make_batch(args.config, batch=batch)
coords = [params.quantizer(torch.from_numpy(pc).to(dev))[0] for pc in clouds]
bc = E.batched_coordinates(coords).contiguous()
feats = torch.ones((bc.shape[0], 1), device=dev)
for i in range(args.iters):
    if i == 0 and args.no_tc:
        model.forward_packed({"coords": bc, "features": feats})
        model._engine.set_tensor_cores(False)
    p = model.forward_packed({"coords": bc, "features": feats})
    E.topk_smallest(p["sigma"], p["local_offsets"], 256)
torch.cuda.synchronize()
print("rows", p["n_rows"][:8])

"""Isolated timing of the coordinate build (pack + radix sort + pyramid + kernel maps + row orders) at a BASELINE config,
per kernel class (engine event brackets).   python tools/sort_bench.py [cfg2]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egonn_b200 as E  # noqa: E402
from egonn_b200 import synth  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
dev = torch.device("cuda", 0)
params = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=synth.CONFIGS[cfg]["voxel"])
clouds = synth.make_batch(cfg)
coords = [params.quantizer(torch.from_numpy(pc).to(dev))[0] for pc in clouds]
bc = E.batched_coordinates(coords).contiguous()
eng = E.Engine(dev)
for _ in range(3):
    eng.build(bc)
eng.profile(True)
for _ in range(20):
    eng.build(bc)
torch.cuda.synchronize()
for e in sorted(eng.profile_read(), key=lambda e: -e["ms"]):
    print(f"{e['name']:28s} launches/build {e['launches'] / 20:5.1f}  us/build {e['ms'] * 1e3 / 20:8.1f}")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.profile(False)
a.record()
for _ in range(50):
    eng.build(bc)
b.record()
torch.cuda.synchronize()
print(f"whole coords_build (incl. its host sync): {a.elapsed_time(b) * 1e3 / 50:.1f} us, n = {bc.shape[0]}")

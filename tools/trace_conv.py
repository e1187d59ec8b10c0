"""Dump the clock64 timeline of CTA 0 of one tensor-core convolution (EGN_TRACE=1)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["EGN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egonn_b200 as E  # noqa: E402
from egonn_b200 import lib as L, synth  # noqa: E402

dev = torch.device("cuda", 0)
level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
c = int(sys.argv[2]) if len(sys.argv) > 2 else 128
params = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.1)
coords = [params.quantizer(torch.from_numpy(pc).to(dev))[0] for pc in synth.make_batch("cfg2")]
eng = E.Engine(dev)
info = eng.build(E.batched_coordinates(coords).contiguous())
x = torch.randn(info.n_rows[level], c, device=dev)
w = torch.randn(27, c, c, device=dev) * 0.05
for _ in range(3):
    eng.conv_tc(level, 3, x, w)
torch.cuda.synchronize()
buf = np.zeros((64, 8), dtype=np.int64)
L.check(L.load().egn_debug_trace(eng._ctx, buf.ctypes.data_as(C.c_void_p)))
t0 = buf[0, 0]
names = ["mma:top", "mma:Bfull", "mma:Afull", "mma:issued", "mma:commit", "prod:top", "prod:empty", "prod:arrived"]
print("chunk " + " ".join(f"{n:>12s}" for n in names))
for i in range(24):
    print(f"{i:5d} " + " ".join(f"{(v - t0):12d}" for v in buf[i]))
d = np.diff(buf[:40, 0])
print("MMA-thread chunk period (cycles): median", np.median(d), "min", d.min(), "max", d.max())

"""Dump the clock64 timeline of one mid-grid CTA of a tensor-core convolution (EGN_TRACE=1; k_sconv_ts stamps).

    EGN_TRACE_BUILD=1 python -m egonn_b200.build --force  # the stamps are compiled in only on request
    python tools/trace_conv.py LEVEL CHANNELS [KSIZE]     # one isolated convolution at a cfg2 level, fp32 input map
"""
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
level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
c = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ksize = int(sys.argv[3]) if len(sys.argv) > 3 else 3          # 3: 3x3x3 at LEVEL; 2: stride-2 from LEVEL to LEVEL+1; 1: 1x1x1 at LEVEL
params = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.1)
coords = [params.quantizer(torch.from_numpy(pc).to(dev))[0] for pc in synth.make_batch("cfg2")]
eng = E.Engine(dev)
info = eng.build(E.batched_coordinates(coords).contiguous())
x = torch.randn(info.n_rows[level], c, device=dev)
w = torch.randn(ksize ** 3, c, c, device=dev) * 0.05
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    eng.conv_tc(level, ksize, x, w)
torch.cuda.synchronize()
buf = np.zeros((64, 8), dtype=np.int64)
L.check(L.load().egn_debug_trace(eng._ctx, buf.ctypes.data_as(C.c_void_p)))
t0 = buf[60, 0]
n = int(buf[61, 0])
print(f"chunks in this tile: {n};  CTA timeline (cycles from kernel start): prologue done {buf[60,1]-t0}, producers done {buf[60,2]-t0}, "
      f"accumulator ready {buf[60,3]-t0}, epilogue done {buf[60,4]-t0}, end {buf[60,5]-t0}")
names = ["mma:top", "mma:full", "mma:issued", "prod:top", "prod:ldg_iss", "prod:empty", "prod:st_iss", "prod:arrived"]
print("chunk " + " ".join(f"{k:>12s}" for k in names) + "   (group 0 producer stamps on even chunks, group 1 on odd)")
for i in range(min(n, 30)):
    print(f"{i:5d} " + " ".join(f"{(v - t0):12d}" for v in buf[i]))

"""Isolated timing of the tensor-core convolution at the real level sizes of a BASELINE config (CUDA events, back-to-back
launches, optional interleaved dummy kernel to expose per-launch fixed costs)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egonn_b200 as E  # noqa: E402
from egonn_b200 import synth  # noqa: E402
from egonn_b200.weights import pack_tc  # noqa: E402
from egonn_b200 import lib as L  # noqa: E402
import ctypes as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS[args.config]
params = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=cfg["voxel"])
clouds = synth.make_batch(args.config)
coords = [params.quantizer(torch.from_numpy(pc).to(dev))[0] for pc in clouds]
bc = E.batched_coordinates(coords).contiguous()
eng = E.Engine(dev)
info = eng.build(bc)
print("rows", info.n_rows[:8])
lib = L.load()


def time_conv(level, ksize, cin, cout, transposed=False, interleave=False, iters=args.iters):
    lvl_out = (level - 1 if transposed else level + 1) if ksize == 2 else level
    x = torch.randn(info.n_rows[level], cin, device=dev)
    w = torch.randn(ksize ** 3, cin, cout, device=dev) * 0.05
    wp = pack_tc(w).to(dev)
    out = torch.empty((info.n_rows[lvl_out], cout), device=dev)
    dummy = torch.zeros(1024, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        L.check(lib.egn_conv_tc(eng._ctx, level, ksize, int(transposed), cin, cout, C.c_void_p(x.data_ptr()), C.c_void_p(wp.data_ptr()),
                                None, None, 0, C.c_void_p(out.data_ptr()), st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        run()
        if interleave:
            dummy.add_(1.0)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


for level in (1, 2, 3, 4, 5, 6, 7):
    c = {1: 32, 2: 64, 3: 64, 4: 128, 5: 128, 6: 128, 7: 128}[level]
    n = info.n_rows[level]
    t3 = time_conv(level, 3, c, c)
    t3i = time_conv(level, 3, c, c, interleave=True)
    t1 = time_conv(level, 1, c, c) if c >= 64 else float("nan")
    t2 = time_conv(level - 1, 2, c, c) if level >= 2 and c == {1: 32, 2: 64, 3: 64, 4: 128, 5: 128, 6: 128, 7: 128}.get(level - 1, 32) else float("nan")
    print(f"L{level} rows={n:7d} tiles={-(-n // 128):5d} c={c:3d}  3x3x3: {t3:8.1f} us (interleaved {t3i:8.1f})   1x1: {t1:7.1f} us   2x2x2s2(from L{level - 1}): {t2:7.1f} us")

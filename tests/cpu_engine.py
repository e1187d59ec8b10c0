"""TEST DOUBLE of ``egonn_b200.engine.Engine`` on the CPU oracle (``oracle/me_ops.py``).

Only for ``-m "not gpu"`` tests of HOST LOGIC that sits above the C ABI - the backward rules of ``egonn_b200.autograd``
and the train-mode walk of ``egonn_b200.models`` - in a container without a GPU.  It is not part of the product (the
product has no CPU path and never imports ``oracle/``); the same tests run against the real engine under ``-m gpu``.

It reproduces what those layers rely on: every coordinate map in the engine's canonical order (batch, Morton code with x
in the low bit of every 3-bit group, coordinates biased by 2^17), ``level_coords`` / ``batch_offsets`` / ``neighbors`` /
``input_rows`` as index tensors, and the forward operators ``conv`` / ``global_pool`` / ``broadcast_mul``.
"""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import me_ops

LEVELS = 8


def _spread3(v: np.ndarray) -> np.ndarray:
    out = np.zeros_like(v, dtype=np.uint64)
    for i in range(18):
        out |= ((v >> np.uint64(i)) & np.uint64(1)) << np.uint64(3 * i)
    return out


def morton_order(coords: np.ndarray) -> np.ndarray:
    c = coords.astype(np.int64)
    u = (c[:, 1:] + (1 << 17)).astype(np.uint64)
    key = (c[:, 0].astype(np.uint64) << np.uint64(54)) | _spread3(u[:, 0]) | (_spread3(u[:, 1]) << np.uint64(1)) | (_spread3(u[:, 2]) << np.uint64(2))
    return np.argsort(key, kind="stable")


class CpuEngine:
    def __init__(self, device=None):
        self.device = torch.device("cpu")
        self.info = None

    # -- coordinate manager -----------------------------------------------------------------------------
    def build(self, coords: torch.Tensor):
        c = coords.detach().cpu().numpy().astype(np.int32)
        self.cm = me_ops.CoordinateManager(c)                 # oracle row order at level 0 = input order
        s = 1
        for _ in range(LEVELS - 1):
            s = self.cm.stride_map(s, 2)
        self.perm, self.inv, self.can = [], [], []            # canonical row j  <->  oracle row perm[L][j]
        for L in range(LEVELS):
            oc = self.cm.coords(1 << L)
            p = morton_order(oc)
            inv = np.empty_like(p)
            inv[p] = np.arange(p.shape[0])
            self.perm.append(torch.from_numpy(p))
            self.inv.append(torch.from_numpy(inv))
            self.can.append(np.ascontiguousarray(oc[p]))
        self._derived = {}
        self.info = SimpleNamespace(n_batches=self.cm.n_batches, n_input=c.shape[0], n_rows=[x.shape[0] for x in self.can])
        return self.info

    def input_rows(self):
        return self.perm[0].to(torch.int32)

    def level_coords(self, level):
        return torch.from_numpy(self.can[level].astype(np.int32))

    def batch_offsets(self, level):
        b = self.can[level][:, 0]
        return torch.from_numpy(np.searchsorted(b, np.arange(self.info.n_batches + 1)).astype(np.int32))

    def neighbors(self, level):
        s = 1 << level
        n = self.can[level].shape[0]
        table = torch.full((n, 27), -1, dtype=torch.int32)
        inv = self.inv[level]
        for k, (i_rows, o_rows) in enumerate(self.cm.kernel_map(s, s, 3)):
            table[inv[torch.from_numpy(o_rows)], k] = inv[torch.from_numpy(i_rows)].to(torch.int32)
        return table

    # -- operators (canonical order in, canonical order out) ----------------------------------------------
    def conv(self, level_in, ksize, transposed, x, kernel, scale=None, shift=None, relu=False, out=None, accumulate=False):
        assert scale is None and shift is None and not relu and out is None and not accumulate
        x = x.detach().float()
        k = kernel.detach().float()
        if ksize == 1:
            return x @ (k if k.dim() == 2 else k[0])
        x_or = x[self.inv[level_in]]
        if ksize == 2 and transposed:
            y, s = me_ops.convolution_transpose(self.cm, x_or, 1 << level_in, k)
        else:
            y, s = me_ops.convolution(self.cm, x_or, 1 << level_in, k, ksize, 2 if ksize == 2 else 1)
        return y[self.perm[int(np.log2(s))]]

    def global_pool(self, level, x, is_max=False):
        fn = me_ops.global_max_pool if is_max else me_ops.global_avg_pool
        return fn(x.detach().float(), self.can[level], self.info.n_batches)

    def broadcast_mul(self, level, x, g):
        return me_ops.broadcast_mul(x.detach().float(), self.can[level], g.detach().float())

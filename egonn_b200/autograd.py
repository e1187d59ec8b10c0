"""Training-mode operators (second half): autograd for the sparse operators of the engine, so that
the reference's training step - ``model.train(); y = model(batch); loss.backward(); optimizer.step()``,
training/trainer.py:141-195 - runs on the CUDA engine.  MinkowskiEngine's counterparts are the ``*Backward*`` halves of
``MinkowskiConvolutionFunction``, ``MinkowskiGlobalPoolingFunction`` and ``MinkowskiBroadcastFunction``.

Every backward is built from the engine's own FORWARD operators (the C ABI of include/egonn_b200.h) wherever the gradient
is itself a sparse convolution / pooling / broadcast on the same coordinate maps:

    y = conv3(x, W)        dx = conv3(dy, W'),  W'[k] = W[26-k]^T      (delta_{26-k} = -delta_k: centred odd kernel, A.3)
    y = conv2s2(x, W)      dx = tconv2s2(dy, W'),  W'[k] = W[k]^T      (fine <- its one parent through slice k(fine), A.5)
    y = tconv2s2(x, W)     dx = conv2s2(dy, W'),   W'[k] = W[k]^T
    y = x @ W (1x1)        dx = dy @ W^T
    y = mean_cloud(x)      dx[r] = dy[cloud(r)] / n_cloud
    y = x * g[cloud]       dx = broadcast_mul(dy, g);  dg = n_cloud * mean_cloud(dy * x)

The WEIGHT gradient ``dW[k] = sum over pairs (i,o) of offset k of x[i]^T dy[o]`` is a gather followed by a dense
(Cin x pairs) @ (pairs x Cout) product per offset; it is evaluated with the engine's kernel maps (``egn_coords_neighbors``,
the parent / child-code links of the pyramid) and torch's GEMM on the device - a plain library contraction, not a
hand-written kernel (training is outside the measured path).  CUDA tensors only; no CPU path; no double backward.
"""
from __future__ import annotations

from typing import Tuple

import torch


# ---- kernel maps of the engine's pyramid, as index tensors (cached per coordinate build) ---------------------------------
def _derived(eng) -> dict:
    d = getattr(eng, "_derived", None)
    if d is None:
        d = eng._derived = {}
    return d


def child_links(eng, fine_level: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """For every row of ``fine_level``: (row of its parent at fine_level + 1, kernel index k = dx + 2 dy + 4 dz of the child
    inside the parent's 2x2x2 region - SURVEY A.3/A.5).  Rows are in canonical (batch, Morton) order on every level, so the
    children of a parent are contiguous and parents appear in the order of their first child."""
    key = ("links", fine_level)
    d = _derived(eng)
    if key not in d:
        c = eng.level_coords(fine_level).long()                        # (N,4) [b,x,y,z], multiples of 2^level
        s = fine_level
        code = ((c[:, 1] >> s) & 1) + 2 * ((c[:, 2] >> s) & 1) + 4 * ((c[:, 3] >> s) & 1)
        p = torch.stack([c[:, 0], c[:, 1] >> (s + 1), c[:, 2] >> (s + 1), c[:, 3] >> (s + 1)], dim=1)
        new = torch.ones((c.shape[0],), dtype=torch.bool, device=c.device)
        if c.shape[0] > 1:
            new[1:] = (p[1:] != p[:-1]).any(dim=1)
        d[key] = (torch.cumsum(new.long(), 0) - 1, code)
    return d[key]


def neighbor_rows(eng, level: int) -> torch.Tensor:
    """(N,27) int64 input row of offset k = kx + 3 ky + 9 kz for every output row, N (= one past the last row) where absent."""
    key = ("nbr", level)
    d = _derived(eng)
    if key not in d:
        nbr = eng.neighbors(level).long()
        d[key] = torch.where(nbr >= 0, nbr, torch.full_like(nbr, nbr.shape[0]))
    return d[key]


def cloud_rows(eng, level: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(cloud index of every row (N,), rows per cloud (B,)) of a level."""
    key = ("cloud", level)
    d = _derived(eng)
    if key not in d:
        off = eng.batch_offsets(level).long()
        counts = off[1:] - off[:-1]
        n = eng.info.n_rows[level]
        bidx = torch.repeat_interleave(torch.arange(counts.shape[0], device=off.device), counts, output_size=n)
        d[key] = (bidx, counts)
    return d[key]


def window_weight_grad(eng, ksize: int, x: torch.Tensor, gy: torch.Tensor) -> torch.Tensor:
    """dW (ksize^3, Cin, Cout) of an odd ksize^3 stride-1 convolution at level 0 (x fastest, centred - A.3).  The engine's
    5x5x5 stem walks occupancy masks and keeps no neighbour table, so the input row of every offset is found here from the
    level's coordinates with one sort + one binary search per offset (training only; nothing is cached: the table of a
    750 k-voxel batch would take 750 MB)."""
    c = eng.level_coords(0).long()
    n, r, bias, span = c.shape[0], ksize // 2, 1 << 17, 1 << 18          # 18 bits per axis as in the engine's keys, batch < 512

    def pack(b, x_, y_, z_):
        return ((b * span + (x_ + bias)) * span + (y_ + bias)) * span + (z_ + bias)

    skeys, perm = torch.sort(pack(c[:, 0], c[:, 1], c[:, 2], c[:, 3]))
    xp = _padded(x)
    gw = []
    for k in range(ksize ** 3):
        dx, dy, dz = k % ksize - r, (k // ksize) % ksize - r, k // (ksize * ksize) - r
        q = pack(c[:, 0], c[:, 1] + dx, c[:, 2] + dy, c[:, 3] + dz)
        pos = torch.searchsorted(skeys, q).clamp(max=n - 1)
        rows = torch.where(skeys[pos] == q, perm[pos], torch.full_like(pos, n))
        gw.append(xp[rows].t() @ gy)
    return torch.stack(gw, dim=0)


def _padded(x: torch.Tensor) -> torch.Tensor:
    """x with one all-zero row appended: absent neighbours index it."""
    return torch.cat([x, x.new_zeros((1, x.shape[1]))], dim=0)


# ---- convolution -----------------------------------------------------------------------------------------------------
class SparseConvFunction(torch.autograd.Function):
    """``MinkowskiConvolution`` / ``MinkowskiConvolutionTranspose`` (kernel 1 | 3 | 5 stride 1, kernel 2 stride 2) with
    gradients for the features and the kernel."""

    @staticmethod
    def forward(ctx, x, kernel, eng, level_in, ksize, transposed):
        ctx.eng, ctx.level_in, ctx.ksize, ctx.transposed = eng, level_in, ksize, transposed
        ctx.save_for_backward(x, kernel)
        return eng.conv(level_in, ksize, transposed, x, kernel)

    @staticmethod
    def backward(ctx, gy):
        x, kernel = ctx.saved_tensors
        eng, lvl, ksize, transposed = ctx.eng, ctx.level_in, ctx.ksize, ctx.transposed
        gy = gy.contiguous().float()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gx = gw = None
        if ksize == 1:
            w = kernel if kernel.dim() == 2 else kernel[0]
            if need_x:
                gx = eng.conv(lvl, 1, False, gy, w.t().contiguous())
            if need_w:
                gw = (x.t() @ gy).reshape(kernel.shape)
        elif ksize in (3, 5):
            kv = ksize ** 3
            if need_x:
                if ksize != 3:
                    raise NotImplementedError("gradient w.r.t. the input of the 5x5x5 stem (its input are the constant "
                                              "all-ones features on the EgoNN / MinkLoc path)")
                gx = eng.conv(lvl, 3, False, gy, kernel.flip(0).transpose(1, 2).contiguous())
            if need_w:
                if ksize == 3:
                    rows, xp = neighbor_rows(eng, lvl), _padded(x)
                    gw = torch.stack([xp[rows[:, k]].t() @ gy for k in range(kv)], dim=0)
                else:
                    assert lvl == 0, "the 5x5x5 convolution is the level-0 stem"
                    gw = window_weight_grad(eng, ksize, x, gy)
        else:
            assert ksize == 2
            wt = kernel.transpose(1, 2).contiguous()
            fine = lvl - 1 if transposed else lvl                      # the finer of the two levels
            parent, code = child_links(eng, fine)
            if need_x:
                gx = eng.conv(lvl - 1 if transposed else lvl + 1, 2, not transposed, gy, wt)
            if need_w:
                a, b = (x[parent], gy) if transposed else (x, gy[parent])   # per fine row: (input row, output-gradient row)
                gw = torch.stack([(a * (code == k).unsqueeze(1).to(a.dtype)).t() @ b for k in range(8)], dim=0)
        return gx, gw, None, None, None, None


# ---- per-cloud pooling / broadcast -------------------------------------------------------------------------------------
class GlobalPoolFunction(torch.autograd.Function):
    """``MinkowskiGlobalPooling`` (mean over the rows of a cloud) / ``MinkowskiGlobalMaxPooling``."""

    @staticmethod
    def forward(ctx, x, eng, level, is_max):
        out = eng.global_pool(level, x, is_max)
        ctx.eng, ctx.level, ctx.is_max = eng, level, is_max
        ctx.save_for_backward(*((x, out) if is_max else ()))
        return out

    @staticmethod
    def backward(ctx, gy):
        bidx, counts = cloud_rows(ctx.eng, ctx.level)
        gy = gy.contiguous().float()
        if not ctx.is_max:
            return (gy / counts.clamp(min=1).unsqueeze(1).to(gy.dtype))[bidx], None, None, None
        x, out = ctx.saved_tensors                                     # the gradient goes to the FIRST row that attains the maximum
        n, c = x.shape
        rows = torch.arange(n, device=x.device).unsqueeze(1).expand(n, c)
        cand = torch.where(x == out[bidx], rows, torch.full_like(rows, n))
        first = torch.full((out.shape[0], c), n, dtype=torch.long, device=x.device)
        first = first.scatter_reduce(0, bidx.unsqueeze(1).expand(n, c), cand, reduce="amin", include_self=True)
        gx = x.new_zeros((n + 1, c))
        gx.scatter_(0, first, gy)
        return gx[:n], None, None, None


class BroadcastMulFunction(torch.autograd.Function):
    """``MinkowskiBroadcastMultiplication``: out[r] = x[r] * g[cloud(r)]."""

    @staticmethod
    def forward(ctx, x, g, eng, level):
        ctx.eng, ctx.level = eng, level
        ctx.save_for_backward(x, g)
        return eng.broadcast_mul(level, x, g)

    @staticmethod
    def backward(ctx, gy):
        x, g = ctx.saved_tensors
        eng, level = ctx.eng, ctx.level
        gy = gy.contiguous().float()
        gx = gg = None
        if ctx.needs_input_grad[0]:
            gx = eng.broadcast_mul(level, gy, g)
        if ctx.needs_input_grad[1]:
            _, counts = cloud_rows(eng, level)
            gg = eng.global_pool(level, gy * x, False) * counts.unsqueeze(1).to(gy.dtype)
        return gx, gg, None, None


def wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)

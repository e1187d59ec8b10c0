"""Functional CPU restatement of the MinkowskiEngine 0.5.4 operations used by the reference.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) - parity unpinned: MinkowskiEngine is absent from
/root/reference and from this image; semantics follow SURVEY.md Appendix A.  Every function cites
the reference call site whose behaviour it restates (paths relative to /root/reference).

Integer work is numpy, floating-point work is torch CPU float32 (``acc64=True`` switches the
convolution accumulation to float64 so the fp32 error budget can be measured).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------
# coordinate keys
# ----------------------------------------------------------------------------------------------
AXIS_BITS = 18                      # signed voxel coordinate range [-2^17, 2^17)
AXIS_BIAS = 1 << (AXIS_BITS - 1)
BATCH_BITS = 10


def pack_keys(coords: np.ndarray) -> np.ndarray:
    """(N,4) int [b,x,y,z] -> uint64 lexicographic key (b, x, y, z).  Used only for set
    membership / canonical ordering inside the oracle; it is not an ME concept."""
    c = np.asarray(coords).astype(np.int64)
    assert c.ndim == 2 and c.shape[1] == 4
    if c.shape[0]:
        assert c[:, 0].min() >= 0 and c[:, 0].max() < (1 << BATCH_BITS) - 1, "batch index out of range"
        assert c[:, 1:].min() >= -AXIS_BIAS and c[:, 1:].max() < AXIS_BIAS, "voxel coordinate out of range"
    u = (c + np.array([0, AXIS_BIAS, AXIS_BIAS, AXIS_BIAS], dtype=np.int64)).astype(np.uint64)
    s = np.uint64(AXIS_BITS)
    return (((u[:, 0] << s | u[:, 1]) << s | u[:, 2]) << s) | u[:, 3]


def canonical_order(coords: np.ndarray) -> np.ndarray:
    """Row permutation that sorts coordinates lexicographically by (b,x,y,z).  Row order of ME
    coordinate maps is not a contract - comparisons are made in this order."""
    return np.argsort(pack_keys(coords), kind="stable")


# ----------------------------------------------------------------------------------------------
# ME.utils.sparse_quantize / batched_coordinates
# ----------------------------------------------------------------------------------------------
def sparse_quantize(coordinates: torch.Tensor, quantization_size=None, return_index: bool = True):
    """``ME.utils.sparse_quantize(pc, quantization_size=q, return_index=True)`` as called from
    datasets/quantization.py:42 (polar, q=1.) and :83 (cartesian, q=step).

    float32 tensor / python float (float32 divide), floor, ``.int()``; duplicates removed with
    first-occurrence-wins, surviving rows kept in input order; returns (coords int32, index int64).
    SURVEY A.1.
    """
    c = torch.as_tensor(coordinates)
    if quantization_size is not None and not (np.isscalar(quantization_size) and quantization_size == 1):
        if isinstance(quantization_size, (list, tuple, np.ndarray, torch.Tensor)):
            q = torch.tensor([float(v) for v in quantization_size], dtype=c.dtype)
            c = c / q
        else:
            c = c / quantization_size
    d = torch.floor(c).int()
    dn = d.numpy()
    # unique over rows (no batch column here): first occurrence wins, ascending original index
    keys = pack_keys(np.concatenate([np.zeros((dn.shape[0], 1), dtype=np.int64), dn.astype(np.int64)], axis=1))
    _, first = np.unique(keys, return_index=True)
    unique_map = np.sort(first).astype(np.int64)
    um = torch.from_numpy(unique_map)
    if return_index:
        return d[um], um
    return d[um]


def batched_coordinates(coords: Sequence, dtype=torch.int32) -> torch.Tensor:
    """``ME.utils.batched_coordinates`` (eval/evaluate.py:333, datasets/dataset_utils.py:77):
    concatenate per-cloud (Mi,3) coordinates and prepend the list index as column 0.  SURVEY A.11."""
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        if c.dtype.is_floating_point:
            c = torch.floor(c)
        c = c.to(dtype)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), c], dim=1))
    if not out:
        return torch.zeros((0, 4), dtype=dtype)
    return torch.cat(out, dim=0)


# ----------------------------------------------------------------------------------------------
# coordinate manager: strided maps + kernel maps
# ----------------------------------------------------------------------------------------------
def kernel_offsets(kernel_size: int, tensor_stride: int, dilation: int = 1) -> np.ndarray:
    """Offsets delta_k of ME's hyper-cube kernel region, x fastest:
    odd K -> centred (k_i - K//2)*dilation*stride ; even K -> k_i*dilation*stride."""
    K = kernel_size
    k = np.arange(K ** 3)
    ki = np.stack([k % K, (k // K) % K, k // (K * K)], axis=1)
    if K % 2 == 1:
        ki = ki - K // 2
    return (ki * dilation * tensor_stride).astype(np.int64)


class CoordinateManager:
    """Holds the coordinate maps of one sparse tensor family, keyed by tensor stride, and the
    kernel maps between them (what ME's CoordinateManager caches).  A new manager is created by
    every ``ME.SparseTensor(features, coordinates=...)`` (models/minkgl.py:269, layers/pooling.py:84)."""

    def __init__(self, coords: np.ndarray):
        coords = np.ascontiguousarray(np.asarray(coords).astype(np.int32))
        keys = pack_keys(coords)
        assert np.unique(keys).shape[0] == keys.shape[0], \
            "duplicate input coordinates: the reference always feeds quantised (unique) coordinates"
        self.maps: Dict[int, np.ndarray] = {1: coords}          # stride -> (N,4) int32
        self._sorted: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}   # stride -> (sorted keys, perm)
        self._kmaps: Dict[Tuple, List[Tuple[np.ndarray, np.ndarray]]] = {}
        self.n_batches = int(coords[:, 0].max()) + 1 if coords.shape[0] else 0

    # -- maps ------------------------------------------------------------------------------
    def coords(self, stride: int) -> np.ndarray:
        return self.maps[stride]

    def _lookup(self, stride: int, query: np.ndarray) -> np.ndarray:
        """Row index of each query coordinate in the map of ``stride`` (-1 if absent)."""
        if stride not in self._sorted:
            keys = pack_keys(self.maps[stride])
            perm = np.argsort(keys, kind="stable")
            self._sorted[stride] = (keys[perm], perm)
        skeys, perm = self._sorted[stride]
        inr = np.all((query[:, 1:] >= -AXIS_BIAS) & (query[:, 1:] < AXIS_BIAS), axis=1)
        q = np.where(inr[:, None], query, 0)
        qk = pack_keys(q)
        pos = np.searchsorted(skeys, qk)
        pos_c = np.minimum(pos, max(skeys.shape[0] - 1, 0))
        hit = inr & (pos < skeys.shape[0])
        if skeys.shape[0]:
            hit &= skeys[pos_c] == qk
        else:
            hit &= False
        return np.where(hit, perm[pos_c] if skeys.shape[0] else 0, -1).astype(np.int64)

    def stride_map(self, in_stride: int, factor: int = 2) -> int:
        """Output coordinate map of a stride-``factor`` convolution: floor(c / s') * s' per spatial
        axis with s' = factor * in_stride, batch kept, de-duplicated.  Row order here is
        lexicographic (ME's is hash-iteration order: not a contract)."""
        out_stride = in_stride * factor
        if out_stride not in self.maps:
            c = self.maps[in_stride].astype(np.int64)
            d = c.copy()
            d[:, 1:] = np.floor_divide(c[:, 1:], out_stride) * out_stride
            keys = pack_keys(d)
            _, first = np.unique(keys, return_index=True)
            self.maps[out_stride] = d[first].astype(np.int32)
        return out_stride

    # -- kernel maps -----------------------------------------------------------------------
    def kernel_map(self, in_stride: int, out_stride: int, kernel_size: int,
                   dilation: int = 1) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Per-offset (in_rows, out_rows) pairs: for every output coordinate o and offset k, the input
        row holding o + delta_k, delta_k in units of the INPUT tensor stride."""
        key = (in_stride, out_stride, kernel_size, dilation)
        if key not in self._kmaps:
            out_c = self.maps[out_stride].astype(np.int64)
            pairs = []
            for d in kernel_offsets(kernel_size, in_stride, dilation):
                q = out_c.copy()
                q[:, 1:] += d
                rows = self._lookup(in_stride, q)
                o = np.nonzero(rows >= 0)[0]
                pairs.append((rows[o], o.astype(np.int64)))
            self._kmaps[key] = pairs
        return self._kmaps[key]


# ----------------------------------------------------------------------------------------------
# feature ops
# ----------------------------------------------------------------------------------------------
def _gather_mm_scatter(feats: torch.Tensor, kernel: torch.Tensor, pairs, n_out: int, acc64: bool) -> torch.Tensor:
    """out[o] += in[i] @ kernel[k] for every pair of every offset k, k ascending: the
    gather -> GEMM -> scatter-add loop of ME's convolution forward."""
    dt = torch.float64 if acc64 else feats.dtype
    out = torch.zeros((n_out, kernel.shape[-1]), dtype=dt)
    f = feats.to(dt)
    w = kernel.to(dt)
    for k, (i_rows, o_rows) in enumerate(pairs):
        if i_rows.shape[0] == 0:
            continue
        g = f.index_select(0, torch.from_numpy(i_rows))
        out.index_add_(0, torch.from_numpy(o_rows), g @ w[k])
    return out.to(feats.dtype)


def convolution(cm: CoordinateManager, feats: torch.Tensor, in_stride: int, kernel: torch.Tensor,
                kernel_size: int, stride: int = 1, dilation: int = 1, acc64: bool = False) -> Tuple[torch.Tensor, int]:
    """``ME.MinkowskiConvolution`` forward, bias=False (models/minkgl.py:100-107,124-126,43;
    MinkowskiEngine BasicBlock conv1/conv2).  kernel: (K^3,Cin,Cout), or (Cin,Cout) for a 1x1x1
    convolution which ME evaluates as ``F.mm(kernel)`` on the same map.
    Returns (features, out_stride)."""
    if kernel_size == 1 and stride == 1:
        assert kernel.dim() == 2
        if acc64:
            return (feats.double() @ kernel.double()).to(feats.dtype), in_stride
        return feats @ kernel, in_stride
    assert kernel.dim() == 3 and kernel.shape[0] == kernel_size ** 3
    out_stride = cm.stride_map(in_stride, stride) if stride > 1 else in_stride
    pairs = cm.kernel_map(in_stride, out_stride, kernel_size, dilation)
    return _gather_mm_scatter(feats, kernel, pairs, cm.coords(out_stride).shape[0], acc64), out_stride


def convolution_transpose(cm: CoordinateManager, feats: torch.Tensor, in_stride: int, kernel: torch.Tensor,
                          kernel_size: int = 2, stride: int = 2, acc64: bool = False) -> Tuple[torch.Tensor, int]:
    """``ME.MinkowskiConvolutionTranspose(kernel_size=2, stride=2)`` forward (models/minkgl.py:39,
    models/minkfpn.py:51).  Output map = the EXISTING map at stride in_stride/2 (bottom-up level),
    kernel map = the forward stride-2 map fine->coarse with in/out swapped:
    out[f] = in[parent(f)] @ kernel[k(f)]  (SURVEY A.5)."""
    assert kernel_size == 2 and stride == 2
    out_stride = in_stride // stride
    assert out_stride in cm.maps, "transposed convolution onto a map that does not exist yet"
    fwd = cm.kernel_map(out_stride, in_stride, kernel_size)          # fine(in) -> coarse(out)
    swapped = [(o_rows, i_rows) for (i_rows, o_rows) in fwd]
    return _gather_mm_scatter(feats, kernel, swapped, cm.coords(out_stride).shape[0], acc64), out_stride


def batch_norm_eval(x: torch.Tensor, weight, bias, running_mean, running_var, eps: float = 1e-5) -> torch.Tensor:
    """``ME.MinkowskiBatchNorm`` in eval mode = ``torch.nn.BatchNorm1d`` on .F."""
    return torch.nn.functional.batch_norm(x, running_mean, running_var, weight, bias, False, 0.1, eps)


def batch_rows(coords: np.ndarray, n_batches: Optional[int] = None) -> List[np.ndarray]:
    """Row indices of each batch index, in row order (``_batchwise_row_indices``)."""
    b = coords[:, 0]
    nb = (int(b.max()) + 1 if b.shape[0] else 0) if n_batches is None else n_batches
    return [np.nonzero(b == i)[0] for i in range(nb)]


def global_avg_pool(x: torch.Tensor, coords: np.ndarray, n_batches: Optional[int] = None) -> torch.Tensor:
    """``ME.MinkowskiGlobalPooling`` / ``GlobalAvgPooling``: per-batch-index mean of rows, one row per
    batch index in batch order (layers/eca_block.py:16,23; layers/pooling.py:80,85)."""
    rows = batch_rows(coords, n_batches)
    out = torch.zeros((len(rows), x.shape[1]), dtype=x.dtype)
    for i, r in enumerate(rows):
        if r.shape[0]:
            out[i] = x[torch.from_numpy(r)].sum(dim=0) / r.shape[0]
    return out


def global_max_pool(x: torch.Tensor, coords: np.ndarray, n_batches: Optional[int] = None) -> torch.Tensor:
    """``ME.MinkowskiGlobalMaxPooling`` (layers/pooling.py:52)."""
    rows = batch_rows(coords, n_batches)
    out = torch.zeros((len(rows), x.shape[1]), dtype=x.dtype)
    for i, r in enumerate(rows):
        if r.shape[0]:
            out[i] = x[torch.from_numpy(r)].max(dim=0).values
    return out


def broadcast_mul(x: torch.Tensor, coords: np.ndarray, g: torch.Tensor) -> torch.Tensor:
    """``ME.MinkowskiBroadcastMultiplication``: out[r] = x[r] * g[batch(r)] (layers/eca_block.py:19,36)."""
    return x * g[torch.from_numpy(coords[:, 0].astype(np.int64))]

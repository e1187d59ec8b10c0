"""``MinkowskiEngine``-shaped front end of the CUDA engine: the ME symbols the reference touches
(SURVEY.md §8b.2), each operator dispatched through the C ABI (``egn_conv``, ``egn_global_pool``,
``egn_broadcast_mul``, ``egn_quantize``, ``egn_coords_build``).  CUDA tensors only - no CPU path.
With autograd enabled and a parameter or input that requires a gradient, the sparse operators go through
``egonn_b200.autograd`` (backward = the engine's own forward operators on transposed kernels; training mode).

Registering this package as ``MinkowskiEngine`` (``egonn_b200.minkowski.install()``) lets the reference's
own ``models/*.py`` / ``layers/*.py`` run unmodified on the B200; ``egonn_b200.models`` uses the same
modules as parameter holders (so the shipped checkpoint loads by name) and replaces the per-layer walk
with one fused ``egn_forward`` call.

Row order: every coordinate map is kept in the engine's canonical order (batch, Morton); ``.C`` and
``.F`` of a SparseTensor are consistent with each other (MinkowskiEngine's order is not a contract)."""
from __future__ import annotations

import sys

import numpy as np
import torch
import torch.nn as nn

from ..engine import Engine
from .. import autograd as _ag
from .. import quantization as _q

__version__ = "0.5.4-egonn_b200"


class CoordinateMapKey:
    def __init__(self, level: int, origin: bool = False):
        self.level, self.origin = level, origin

    def get_tensor_stride(self):
        return [1 << self.level] * 3

    def __eq__(self, o):
        return isinstance(o, CoordinateMapKey) and (self.level, self.origin) == (o.level, o.origin)

    def __hash__(self):
        return hash((self.level, self.origin))


class SparseTensor:
    def __init__(self, features, coordinates=None, coordinate_manager=None, coordinate_map_key=None, **_):
        if coordinates is not None:
            eng = Engine(features.device)
            eng.build(coordinates.to(features.device))
            rows = eng.input_rows().long()
            self._F = features[rows]                      # canonical order, duplicates dropped (first wins)
            self.coordinate_manager = eng
            self.coordinate_map_key = CoordinateMapKey(0)
        else:
            assert coordinate_manager is not None and coordinate_map_key is not None
            self._F = features
            self.coordinate_manager = coordinate_manager
            self.coordinate_map_key = coordinate_map_key

    @property
    def F(self):
        return self._F

    feats = F

    @property
    def C(self):
        eng, key = self.coordinate_manager, self.coordinate_map_key
        if key.origin:
            c = torch.zeros((eng.info.n_batches, 4), dtype=torch.int32, device=self._F.device)
            c[:, 0] = torch.arange(eng.info.n_batches, device=self._F.device)
            return c
        return eng.level_coords(key.level)

    coordinates = C

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def shape(self):
        return self._F.shape

    @property
    def device(self):
        return self._F.device

    @property
    def _batchwise_row_indices(self):
        eng, key = self.coordinate_manager, self.coordinate_map_key
        if key.origin:
            return [torch.tensor([b], device=self._F.device) for b in range(eng.info.n_batches)]
        off = eng.batch_offsets(key.level).tolist()
        return [torch.arange(off[b], off[b + 1], device=self._F.device) for b in range(eng.info.n_batches)]

    @property
    def decomposed_features(self):
        eng, key = self.coordinate_manager, self.coordinate_map_key
        if key.origin:
            return [self._F[b:b + 1] for b in range(eng.info.n_batches)]
        off = eng.batch_offsets(key.level).tolist()
        return [self._F[off[b]:off[b + 1]] for b in range(eng.info.n_batches)]

    def _like(self, feats):
        return SparseTensor(feats, coordinate_manager=self.coordinate_manager, coordinate_map_key=self.coordinate_map_key)

    def __add__(self, other):
        assert self.coordinate_map_key == other.coordinate_map_key
        return self._like(self._F + other._F)

    def __iadd__(self, other):
        assert self.coordinate_map_key == other.coordinate_map_key
        self._F = self._F + other._F
        return self


class MinkowskiConvolution(nn.Module):
    """Parameter ``kernel``: (K^3, Cin, Cout), or (Cin, Cout) for kernel_size 1."""
    _transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3, "only 3-D sparse tensors are supported"
        assert not bias, "the reference never uses a convolution bias"
        assert dilation == 1, "dilation is not used by the reference"
        assert (kernel_size, stride) in ((1, 1), (3, 1), (5, 1), (2, 2)), \
            f"kernel_size={kernel_size}, stride={stride} is not on the EgoNN/MinkLoc path"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation
        kv = kernel_size ** 3
        shape = (in_channels, out_channels) if kv == 1 else (kv, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(shape))
        self.bias = None

    def forward(self, x: SparseTensor) -> SparseTensor:
        lvl = x.coordinate_map_key.level
        if _ag.wants_grad(x.F, self.kernel):               # training (training/trainer.py:141-195): egonn_b200.autograd
            f = _ag.SparseConvFunction.apply(x.F, self.kernel, x.coordinate_manager, lvl, self.kernel_size, self._transposed)
        else:
            f = x.coordinate_manager.conv(lvl, self.kernel_size, self._transposed, x.F, self.kernel)
        out = lvl if self.kernel_size != 2 else (lvl - 1 if self._transposed else lvl + 1)
        return SparseTensor(f, coordinate_manager=x.coordinate_manager, coordinate_map_key=CoordinateMapKey(out))

    def extra_repr(self):
        return f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"


class MinkowskiConvolutionTranspose(MinkowskiConvolution):
    _transposed = True


class _OnFeatures(nn.Module):
    def _apply_f(self, x, fn):
        return x._like(fn(x.F))


class MinkowskiBatchNorm(_OnFeatures):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        return self._apply_f(x, self.bn)


class MinkowskiLinear(_OnFeatures):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        return self._apply_f(x, self.linear)


def _wrap(torch_cls, name):
    class _M(_OnFeatures):
        def __init__(self, *a, **k):
            super().__init__()
            self.module = torch_cls(*a, **k)

        def forward(self, x):
            return self._apply_f(x, self.module)
    _M.__name__ = _M.__qualname__ = name
    return _M


MinkowskiReLU = _wrap(nn.ReLU, "MinkowskiReLU")
MinkowskiSigmoid = _wrap(nn.Sigmoid, "MinkowskiSigmoid")
MinkowskiTanh = _wrap(nn.Tanh, "MinkowskiTanh")
MinkowskiSoftplus = _wrap(nn.Softplus, "MinkowskiSoftplus")


class MinkowskiGlobalPooling(nn.Module):
    _max = False

    def __init__(self, *_, **__):
        super().__init__()

    def forward(self, x: SparseTensor) -> SparseTensor:
        if _ag.wants_grad(x.F):
            f = _ag.GlobalPoolFunction.apply(x.F, x.coordinate_manager, x.coordinate_map_key.level, self._max)
        else:
            f = x.coordinate_manager.global_pool(x.coordinate_map_key.level, x.F, self._max)
        return SparseTensor(f, coordinate_manager=x.coordinate_manager, coordinate_map_key=CoordinateMapKey(0, origin=True))


MinkowskiGlobalAvgPooling = MinkowskiGlobalPooling


class MinkowskiGlobalMaxPooling(MinkowskiGlobalPooling):
    _max = True


class MinkowskiAvgPooling(nn.Module):
    """Constructed by models/resnet.py:53, never executed on the MinkFPN path."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, dimension=None):
        super().__init__()

    def forward(self, x):
        raise NotImplementedError("MinkowskiAvgPooling is not on the EgoNN/MinkLoc forward path")


class MinkowskiBroadcastMultiplication(nn.Module):
    def forward(self, x: SparseTensor, y: SparseTensor) -> SparseTensor:
        if _ag.wants_grad(x.F, y.F):
            return x._like(_ag.BroadcastMulFunction.apply(x.F, y.F, x.coordinate_manager, x.coordinate_map_key.level))
        return x._like(x.coordinate_manager.broadcast_mul(x.coordinate_map_key.level, x.F, y.F))


class _Functional:
    @staticmethod
    def normalize(x: SparseTensor, *a, **k):
        return x._like(torch.nn.functional.normalize(x.F, *a, **k))

    @staticmethod
    def relu(x: SparseTensor):
        return x._like(torch.relu(x.F))


MinkowskiFunctional = _Functional()


class _Utils:
    @staticmethod
    def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                        return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
        assert features is None and labels is None and not return_inverse and not return_maps_only, \
            "only the (coordinates, return_index, quantization_size) form used by datasets/quantization.py is supported"
        c = torch.as_tensor(coordinates).float()
        if isinstance(quantization_size, (list, tuple, np.ndarray, torch.Tensor)):
            # per-axis steps: ME divides the (N,3) coordinates by the step vector (one f32 divide per element) and floors;
            # egn_quantize's cartesian mode divides all axes by step[0], so the divide is done here, the floor there
            q = torch.as_tensor([float(v) for v in quantization_size], dtype=torch.float32, device=c.device)
            assert q.numel() == c.shape[1], "one quantisation step per coordinate axis"
            c, step = c / q, 1.0
        else:
            step = 1.0 if quantization_size is None else float(quantization_size)
        coords, ndx = _q._quantize_on_gpu(c, step, polar=False)
        return (coords, ndx) if return_index else coords

    batched_coordinates = staticmethod(_q.batched_coordinates)

    @staticmethod
    def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
        vol = tensor.shape[0] if tensor.dim() == 3 else 1
        fan = (tensor.shape[-2] if mode == "fan_in" else tensor.shape[-1]) * vol
        gain = nn.init.calculate_gain(nonlinearity, a)
        with torch.no_grad():
            return tensor.normal_(0, gain / np.sqrt(fan))


utils = _Utils()

from . import modules  # noqa: E402,F401


def install():
    """Make ``import MinkowskiEngine`` resolve to this package (for running the reference's own model code)."""
    me = sys.modules[__name__]
    sys.modules["MinkowskiEngine"] = me
    sys.modules["MinkowskiEngine.modules"] = modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = modules.resnet_block
    return me

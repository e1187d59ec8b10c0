"""``MinkowskiEngine``-shaped CPU shim on top of ``oracle.me_ops``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) - parity unpinned.  It exports exactly the ME
symbols the reference touches so that the UNMODIFIED reference files
``models/model_factory.py``, ``models/minkgl.py``, ``models/minkfpn.py``, ``layers/eca_block.py``,
``layers/pooling.py``, ``datasets/quantization.py`` import and run on the CPU when
``oracle/me_shim`` is put on ``sys.path`` (see ``oracle/ref_import.py``).  It is how
``tests/golden/make_golden.py`` produces golden vectors from the reference's own graph code.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from oracle import me_ops

__version__ = "0.5.4-oracle-shim"


class CoordinateMapKey:
    def __init__(self, stride: int, tag: str = ""):
        self.stride, self.tag = stride, tag

    def get_tensor_stride(self):
        return [self.stride] * 3

    def __eq__(self, o):
        return isinstance(o, CoordinateMapKey) and (self.stride, self.tag) == (o.stride, o.tag)

    def __hash__(self):
        return hash((self.stride, self.tag))


ORIGIN = "origin"


class SparseTensor:
    """ME.SparseTensor subset: .F .C .shape .tensor_stride .coordinate_manager .coordinate_map_key
    .decomposed_features ._batchwise_row_indices, + and += on a shared map."""

    def __init__(self, features, coordinates=None, coordinate_manager=None, coordinate_map_key=None,
                 tensor_stride=1, **_):
        self._F = features
        if coordinates is not None:
            c = coordinates.detach().cpu().numpy() if isinstance(coordinates, torch.Tensor) else np.asarray(coordinates)
            self.coordinate_manager = me_ops.CoordinateManager(c)
            self.coordinate_map_key = CoordinateMapKey(1)
        else:
            assert coordinate_manager is not None and coordinate_map_key is not None
            self.coordinate_manager = coordinate_manager
            self.coordinate_map_key = coordinate_map_key

    # -- accessors ---------------------------------------------------------------------------
    @property
    def F(self):
        return self._F

    @property
    def feats(self):
        return self._F

    def _coords_np(self):
        if self.coordinate_map_key.tag == ORIGIN:
            nb = self.coordinate_manager.n_batches
            c = np.zeros((nb, 4), dtype=np.int32)
            c[:, 0] = np.arange(nb)
            return c
        return self.coordinate_manager.coords(self.coordinate_map_key.stride)

    @property
    def C(self):
        return torch.from_numpy(self._coords_np())

    @property
    def coordinates(self):
        return self.C

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
        return [torch.from_numpy(r) for r in me_ops.batch_rows(self._coords_np(), self.coordinate_manager.n_batches)]

    @property
    def decomposed_features(self):
        return [self._F[r] for r in self._batchwise_row_indices]

    @property
    def decomposed_coordinates(self):
        c = self.C
        return [c[r, 1:] for r in self._batchwise_row_indices]

    def _like(self, feats):
        return SparseTensor(feats, coordinate_manager=self.coordinate_manager,
                            coordinate_map_key=self.coordinate_map_key)

    def __add__(self, other):
        assert self.coordinate_map_key == other.coordinate_map_key
        return self._like(self._F + other._F)

    def __iadd__(self, other):
        assert self.coordinate_map_key == other.coordinate_map_key
        self._F = self._F + other._F
        return self


# ------------------------------------------------------------------------------------------------
class MinkowskiConvolution(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3 and not bias
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = kernel_size, stride, dilation
        kv = kernel_size ** 3
        shape = (in_channels, out_channels) if kv == 1 else (kv, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(shape))
        self.bias = None

    def forward(self, x: SparseTensor) -> SparseTensor:
        f, s = me_ops.convolution(x.coordinate_manager, x.F, x.coordinate_map_key.stride, self.kernel,
                                  self.kernel_size, self.stride, self.dilation)
        return SparseTensor(f, coordinate_manager=x.coordinate_manager, coordinate_map_key=CoordinateMapKey(s))


class MinkowskiConvolutionTranspose(MinkowskiConvolution):
    def forward(self, x: SparseTensor) -> SparseTensor:
        f, s = me_ops.convolution_transpose(x.coordinate_manager, x.F, x.coordinate_map_key.stride, self.kernel,
                                            self.kernel_size, self.stride)
        return SparseTensor(f, coordinate_manager=x.coordinate_manager, coordinate_map_key=CoordinateMapKey(s))


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
MinkowskiDropout = _wrap(nn.Dropout, "MinkowskiDropout")


class MinkowskiGlobalPooling(nn.Module):
    """Per-batch mean, output on the origin map in batch order."""
    _max = False

    def __init__(self, *_, **__):
        super().__init__()

    def forward(self, x: SparseTensor) -> SparseTensor:
        fn = me_ops.global_max_pool if self._max else me_ops.global_avg_pool
        f = fn(x.F, x._coords_np(), x.coordinate_manager.n_batches)
        return SparseTensor(f, coordinate_manager=x.coordinate_manager,
                            coordinate_map_key=CoordinateMapKey(0, ORIGIN))


MinkowskiGlobalAvgPooling = MinkowskiGlobalPooling
MinkowskiGlobalSumPooling = None  # not used by the reference


class MinkowskiGlobalMaxPooling(MinkowskiGlobalPooling):
    _max = True


class MinkowskiAvgPooling(nn.Module):
    """Constructor only (models/resnet.py:53 builds one that the MinkFPN forward never calls)."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, dimension=None):
        super().__init__()

    def forward(self, x):
        raise NotImplementedError("MinkowskiAvgPooling is constructed but never executed by the reference path")


class MinkowskiBroadcastMultiplication(nn.Module):
    def forward(self, x: SparseTensor, y: SparseTensor) -> SparseTensor:
        return x._like(me_ops.broadcast_mul(x.F, x._coords_np(), y.F))


class _Functional:
    @staticmethod
    def normalize(x: SparseTensor, *a, **k):
        return x._like(torch.nn.functional.normalize(x.F, *a, **k))

    @staticmethod
    def relu(x: SparseTensor):
        return x._like(torch.relu(x.F))


MinkowskiFunctional = _Functional()


class _Utils:
    sparse_quantize = staticmethod(me_ops.sparse_quantize)
    batched_coordinates = staticmethod(me_ops.batched_coordinates)

    @staticmethod
    def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
        # ME: fan_in = Cin * volume, fan_out = Cout * volume for a (K,Cin,Cout) kernel.
        vol = tensor.shape[0] if tensor.dim() == 3 else 1
        fan = (tensor.shape[-2] if mode == "fan_in" else tensor.shape[-1]) * vol
        gain = nn.init.calculate_gain(nonlinearity, a)
        with torch.no_grad():
            return tensor.normal_(0, gain / np.sqrt(fan))


utils = _Utils()

from . import modules  # noqa: E402,F401

"""EgoNN model API on the CUDA engine: ``model_factory(model_params) -> nn.Module`` whose
``forward(batch)`` returns the same dict as the reference (models/model_factory.py:12-78,
models/minkgl.py:228-315) and whose ``state_dict`` has the reference's key names and shapes
(SURVEY.md Appendix B), so ``weights/model_egonn_20210916_1104.pth`` loads unchanged.

Two execution paths share the parameters:
  * ``forward`` in ``eval()`` mode - ONE ``egn_forward`` call: the whole network scheduled by the C++ engine
                            (eval-mode BatchNorm folded into the convolution epilogues).
  * ``forward_layerwise`` - the reference's layer-by-layer walk on ``egonn_b200.minkowski`` operators
                            (each a C-ABI call); used to cross-check the fused path, and - with autograd enabled - it is
                            what ``forward`` runs in ``train()`` mode: batch-statistics BatchNorm and gradients through
                            ``egonn_b200.autograd`` (the training step of training/trainer.py:141-195).
Both are CUDA-only.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import minkowski as ME
from . import weights as W
from .engine import Engine
from .minkowski.modules.resnet_block import BasicBlock
from .quantization import Quantizer


# ---- layers/eca_block.py -------------------------------------------------------------------------------
class ECALayer(nn.Module):
    """Efficient channel attention gate (layers/eca_block.py:11-36)."""

    def __init__(self, channels, gamma=2, b=1):
        super().__init__()
        t = int(abs((np.log2(channels) + b) / gamma))
        k = t if t % 2 else t + 1
        self.avg_pool = ME.MinkowskiGlobalPooling()
        self.conv = nn.Conv1d(1, 1, kernel_size=k, padding=(k - 1) // 2, bias=False)
        self.sigmoid = nn.Sigmoid()
        self.broadcast_mul = ME.MinkowskiBroadcastMultiplication()

    def forward(self, x):
        pooled = self.avg_pool(x)
        gate = self.sigmoid(self.conv(pooled.F.unsqueeze(1)).squeeze(1))
        return self.broadcast_mul(x, pooled._like(gate))


class ECABasicBlock(BasicBlock):
    """Residual block with the ECA gate before the skip addition (layers/eca_block.py:39-73)."""

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, dimension=3):
        super().__init__(inplanes, planes, stride=stride, dilation=dilation, downsample=downsample, dimension=dimension)
        self.eca = ECALayer(planes, gamma=2, b=1)

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.eca(self.norm2(self.conv2(out)))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out)


# ---- layers/pooling.py -----------------------------------------------------------------------------------
class GeM(nn.Module):
    """layers/pooling.py:72-86."""

    def __init__(self, input_dim, p=3, eps=1e-6):
        super().__init__()
        self.input_dim = self.output_dim = input_dim
        self.p = nn.Parameter(torch.ones(1) * p)
        self.eps = eps
        self.f = ME.MinkowskiGlobalAvgPooling()

    def forward(self, x):
        t = self.f(x._like(x.F.clamp(min=self.eps).pow(self.p)))
        return t.F.pow(1. / self.p)


class MAC(nn.Module):
    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = self.output_dim = input_dim
        self.f = ME.MinkowskiGlobalMaxPooling()

    def forward(self, x):
        return self.f(x).F


class SPoC(nn.Module):
    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = self.output_dim = input_dim
        self.f = ME.MinkowskiGlobalAvgPooling()

    def forward(self, x):
        return self.f(x).F


class PoolingWrapper(nn.Module):
    """layers/pooling.py:13-43 (NetVLAD variants are outside the EgoNN path and are rejected)."""

    def __init__(self, pool_method, in_dim, output_dim):
        super().__init__()
        self.pool_method, self.in_dim, self.output_dim = pool_method, in_dim, output_dim
        table = {"MAC": MAC, "SPoC": SPoC, "GeM": GeM}
        if pool_method not in table:
            raise NotImplementedError("Unknown pooling method: {}".format(pool_method))
        assert in_dim == output_dim
        self.pooling = table[pool_method](input_dim=in_dim)

    def forward(self, x):
        return self.pooling(x)


# ---- models/minkgl.py ------------------------------------------------------------------------------------
class MinkHead(nn.Module):
    """Top-down FPN head (models/minkgl.py:14-60)."""

    def __init__(self, in_levels: List[int], in_channels: List[int], out_channels: int):
        assert len(in_levels) > 0 and len(in_levels) == len(in_channels)
        super().__init__()
        self.in_levels, self.in_channels, self.out_channels = in_levels, in_channels, out_channels
        self.min_level, self.max_level = min(in_levels), max(in_levels)
        assert self.min_level > 0
        self.in_d = dict(zip(in_levels, in_channels))
        self.conv1x1 = nn.ModuleDict()
        self.tconv = nn.ModuleDict()
        for lv in range(self.min_level + 1, self.max_level + 1):
            self.tconv[str(lv)] = ME.MinkowskiConvolutionTranspose(out_channels, out_channels, kernel_size=2, stride=2, dimension=3)
        for lv, ch in self.in_d.items():
            self.conv1x1[str(lv)] = ME.MinkowskiConvolution(ch, out_channels, kernel_size=1, stride=1, dimension=3)

    def forward(self, x: Dict[int, ME.SparseTensor]):
        y = self.conv1x1[str(self.max_level)](x[self.max_level])
        for lv in range(self.max_level - 1, self.min_level - 1, -1):
            y = self.tconv[str(lv + 1)](y)
            if lv in self.in_d:
                y = y + self.conv1x1[str(lv)](x[lv])
        assert y.shape[1] == self.out_channels
        return y


class MinkTrunk(nn.Module):
    """Bottom-up trunk (models/minkgl.py:68-153)."""

    def __init__(self, in_channels: int, planes: List[int], layers: List[int] = None, conv0_kernel_size: int = 5,
                 block=BasicBlock, min_out_level: int = 1):
        super().__init__()
        self.in_channels, self.planes = in_channels, planes
        self.layers = [1] * len(planes) if layers is None else layers
        assert len(self.layers) == len(planes) and min_out_level <= len(planes)
        assert all(1 <= n <= 4 for n in self.layers), "the CUDA engine schedules 1..4 blocks per level"
        self.conv0_kernel_size, self.block, self.min_out_level = conv0_kernel_size, block, min_out_level
        self.num_bottom_up = len(planes)
        self.init_dim = planes[0]
        self.convs, self.bn, self.blocks = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        self.inplanes = planes[0]
        self.convs["0"] = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=conv0_kernel_size, dimension=3)
        self.bn["0"] = ME.MinkowskiBatchNorm(self.inplanes)
        for i, (plane, n_blocks) in enumerate(zip(planes, self.layers), start=1):
            self.convs[str(i)] = ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2, dimension=3)
            self.bn[str(i)] = ME.MinkowskiBatchNorm(self.inplanes)
            self.blocks[str(i)] = self._make_layer(block, plane, n_blocks)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.weight_initialization()

    def weight_initialization(self):
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                ME.MinkowskiConvolution(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, dimension=3),
                ME.MinkowskiBatchNorm(planes * block.expansion))
        seq = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=downsample, dimension=3)]
        self.inplanes = planes * block.expansion
        seq += [block(self.inplanes, planes, stride=1, dilation=dilation, dimension=3) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def forward(self, x):
        y = {}
        x = self.relu(self.bn["0"](self.convs["0"](x)))
        for i in range(1, len(self.layers) + 1):
            x = self.relu(self.bn[str(i)](self.convs[str(i)](x)))
            x = self.blocks[str(i)](x)
            if i >= self.min_out_level:
                y[i] = x
        return y


def _regressor(in_channels, hidden, out, act):
    return nn.Sequential(ME.MinkowskiLinear(in_channels, hidden), ME.MinkowskiReLU(inplace=True),
                         ME.MinkowskiLinear(hidden, out), *([act] if act is not None else []))


class KeypointRegressor(nn.Module):
    """models/minkgl.py:175-185."""

    def __init__(self, in_channels: int, reduction: int = 2):
        super().__init__()
        self.net = _regressor(in_channels, in_channels // reduction, 3, ME.MinkowskiTanh())

    def forward(self, x):
        return self.net(x)


class SigmaRegressor(nn.Module):
    """models/minkgl.py:188-204 (the reference's lower bound is commented out there, so none here)."""

    def __init__(self, in_channels: int, reduction: int = 2):
        super().__init__()
        self.sigma_lower_bound = 1e-6
        self.net = _regressor(in_channels, in_channels // reduction, 1, ME.MinkowskiSoftplus())

    def forward(self, x):
        return self.net(x)


class DescriptorDecoder(nn.Module):
    """models/minkgl.py:207-225."""

    def __init__(self, in_channels: int, out_channels: int, normalize=True):
        super().__init__()
        self.normalize = normalize
        self.net = _regressor(in_channels, out_channels + (in_channels - out_channels) // 2, out_channels, None)

    def forward(self, x):
        x = self.net(x)
        return ME.MinkowskiFunctional.normalize(x) if self.normalize else x


class _EngineModel(nn.Module):
    """Shared host plumbing of the engine-backed models: one engine context per (device, stream), weights packed
    lazily and re-packed when a parameter / buffer changes."""

    def _init_engine_state(self):
        self.l2_resident_weights = True   # keep the weight blob in a persisting L2 window (egn_weights_resident)
        self._engine: Optional[Engine] = None
        self._engines: Dict = {}
        self._packed = None               # (signature, blob on device, Net)

    def _signature(self, device):
        return (str(device), bool(getattr(self, "ignore_keypoint_regressor", False)),
                tuple((id(p), p._version) for p in self.parameters()), tuple((id(b), b._version) for b in self.buffers()))

    def _pack(self, device):
        sig = self._signature(device)
        if self._packed is None or self._packed[0] != sig:
            sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            blob, net = self._pack_weights(sd)
            self._packed = (sig, blob.to(device), net)
            if self.l2_resident_weights:
                for eng in self._engines.values():
                    eng.weights_resident(self._packed[1])
        return self._packed[1], self._packed[2]

    def _engine_for(self, device) -> Engine:
        """One engine context (coordinate manager + scratch arenas) per (device, CUDA stream): running the model under
        different ``torch.cuda.stream`` contexts gives independent contexts whose forwards overlap on the GPU."""
        key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = Engine(device)
            if self._packed is not None and self.l2_resident_weights:
                eng.weights_resident(self._packed[1])
        self._engine = eng                                    # the engine of the most recent call (taps, counters)
        return eng


class MinkGL(_EngineModel):
    """models/minkgl.py:228-315, executed by the CUDA engine."""

    def __init__(self, trunk: MinkTrunk, local_head: MinkHead = None, local_descriptor_size: int = None,
                 local_normalize: bool = True, global_head: MinkHead = None, global_descriptor_size: int = None,
                 global_pool_method: str = "GeM", global_normalize: bool = False, quantizer: Quantizer = None):
        assert quantizer is not None
        super().__init__()
        assert global_pool_method == "GeM" and not global_normalize and local_normalize, \
            "only the shipped egonn configuration (GeM, un-normalised global, normalised local) is scheduled"
        self.trunk, self.global_head = trunk, global_head
        self.global_pool_method = global_pool_method
        self.global_channels = global_head.out_channels
        self.global_pooling = PoolingWrapper(global_pool_method, self.global_channels, self.global_channels)
        self.global_normalize = global_normalize
        self.global_descriptor_size = global_descriptor_size
        self.global_descriptor_decoder = DescriptorDecoder(self.global_channels, global_descriptor_size, normalize=False)
        self.local_head = local_head
        if local_head is not None:
            self.local_descriptor_size, self.local_normalize = local_descriptor_size, local_normalize
            c = local_head.out_channels
            self.local_keypoint_regressor = KeypointRegressor(c, reduction=2)
            self.local_sigma_regressor = SigmaRegressor(c, reduction=2)
            self.local_descriptor_decoder = DescriptorDecoder(c, local_descriptor_size, normalize=local_normalize)
        self.quantizer = quantizer
        self.ignore_keypoint_regressor = False
        self._init_engine_state()
        self.last: Dict = {}      # extras of the last forward: local coordinates, batch offsets, level sizes

    # -- weights -> engine ------------------------------------------------------------------------------------
    def _pack_weights(self, sd):
        return W.pack_egonn(sd, self.quantizer.describe(), global_levels=self.global_head.in_levels,
                            local_levels=self.local_head.in_levels if self.local_head is not None else (),
                            ignore_keypoint_regressor=self.ignore_keypoint_regressor, bn_eps=self.trunk.bn["0"].bn.eps)


    # -- fused path ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_packed(self, batch: Dict[str, torch.Tensor], disable_global_head=False, disable_local_head=False) -> Dict:
        """Same computation as ``forward`` but local outputs stay packed: (n,128)/(n,3)/(n,1) tensors in canonical
        row order + ``local_coords`` (n,4) + ``local_offsets`` (B+1) on the device - no per-cloud Python lists."""
        assert not self.training, "the fused path folds eval-mode BatchNorm: call model.eval() (train mode runs through forward())"
        coords, feats = batch["coords"], batch["features"]
        if not coords.is_cuda or not feats.is_cuda:
            raise RuntimeError("egonn_b200 has no CPU path: move batch['coords'] and batch['features'] to a CUDA device")
        eng = self._engine_for(coords.device)
        blob, net = self._pack(coords.device)
        with torch.cuda.device(coords.device):
            info = eng.build(coords)
            want_local = self.local_head is not None and not disable_local_head
            out = eng.forward(net, blob, feats, want_global=not disable_global_head, want_local=want_local)
            if want_local:
                lvl = out["local_level"]
                out["local_coords"] = eng.level_coords(lvl)
                out["local_offsets"] = eng.batch_offsets(lvl)
        out["n_rows"] = info.n_rows
        return out

    @torch.no_grad()
    def forward_points(self, points: torch.Tensor, cloud_offsets: torch.Tensor, disable_global_head=False,
                       disable_local_head=False) -> Dict:
        """Fused ingest + forward for the caller sequence of eval/evaluate.py:331-338: raw points of B clouds
        concatenated ((n,3) f32, CUDA) + first-point offsets ((B+1,) int32, CUDA) -> the packed outputs of
        ``forward_packed`` on the quantised, batched, all-ones-feature input - one sort, one host sync."""
        assert not self.training, "the fused path folds eval-mode BatchNorm: call model.eval() (train mode runs through forward())"
        if not points.is_cuda:
            raise RuntimeError("egonn_b200 has no CPU path: move the points to a CUDA device")
        eng = self._engine_for(points.device)
        blob, net = self._pack(points.device)
        q = self.quantizer.describe()
        with torch.cuda.device(points.device):
            info = eng.build_points(points, cloud_offsets, q["step"], q["coordinates"] == "polar")
            want_local = self.local_head is not None and not disable_local_head
            out = eng.forward(net, blob, None, want_global=not disable_global_head, want_local=want_local)
            if want_local:
                lvl = out["local_level"]
                out["local_coords"] = eng.level_coords(lvl)
                out["local_offsets"] = eng.batch_offsets(lvl)
        out["n_rows"] = info.n_rows
        return out

    def forward(self, batch: Dict[str, torch.Tensor], disable_global_head: bool = False, disable_local_head: bool = False):
        if self.training:                                           # training/trainer.py:124-126,160-162: layer walk with autograd
            return self._walk(batch, disable_global_head, disable_local_head)
        with torch.no_grad():
            return self._forward_eval(batch, disable_global_head, disable_local_head)

    def _forward_eval(self, batch, disable_global_head, disable_local_head):
        p = self.forward_packed(batch, disable_global_head, disable_local_head)
        y = {}
        if "global" in p:
            assert p["global"].dim() == 2 and p["global"].shape[1] == self.global_descriptor_size
            y["global"] = p["global"]
        if "descriptors" in p:
            off = p["local_offsets"].tolist()                       # the only host sync of the call
            cut = lambda t: [t[off[b]:off[b + 1]] for b in range(len(off) - 1)]
            y["descriptors"], y["keypoints"], y["sigma"] = cut(p["descriptors"]), cut(p["keypoints"]), cut(p["sigma"])
            self.last = {"local_coords": cut(p["local_coords"]), "n_rows": p["n_rows"]}
        return y

    # -- reference-style layer walk (cross-check path) --------------------------------------------------------------
    @torch.no_grad()
    def forward_layerwise(self, batch, disable_global_head=False, disable_local_head=False):
        return self._walk(batch, disable_global_head, disable_local_head)

    def _walk(self, batch, disable_global_head=False, disable_local_head=False):
        """models/minkgl.py:267-315 operator by operator; a new coordinate manager (engine context) per call, as
        ``ME.SparseTensor(features, coordinates=...)`` creates one - several forwards may be alive before one backward
        (training/trainer.py:178-188).  CPU tensors are rejected by the engine (no CPU path)."""
        x = ME.SparseTensor(batch["features"], coordinates=batch["coords"])
        x = self.trunk(x)
        y = {}
        if not disable_global_head:
            g = self.global_descriptor_decoder(self.global_head(x))
            y["global"] = self.global_pooling(g)
        if self.local_head is not None and not disable_local_head:
            xl = self.local_head(x)
            y["descriptors"] = self.local_descriptor_decoder(xl).decomposed_features
            kp = self.local_keypoint_regressor(xl)
            off = torch.zeros_like(kp.F) if self.ignore_keypoint_regressor else kp.F
            pos = self.quantizer.keypoint_position(kp.C[:, 1:], kp.tensor_stride, off)
            y["keypoints"] = [pos[ndx] for ndx in kp._batchwise_row_indices]
            y["sigma"] = self.local_sigma_regressor(xl).decomposed_features
        return y

    def print_info(self):
        n = sum(p.nelement() for p in self.parameters())
        print(f"Model class: {type(self).__name__}   # parameters: {n / 1000:.1f}k   (egonn_b200 CUDA engine)")


# ---- models/minkfpn.py, models/minkloc.py, third_party/minkloc3d/minkloc.py ------------------------------------------
class MinkFPN(nn.Module):
    """Parameter tree of the reference's MinkFPN (models/minkfpn.py:9-93 on top of models/resnet.py:31-117): bottom-up
    conv0 + n x (2x2x2 stride-2 conv, BN, block), top-down 1x1 laterals + transposed convolutions.  The module order
    (convs, bn, blocks, tconvs, conv1x1, conv0, bn0) reproduces the reference's state_dict key order."""

    def __init__(self, in_channels, out_channels, num_top_down=1, conv0_kernel_size=5, block=BasicBlock,
                 layers=(1, 1, 1), planes=(32, 64, 64)):
        super().__init__()
        assert len(layers) == len(planes) and 1 <= len(layers) and 0 <= num_top_down <= len(layers)
        assert all(1 <= n <= 4 for n in layers), "the CUDA engine schedules 1..4 blocks per level"
        self.num_bottom_up, self.num_top_down = len(layers), num_top_down
        self.conv0_kernel_size, self.block, self.layers, self.planes = conv0_kernel_size, block, layers, planes
        self.lateral_dim, self.init_dim = out_channels, planes[0]
        self.convs, self.bn, self.blocks = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.tconvs, self.conv1x1 = nn.ModuleList(), nn.ModuleList()
        self.inplanes = planes[0]
        conv0 = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=conv0_kernel_size, dimension=3)
        bn0 = ME.MinkowskiBatchNorm(self.inplanes)
        for plane, n_blocks in zip(planes, layers):
            self.convs.append(ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2, dimension=3))
            self.bn.append(ME.MinkowskiBatchNorm(self.inplanes))
            self.blocks.append(self._make_layer(block, plane, n_blocks))
        for i in range(num_top_down):
            self.conv1x1.append(ME.MinkowskiConvolution(planes[-1 - i], out_channels, kernel_size=1, stride=1, dimension=3))
            self.tconvs.append(ME.MinkowskiConvolutionTranspose(out_channels, out_channels, kernel_size=2, stride=2, dimension=3))
        last = planes[-1 - num_top_down] if num_top_down < self.num_bottom_up else planes[0]
        self.conv1x1.append(ME.MinkowskiConvolution(last, out_channels, kernel_size=1, stride=1, dimension=3))
        self.conv0, self.bn0 = conv0, bn0
        self.relu = ME.MinkowskiReLU(inplace=True)
        for m in self.modules():                                      # models/resnet.py:72-79
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _make_layer(self, block, planes, blocks=1):                   # models/resnet.py:81-97
        downsample = None
        if self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                ME.MinkowskiConvolution(self.inplanes, planes * block.expansion, kernel_size=1, stride=1, dimension=3),
                ME.MinkowskiBatchNorm(planes * block.expansion))
        seq = [block(self.inplanes, planes, stride=1, dilation=1, downsample=downsample, dimension=3)]
        self.inplanes = planes * block.expansion
        seq += [block(self.inplanes, planes, stride=1, dilation=1, dimension=3) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def forward(self, x):                                             # layer-wise operator walk (models/minkfpn.py:65-93)
        maps = []
        x = self.relu(self.bn0(self.conv0(x)))
        if self.num_top_down == self.num_bottom_up:
            maps.append(x)
        for ndx, (conv, bn, block) in enumerate(zip(self.convs, self.bn, self.blocks)):
            x = block(self.relu(bn(conv(x))))
            if self.num_bottom_up - 1 - self.num_top_down <= ndx < len(self.convs) - 1:
                maps.append(x)
        x = self.conv1x1[0](x)
        for ndx, tconv in enumerate(self.tconvs):
            x = tconv(x)
            x = x + self.conv1x1[ndx + 1](maps[-ndx - 1])
        return x


class _GlobalOnlyModel(_EngineModel):
    """MinkFPN backbone + global pooling, executed by the CUDA engine (models/minkloc.py:44-61)."""

    def _finish_init(self, quantizer):
        self.quantizer = quantizer
        self._init_engine_state()

    def _quant_desc(self):
        return self.quantizer.describe() if self.quantizer is not None else {"coordinates": "cartesian", "step": 1.0}

    def forward(self, batch, disable_local_head: bool = True):
        assert disable_local_head, "this model has only the global head"
        if self.training:                                           # layer walk with autograd (egonn_b200.autograd)
            return self._walk(batch)
        with torch.no_grad():
            return self._forward_eval(batch["coords"], batch["features"])

    def _forward_eval(self, coords, feats):
        if not coords.is_cuda or not feats.is_cuda:
            raise RuntimeError("egonn_b200 has no CPU path: move batch['coords'] and batch['features'] to a CUDA device")
        eng = self._engine_for(coords.device)
        blob, net = self._pack(coords.device)
        with torch.cuda.device(coords.device):
            eng.build(coords)
            out = eng.forward(net, blob, feats, want_global=True, want_local=False)
        x = out["global"]
        assert x.dim() == 2 and x.shape[1] == self.output_dim
        return {"global": x}

    @torch.no_grad()
    def forward_layerwise(self, batch):
        return self._walk(batch)

    def _walk(self, batch):
        x = self.backbone(ME.SparseTensor(batch["features"], coordinates=batch["coords"]))
        return {"global": self.pooling(x)}


class MinkLoc(_GlobalOnlyModel):
    """models/minkloc.py:13-61."""

    def __init__(self, in_channels, feature_size, output_dim, planes, layers, num_top_down, conv0_kernel_size,
                 block="BasicBlock", pooling_method="GeM", quantizer=None):
        super().__init__()
        blocks = {"BasicBlock": BasicBlock, "ECABasicBlock": ECABasicBlock}
        if block not in blocks:
            raise NotImplementedError("Unsupported network block: {}".format(block))
        assert in_channels == 1
        self.in_channels, self.feature_size, self.output_dim, self.block = in_channels, feature_size, output_dim, block
        self.pooling_method = pooling_method
        self.backbone = MinkFPN(in_channels=in_channels, out_channels=feature_size, num_top_down=num_top_down,
                                conv0_kernel_size=conv0_kernel_size, block=blocks[block], layers=layers, planes=planes)
        self.pooling = PoolingWrapper(pool_method=pooling_method, in_dim=feature_size, output_dim=output_dim)
        self.pooled_feature_size = self.pooling.output_dim
        self._finish_init(quantizer)

    def _pack_weights(self, sd):
        return W.pack_minkfpn(sd, self._quant_desc(), self.backbone.num_top_down, pool_method=self.pooling_method,
                              pool_key="pooling.pooling.p", bn_eps=self.backbone.bn0.bn.eps)


class MinkLoc3D(_GlobalOnlyModel):
    """third_party/minkloc3d/minkloc.py:9-58 (MinkFPN 32/64/64, one top-down step, 256 channels, its own GeM)."""

    def __init__(self, quantizer=None):
        super().__init__()
        self.feature_size = self.output_dim = 256
        self.backbone = MinkFPN(in_channels=1, out_channels=self.feature_size, num_top_down=1, conv0_kernel_size=5,
                                layers=[1, 1, 1], planes=[32, 64, 64])
        self.pooling = GeM(input_dim=self.feature_size)
        self._finish_init(quantizer)

    def _pack_weights(self, sd):
        return W.pack_minkfpn(sd, self._quant_desc(), 1, pool_method="GeM", pool_key="pooling.p", bn_eps=self.backbone.bn0.bn.eps)


# ---- models/model_factory.py -----------------------------------------------------------------------------------
def create_egonn_model(model_params):
    """models/model_factory.py:31-78."""
    if model_params.model != "egonn":
        raise NotImplementedError(f"Unknown model: {model_params.model}")
    planes = [32, 64, 64, 128, 128, 128, 128]
    layers = [1] * 7
    global_in_levels, global_map_channels, global_descriptor_size = [5, 6, 7], 128, 256
    local_in_levels, local_map_channels, local_descriptor_size = [3, 4], 64, 128
    head_global = MinkHead(global_in_levels, [planes[i - 1] for i in global_in_levels], global_map_channels)
    head_local = MinkHead(local_in_levels, [planes[i - 1] for i in local_in_levels], local_map_channels)
    min_out_level = min(len(planes), min(global_in_levels), min(local_in_levels))
    trunk = MinkTrunk(in_channels=1, planes=planes, layers=layers, conv0_kernel_size=5, block=ECABasicBlock,
                      min_out_level=min_out_level)
    return MinkGL(trunk, local_head=head_local, local_descriptor_size=local_descriptor_size, local_normalize=True,
                  global_head=head_global, global_descriptor_size=global_descriptor_size, global_pool_method="GeM",
                  global_normalize=False, quantizer=model_params.quantizer)


def model_factory(model_params):
    """models/model_factory.py:12-28: 'MinkLoc', 'MinkLoc3D' and the 'egonn' family, all scheduled by the CUDA engine."""
    in_channels = 1
    if model_params.model == "MinkLoc":
        return MinkLoc(in_channels=in_channels, feature_size=model_params.feature_size, output_dim=model_params.output_dim,
                       planes=model_params.planes, layers=model_params.layers, num_top_down=model_params.num_top_down,
                       conv0_kernel_size=model_params.conv0_kernel_size, block=model_params.block,
                       pooling_method=model_params.pooling, quantizer=model_params.quantizer)
    if model_params.model == "MinkLoc3D":
        return MinkLoc3D(quantizer=model_params.quantizer)
    if "egonn" in (model_params.model or ""):
        return create_egonn_model(model_params)
    raise NotImplementedError("Model not implemented: {}".format(model_params.model))

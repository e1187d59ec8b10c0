"""CPU restatement of the EgoNN forward (``MinkGL.forward``, models/minkgl.py:267-315) on top of
``oracle.me_ops`` - no reference import, so it also runs on the GPU box as the parity checker.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) - parity unpinned.  The graph restated here is
checked against the UNMODIFIED reference graph code executed on ``oracle/me_shim`` through the
golden vectors in ``tests/golden`` (tests/test_oracle_golden.py).

Weights come in as the reference ``state_dict`` (key names of SURVEY.md Appendix B).  All outputs are
returned in CANONICAL order (rows sorted lexicographically by (b,x,y,z)) together with their
coordinates, because ME row order is not a contract.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle import me_ops

# models/model_factory.py:37-49 - the 'egonn' architecture
EGONN = dict(planes=[32, 64, 64, 128, 128, 128, 128], conv0_kernel_size=5,
             global_levels=[5, 6, 7], global_channels=128, global_descriptor_size=256,
             local_levels=[3, 4], local_channels=64, local_descriptor_size=128)


def _bn(sd, prefix, x):
    return me_ops.batch_norm_eval(x, sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"],
                                  sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"])


def eca_layer(sd, prefix, x, coords, n_batches):
    """layers/eca_block.py:21-36: per-cloud mean -> Conv1d over channels (zero pad, no bias) ->
    sigmoid -> broadcast multiply."""
    y = me_ops.global_avg_pool(x, coords, n_batches)                       # (B,C)
    w = sd[prefix + ".conv.weight"]                                        # (1,1,k)
    k = w.shape[-1]
    y = F.conv1d(y.unsqueeze(1), w, padding=(k - 1) // 2).squeeze(1)       # (B,C)
    y = torch.sigmoid(y)
    return me_ops.broadcast_mul(x, coords, y)


def eca_basic_block(sd, prefix, cm, x, stride, n_batches, acc64=False, eca=True):
    """layers/eca_block.py:56-73 (ECABasicBlock.forward); ``eca=False`` gives ME's plain BasicBlock."""
    coords = cm.coords(stride)
    out, _ = me_ops.convolution(cm, x, stride, sd[prefix + ".conv1.kernel"], 3, acc64=acc64)
    out = torch.relu(_bn(sd, prefix + ".norm1", out))
    out, _ = me_ops.convolution(cm, out, stride, sd[prefix + ".conv2.kernel"], 3, acc64=acc64)
    out = _bn(sd, prefix + ".norm2", out)
    if eca:
        out = eca_layer(sd, prefix + ".eca", out, coords, n_batches)
    residual = x
    if prefix + ".downsample.0.kernel" in sd:
        residual, _ = me_ops.convolution(cm, x, stride, sd[prefix + ".downsample.0.kernel"], 1, acc64=acc64)
        residual = _bn(sd, prefix + ".downsample.1", residual)
    return torch.relu(out + residual)


def trunk(sd, cm, feats, n_levels, n_batches, acc64=False, keep=None):
    """MinkTrunk.forward, models/minkgl.py:136-153."""
    x, _ = me_ops.convolution(cm, feats, 1, sd["trunk.convs.0.kernel"], int(round(sd["trunk.convs.0.kernel"].shape[0] ** (1 / 3))),
                              acc64=acc64)
    x = torch.relu(_bn(sd, "trunk.bn.0", x))
    if keep is not None:
        keep["conv0"] = x
    y = {}
    stride = 1
    for i in range(1, n_levels + 1):
        x, stride = me_ops.convolution(cm, x, stride, sd[f"trunk.convs.{i}.kernel"], 2, stride=2, acc64=acc64)
        x = torch.relu(_bn(sd, f"trunk.bn.{i}", x))
        if keep is not None:
            keep[f"down{i}"] = x
        j = 0
        while f"trunk.blocks.{i}.{j}.conv1.kernel" in sd:              # layers[i] blocks per level (models/minkgl.py:121-134)
            x = eca_basic_block(sd, f"trunk.blocks.{i}.{j}", cm, x, stride, n_batches, acc64=acc64,
                                eca=f"trunk.blocks.{i}.{j}.eca.conv.weight" in sd)
            j += 1
        y[i] = x
    return y


def head(sd, prefix, cm, x: Dict[int, torch.Tensor], levels: List[int], acc64=False):
    """MinkHead.forward, models/minkgl.py:46-60."""
    lo, hi = min(levels), max(levels)
    y, _ = me_ops.convolution(cm, x[hi], 1 << hi, sd[f"{prefix}.conv1x1.{hi}.kernel"], 1, acc64=acc64)
    for level in range(hi - 1, lo - 1, -1):
        y, _ = me_ops.convolution_transpose(cm, y, 1 << (level + 1), sd[f"{prefix}.tconv.{level + 1}.kernel"], acc64=acc64)
        if level in levels:
            lat, _ = me_ops.convolution(cm, x[level], 1 << level, sd[f"{prefix}.conv1x1.{level}.kernel"], 1, acc64=acc64)
            y = y + lat
    return y


def _mlp(sd, prefix, x):
    x = F.linear(x, sd[prefix + ".net.0.linear.weight"], sd[prefix + ".net.0.linear.bias"])
    x = torch.relu(x)
    return F.linear(x, sd[prefix + ".net.2.linear.weight"], sd[prefix + ".net.2.linear.bias"])


def gem(x, coords, p, n_batches, eps=1e-6):
    """layers/pooling.py:82-86."""
    t = x.clamp(min=eps).pow(p)
    t = me_ops.global_avg_pool(t, coords, n_batches)
    return t.pow(1.0 / p)


def keypoint_position(coords_xyz: torch.Tensor, stride: int, kp_offset: Optional[torch.Tensor], quant: dict):
    """datasets/quantization.py:60-72 (polar) and :93-103 (cartesian)."""
    if quant["coordinates"] == "cartesian":
        q = quant["step"]
        centres = (coords_xyz + 0.5) * q
        size = torch.tensor([stride] * 3, dtype=torch.float) * q
        return centres if kp_offset is None else centres + kp_offset * size / 2.
    qs = torch.tensor(quant["step"], dtype=torch.float)
    centres = (coords_xyz + 0.5) * qs
    size = torch.tensor([stride] * 3, dtype=torch.float) * qs
    kp = centres + kp_offset * size / 2.
    theta = np.pi * (kp[:, 0] - 180.) / 180.
    return torch.stack([torch.cos(theta) * kp[:, 1], torch.sin(theta) * kp[:, 1], kp[:, 2]], dim=1)


def quantize(pc: torch.Tensor, quant: dict):
    """datasets/quantization.py:29-44 (PolarQuantizer.__call__), :79-85 (CartesianQuantizer.__call__)."""
    if quant["coordinates"] == "cartesian":
        return me_ops.sparse_quantize(pc, quantization_size=quant["step"], return_index=True)
    theta = 180. + torch.atan2(pc[:, 1], pc[:, 0]) * 180. / np.pi
    dist = torch.sqrt(pc[:, 0] ** 2 + pc[:, 1] ** 2)
    polar = torch.stack([theta, dist, pc[:, 2]], dim=1)
    polar = polar / torch.tensor(quant["step"], dtype=torch.float)
    return me_ops.sparse_quantize(polar, quantization_size=1., return_index=True)


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], coords, feats: torch.Tensor, quant: dict, arch: dict = EGONN,
            acc64: bool = False, keep_intermediates: bool = False, ignore_keypoint_regressor: bool = False):
    """MinkGL.forward (models/minkgl.py:267-315) for batch {'coords': (N,4) int32, 'features': (N,1) f32}.

    Returns a dict with, in canonical row order:
      'global' (B,256); 'coords_L3' (n,4) int32; 'descriptors' (n,128); 'keypoints' (n,3); 'sigma' (n,1);
      per-cloud lists 'descriptors_list' / 'keypoints_list' / 'sigma_list' (what the reference returns);
      'levels': {L: coords (N_L,4)} and, if keep_intermediates, 'features': {name: (coords, F)}.
    """
    c = coords.detach().cpu().numpy() if isinstance(coords, torch.Tensor) else np.asarray(coords)
    feats = feats.float()
    cm = me_ops.CoordinateManager(c)
    nb = cm.n_batches
    n_levels = len(arch["planes"])
    keep = {} if keep_intermediates else None
    x = trunk(sd, cm, feats, n_levels, nb, acc64=acc64, keep=keep)
    out = {}

    # global head -> decoder -> GeM  (models/minkgl.py:273-286)
    gl = arch["global_levels"]
    xg_head = head(sd, "global_head", cm, x, gl, acc64=acc64)
    xg = _mlp(sd, "global_descriptor_decoder", xg_head)
    cg = cm.coords(1 << min(gl))
    out["global"] = gem(xg, cg, sd["global_pooling.pooling.p"], nb)

    # local head (models/minkgl.py:288-308)
    ll = arch["local_levels"]
    sl = 1 << min(ll)
    xl = head(sd, "local_head", cm, x, ll, acc64=acc64)
    desc = F.normalize(_mlp(sd, "local_descriptor_decoder", xl), p=2, dim=1, eps=1e-12)
    kp_off = torch.tanh(_mlp(sd, "local_keypoint_regressor", xl))
    sigma = F.softplus(_mlp(sd, "local_sigma_regressor", xl))
    cl = cm.coords(sl)
    cxyz = torch.from_numpy(cl[:, 1:].astype(np.int64))
    # models/minkgl.py:296-302: the ablation switch feeds zeros instead of the regressed offset
    kp = keypoint_position(cxyz, sl, torch.zeros_like(kp_off) if ignore_keypoint_regressor else kp_off, quant)

    order = me_ops.canonical_order(cl)
    ot = torch.from_numpy(order)
    out["coords_L3"] = cl[order]
    out["descriptors"], out["keypoints"], out["sigma"] = desc[ot], kp[ot], sigma[ot]
    rows = me_ops.batch_rows(out["coords_L3"], nb)
    for name in ("descriptors", "keypoints", "sigma"):
        out[name + "_list"] = [out[name][torch.from_numpy(r)] for r in rows]

    out["levels"] = {}
    for L in range(0, n_levels + 1):
        cc = cm.coords(1 << L)
        out["levels"][L] = cc[me_ops.canonical_order(cc)]
    if keep_intermediates:
        fe = {}
        o0 = me_ops.canonical_order(cm.coords(1))
        fe["conv0"] = keep["conv0"][torch.from_numpy(o0)]
        for i in range(1, n_levels + 1):
            oi = torch.from_numpy(me_ops.canonical_order(cm.coords(1 << i)))
            fe[f"down{i}"] = keep[f"down{i}"][oi]
            fe[f"block{i}"] = x[i][oi]
        fe["global_map"] = xg[torch.from_numpy(me_ops.canonical_order(cg))]            # decoder output (what GeM pools)
        fe["global_head_map"] = xg_head[torch.from_numpy(me_ops.canonical_order(cg))]  # MinkHead output (engine tap 3)
        fe["local_map"] = xl[ot]
        out["features"] = fe
    return out


@torch.no_grad()
def forward_minkloc(sd: Dict[str, torch.Tensor], coords, feats: torch.Tensor, num_top_down: int = 1,
                    pool_method: str = "GeM", pool_key: str = "pooling.p", acc64: bool = False):
    """MinkLoc.forward / MinkLoc3D.forward (models/minkloc.py:44-61, third_party/minkloc3d/minkloc.py:19-31) with the
    MinkFPN backbone (models/minkfpn.py:65-93).  Returns {'global': (B,C), 'map': (coords, F) of the FPN output}."""
    c = coords.detach().cpu().numpy() if isinstance(coords, torch.Tensor) else np.asarray(coords)
    cm = me_ops.CoordinateManager(c)
    nb = cm.n_batches
    b = "backbone"
    n_levels = 0
    while f"{b}.convs.{n_levels}.kernel" in sd:
        n_levels += 1
    k0 = sd[f"{b}.conv0.kernel"]
    x, _ = me_ops.convolution(cm, feats.float(), 1, k0, int(round(k0.shape[0] ** (1 / 3))), acc64=acc64)
    x = torch.relu(_bn(sd, f"{b}.bn0", x))
    maps = []
    if num_top_down == n_levels:
        maps.append((1, x))
    stride = 1
    for ndx in range(n_levels):
        x, stride = me_ops.convolution(cm, x, stride, sd[f"{b}.convs.{ndx}.kernel"], 2, stride=2, acc64=acc64)
        x = torch.relu(_bn(sd, f"{b}.bn.{ndx}", x))
        j = 0
        while f"{b}.blocks.{ndx}.{j}.conv1.kernel" in sd:               # layers[ndx] blocks per level (models/resnet.py:81-97)
            x = eca_basic_block(sd, f"{b}.blocks.{ndx}.{j}", cm, x, stride, nb, acc64=acc64,
                                eca=f"{b}.blocks.{ndx}.{j}.eca.conv.weight" in sd)
            j += 1
        if n_levels - 1 - num_top_down <= ndx < n_levels - 1:
            maps.append((stride, x))
    x, _ = me_ops.convolution(cm, x, stride, sd[f"{b}.conv1x1.0.kernel"], 1, acc64=acc64)
    for ndx in range(num_top_down):
        x, stride = me_ops.convolution_transpose(cm, x, stride, sd[f"{b}.tconvs.{ndx}.kernel"], acc64=acc64)
        ms, mf = maps[-ndx - 1]
        assert ms == stride
        lat, _ = me_ops.convolution(cm, mf, stride, sd[f"{b}.conv1x1.{ndx + 1}.kernel"], 1, acc64=acc64)
        x = x + lat
    cc = cm.coords(stride)
    if pool_method == "GeM":
        g = gem(x, cc, sd[pool_key], nb)
    elif pool_method == "SPoC":
        g = me_ops.global_avg_pool(x, cc, nb)
    else:
        g = me_ops.global_max_pool(x, cc, nb)
    order = me_ops.canonical_order(cc)
    return {"global": g, "map": (cc[order], x[torch.from_numpy(order)])}

"""Pack a reference ``state_dict`` (key names of SURVEY.md Appendix B) into the single f32 device blob +
``egn_net`` descriptor that ``egn_forward`` consumes.  Eval-mode BatchNorm is folded to scale/shift
(MinkowskiBatchNorm == torch.nn.BatchNorm1d, eps from the module), ``nn.Linear`` weights are
transposed to the (Cin, Cout) layout of the convolution kernels and their bias becomes the epilogue shift."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import lib as L


class _Blob:
    def __init__(self):
        self.chunks: List[torch.Tensor] = []
        self.n = 0

    def add(self, t: torch.Tensor) -> int:
        t = t.detach().to(torch.float32).cpu().contiguous().reshape(-1)
        off = self.n
        pad = (-t.numel()) % 4                      # keep every tensor 16-byte aligned for float4 loads
        self.chunks.append(t)
        if pad:
            self.chunks.append(torch.zeros(pad))
        self.n += t.numel() + pad
        return off

    def tensor(self) -> torch.Tensor:
        return torch.cat(self.chunks) if self.chunks else torch.zeros(4)


def _bn_fold(sd, prefix, eps=1e-5):
    w, b = sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float()
    m, v = sd[prefix + ".running_mean"].float(), sd[prefix + ".running_var"].float()
    scale = w / torch.sqrt(v + eps)
    return scale, b - m * scale


TC_SHAPES = {27: {(32, 32), (32, 64), (64, 64), (64, 128), (128, 128)}, 8: {(32, 32), (64, 64), (128, 128)},
             1: {(32, 64), (64, 32), (64, 64), (64, 128), (128, 64), (128, 128), (128, 256), (256, 256)},
             125: {(1, 32)}}          # conv0 5x5x5 with all-ones features: presence matrix x kernel (conv0_tc.cu)


def _perm32() -> torch.Tensor:
    """K position p = 16*mm + 4*j + e (mm in 0..1, j in 0..3, e in 0..3) of a 32-channel block holds channel 8*j + 4*mm + e:
    the four lanes that feed one row of tcgen05.st.16x256b then read the row's 128-byte line as four contiguous 32-byte
    pieces (one 256-bit load each, one L1 wavefront per line) instead of eight 16-byte pieces in two instructions."""
    p = torch.arange(32)
    mm, j, e = p // 16, (p % 16) // 4, p % 4
    return 8 * j + 4 * mm + e


def pack_tc(kernel: torch.Tensor, perm32: Optional[bool] = None) -> torch.Tensor:
    """(K, Cin, Cout) f32 kernel -> the tensor-core image of ``egn_conv_tc`` as a flat bf16 tensor:
    [ceil(K*Cin/64)][hi|lo][Cout][64] with the 16-byte groups of row n XOR-swizzled by (n & 7) - byte for byte the
    SWIZZLE_128B K-major shared-memory tile that ``tcgen05.mma`` reads, so the kernel fetches a chunk with one
    bulk copy.  hi = bf16(w), lo = bf16(w - hi): w ~ hi + lo to 2^-17 relative."""
    K, cin, cout = kernel.shape
    flat = kernel.detach().to(torch.float32).cpu().reshape(K * cin, cout)
    if perm32 is None:
        perm32 = cin % 32 == 0                                                # gathered feature rows; conv0's offset axis stays natural
    if perm32:
        assert cin % 32 == 0
        flat = flat.reshape(-1, 32, cout)[:, _perm32(), :].reshape(K * cin, cout)
    nch = (K * cin + 63) // 64
    pad = nch * 64 - K * cin
    if pad:
        flat = torch.cat([flat, torch.zeros(pad, cout)], dim=0)
    hi = flat.to(torch.bfloat16)
    lo = (flat - hi.to(torch.float32)).to(torch.bfloat16)
    n = torch.arange(cout)
    g = torch.arange(8)
    src_group = g[None, :] ^ (n[:, None] & 7)                                  # stored group g' holds source group g' ^ (n & 7)
    images = []
    for part in (hi, lo):
        t = part.reshape(nch, 64, cout).permute(0, 2, 1).reshape(nch, cout, 8, 8)   # [chunk][n][group][elem]
        idx = src_group[None, :, :, None].expand(nch, cout, 8, 8)
        images.append(torch.gather(t, 2, idx))
    return torch.stack(images, dim=1).reshape(-1).contiguous()                 # [nch][2][cout][64]


def _conv(blob, kernel, bn=None) -> L.Layer:
    k = kernel if kernel.dim() == 3 else kernel.unsqueeze(0)
    lay = L.Layer(cin=k.shape[1], cout=k.shape[2], w=blob.add(k), scale=-1, shift=-1, wtc=-1)
    if bn is not None:
        lay.scale, lay.shift = blob.add(bn[0]), blob.add(bn[1])
    if (k.shape[1], k.shape[2]) in TC_SHAPES.get(k.shape[0], ()):
        lay.wtc = blob.add(pack_tc(k).view(torch.float32))                     # bf16 pairs stored in f32 slots
    return lay


def _dense(blob, w_io: torch.Tensor, bias: Optional[torch.Tensor], pad_in: int = 0, pad_out: int = 0) -> L.Layer:
    """(in, out) matrix [+ bias] -> row-wise layer, optionally zero-padded to (pad_in, pad_out) so that it has a
    tensor-core instance (the padded hidden channels are exact zeros: relu(0 * x + 0) = 0 feeds zero weight rows)."""
    cin, cout = w_io.shape
    pi, po = max(pad_in, cin), max(pad_out, cout)
    w = torch.zeros((pi, po), dtype=torch.float32)
    w[:cin, :cout] = w_io
    lay = L.Layer(cin=pi, cout=po, w=blob.add(w), scale=-1, shift=-1, wtc=-1)
    if bias is not None:
        b = torch.zeros(po, dtype=torch.float32)
        b[:cout] = bias
        lay.shift = blob.add(b)
    if (pi, po) in TC_SHAPES.get(1, ()):
        lay.wtc = blob.add(pack_tc(w.unsqueeze(0)).view(torch.float32))
    return lay


def _linear(blob, sd, prefix, pad_in: int = 0, pad_out: int = 0) -> L.Layer:
    w = sd[prefix + ".linear.weight"].detach().to(torch.float32).cpu()     # (out, in)
    bias = sd.get(prefix + ".linear.bias")
    if pad_in or pad_out:
        return _dense(blob, w.t(), None if bias is None else bias.detach().to(torch.float32).cpu(), pad_in, pad_out)
    lay = L.Layer(cin=w.shape[1], cout=w.shape[0], w=blob.add(w.t()), scale=-1, shift=-1, wtc=-1)
    if bias is not None:
        lay.shift = blob.add(bias)
    return lay


def _extra_blocks(blob, net, sd, lv: int, prefix: str, bn_eps: float):
    """Blocks 1.. of level ``lv`` (``{prefix}.{j}``, j >= 1): same channels in and out, identity residual."""
    j = 1
    while f"{prefix}.{j}.conv1.kernel" in sd:
        assert j <= L.EGN_MAX_EXTRA_BLOCKS, f"at most {L.EGN_MAX_EXTRA_BLOCKS + 1} blocks per level"
        p = f"{prefix}.{j}"
        assert p + ".downsample.0.kernel" not in sd, "only the first block of a level may change the channel count"
        net.xconv1[lv][j - 1] = _conv(blob, sd[p + ".conv1.kernel"], _bn_fold(sd, p + ".norm1.bn", bn_eps))
        net.xconv2[lv][j - 1] = _conv(blob, sd[p + ".conv2.kernel"], _bn_fold(sd, p + ".norm2.bn", bn_eps))
        if p + ".eca.conv.weight" in sd:
            w = sd[p + ".eca.conv.weight"].reshape(-1)
            net.xeca_k[lv][j - 1], net.xeca_w[lv][j - 1] = w.numel(), blob.add(w)
        j += 1
    net.n_extra[lv] = j - 1


def _head(blob, sd, prefix, levels, out_channels) -> L.Head:
    h = L.Head()
    h.n_levels = len(levels)
    for i, lv in enumerate(sorted(levels)):
        h.levels[i] = lv
    h.out_channels = out_channels
    for lv in levels:
        h.conv1x1[lv] = _conv(blob, sd[f"{prefix}.conv1x1.{lv}.kernel"])
    for lv in range(min(levels) + 1, max(levels) + 1):
        h.tconv[lv] = _conv(blob, sd[f"{prefix}.tconv.{lv}.kernel"])
    return h


def pack_egonn(sd: Dict[str, torch.Tensor], quantizer_desc: dict, global_levels=(5, 6, 7), local_levels=(3, 4),
               ignore_keypoint_regressor: bool = False, bn_eps: float = 1e-5):
    """state_dict of models/minkgl.py:MinkGL (built by models/model_factory.py:31-78) -> (blob cpu f32, Net)."""
    blob = _Blob()
    net = L.Net()
    n_levels = 0
    while f"trunk.convs.{n_levels + 1}.kernel" in sd:
        n_levels += 1
    net.n_levels = n_levels
    k0 = sd["trunk.convs.0.kernel"]
    net.conv0_ksize = int(round(k0.shape[0] ** (1.0 / 3.0)))
    net.conv0 = _conv(blob, k0, _bn_fold(sd, "trunk.bn.0.bn", bn_eps))
    for lv in range(1, n_levels + 1):
        net.down[lv] = _conv(blob, sd[f"trunk.convs.{lv}.kernel"], _bn_fold(sd, f"trunk.bn.{lv}.bn", bn_eps))
        p = f"trunk.blocks.{lv}.0"
        _extra_blocks(blob, net, sd, lv, f"trunk.blocks.{lv}", bn_eps)
        net.conv1[lv] = _conv(blob, sd[p + ".conv1.kernel"], _bn_fold(sd, p + ".norm1.bn", bn_eps))
        net.conv2[lv] = _conv(blob, sd[p + ".conv2.kernel"], _bn_fold(sd, p + ".norm2.bn", bn_eps))
        if p + ".downsample.0.kernel" in sd:
            net.res[lv] = _conv(blob, sd[p + ".downsample.0.kernel"], _bn_fold(sd, p + ".downsample.1.bn", bn_eps))
        if p + ".eca.conv.weight" in sd:
            w = sd[p + ".eca.conv.weight"].reshape(-1)
            net.eca_k[lv], net.eca_w[lv] = w.numel(), blob.add(w)
    gc = sd[f"global_head.conv1x1.{max(global_levels)}.kernel"].shape[-1]
    net.global_head = _head(blob, sd, "global_head", list(global_levels), gc)
    net.global_mlp[0] = _linear(blob, sd, "global_descriptor_decoder.net.0")
    net.global_mlp[1] = _linear(blob, sd, "global_descriptor_decoder.net.2")
    # tensor-core form of the global decoder (models/minkgl.py:207-225, 128 -> 192 -> 256): hidden width padded 192 -> 256
    ghid, gout = net.global_mlp[0].cout, net.global_mlp[1].cout
    if (gc, 256) in TC_SHAPES[1] and ghid <= 256 and gout == 256:
        net.global_mlp[0] = _linear(blob, sd, "global_descriptor_decoder.net.0", pad_out=256)
        net.global_mlp[1] = _linear(blob, sd, "global_descriptor_decoder.net.2", pad_in=256)
    net.pool_method = 0
    net.gem_p = float(sd["global_pooling.pooling.p"].reshape(-1)[0])
    net.gem_eps = 1e-6
    if local_levels and f"local_head.conv1x1.{max(local_levels)}.kernel" in sd:
        lc = sd[f"local_head.conv1x1.{max(local_levels)}.kernel"].shape[-1]
        net.local_head = _head(blob, sd, "local_head", list(local_levels), lc)
        for name, dst in (("local_descriptor_decoder", net.desc_mlp), ("local_keypoint_regressor", net.kp_mlp),
                          ("local_sigma_regressor", net.sigma_mlp)):
            dst[0] = _linear(blob, sd, name + ".net.0")
            dst[1] = _linear(blob, sd, name + ".net.2")
        # tensor-core forms of the per-voxel MLPs (models/minkgl.py:175-225): hidden widths padded to a channel count with
        # a tcgen05 instance (96 -> 128), and the two regressors that read the same map fused into one layer pair
        hid = net.desc_mlp[0].cout
        if (lc, 128) in TC_SHAPES[1] and hid <= 128 and net.desc_mlp[1].cout == 128:
            net.desc_mlp[0] = _linear(blob, sd, "local_descriptor_decoder.net.0", pad_out=128)
            net.desc_mlp[1] = _linear(blob, sd, "local_descriptor_decoder.net.2", pad_in=128)
        kw0, sw0 = sd["local_keypoint_regressor.net.0.linear.weight"], sd["local_sigma_regressor.net.0.linear.weight"]
        kw1, sw1 = sd["local_keypoint_regressor.net.2.linear.weight"], sd["local_sigma_regressor.net.2.linear.weight"]
        hk, hs = kw0.shape[0], sw0.shape[0]
        if kw1.shape[0] == 3 and sw1.shape[0] == 1 and (lc, hk + hs) in TC_SHAPES[1]:
            f32 = lambda t: t.detach().to(torch.float32).cpu()
            w0 = torch.cat([f32(kw0).t(), f32(sw0).t()], dim=1)                                    # (lc, hk + hs)
            b0 = torch.cat([f32(sd["local_keypoint_regressor.net.0.linear.bias"]), f32(sd["local_sigma_regressor.net.0.linear.bias"])])
            w1 = torch.zeros((hk + hs, 4))
            w1[:hk, :3] = f32(kw1).t()
            w1[hk:, 3:] = f32(sw1).t()
            b1 = torch.cat([f32(sd["local_keypoint_regressor.net.2.linear.bias"]), f32(sd["local_sigma_regressor.net.2.linear.bias"])])
            net.kpsig_mlp[0] = _dense(blob, w0, b0)
            net.kpsig_mlp[1] = _dense(blob, w1, b1, pad_out=32 if (hk + hs, 32) in TC_SHAPES[1] else 0)   # 3+1 outputs in a 32-wide tile
    net.polar = 1 if quantizer_desc["coordinates"] == "polar" else 0
    step = quantizer_desc["step"]
    step = list(step) if isinstance(step, (list, tuple)) else [step, step, step]
    for i in range(3):
        net.quant_step[i] = float(step[i])
    net.ignore_keypoint_regressor = int(ignore_keypoint_regressor)
    return blob.tensor(), net


def pack_minkfpn(sd: Dict[str, torch.Tensor], quantizer_desc: dict, num_top_down: int, pool_method: str = "GeM",
                 backbone: str = "backbone", pool_key: Optional[str] = None, bn_eps: float = 1e-5):
    """state_dict of models/minkloc.py:MinkLoc / third_party/minkloc3d/minkloc.py:MinkLoc3D (MinkFPN backbone,
    models/minkfpn.py:9-93, + global pooling) -> (blob cpu f32, Net).  MinkFPN numbers its ModuleLists from 0:
    convs[i]/bn[i]/blocks[i] build level i+1; conv1x1[0] sits on the top level T, conv1x1[j] on level T-j,
    tconvs[j] maps level T-j -> T-j-1 (models/minkfpn.py:48-63,84-91)."""
    blob = _Blob()
    net = L.Net()
    b = backbone
    n_levels = 0
    while f"{b}.convs.{n_levels}.kernel" in sd:
        n_levels += 1
    assert 1 <= n_levels < L.EGN_MAX_LEVELS and 0 <= num_top_down <= n_levels
    net.n_levels = n_levels
    k0 = sd[f"{b}.conv0.kernel"]
    net.conv0_ksize = int(round(k0.shape[0] ** (1.0 / 3.0)))
    net.conv0 = _conv(blob, k0, _bn_fold(sd, f"{b}.bn0.bn", bn_eps))
    for lv in range(1, n_levels + 1):
        i = lv - 1
        net.down[lv] = _conv(blob, sd[f"{b}.convs.{i}.kernel"], _bn_fold(sd, f"{b}.bn.{i}.bn", bn_eps))
        p = f"{b}.blocks.{i}.0"
        _extra_blocks(blob, net, sd, lv, f"{b}.blocks.{i}", bn_eps)
        net.conv1[lv] = _conv(blob, sd[p + ".conv1.kernel"], _bn_fold(sd, p + ".norm1.bn", bn_eps))
        net.conv2[lv] = _conv(blob, sd[p + ".conv2.kernel"], _bn_fold(sd, p + ".norm2.bn", bn_eps))
        if p + ".downsample.0.kernel" in sd:
            net.res[lv] = _conv(blob, sd[p + ".downsample.0.kernel"], _bn_fold(sd, p + ".downsample.1.bn", bn_eps))
        if p + ".eca.conv.weight" in sd:
            w = sd[p + ".eca.conv.weight"].reshape(-1)
            net.eca_k[lv], net.eca_w[lv] = w.numel(), blob.add(w)
    top = n_levels
    levels = list(range(top - num_top_down, top + 1))
    h = L.Head()
    h.n_levels = len(levels)
    for i, lv in enumerate(levels):
        h.levels[i] = lv
    h.out_channels = sd[f"{b}.conv1x1.0.kernel"].shape[-1]
    for j in range(num_top_down + 1):
        h.conv1x1[top - j] = _conv(blob, sd[f"{b}.conv1x1.{j}.kernel"])
    for j in range(num_top_down):
        h.tconv[top - j] = _conv(blob, sd[f"{b}.tconvs.{j}.kernel"])
    net.global_head = h
    net.pool_method = {"GeM": 0, "SPoC": 1, "MAC": 2}[pool_method]
    net.gem_p, net.gem_eps = 3.0, 1e-6
    if pool_method == "GeM":
        net.gem_p = float(sd[pool_key or "pooling.pooling.p"].reshape(-1)[0])
    net.polar = 1 if quantizer_desc["coordinates"] == "polar" else 0
    step = quantizer_desc["step"]
    step = list(step) if isinstance(step, (list, tuple)) else [step, step, step]
    for i in range(3):
        net.quant_step[i] = float(step[i])
    return blob.tensor(), net

"""Host logic of egonn_b200/weights.py on the CPU: the tensor-core weight image, the padded / fused per-voxel MLP layers and the
BatchNorm fold are checked against plain torch arithmetic (no GPU, no C ABI compute)."""
import numpy as np
import torch

from egonn_b200 import weights as W
from egonn_b200 import lib as L


def _unpack_tc(img: torch.Tensor, K: int, cin: int, cout: int):
    """inverse of pack_tc: flat bf16 [chunk][hi|lo][cout][64] with 16-byte groups XOR-swizzled by (n & 7) -> (hi, lo) (K*cin, cout)."""
    nch = (K * cin + 63) // 64
    t = img.view(nch, 2, cout, 8, 8)
    n = torch.arange(cout)
    g = torch.arange(8)
    src_group = g[None, :] ^ (n[:, None] & 7)                       # stored group g' holds source group g' ^ (n & 7)
    out = torch.empty_like(t)
    idx = src_group[None, None, :, :, None].expand(nch, 2, cout, 8, 8)
    out.scatter_(3, idx, t)                                         # undo the gather
    flat = out.reshape(nch, 2, cout, 64).permute(1, 0, 3, 2).reshape(2, nch * 64, cout)[:, : K * cin]
    return flat[0].float(), flat[1].float()


def test_pack_tc_is_a_swizzled_hi_lo_split():
    torch.manual_seed(0)
    for K, cin, cout in ((27, 32, 32), (8, 64, 64), (1, 64, 128), (125, 1, 32)):
        w = torch.randn(K, cin, cout) * 0.1
        img = W.pack_tc(w)
        assert img.dtype == torch.bfloat16 and img.numel() == ((K * cin + 63) // 64) * 2 * cout * 64
        hi, lo = _unpack_tc(img, K, cin, cout)
        flat = w.reshape(K * cin, cout)
        if cin % 32 == 0:      # gathered feature rows: K position 16*mm + 4*j + e of every 32-channel block holds channel 8*j + 4*mm + e
            p = torch.arange(32)
            perm = 8 * ((p % 16) // 4) + 4 * (p // 16) + p % 4
            assert sorted(perm.tolist()) == list(range(32))
            flat = flat.reshape(-1, 32, cout)[:, perm, :].reshape(K * cin, cout)
        assert torch.equal(hi, flat.to(torch.bfloat16).float())                       # hi = bf16(w)
        assert torch.equal(lo, (flat - hi).to(torch.bfloat16).float())                # lo = bf16(w - hi)
        assert float(((hi + lo) - flat).abs().max() / flat.abs().max()) < 2.0 ** -16  # w ~ hi + lo


def _blob_layer(blob_t, lay):
    w = blob_t[lay.w: lay.w + lay.cin * lay.cout].reshape(lay.cin, lay.cout)
    b = blob_t[lay.shift: lay.shift + lay.cout] if lay.shift >= 0 else torch.zeros(lay.cout)
    return w, b


def test_padded_and_fused_mlp_layers_compute_the_reference_mlps(weights):
    """desc_mlp (64 -> 96 -> 128 padded to 64 -> 128 -> 128) and the fused keypoint/sigma regressor (64 -> 32+32 -> 3+1) give the
    same numbers as the reference's separate nn.Linear stacks (models/minkgl.py:175-225)."""
    blob, net = W.pack_egonn(weights, {"coordinates": "cartesian", "step": 0.3})
    torch.manual_seed(1)
    x = torch.randn(50, 64)

    def ref_mlp(prefix):
        w0, b0 = weights[prefix + ".net.0.linear.weight"], weights[prefix + ".net.0.linear.bias"]
        w1, b1 = weights[prefix + ".net.2.linear.weight"], weights[prefix + ".net.2.linear.bias"]
        return torch.relu(x @ w0.t() + b0) @ w1.t() + b1

    w0, b0 = _blob_layer(blob, net.desc_mlp[0])
    w1, b1 = _blob_layer(blob, net.desc_mlp[1])
    assert (net.desc_mlp[0].cin, net.desc_mlp[0].cout, net.desc_mlp[1].cin, net.desc_mlp[1].cout) == (64, 128, 128, 128)
    assert net.desc_mlp[0].wtc >= 0 and net.desc_mlp[1].wtc >= 0                        # both have a tensor-core image
    got = torch.relu(x @ w0 + b0) @ w1 + b1
    torch.testing.assert_close(got, ref_mlp("local_descriptor_decoder"), rtol=1e-5, atol=1e-6)
    assert torch.all((torch.relu(x @ w0 + b0))[:, 96:] == 0)                            # the padded hidden channels are exact zeros

    w0, b0 = _blob_layer(blob, net.kpsig_mlp[0])
    w1, b1 = _blob_layer(blob, net.kpsig_mlp[1])
    # the 3 + 1 outputs sit in a 32-wide (tensor-core) tile; columns 4.. are exact zeros
    assert (net.kpsig_mlp[0].cin, net.kpsig_mlp[0].cout, net.kpsig_mlp[1].cin, net.kpsig_mlp[1].cout) == (64, 64, 64, 32)
    assert net.kpsig_mlp[0].wtc >= 0 and net.kpsig_mlp[1].wtc >= 0
    got = torch.relu(x @ w0 + b0) @ w1 + b1
    torch.testing.assert_close(got[:, :3], ref_mlp("local_keypoint_regressor"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got[:, 3:4], ref_mlp("local_sigma_regressor"), rtol=1e-5, atol=1e-6)
    assert torch.all(got[:, 4:] == 0)

    # global decoder 128 -> 192 -> 256 padded to 128 -> 256 -> 256 (models/minkgl.py:207-225), both layers on tensor cores
    g0, g1 = net.global_mlp[0], net.global_mlp[1]
    assert (g0.cin, g0.cout, g1.cin, g1.cout) == (128, 256, 256, 256) and g0.wtc >= 0 and g1.wtc >= 0
    xg = torch.randn(40, 128)
    w0, b0 = _blob_layer(blob, g0)
    w1, b1 = _blob_layer(blob, g1)
    p = "global_descriptor_decoder"
    ref = torch.relu(xg @ weights[p + ".net.0.linear.weight"].t() + weights[p + ".net.0.linear.bias"]) @ \
        weights[p + ".net.2.linear.weight"].t() + weights[p + ".net.2.linear.bias"]
    torch.testing.assert_close(torch.relu(xg @ w0 + b0) @ w1 + b1, ref, rtol=1e-5, atol=1e-5)
    assert torch.all(torch.relu(xg @ w0 + b0)[:, 192:] == 0)


def test_batchnorm_fold_matches_torch(weights):
    blob, net = W.pack_egonn(weights, {"coordinates": "cartesian", "step": 0.3})
    lay = net.conv1[3]
    scale = blob[lay.scale: lay.scale + lay.cout]
    shift = blob[lay.shift: lay.shift + lay.cout]
    bn = torch.nn.BatchNorm1d(lay.cout, eps=1e-5).eval()
    p = "trunk.blocks.3.0.norm1.bn"
    bn.load_state_dict({k: weights[f"{p}.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")})
    x = torch.randn(20, lay.cout)
    torch.testing.assert_close(x * scale + shift, bn(x), rtol=1e-5, atol=1e-6)
    assert net.conv0.wtc >= 0 and net.conv0_ksize == 5                                  # conv0 carries its tensor-core image
    assert all(net.conv1[lv].wtc >= 0 and net.conv2[lv].wtc >= 0 and net.down[lv].wtc >= 0 for lv in range(1, 8))

"""pin the oracle's sparse convolutions against torch's DENSE conv3d / conv_transpose3d
on a densified grid - an implementation that shares no code with the sparse path - and check the
invariances the domain offers."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import egonn_oracle, me_ops
from conftest import GOLDEN_CASES, load_golden


def _random_sparse(seed, n=400, extent=12, batches=2, lo=-6):
    rng = np.random.default_rng(seed)
    c = np.concatenate([rng.integers(0, batches, (n, 1)), rng.integers(lo, lo + extent, (n, 3))], axis=1)
    c = np.unique(c, axis=0).astype(np.int32)
    return c[rng.permutation(c.shape[0])]


def _densify(coords, feats, lo, size, batches, stride=1):
    """(N,4),(N,C) -> dense (B,C,Z,Y,X) with voxel index (c - lo) / stride."""
    d = torch.zeros((batches, feats.shape[1], size, size, size), dtype=feats.dtype)
    i = (coords[:, 1:].astype(np.int64) - lo) // stride
    d[coords[:, 0], :, i[:, 2], i[:, 1], i[:, 0]] = feats
    return d


def _dense_weight(kernel, K):
    """(K^3,Cin,Cout), k = kx + K(ky + K kz)  ->  conv3d weight (Cout,Cin,kz,ky,kx)."""
    cin, cout = kernel.shape[1:]
    return kernel.reshape(K, K, K, cin, cout).permute(4, 3, 0, 1, 2).contiguous()


@pytest.mark.parametrize("K,cin,cout", [(3, 5, 7), (5, 1, 4), (3, 8, 8)])
def test_stride1_conv_vs_dense(K, cin, cout):
    coords = _random_sparse(K * 10 + cin)
    torch.manual_seed(0)
    feats = torch.randn(coords.shape[0], cin, dtype=torch.float64)
    kernel = torch.randn(K ** 3, cin, cout, dtype=torch.float64)
    cm = me_ops.CoordinateManager(coords)
    out, s = me_ops.convolution(cm, feats, 1, kernel, K)
    assert s == 1
    lo, size = -6 - K, 12 + 2 * K
    dense = F.conv3d(_densify(coords, feats, lo, size, 2), _dense_weight(kernel, K), padding=K // 2)
    i = coords[:, 1:].astype(np.int64) - lo
    ref = dense[coords[:, 0], :, i[:, 2], i[:, 1], i[:, 0]]
    torch.testing.assert_close(out, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("in_stride", [1, 2, 4])
def test_stride2_conv_and_transpose_vs_dense(in_stride):
    """2x2x2 stride-2 convolution (non-centred even kernel) and its transposed partner
     against dense conv3d(stride=2) / conv_transpose3d(stride=2), negatives included."""
    base = _random_sparse(77 + in_stride, n=500, extent=16, lo=-8)
    cm = me_ops.CoordinateManager(base)
    s = 1
    while s < in_stride:                       # build the pyramid up to in_stride the way the trunk does
        s = cm.stride_map(s)
    coords = cm.coords(in_stride)
    torch.manual_seed(1)
    cin, cout = 6, 5
    feats = torch.randn(coords.shape[0], cin, dtype=torch.float64)
    kernel = torch.randn(8, cin, cout, dtype=torch.float64)
    out, so = me_ops.convolution(cm, feats, in_stride, kernel, 2, stride=2)
    assert so == 2 * in_stride
    oc = cm.coords(so)
    # expected output coordinates = unique(floor(c / s') * s')
    exp = coords.astype(np.int64).copy()
    exp[:, 1:] = np.floor(exp[:, 1:] / so).astype(np.int64) * so
    assert set(map(tuple, np.unique(exp, axis=0))) == set(map(tuple, oc.astype(np.int64)))
    lo = -8 * in_stride if in_stride > 1 else -8
    lo = (lo // so) * so - so
    size = (16 + 8) * max(1, 1) + 8
    size = ((8 - lo) // in_stride + 4) // 2 * 2
    dense_in = _densify(coords, feats, lo, size, 2, stride=in_stride)
    dense = F.conv3d(dense_in, _dense_weight(kernel, 2), stride=2)
    j = (oc[:, 1:].astype(np.int64) - lo) // so
    ref = dense[oc[:, 0], :, j[:, 2], j[:, 1], j[:, 0]]
    torch.testing.assert_close(out, ref, rtol=1e-10, atol=1e-10)

    # transposed: coarse features back onto the existing fine map
    g = torch.randn(oc.shape[0], cout, dtype=torch.float64)
    tk = torch.randn(8, cout, cin, dtype=torch.float64)
    up, su = me_ops.convolution_transpose(cm, g, so, tk)
    assert su == in_stride and up.shape == (coords.shape[0], cin)
    dense_c = _densify(oc, g, lo, size // 2, 2, stride=so)
    # conv_transpose3d weight is (Cin,Cout,kz,ky,kx): out[2j+k] += in[j] @ w[:, :, k]
    wt = tk.reshape(2, 2, 2, cout, cin).permute(3, 4, 0, 1, 2).contiguous()
    dense_up = F.conv_transpose3d(dense_c, wt, stride=2)
    i = (coords[:, 1:].astype(np.int64) - lo) // in_stride
    ref_up = dense_up[coords[:, 0], :, i[:, 2], i[:, 1], i[:, 0]]
    torch.testing.assert_close(up, ref_up, rtol=1e-10, atol=1e-10)


def test_sparse_quantize_first_wins_and_f32_divide():
    pc = torch.tensor([[0.31, 0.0, 0.0], [0.05, 0.0, 0.0], [0.59, 0.0, 0.0], [-0.01, 0.0, 0.0], [0.29, 0.1, 0.2]])
    c, ndx = me_ops.sparse_quantize(pc, quantization_size=0.3)
    assert ndx.tolist() == [0, 1, 3]                       # first occurrence wins, input order kept
    assert c.tolist() == [[1, 0, 0], [0, 0, 0], [-1, 0, 0]]
    assert c.dtype == torch.int32 and ndx.dtype == torch.int64
    # float32 divide (not reciprocal multiply / float64): see SURVEY A.1
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.uniform(-80, 80, (200000, 3)).astype(np.float32))
    q = 0.1
    c32 = torch.floor(x / q).int()
    c, ndx = me_ops.sparse_quantize(x, quantization_size=q)
    assert torch.equal(c, c32[ndx])


def test_forward_invariances(weights):
    """Permutation of input rows, batch independence, translation by multiples of 128 voxels."""
    g = load_golden("mini3_cartesian")
    quant = GOLDEN_CASES["mini3_cartesian"]
    coords = g["coords"]
    f = torch.ones((coords.shape[0], 1))
    base = egonn_oracle.forward(weights, coords, f, quant)
    perm = np.random.default_rng(3).permutation(coords.shape[0])
    p = egonn_oracle.forward(weights, coords[perm], f, quant)
    assert np.array_equal(p["coords_L3"], base["coords_L3"])
    torch.testing.assert_close(p["global"], base["global"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(p["descriptors"], base["descriptors"], rtol=1e-4, atol=1e-5)
    # cloud 1 alone == cloud 1 inside the batch
    one = coords[coords[:, 0] == 1].copy()
    one[:, 0] = 0
    s = egonn_oracle.forward(weights, one, torch.ones((one.shape[0], 1)), quant)
    torch.testing.assert_close(s["global"][0], base["global"][1], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(s["sigma_list"][0], base["sigma_list"][1], rtol=1e-4, atol=1e-5)
    # shift by (128, -256, 384) voxels: identical features, shifted coordinates
    sh = coords.copy()
    sh[:, 1:] += np.array([128, -256, 384], dtype=np.int32)
    t = egonn_oracle.forward(weights, sh, f, quant)
    assert np.array_equal(t["coords_L3"][:, 1:] - np.array([128, -256, 384]), base["coords_L3"][:, 1:])
    torch.testing.assert_close(t["global"], base["global"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(t["descriptors"], base["descriptors"], rtol=1e-5, atol=1e-6)

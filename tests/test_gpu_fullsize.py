"""-m gpu parity at BASELINE.json's FULL BATCH sizes (cfg2 B=16, cfg3 B=64, cfg4 B=256, cfg5 1M points): size-independent
properties of the domain over the WHOLE batch.  The element-wise diff against the CPU oracle at these sizes (1-3 s of
oracle time per cloud, ~1.5 min for cfg5) is tests/test_gpu_oracle_fullsize.py; the properties here cover all clouds of
the batch, which the oracle (16-256 clouds x seconds) would make slow.

  * coordinate maps   : level-L map == unique(floor(c / 2^L) * 2^L) computed independently with torch ops on the
                        device (bit-exact, keyed by coordinate), batch offsets partition the rows
  * kernel maps       : the 27-neighbour table == a binary search of (o + delta_k) in the sorted level keys done
                        with torch (bit-exact, every entry), plus the symmetry nbr[nbr[i,k], 26-k] == i
  * quantisation      : de-duplicated voxels == torch.unique of floor(pc / q) (bit-exact), first-occurrence-wins
  * forward           : batch independence (cloud j alone == cloud j inside the batch), invariance to a shuffle of
                        the input rows and to duplicated points, translation by multiples of 128 voxels,
                        tensor-core path == FP32 CUDA-core path, per-cloud top-k == torch.topk
  * single convolution: linearity conv(a x + b y) == a conv(x) + b conv(y)
Tolerances: integers / indices bit-exact; floating point max|a-b|/max|b| <= 1e-3 (north star), tighter where the
two sides run the same arithmetic."""
import numpy as np
import pytest
import torch

from gpu_common import assert_close_rel

pytestmark = pytest.mark.gpu

FULL = {"cfg2": 16, "cfg3": 64, "cfg4": 256, "cfg5": 1}


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda", 0)


_cache = {}


def _batch(cfg, cuda):
    """(list of per-cloud voxel coords on the device, batched coords, voxel size) at the config's full batch size."""
    if cfg not in _cache:
        import egonn_b200 as E
        from egonn_b200 import synth
        _cache.clear()                                   # one full-size workload resident at a time
        voxel = synth.CONFIGS[cfg]["voxel"]
        q = E.CartesianQuantizer(voxel)
        clouds = synth.make_batch(cfg, batch=FULL[cfg])
        pts = [torch.from_numpy(pc).to(cuda) for pc in clouds]
        coords = [q(p)[0] for p in pts]
        _cache[cfg] = (pts, coords, E.batched_coordinates(coords).contiguous(), voxel)
    return _cache[cfg]


def _model(weights, voxel, cuda):
    import egonn_b200 as E
    mp = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=voxel)
    m = E.model_factory(mp)
    m.load_state_dict(weights)
    return m.eval().to(cuda)


def _lexkey(c: torch.Tensor) -> torch.Tensor:
    """(n,4) int32 [b,x,y,z] -> int64 key whose order is lexicographic (b,x,y,z); coordinates in [-2^17, 2^17)."""
    c = c.long()
    bias = 1 << 17
    return (((c[:, 0] << 18 | (c[:, 1] + bias)) << 18 | (c[:, 2] + bias)) << 18) | (c[:, 3] + bias)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", list(FULL))
def test_quantise_full_size_vs_torch_unique(cfg, cuda):
    """ME.utils.sparse_quantize semantics at full size: voxel set == unique(floor(pc / q)), the returned index names
    the FIRST point of every voxel, output keeps input order (datasets/quantization.py:79-85)."""
    import egonn_b200 as E
    pts, coords, _, voxel = _batch(cfg, cuda)
    q = E.CartesianQuantizer(voxel)
    for p in pts[:4]:
        c, ndx = q(p)
        d = torch.floor(p / torch.tensor(voxel, device=cuda)).int()        # tensor divisor: a true IEEE f32 divide (tensor / python
        #                                                                    scalar is a multiply by the reciprocal on CUDA)
        assert torch.equal(d[ndx], c)                                      # the index points at a member of the voxel
        assert torch.all(ndx[1:] > ndx[:-1])                               # input order kept
        key = _lexkey(torch.cat([torch.zeros_like(d[:, :1]), d], 1))
        uk, inv = torch.unique(key, return_inverse=True)
        assert uk.numel() == c.shape[0]                                    # same number of voxels
        first = torch.full((uk.numel(),), p.shape[0], dtype=torch.long, device=cuda)
        first.scatter_reduce_(0, inv, torch.arange(p.shape[0], device=cuda), reduce="amin")
        assert torch.equal(torch.sort(first).values, ndx)                  # first occurrence wins


@pytest.mark.parametrize("cfg", list(FULL))
def test_pyramid_and_kernel_maps_full_size(cfg, cuda):
    import egonn_b200 as E
    _, _, bc, _ = _batch(cfg, cuda)
    eng = E.Engine(cuda)
    info = eng.build(bc)
    assert info.n_batches == FULL[cfg] and info.n_rows[0] == bc.shape[0]
    rows = eng.input_rows().long()
    assert torch.equal(bc[rows], eng.level_coords(0))                      # canonical row r came from input row rows[r]
    for L in range(0, 8):
        cl = eng.level_coords(L)
        s = 1 << L
        exp = torch.unique(_lexkey(torch.cat([bc[:, :1], torch.div(bc[:, 1:], s, rounding_mode="floor") * s], 1)))
        got = _lexkey(cl)
        assert got.numel() == exp.numel() and torch.equal(torch.sort(got).values, exp), f"level {L} coordinate set"
        off = eng.batch_offsets(L).long()
        assert off[0] == 0 and off[-1] == cl.shape[0] and torch.all(off[1:] >= off[:-1])
        b = cl[:, 0].long()
        assert torch.equal(torch.searchsorted(b, torch.arange(info.n_batches + 1, device=cuda)), off)
        if L == 0:
            continue
        # 27-neighbour table vs an independent binary search (k = kx + 3 ky + 9 kz, offsets in units of the stride)
        nbr = eng.neighbors(L).long()
        sk, order = torch.sort(got)
        n = cl.shape[0]
        for k in range(27):
            d = torch.tensor([0, (k % 3 - 1) * s, ((k // 3) % 3 - 1) * s, (k // 9 - 1) * s], dtype=torch.int32, device=cuda)
            qk = _lexkey(cl + d)
            pos = torch.searchsorted(sk, qk).clamp_max(n - 1)
            hit = sk[pos] == qk
            exp_k = torch.where(hit, order[pos], torch.full_like(pos, -1))
            assert torch.equal(nbr[:, k], exp_k), f"level {L} offset {k}"
        assert torch.equal(nbr[:, 13], torch.arange(n, device=cuda))
        i, k = torch.nonzero(nbr >= 0, as_tuple=True)
        assert torch.equal(nbr[nbr[i, k], 26 - k], i), f"level {L} symmetry"


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_forward_invariances_full_size(cfg, cuda, weights):
    import egonn_b200 as E
    pts, coords, bc, voxel = _batch(cfg, cuda)
    model = _model(weights, voxel, cuda)
    ones = lambda n: torch.ones((n, 1), device=cuda)
    full = model.forward_packed({"coords": bc, "features": ones(bc.shape[0])})
    off = full["local_offsets"].long()
    for k in ("global", "descriptors", "keypoints", "sigma"):
        assert torch.isfinite(full[k]).all(), k
    np.testing.assert_allclose(full["descriptors"].norm(dim=1).cpu().numpy(), 1.0, rtol=1e-5)      # F.normalize
    assert (full["sigma"] > 0).all()                                                               # softplus
    # keypoints stay inside their supervoxel: centre +- stride*q/2 (tanh offset), datasets/quantization.py:93-103
    centre = (full["local_coords"][:, 1:].float() + 0.5) * voxel
    assert ((full["keypoints"] - centre).abs() <= 4 * voxel * (1 + 1e-5) + 1e-4).all()

    # (1) batch independence: cloud j alone == cloud j inside the batch (not bit-identical: the slice partition of the
    #     per-cloud pooling sums, hence their fp32 summation order, depends on the batch composition -> 5e-5)
    for j in (0, FULL[cfg] - 1):
        cj = E.batched_coordinates([coords[j]])
        single = model.forward_packed({"coords": cj, "features": ones(cj.shape[0])})
        assert torch.equal(single["local_coords"][:, 1:], full["local_coords"][off[j]:off[j + 1], 1:])
        assert_close_rel(single["global"][0], full["global"][j], 5e-5, f"global of cloud {j} alone")
        for k in ("descriptors", "keypoints", "sigma"):
            assert_close_rel(single[k], full[k][off[j]:off[j + 1]], 5e-5, f"{k} of cloud {j} alone")

    # (2) shuffled input rows + duplicated rows: canonical order makes the result identical
    g = torch.Generator(device="cpu").manual_seed(3)
    perm = torch.randperm(bc.shape[0], generator=g).to(cuda)
    dup = torch.cat([bc[perm], bc[perm[:1000]]], 0)
    sh = model.forward_packed({"coords": dup, "features": ones(dup.shape[0])})
    assert torch.equal(sh["local_coords"], full["local_coords"])
    for k in ("global", "descriptors", "keypoints", "sigma"):
        assert_close_rel(sh[k], full[k], 1e-6, f"{k} under a row shuffle")

    # (3) translation by multiples of 128 voxels: features identical, keypoints shifted
    t = torch.tensor([0, 256, -128, 128], dtype=torch.int32, device=cuda)
    tr = model.forward_packed({"coords": bc + t, "features": ones(bc.shape[0])})
    # (the canonical Morton row order changes under translation: compare keyed by coordinate)
    o1, o0 = torch.argsort(_lexkey(tr["local_coords"] - t)), torch.argsort(_lexkey(full["local_coords"]))
    assert torch.equal(tr["local_coords"][o1], full["local_coords"][o0] + t)
    assert_close_rel(tr["global"], full["global"], 5e-5, "global under translation")
    assert_close_rel(tr["descriptors"][o1], full["descriptors"][o0], 5e-5, "descriptors under translation")
    assert_close_rel(tr["sigma"][o1], full["sigma"][o0], 5e-5, "sigma under translation")
    assert_close_rel(tr["keypoints"][o1] - t[1:].float() * voxel, full["keypoints"][o0], 1e-4, "keypoints under translation")

    # (4) tensor-core (bf16x3 split) path == FP32 CUDA-core path
    model._engine.set_tensor_cores(False)
    f32 = model.forward_packed({"coords": bc, "features": ones(bc.shape[0])})
    model._engine.set_tensor_cores(True)
    for k in ("global", "descriptors", "keypoints", "sigma"):
        assert_close_rel(full[k], f32[k], 1e-4, f"{k}: tensor cores vs FP32 path")

    # (5) keypoint selection == torch.topk(sigma, k, largest=False) per cloud (eval/evaluate.py:352-361)
    idx = E.topk_smallest(full["sigma"], full["local_offsets"], 256).long()
    for j in (0, FULL[cfg] // 2, FULL[cfg] - 1):
        seg = full["sigma"][off[j]:off[j + 1], 0]
        kk = min(256, seg.numel())
        exp = torch.sort(seg, stable=True).indices[:kk]
        assert torch.equal(idx[j, :kk], exp)

    # (6) fused raw-point ingest == staged quantise -> batch -> forward
    starts = torch.tensor(np.cumsum([0] + [p.shape[0] for p in pts]), dtype=torch.int32, device=cuda)
    fp = model.forward_points(torch.cat(pts, 0), starts)
    assert torch.equal(fp["local_coords"], full["local_coords"])
    for k in ("global", "descriptors", "keypoints", "sigma"):
        assert_close_rel(fp[k], full[k], 1e-6, f"{k}: fused ingest")


def test_single_cloud_1m_points_forward(cuda, weights):
    """cfg5: one dense 1M-point map tile.  TC path vs FP32 path, and a row shuffle."""
    pts, coords, bc, voxel = _batch("cfg5", cuda)
    model = _model(weights, voxel, cuda)
    feats = torch.ones((bc.shape[0], 1), device=cuda)
    a = model.forward_packed({"coords": bc, "features": feats})
    model._engine.set_tensor_cores(False)
    b = model.forward_packed({"coords": bc, "features": feats})
    model._engine.set_tensor_cores(True)
    for k in ("global", "descriptors", "keypoints", "sigma"):
        assert torch.isfinite(a[k]).all()
        assert_close_rel(a[k], b[k], 1e-4, f"{k}: tensor cores vs FP32 path")
    perm = torch.randperm(bc.shape[0], generator=torch.Generator().manual_seed(1)).to(cuda)
    c = model.forward_packed({"coords": bc[perm], "features": feats})
    assert torch.equal(c["local_coords"], a["local_coords"])
    assert_close_rel(c["global"], a["global"], 1e-6, "global under a row shuffle")


@pytest.mark.parametrize("ksize,cin,cout,level", [(3, 32, 32, 1), (3, 64, 128, 4), (2, 64, 64, 2), (3, 128, 128, 5)])
def test_conv_linearity_full_size(ksize, cin, cout, level, cuda):
    import egonn_b200 as E
    _, _, bc, _ = _batch("cfg2", cuda)
    eng = E.Engine(cuda)
    info = eng.build(bc)
    torch.manual_seed(level)
    n = info.n_rows[level]
    x, y = torch.randn(n, cin, device=cuda), torch.randn(n, cin, device=cuda)
    w = torch.randn(ksize ** 3, cin, cout, device=cuda) / np.sqrt(cin * 4.0)
    lhs = eng.conv_tc(level, ksize, 0.7 * x - 1.3 * y, w)
    rhs = 0.7 * eng.conv_tc(level, ksize, x, w) - 1.3 * eng.conv_tc(level, ksize, y, w)
    assert_close_rel(lhs, rhs, 5e-5, "linearity")
    assert_close_rel(lhs, eng.conv(level, ksize, False, 0.7 * x - 1.3 * y, w), 5e-5, "tc vs fp32")

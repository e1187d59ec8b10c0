"""Rows f2 / f3 of SURVEY 8f: raw-scan point filter and local-descriptor matching.  CPU part: known-answer checks of the
oracle restatement; GPU part: the device kernels through the C ABI against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import eval_ops


def _records(n=20000, seed=0):
    rng = np.random.default_rng(seed)
    r = rng.normal(0, 20, size=(n, 4)).astype(np.float32)
    r[:, 2] = rng.uniform(-3.0, 4.0, size=n).astype(np.float32)
    r[rng.choice(n, 300, replace=False), :3] = 0.0                    # missing returns
    r[rng.choice(n, 50, replace=False), :3] = np.float32(5e-9)         # inside np.isclose's atol
    r[rng.choice(n, 50, replace=False), 2] = np.float32(-1.5)          # exactly on the ground level: dropped (z > level)
    return r


def test_oracle_filter_known_answers():
    r = np.array([[0, 0, 0, 9], [1, 2, 3, 9], [1, 2, -1.5, 9], [1, 2, -1.4999, 9], [0, 0, 1e-9, 9], [0, 0, 1e-7, 9]], dtype=np.float32)
    out = eval_ops.filter_points(r)
    assert out.tolist() == [[1, 2, 3], [1, 2, np.float32(-1.4999)], [0, 0, np.float32(1e-7)]]
    assert eval_ops.filter_points(r, remove_zero_points=False, remove_ground_plane=False).shape == (6, 3)
    assert eval_ops.filter_points(r, remove_ground_plane=False).shape == (4, 3)
    assert eval_ops.filter_points(r, ground_plane_level=-0.9).shape == (2, 3)     # MulRan level


def test_oracle_match_known_answers():
    a = np.array([[0, 0], [10, 0], [0, 10], [5, 5]], dtype=np.float32)
    b = np.array([[0.1, 0], [9, 0], [9.5, 0.2]], dtype=np.float32)
    idx, dist = eval_ops.match_mutual(a, b)
    # a1 -> b2 (0.54) and b2 -> a1; b1 -> a1 too but a1's nearest is b2, so b1 stays unmatched; a2, a3 -> b0/b2, not mutual
    assert idx.tolist() == [0, 2, -1, -1]
    np.testing.assert_allclose(dist[:2], [0.1, np.hypot(0.5, 0.2)], rtol=1e-6)
    assert eval_ops.match_mutual(a, b, mutual=False)[0].tolist() == [0, 2, 0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [(True, True, -1.5), (True, False, 0.0), (False, True, -0.9), (False, False, 0.0)])
def test_filter_points_matches_oracle(flags):
    import egonn_b200 as E
    dev = torch.device("cuda", 0)
    r = _records()
    for rec in (r, np.ascontiguousarray(r[:, :3]), r[:1], r[:2049]):
        got = E.filter_points(torch.from_numpy(rec).to(dev), *flags).cpu().numpy()
        exp = eval_ops.filter_points(rec, *flags)
        assert got.shape == exp.shape and np.array_equal(got, exp)     # bit-exact, order kept


@pytest.mark.gpu
def test_filter_then_quantise_pipeline_matches_oracle():
    """loader -> quantizer -> batched_coordinates as in eval/evaluate.py:315-333, all on the device."""
    import egonn_b200 as E
    from oracle import egonn_oracle
    dev = torch.device("cuda", 0)
    r = _records(50000, seed=3)
    pts = E.filter_points(torch.from_numpy(r).to(dev), True, True, -1.5)
    c, ndx = E.CartesianQuantizer(0.3)(pts)
    ref_pts = eval_ops.filter_points(r)
    ref_c, ref_ndx = egonn_oracle.quantize(torch.from_numpy(ref_pts), {"coordinates": "cartesian", "step": 0.3})
    assert torch.equal(c.cpu(), ref_c) and torch.equal(ndx.cpu(), ref_ndx)


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb,dim", [(256, 256, 128), (128, 300, 128), (1, 7, 32), (513, 64, 256)])
def test_match_descriptors_matches_oracle(na, nb, dim):
    import egonn_b200 as E
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(na + nb)
    a = rng.normal(size=(na, dim)).astype(np.float32)
    b = rng.normal(size=(nb, dim)).astype(np.float32)
    k = min(na, nb) // 2
    b[:k] = a[:k] + 0.01 * rng.normal(size=(k, dim)).astype(np.float32)   # true correspondences
    if nb > 4:
        b[-1] = b[-2]                                                       # an exact tie: the lower row wins
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d2 = ((a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)) ** 2).sum(-1)
    # plain nearest neighbour, both directions: the oracle's choice, or (fp32 vs fp64 near-tie) one that is as near to 1e-6
    ab, dist = E.match_descriptors(ta, tb, mutual=False)
    ba, _ = E.match_descriptors(tb, ta, mutual=False)
    ab, ba = ab.cpu().numpy(), ba.cpu().numpy()
    ridx, rdist = eval_ops.match_mutual(a, b, mutual=False)
    same = ab == ridx
    assert same.mean() > 0.99
    assert np.all(d2[np.arange(na), ab] <= d2.min(axis=1) * (1 + 1e-6) + 1e-12)
    assert np.all(d2[ba, np.arange(nb)] <= d2.min(axis=0) * (1 + 1e-6) + 1e-12)
    np.testing.assert_allclose(dist.cpu().numpy()[same], rdist[same], rtol=1e-4, atol=1e-6)
    if nb > 4:
        assert not np.any(ab == nb - 1) or not np.allclose(b[-1], b[-2])        # exact tie: the lower row wins
    # mutual filter: exactly the mutual subset of the device's own nearest neighbours, and equal to the oracle's
    m, _ = E.match_descriptors(ta, tb, mutual=True)
    m = m.cpu().numpy()
    assert np.array_equal(m, np.where(ba[ab] == np.arange(na), ab, -1))
    rm, _ = eval_ops.match_mutual(a, b, mutual=True)
    assert (m == rm).mean() > 0.99
    if k:
        assert (m[:k] == np.arange(k)).all() and (rm[:k] == np.arange(k)).all()     # the planted correspondences

"""-m gpu: the pipelined extraction API (`egonn_b200.Extractor`) - the caller loop of eval/evaluate.py:327-350 / :454-466 for a
stream of batches - gives, batch for batch and in submission order, what the direct calls give: `model.forward_points` +
`topk_smallest` + per-cloud selection (eval/evaluate.py:339-361)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda", 0)


def _direct(model, clouds, k, cuda):
    import egonn_b200 as E
    pts = torch.cat([torch.as_tensor(c) for c in clouds]).to(cuda)
    off = torch.tensor(np.cumsum([0] + [c.shape[0] for c in clouds]), dtype=torch.int32, device=cuda)
    p = model.forward_points(pts, off)
    idx = E.topk_smallest(p["sigma"], p["local_offsets"], k).long().cpu()
    lo = p["local_offsets"].cpu().long()
    kp, ds = p["keypoints"].cpu(), p["descriptors"].cpu()
    out_kp = torch.zeros((len(clouds), k, 3))
    out_ds = torch.zeros((len(clouds), k, ds.shape[1]))
    for b in range(len(clouds)):
        sel = idx[b][idx[b] >= 0] + lo[b]
        out_kp[b, : sel.numel()] = kp[sel]
        out_ds[b, : sel.numel()] = ds[sel]
    return p["global"].cpu(), out_kp, out_ds, (lo[1:] - lo[:-1]).clamp(max=k)


def test_extractor_matches_direct_calls_in_order(cuda, weights):
    import egonn_b200 as E
    from egonn_b200 import synth
    mp = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.3)
    model = E.model_factory(mp)
    model.load_state_dict(weights)
    model = model.eval().to(cuda)
    K = 64
    rng = np.random.default_rng(3)
    batches = []
    for i in range(9):                                          # varied batch sizes and cloud sizes, one cloud with < K keypoints
        nb = int(rng.integers(1, 4))
        batches.append([synth.uniform_cloud(int(rng.integers(300, 6000)), 100 * i + j) for j in range(nb)])
    batches[4] = [synth.uniform_cloud(40, 777), synth.uniform_cloud(3000, 778)]
    expect = [_direct(model, b, K, cuda) for b in batches]
    ext = E.Extractor(model, streams=3, topk=K)
    mixed = [E.stage_batch(b) if i % 2 else b for i, b in enumerate(batches)]          # plain lists and pre-staged batches
    got = list(ext.extract(iter(mixed)))
    assert len(got) == len(batches)
    for i, (r, e) in enumerate(zip(got, expect)):
        assert torch.equal(r["global"], e[0]), f"batch {i}: global"
        assert torch.equal(r["n_keypoints"].long(), e[3]), f"batch {i}: keypoint counts"
        assert torch.equal(r["keypoints"], e[1]), f"batch {i}: keypoints"
        assert torch.equal(r["descriptors"], e[2]), f"batch {i}: descriptors"
    # a second pass reuses the contexts and staging slots
    again = list(ext.extract(iter(mixed[:3])))
    assert all(torch.equal(a["global"], e[0]) for a, e in zip(again, expect))
    # errors surface in order, at the failing batch
    bad = [batches[0], [np.full((10, 3), 1e9, dtype=np.float32)], batches[1]]
    it = ext.extract(iter(bad))
    assert torch.equal(next(it)["global"], expect[0][0])
    with pytest.raises(Exception):
        next(it)


def test_extractor_growing_batches_reallocate_staging_safely(cuda, weights):
    """Regression: when a batch needs a larger device staging buffer, the caching allocator may hand the copy stream a block
    that kernels still queued on the compute stream write (outputs of the previous batch, already freed on the host); the
    extractor orders the copy stream behind the compute stream on (re)allocation.  Batches grow so that every slot
    reallocates several times while work is in flight."""
    import egonn_b200 as E
    from egonn_b200 import synth
    mp = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.3)
    model = E.model_factory(mp)
    model.load_state_dict(weights)
    model = model.eval().to(cuda)
    sizes = [2000, 2500, 9000, 12000, 40000, 52000, 160000, 200000, 600000, 700000]
    batches = [[synth.uniform_cloud(n, 10 + i)] for i, n in enumerate(sizes)]
    expect = [_direct(model, b, 32, cuda)[0] for b in batches]
    for rep in range(3):
        ext = E.Extractor(model, streams=2, topk=32)                   # fresh slots: the growth happens again
        got = [r["global"] for r in ext.extract(iter(batches))]
        ext.close()
        for i, (g, e) in enumerate(zip(got, expect)):
            assert torch.equal(g, e), f"repetition {rep}, batch {i}"

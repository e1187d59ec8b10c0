"""Host logic of the training path (second half) without a GPU: ``egonn_b200.autograd``'s backward rules and
the train-mode walk of ``egonn_b200.models`` run on a CPU TEST DOUBLE of the engine (tests/cpu_engine.py, built on the
oracle) and must reproduce the gradients that torch's autograd gives the UNMODIFIED reference graph on the oracle shim
(tests/golden/train_mini3.npz from tests/golden/make_golden_train.py).  The same comparison runs on the CUDA engine in
tests/test_training_step_gpu.py."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, train_fixtures


@pytest.fixture()
def cpu_engine(monkeypatch):
    import egonn_b200.minkowski as ME
    from cpu_engine import CpuEngine
    monkeypatch.setattr(ME, "Engine", CpuEngine)
    return CpuEngine


def _model(weights, step):
    import egonn_b200 as E
    mp = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=step)
    m = E.model_factory(mp)
    m.load_state_dict(weights)
    return m


def test_training_step_matches_reference_autograd(cpu_engine, weights):
    import train_case
    model = _model(weights, train_case.QUANT["step"])
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])
    loss = train_case.step(model, coords)
    for fixture in train_fixtures():                               # + the real-MinkowskiEngine fixture once it exists
        golden = dict(np.load(os.path.join(GOLDEN, fixture)))
        real = fixture.endswith("_me.npz")                         # real ME rounds differently: the GPU test's bars
        r = train_case.compare(model, loss, golden, *((5e-2, 1e-3, 1e-4) if real else (2e-4, 2e-5, 1e-5)), f"CPU test double vs {fixture}")
        print("\n[training step, CPU double vs %s] gradients: worst %s %.2e, median %.2e; forward: worst %s %.2e"
              % (fixture, *r["worst_grad"], r["median_grad"], *r["worst_forward"]))


def _leaf(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).requires_grad_(True)


@pytest.mark.parametrize("ksize,transposed,level,cin,cout", [(3, False, 1, 8, 12), (3, False, 2, 16, 4), (2, False, 0, 4, 8),
                                                           (2, False, 2, 8, 8), (2, True, 3, 8, 4), (2, True, 1, 4, 12),
                                                           (1, False, 2, 12, 8), (5, False, 0, 1, 8)])
def test_conv_backward_rules_against_torch_autograd(cpu_engine, ksize, transposed, level, cin, cout):
    """Each backward rule of SparseConvFunction against autograd through the oracle's gather -> mm -> index_add convolution
    (negative coordinates included: mini3 is centred on the sensor)."""
    from egonn_b200.autograd import SparseConvFunction
    from oracle import me_ops
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])
    eng = cpu_engine()
    info = eng.build(coords)
    n_in = info.n_rows[level]
    x = _leaf(n_in, cin, seed=1)
    shape = (cin, cout) if ksize == 1 else (ksize ** 3, cin, cout)
    w = _leaf(*shape, seed=2)
    y = SparseConvFunction.apply(x, w, eng, level, ksize, transposed)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3))
    if ksize == 5:                                             # the stem: gradient for the kernel only
        x = x.detach()
        y = SparseConvFunction.apply(x, w, eng, level, ksize, transposed)
        (gw,) = torch.autograd.grad(y, [w], gy)
        gx = None
    else:
        gx, gw = torch.autograd.grad(y, [x, w], gy)
    # reference: autograd through the oracle operator in ITS row order
    xr = x.detach()[eng.inv[level]].clone().requires_grad_(ksize != 5)
    wr = w.detach().clone().requires_grad_(True)
    if ksize == 1:
        yr, s = xr @ wr, 1 << level
    elif ksize == 2 and transposed:
        yr, s = me_ops.convolution_transpose(eng.cm, xr, 1 << level, wr)
    else:
        yr, s = me_ops.convolution(eng.cm, xr, 1 << level, wr, ksize, 2 if ksize == 2 else 1)
    lo = int(np.log2(s))
    gyr = gy[eng.inv[lo]]
    assert torch.allclose(y[eng.inv[lo]], yr, atol=1e-4)
    if ksize == 5:
        (gwr,) = torch.autograd.grad(yr, [wr], gyr)
    else:
        gxr, gwr = torch.autograd.grad(yr, [xr, wr], gyr)
        assert float((gx[eng.inv[level]] - gxr).abs().max()) <= 1e-4 * max(1.0, float(gxr.abs().max()))
    assert float((gw - gwr).abs().max()) <= 1e-4 * max(1.0, float(gwr.abs().max()))


@pytest.mark.parametrize("is_max", [False, True])
def test_pool_and_broadcast_backward_rules(cpu_engine, is_max):
    from egonn_b200.autograd import BroadcastMulFunction, GlobalPoolFunction
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])
    eng = cpu_engine()
    info = eng.build(coords)
    level, c = 2, 8
    n, nb = info.n_rows[level], info.n_batches
    bidx = torch.from_numpy(eng.can[level][:, 0].astype(np.int64))
    x = _leaf(n, c, seed=4)
    gy = torch.randn((nb, c), generator=torch.Generator().manual_seed(5))
    (gx,) = torch.autograd.grad(GlobalPoolFunction.apply(x, eng, level, is_max), [x], gy)
    xr = x.detach().clone().requires_grad_(True)
    rows = [xr[bidx == b] for b in range(nb)]
    ref = torch.stack([r.max(dim=0).values if is_max else r.mean(dim=0) for r in rows])
    (gxr,) = torch.autograd.grad(ref, [xr], gy)
    assert torch.allclose(gx, gxr, atol=1e-6)
    g = _leaf(nb, c, seed=6)
    gy2 = torch.randn((n, c), generator=torch.Generator().manual_seed(7))
    gx2, gg2 = torch.autograd.grad(BroadcastMulFunction.apply(x, g, eng, level), [x, g], gy2)
    gr = g.detach().clone().requires_grad_(True)
    gx2r, gg2r = torch.autograd.grad(xr * gr[bidx], [xr, gr], gy2)
    assert torch.allclose(gx2, gx2r, atol=1e-6) and torch.allclose(gg2, gg2r, atol=1e-4)


def test_eval_mode_still_takes_the_fused_path(weights):
    """model.eval() must not reach the layer walk: on a box without CUDA the fused path fails loudly (no CPU fallback)."""
    import egonn_b200 as E
    model = _model(weights, 0.4).eval()
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])
    if torch.cuda.is_available():
        pytest.skip("CPU-box check")
    with pytest.raises(Exception, match="no CPU path|CUDA"):
        model({"coords": coords, "features": torch.ones((coords.shape[0], 1))})
    with pytest.raises(Exception, match="no CPU path|CUDA|no CUDA"):
        model.train()({"coords": coords, "features": torch.ones((coords.shape[0], 1))})     # the real engine refuses CPU too


def test_unmodified_reference_graph_trains_on_the_front_end(reference_on_front_end, weights):
    """INTEGRATION.md path 2 in training mode: the reference's OWN ``models/model_factory.py`` / ``models/minkgl.py`` /
    ``layers/*.py`` (imported unmodified from /root/reference, skipped where it is absent) on ``egonn_b200.minkowski``
    registered as ``MinkowskiEngine`` - engine replaced by the CPU double here - gives the fixture's gradients."""
    import tempfile
    import train_case
    model_factory, ModelParams = reference_on_front_end
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("[MODEL]\nmodel = egonn\ncoordinates = cartesian\nquantization_step = %s\n" % train_case.QUANT["step"])
    model = model_factory(ModelParams(f.name))
    os.unlink(f.name)
    model.load_state_dict(weights)
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])
    loss = train_case.step(model, coords)
    golden = dict(np.load(os.path.join(GOLDEN, "train_mini3.npz")))
    r = train_case.compare(model, loss, golden, 2e-4, 2e-5, 1e-5, "reference graph on the front end (CPU double)")
    print("\n[reference graph, front end + CPU double] gradients: worst %s %.2e" % r["worst_grad"])

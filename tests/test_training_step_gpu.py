"""-m gpu: ONE TRAINING STEP on the CUDA engine (second half; training/trainer.py:141-195):
``model.train()``, two forwards alive before one backward, ``loss.backward()``, ``optimizer.step()``, back to ``eval()``.

Gradients of every parameter, the BatchNorm running statistics and the loss are compared with the fixture written by the
UNMODIFIED reference graph under torch autograd on the CPU shim (tests/golden/train_mini3.npz,
tests/golden/make_golden_train.py).  A second case runs a MinkLoc model (BasicBlock + SPoC pooling) on the engine and on
the CPU test double of the engine (tests/cpu_engine.py, oracle-backed) side by side; a third checks the per-cloud pooling
(mean / max) and broadcast rules on the engine against torch autograd on the same device."""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, train_fixtures

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda", 0)


def _lex(c: torch.Tensor) -> torch.Tensor:
    """Row order that sorts (N,4) coordinates lexicographically: engine and double are compared keyed by coordinate."""
    c = c.cpu().long()
    key = ((c[:, 0] * (1 << 18) + c[:, 1] + (1 << 17)) * (1 << 18) + c[:, 2] + (1 << 17)) * (1 << 18) + c[:, 3] + (1 << 17)
    return torch.argsort(key)


def _field(coords: torch.Tensor, channels: int, seed: int) -> torch.Tensor:
    """A deterministic feature field: row values depend on the row's coordinate only (not on the row order)."""
    g = torch.Generator().manual_seed(seed)
    m = torch.randn((4, channels), generator=g, dtype=torch.float64)
    return torch.sin(coords.cpu().double() @ m * 0.37 + 0.1).float().to(coords.device)


@pytest.mark.parametrize("ksize,transposed,level,cin,cout", [(3, False, 1, 32, 32), (3, False, 2, 32, 64), (3, False, 4, 128, 128),
                                                           (2, False, 0, 32, 32), (2, False, 3, 64, 64), (2, True, 4, 64, 64),
                                                           (2, True, 6, 128, 128), (1, False, 3, 64, 96), (5, False, 0, 1, 32)])
def test_conv_backward_rules_engine_vs_cpu_double(cuda, ksize, transposed, level, cin, cout):
    """Single-operator gradients (well conditioned): SparseConvFunction on the engine == the same rules on the CPU double
    (which tests/test_autograd_cpu.py holds against torch autograd through the oracle), keyed by coordinate."""
    from cpu_engine import CpuEngine
    from egonn_b200.autograd import SparseConvFunction
    from egonn_b200.engine import Engine
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])
    out_level = level if ksize != 2 else (level - 1 if transposed else level + 1)
    shape = (cin, cout) if ksize == 1 else (ksize ** 3, cin, cout)
    w0 = torch.randn(shape, generator=torch.Generator().manual_seed(5)) / (cin * ksize ** 1.5)
    res = []
    for eng, dev in ((Engine(cuda), cuda), (CpuEngine(), torch.device("cpu"))):
        eng.build(coords.to(dev))
        ci, co = eng.level_coords(level), eng.level_coords(out_level)
        x = (torch.ones((ci.shape[0], 1), device=dev) if cin == 1 else _field(ci, cin, 1)).requires_grad_(ksize != 5)
        w = w0.to(dev).requires_grad_(True)
        y = SparseConvFunction.apply(x, w, eng, level, ksize, transposed)
        gy = _field(co, cout, 2)
        if ksize == 5:
            (gw,) = torch.autograd.grad(y, [w], gy)
            gx = torch.zeros_like(x)
        else:
            gx, gw = torch.autograd.grad(y, [x, w], gy)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        oi, oo = _lex(ci), _lex(co)
        assert torch.equal(ci.cpu()[oi], res[0][3]) if res else True
        res.append((y.detach().cpu()[oo], gx.detach().cpu()[oi], gw.detach().cpu(), ci.cpu()[oi]))
    for name, a, b in zip(("y", "dx", "dW"), res[0][:3], res[1][:3]):
        scale = max(float(b.abs().max()), 1e-30)
        err = float((a - b).abs().max()) / scale
        assert err <= 1e-4, f"{name}: engine vs double rel err {err:.3e}"


@pytest.mark.parametrize("is_max", [False, True])
def test_pool_and_broadcast_backward_rules_on_the_engine(cuda, is_max):
    from egonn_b200.autograd import BroadcastMulFunction, GlobalPoolFunction
    from egonn_b200.engine import Engine
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"]).to(cuda)
    eng = Engine(cuda)
    info = eng.build(coords)
    level, c = 2, 64
    n, nb = info.n_rows[level], info.n_batches
    bidx = eng.level_coords(level)[:, 0].long()
    gen = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn((n, c), generator=gen).to(cuda).requires_grad_(True)
    gy = torch.randn((nb, c), generator=gen).to(cuda)
    (gx,) = torch.autograd.grad(GlobalPoolFunction.apply(x, eng, level, is_max), [x], gy)
    xr = x.detach().clone().requires_grad_(True)
    rows = [xr[bidx == b] for b in range(nb)]
    ref = torch.stack([r.max(dim=0).values if is_max else r.mean(dim=0) for r in rows])
    (gxr,) = torch.autograd.grad(ref, [xr], gy)
    assert float((gx - gxr).abs().max()) <= 1e-5 * float(gxr.abs().max())
    g = torch.randn((nb, c), generator=gen).to(cuda).requires_grad_(True)
    gy2 = torch.randn((n, c), generator=gen).to(cuda)
    gx2, gg2 = torch.autograd.grad(BroadcastMulFunction.apply(x, g, eng, level), [x, g], gy2)
    gr = g.detach().clone().requires_grad_(True)
    gx2r, gg2r = torch.autograd.grad(xr * gr[bidx], [xr, gr], gy2)
    assert float((gx2 - gx2r).abs().max()) <= 1e-5 * float(gx2r.abs().max())
    assert float((gg2 - gg2r).abs().max()) <= 1e-4 * float(gg2r.abs().max())


def test_training_step_matches_reference_autograd(cuda, weights):
    import egonn_b200 as E
    import train_case
    mp = E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=train_case.QUANT["step"])
    model = E.model_factory(mp)
    model.load_state_dict(weights)
    model = model.to(cuda)
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"]).to(cuda)
    batch = train_case.batches(coords)[0]

    before = model.eval()(batch)["global"].clone()                  # fused path, checkpoint weights
    loss = train_case.step(model, coords)                           # train(): layer walk + egonn_b200.autograd
    torch.cuda.synchronize()
    for fixture in train_fixtures():                                # + the real-MinkowskiEngine fixture once it exists
        golden = dict(np.load(os.path.join(GOLDEN, fixture)))
        r = train_case.compare(model, loss, golden, 5e-2, 1e-3, 1e-4, f"CUDA engine vs {fixture}")     # bars: see train_case.compare
        print("\n[training step, CUDA engine vs %s] loss %.6f (fixture %.6f); gradients: worst %s %.2e, median %.2e; forward: worst %s %.2e"
              % (fixture, loss, float(golden["loss"]), *r["worst_grad"], r["median_grad"], *r["worst_forward"]))

    # optimizer step, then inference again: the fused path must pick up the new parameters and running statistics
    opt = torch.optim.SGD(model.parameters(), lr=1e-7)          # gradients reach 4e3 with this loss: a small, finite update
    opt.step()
    after = model.eval()(batch)["global"]
    walk = model.forward_layerwise(batch)["global"]
    torch.cuda.synchronize()
    assert torch.isfinite(after).all()
    assert float((after - before).abs().max()) > 0.0, "the fused path still runs the old weights"
    assert float((after - walk).abs().max()) <= 1e-4 * float(walk.abs().max()), "fused path != layer walk after the update"


def test_minkloc_training_step_engine_vs_cpu_double(cuda, monkeypatch):
    import egonn_b200 as E
    import egonn_b200.minkowski as ME
    from cpu_engine import CpuEngine
    torch.manual_seed(3)
    mp = E.ModelParams.from_dict(model="MinkLoc", coordinates="cartesian", quantization_step=0.4, block="BasicBlock",
                                 pooling="SPoC", feature_size=256, output_dim=256)
    ref = E.model_factory(mp)
    with torch.no_grad():
        for name, b in ref.named_buffers():
            if name.endswith("running_var"):
                b.uniform_(0.5, 1.5)
    model = copy.deepcopy(ref).to(cuda)
    coords = torch.from_numpy(load_golden("mini3_cartesian")["coords"])

    def step(m, c):
        m.train()
        m.zero_grad(set_to_none=True)
        g = m({"coords": c, "features": torch.ones((c.shape[0], 1), device=c.device)})["global"]
        loss = (g ** 2).sum() + g.sum()
        loss.backward()
        return float(loss.detach())

    loss_gpu = step(model, coords.to(cuda))
    torch.cuda.synchronize()
    monkeypatch.setattr(ME, "Engine", CpuEngine)                    # same model classes, engine replaced by the oracle-backed double
    loss_cpu = step(ref, coords)
    assert abs(loss_gpu - loss_cpu) <= 1e-4 * abs(loss_cpu)
    errs = {}
    for (name, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and q.grad is not None, name
        errs[name] = float((p.grad.cpu() - q.grad).abs().max()) / max(float(q.grad.abs().max()), 1e-30)
    worst = max(errs.items(), key=lambda t: t[1])
    median = float(np.median(list(errs.values())))
    print("\n[MinkLoc training step] engine vs CPU double: worst %s %.2e, median %.2e" % (*worst, median))
    assert worst[1] <= 5e-2 and median <= 1e-3, (worst, median)     # two bars: see train_case.compare

"""The training-step case shared by tests/golden/make_golden_train.py (reference graph on the CPU shim), tests/
test_autograd_cpu.py (egonn_b200 host logic on the CPU test double) and tests/test_training_step_gpu.py (the CUDA engine):
batches, a loss that touches all four outputs and does not depend on the row order inside a cloud, and the sampling of
large gradient tensors for the fixture."""
import numpy as np
import torch

QUANT = dict(coordinates="cartesian", step=0.4)          # the mini3_cartesian golden case
SHIFT = (8, -16, 24)                                     # second batch = the same clouds translated (different BN statistics
FULL = 8192                                              # after the stride-2 levels: 8 and 24 are not multiples of 128)
SAMPLE = 4096


def batches(coords: torch.Tensor):
    a = coords
    b = coords + torch.tensor([0, *SHIFT], dtype=coords.dtype, device=coords.device)
    return [{"coords": c.contiguous(), "features": torch.ones((c.shape[0], 1), device=c.device)} for c in (a, b)]


def loss_of(y) -> torch.Tensor:
    g = y["global"]
    dev = g.device
    wd = torch.sin(torch.arange(128, dtype=torch.float32, device=dev))
    wk = torch.tensor([0.3, -0.2, 0.5], device=dev)
    loss = (g ** 2).sum()
    for d, k, s in zip(y["descriptors"], y["keypoints"], y["sigma"]):
        loss = loss + (d * wd).sum() + (k * wk).sum() * 0.1 + 0.5 * (s ** 2).sum()
    return loss


def step(model, coords: torch.Tensor):
    """model.train(); two forwards alive before one backward (training/trainer.py:178-188).  Returns the loss (float)."""
    model.train()
    model.zero_grad(set_to_none=True)
    a, b = batches(coords)
    loss = loss_of(model(a)) + 0.5 * loss_of(model(b))
    loss.backward()
    return float(loss.detach())


def sample(t: torch.Tensor) -> np.ndarray:
    f = t.detach().reshape(-1).cpu().double().numpy()
    if f.shape[0] <= FULL:
        return f
    return f[:: f.shape[0] // SAMPLE][:SAMPLE]


def record(model, loss: float) -> dict:
    out = {"loss": np.float64(loss)}
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        out["g/" + name] = sample(p.grad)
        out["n/" + name] = np.float64(p.grad.detach().double().norm().item())
    for name, b in model.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            out["b/" + name] = b.detach().cpu().double().numpy()
    return out


def compare(model, loss: float, golden: dict, tol_grad: float, tol_median: float, tol_fwd: float, what: str):
    """Against the fixture, per tensor e = max|a-b| / max|b|:
      * forward quantities (the loss, the BatchNorm running statistics after the two train-mode forwards): e <= tol_fwd;
      * every parameter gradient (sampled like the fixture) and its L2 norm: e <= tol_grad, and the MEDIAN over the
        gradient tensors <= tol_median.
    Two bars for the gradients because one training step of this network is ill-conditioned in a heavy-tailed way: the
    upper pyramid levels normalise over a few dozen rows (24 at level 7 of this case) and a pre-activation that changes
    sign moves a whole tensor.  Measured on the CPU (forward operators perturbed by relative noise 1e-6 / 3e-6 / 1e-5, i.e.
    fp32 rounding): median 2e-6 / 9e-6 / 3e-4, worst tensor 5e-3 / 2e-3 / 8e-3.  A wrong backward rule (kernel flip,
    transposition, parent / child link) is off by O(1) in every tensor below it."""
    got = record(model, loss)
    assert set(got) == set(golden), sorted(set(got) ^ set(golden))[:6]
    errs = {}
    for k in sorted(golden):
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(golden[k], dtype=np.float64)
        assert a.shape == b.shape, k
        errs[k] = float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30)
    grads = {k: e for k, e in errs.items() if k.startswith("g/")}
    fwd = {k: e for k, e in errs.items() if k == "loss" or k.startswith("b/")}
    for k, e in fwd.items():
        assert e <= tol_fwd, f"{what}: {k}: rel err {e:.3e} > {tol_fwd:g}"
    for k, e in errs.items():
        if k not in fwd:
            assert e <= tol_grad, f"{what}: {k}: rel err {e:.3e} > {tol_grad:g}"
    median = float(np.median(list(grads.values())))
    assert median <= tol_median, f"{what}: median gradient rel err {median:.3e} > {tol_median:g}"
    worst = max(grads.items(), key=lambda t: t[1])
    return {"worst_grad": worst, "median_grad": median, "worst_forward": max(fwd.items(), key=lambda t: t[1])}

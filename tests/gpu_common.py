"""Helpers shared by the -m gpu parity tests: compare engine outputs (canonical Morton order) with oracle
outputs (lexicographic order) keyed by coordinate."""
import numpy as np
import torch

from oracle import me_ops

# north_star: descriptor tensors within 1e-3 relative fp32; we hold the FP32 path to a tighter bar.
RTOL = 1e-3


def lex_order(coords_t: torch.Tensor) -> np.ndarray:
    return me_ops.canonical_order(coords_t.cpu().numpy())


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close_rel(a, b, tol, what):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: max|a-b|/max|b| = {e:.3e} > {tol:.1e}"
    return e

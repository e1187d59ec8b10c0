"""Mirror of the reference ``datasets/quantization.py`` (Quantizer / PolarQuantizer / CartesianQuantizer)
with the quantisation itself running on the GPU through ``egn_quantize``.

Same names, constructor arguments and return values as the reference (datasets/quantization.py:9-103);
``__call__`` accepts a CPU or CUDA (N,3) float tensor and returns (coords int32 (M,3), index int64 (M,))
on the input's device - ME.utils.sparse_quantize semantics (first occurrence wins, input order)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List

import threading

import numpy as np
import torch

_engines = {}
_engines_lock = threading.Lock()


def _engine(device):
    """One engine context per (device, CUDA stream, host thread): an ``egn_ctx`` holds single-stream scratch (arena, device
    counters, pinned counts), so calls from two streams or two threads must not share one."""
    from .engine import Engine
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream, threading.get_ident())
    with _engines_lock:
        eng = _engines.get(key)
        if eng is None:
            eng = _engines[key] = Engine(torch.device("cuda", idx))
    return eng


def _quantize_on_gpu(pc: torch.Tensor, step, polar: bool):
    assert pc.shape[1] == 3
    if not torch.cuda.is_available():
        raise RuntimeError("egonn_b200 quantises on the GPU and has no CPU path: no CUDA device is visible in this process "
                           "(inside a forked DataLoader worker, quantise in the main process or start workers with 'spawn')")
    src = pc.device
    dev = pc.device if pc.is_cuda else torch.device("cuda", torch.cuda.current_device())
    coords, ndx = _engine(dev).quantize(pc.to(dev), step, polar)
    return coords.to(src), ndx.to(src)


class Quantizer(ABC):
    @abstractmethod
    def __call__(self, pc):
        pass

    @abstractmethod
    def dequantize(self, coords):
        pass

    @abstractmethod
    def keypoint_position(self, supervoxel_centers, stride, kp_offset):
        pass

    @abstractmethod
    def describe(self) -> dict:
        """{'coordinates': 'polar'|'cartesian', 'step': ...} - what the CUDA forward needs to place keypoints."""


class PolarQuantizer(Quantizer):
    """datasets/quantization.py:22-72: sector [deg], ring [m], z [m] steps."""

    def __init__(self, quant_step: List[float]):
        assert len(quant_step) == 3, '3 quantization steps expected: for sector (in degrees), ring and z-coordinate (in meters)'
        self.quant_step = torch.tensor(quant_step, dtype=torch.float)
        self.theta_range = int(360. // self.quant_step[0])

    def __call__(self, pc):
        return _quantize_on_gpu(pc, [float(v) for v in self.quant_step], polar=True)

    def to_cartesian(self, pc):
        theta = np.pi * (pc[:, 0] - 180.) / 180.
        return torch.stack([torch.cos(theta) * pc[:, 1], torch.sin(theta) * pc[:, 1], pc[:, 2]], dim=1)

    def dequantize(self, coords):
        return self.to_cartesian((0.5 + coords) * self.quant_step.to(coords.device))

    def keypoint_position(self, supervoxel_centres, stride, kp_offset):
        device = supervoxel_centres.device
        centres = (supervoxel_centres + 0.5) * self.quant_step.to(device)
        size = torch.tensor(stride, dtype=torch.float, device=device) * self.quant_step.to(device)
        return self.to_cartesian(centres + kp_offset * size / 2.)

    def describe(self):
        return {"coordinates": "polar", "step": [float(v) for v in self.quant_step]}


class CartesianQuantizer(Quantizer):
    """datasets/quantization.py:75-103."""

    def __init__(self, quant_step: float):
        self.quant_step = quant_step

    def __call__(self, pc):
        return _quantize_on_gpu(pc, float(self.quant_step), polar=False)

    def dequantize(self, coords):
        return (0.5 + coords) * self.quant_step

    def keypoint_position(self, supervoxel_centers, stride, kp_offset):
        centres = (supervoxel_centers + 0.5) * self.quant_step
        size = torch.tensor(stride, dtype=torch.float, device=centres.device) * self.quant_step
        return centres + kp_offset * size / 2. if kp_offset is not None else centres

    def describe(self):
        return {"coordinates": "cartesian", "step": float(self.quant_step)}


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """ME.utils.batched_coordinates (eval/evaluate.py:333, datasets/dataset_utils.py:77): concatenate the
    per-cloud (Mi,3) voxel coordinates and prepend the list index as the batch column."""
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        if c.dtype.is_floating_point:
            c = torch.floor(c)
        c = c.to(dtype)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c], dim=1))
    bc = torch.cat(out, dim=0) if out else torch.zeros((0, 4), dtype=dtype)
    return bc if device is None else bc.to(device)

"""egonn_b200 - B200-native sparse-voxel descriptor extraction behind EgoNN's model_factory / MinkGL API.

    from egonn_b200 import ModelParams, model_factory
    params = ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.1)
    model = model_factory(params).eval().cuda()
    model.load_state_dict(torch.load("weights.pth"))
    y = model({"coords": coords_cuda_int32_N4, "features": ones_cuda_N1})

The inference path (``model.eval()``) runs entirely in hand-written sm_100a CUDA kernels behind the C ABI of
``include/egonn_b200.h`` (``egonn_b200/csrc/libegonn_b200.so``).  The training step (``model.train()``) walks
the layers on the same C-ABI operators, with torch's BatchNorm / Linear on the feature matrices and torch's GEMM for the
weight gradients (``egonn_b200.autograd``).  There is no CPU fallback.
"""
from .params import ModelParams  # noqa: F401
from .models import (model_factory, create_egonn_model, MinkGL, MinkTrunk, MinkHead, ECABasicBlock, MinkFPN,  # noqa: F401
                     MinkLoc, MinkLoc3D)
from .quantization import CartesianQuantizer, PolarQuantizer, batched_coordinates  # noqa: F401
from .engine import Engine, topk_smallest, pack_topk, knn_global, match_descriptors, filter_points  # noqa: F401

__version__ = "0.1.0"
from .pipeline import Extractor, StagedBatch, stage_batch  # noqa: F401,E402

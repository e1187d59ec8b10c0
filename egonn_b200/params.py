"""Mirror of the reference ``misc.utils.ModelParams`` (misc/utils.py:11-73): same INI keys, same attributes,
but the quantizer it builds is the GPU-backed one from ``egonn_b200.quantization``.  ``from_dict`` builds the
same object without an INI file."""
from __future__ import annotations

import configparser

from .quantization import CartesianQuantizer, PolarQuantizer


class ModelParams:
    def __init__(self, model_params_path=None, **overrides):
        params = {}
        if model_params_path is not None:
            config = configparser.ConfigParser()
            if not config.read(model_params_path):
                raise FileNotFoundError(model_params_path)
            params = dict(config["MODEL"])
        params.update({k: str(v) for k, v in overrides.items()})
        self.model_params_path = model_params_path
        self.model = params.get("model")
        self.output_dim = int(params.get("output_dim", 256))
        self.coordinates = params.get("coordinates", "polar")
        assert self.coordinates in ["polar", "cartesian"], f"Unsupported coordinates: {self.coordinates}"
        if "polar" in self.coordinates:
            self.quantization_step = [float(e) for e in params["quantization_step"].split(",")]
            assert len(self.quantization_step) == 3, \
                "Expected 3 quantization steps: for sectors (degrees), rings (meters) and z coordinate (meters)"
            self.quantizer = PolarQuantizer(quant_step=self.quantization_step)
        else:
            self.quantization_step = float(params["quantization_step"])
            self.quantizer = CartesianQuantizer(quant_step=self.quantization_step)
        if "MinkLoc" in (self.model or ""):
            self.feature_size = int(params.get("feature_size", 256))
            self.planes = [int(e) for e in params["planes"].split(",")] if "planes" in params else [32, 64, 64]
            self.layers = [int(e) for e in params["layers"].split(",")] if "layers" in params else [1, 1, 1]
            self.num_top_down = int(params.get("num_top_down", 1))
            self.conv0_kernel_size = int(params.get("conv0_kernel_size", 5))
            self.block = params.get("block", "BasicBlock")
            self.pooling = params.get("pooling", "GeM")

    @classmethod
    def from_dict(cls, **kw):
        if isinstance(kw.get("quantization_step"), (list, tuple)):
            kw["quantization_step"] = ",".join(str(v) for v in kw["quantization_step"])
        return cls(None, **kw)

    def print(self):
        print("Model parameters:")
        for k, v in vars(self).items():
            print(f"{k}: {v}")
        print("")

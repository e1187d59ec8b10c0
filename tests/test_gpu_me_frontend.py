"""-m gpu: INTEGRATION.md path 2 - a MinkowskiEngine-style graph running on `egonn_b200.minkowski.install()`.

The reference's forward is written against `import MinkowskiEngine as ME` (models/minkgl.py, layers/eca_block.py,
layers/pooling.py).  After `install()` that import resolves to the engine's front end, every ME operator becoming one
C-ABI call.  Two tests:
  * a test-LOCAL graph (typed here, in the reference's idiom: an ECA block `layers/eca_block.py:8-73`, a two-level
    top-down head `models/minkgl.py:14-60`, a per-voxel MLP `:207-225` and GeM `layers/pooling.py:72-86`) written
    only against `ME.*`, loaded with the shipped checkpoint's tensors, compared with the FUSED engine path
    (`egn_forward` taps / outputs) <= 1e-4 - this runs on the GPU box, where /root/reference does not exist;
  * the UNMODIFIED reference files (`models/model_factory.py:12-78` + everything they import) on `install()`, compared
    with `egonn_b200.model_factory` - skipped where /root/reference is absent."""
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN_CASES, load_golden
from gpu_common import assert_close_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda", 0)


def _engine_model(weights, quant, cuda):
    import egonn_b200 as E
    mp = E.ModelParams.from_dict(model="egonn", coordinates=quant["coordinates"], quantization_step=quant["step"])
    m = E.model_factory(mp)
    m.load_state_dict(weights)
    return m.eval().to(cuda)


def test_me_style_graph_on_installed_frontend(cuda, weights):
    import egonn_b200.minkowski as front
    front.install()
    import MinkowskiEngine as ME                                    # == egonn_b200.minkowski from here on
    from MinkowskiEngine.modules.resnet_block import BasicBlock
    assert ME is front

    class ECALayer(nn.Module):                                      # layers/eca_block.py:8-36, re-typed
        def __init__(self, channels, k_size):
            super().__init__()
            self.avg_pool = ME.MinkowskiGlobalPooling()
            self.conv = nn.Conv1d(1, 1, kernel_size=k_size, padding=(k_size - 1) // 2, bias=False)
            self.sigmoid = nn.Sigmoid()
            self.broadcast_mul = ME.MinkowskiBroadcastMultiplication()

        def forward(self, x):
            y_sparse = self.avg_pool(x)
            y = self.sigmoid(self.conv(y_sparse.F.unsqueeze(-1).transpose(-1, -2)).transpose(-1, -2).squeeze(-1))
            y_sparse = ME.SparseTensor(y, coordinate_manager=y_sparse.coordinate_manager, coordinate_map_key=y_sparse.coordinate_map_key)
            return self.broadcast_mul(x, y_sparse)

    class ECABlock(BasicBlock):                                     # layers/eca_block.py:39-73, re-typed
        def __init__(self, inplanes, planes, k_size, downsample=None):
            super().__init__(inplanes, planes, stride=1, dilation=1, downsample=downsample, dimension=3)
            self.eca = ECALayer(planes, k_size)

        def forward(self, x):
            residual = x
            out = self.relu(self.norm1(self.conv1(x)))
            out = self.eca(self.norm2(self.conv2(out)))
            if self.downsample is not None:
                residual = self.downsample(x)
            out += residual
            return self.relu(out)

    class MiniNet(nn.Module):
        """conv0 -> 4 x (stride-2 conv, BN, ReLU, ECA block) -> local FPN head over levels 3-4 -> descriptor MLP; GeM over
        the level-4 block output: every operator class of the reference's forward, on the checkpoint's own tensors."""

        def __init__(self):
            super().__init__()
            planes = [32, 64, 64, 128]
            self.conv0 = ME.MinkowskiConvolution(1, 32, kernel_size=5, dimension=3)
            self.bn0 = ME.MinkowskiBatchNorm(32)
            self.relu = ME.MinkowskiReLU(inplace=True)
            self.convs, self.bns, self.blocks = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
            inp = 32
            for i, p in enumerate(planes):
                self.convs.append(ME.MinkowskiConvolution(inp, inp, kernel_size=2, stride=2, dimension=3))
                self.bns.append(ME.MinkowskiBatchNorm(inp))
                ds = None
                if inp != p:
                    ds = nn.Sequential(ME.MinkowskiConvolution(inp, p, kernel_size=1, stride=1, dimension=3), ME.MinkowskiBatchNorm(p))
                self.blocks.append(ECABlock(inp, p, 3 if p < 128 else 5, ds))
                inp = p
            self.lat4 = ME.MinkowskiConvolution(128, 64, kernel_size=1, stride=1, dimension=3)
            self.lat3 = ME.MinkowskiConvolution(64, 64, kernel_size=1, stride=1, dimension=3)
            self.tconv4 = ME.MinkowskiConvolutionTranspose(64, 64, kernel_size=2, stride=2, dimension=3)
            self.mlp = nn.Sequential(ME.MinkowskiLinear(64, 96), ME.MinkowskiReLU(inplace=True), ME.MinkowskiLinear(96, 128))
            self.pool = ME.MinkowskiGlobalAvgPooling()

        def forward(self, batch):
            x = ME.SparseTensor(batch["features"], coordinates=batch["coords"])
            x = self.relu(self.bn0(self.conv0(x)))
            ys = []
            for conv, bn, block in zip(self.convs, self.bns, self.blocks):
                x = block(self.relu(bn(conv(x))))
                ys.append(x)
            y = self.tconv4(self.lat4(ys[3])) + self.lat3(ys[2])
            d = ME.MinkowskiFunctional.normalize(self.mlp(y))
            return ys, y, d, self.pool(ys[3])

    net = MiniNet()
    sd = {}
    ck = weights
    sd["conv0.kernel"], pre = ck["trunk.convs.0.kernel"], "trunk.bn.0.bn."
    for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
        sd["bn0.bn." + k] = ck[pre + k]
    for i in range(4):
        L = i + 1
        sd[f"convs.{i}.kernel"] = ck[f"trunk.convs.{L}.kernel"]
        for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            sd[f"bns.{i}.bn.{k}"] = ck[f"trunk.bn.{L}.bn.{k}"]
        for name, v in ck.items():
            if name.startswith(f"trunk.blocks.{L}.0."):
                sd[f"blocks.{i}." + name[len(f"trunk.blocks.{L}.0."):]] = v
    sd["lat4.kernel"], sd["lat3.kernel"] = ck["local_head.conv1x1.4.kernel"], ck["local_head.conv1x1.3.kernel"]
    sd["tconv4.kernel"] = ck["local_head.tconv.4.kernel"]
    for j in (0, 2):
        for k in ("weight", "bias"):
            sd[f"mlp.{j}.linear.{k}"] = ck[f"local_descriptor_decoder.net.{j}.linear.{k}"]
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all("eca.conv" not in m for m in missing), (missing, unexpected)
    assert not missing, missing
    net = net.eval().to(cuda)

    g = load_golden("mini3_cartesian")
    coords = torch.from_numpy(g["coords"]).to(cuda)
    batch = {"coords": coords, "features": torch.ones((coords.shape[0], 1), device=cuda)}
    with torch.no_grad():
        ys, ymap, desc, pooled = net(batch)

    fused = _engine_model(weights, GOLDEN_CASES["mini3_cartesian"], cuda)
    p = fused.forward_packed(batch)
    eng = fused._engine
    for L in (1, 2, 3, 4):                                          # both paths keep canonical row order: compare row for row
        assert torch.equal(ys[L - 1].C, eng.level_coords(L))
        assert_close_rel(ys[L - 1].F, eng.tap(2, L, ys[L - 1].F.shape[1]), 1e-4, f"block{L}")
    assert_close_rel(ymap.F, eng.tap(4, 3, 64), 1e-4, "local head map")
    assert_close_rel(desc.F, p["descriptors"], 1e-4, "local descriptors")
    # per-cloud mean of the level-4 block output through the front end's pooling == the same reduction done with torch
    off = eng.batch_offsets(4).cpu().tolist()
    ref_pool = torch.stack([ys[3].F[off[b]:off[b + 1]].mean(0) for b in range(len(off) - 1)])
    assert_close_rel(pooled.F, ref_pool, 1e-5, "global average pooling")


def test_unmodified_reference_graph_on_installed_frontend(cuda, weights):
    """The reference's own files on the engine front end (skipped on the GPU box: /root/reference is not there)."""
    ref_root = os.environ.get("EGONN_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref_root, "models")):
        pytest.skip("reference sources absent (GPU box)")
    import egonn_b200.minkowski as front
    front.install()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    if "datasets" not in sys.modules or not hasattr(sys.modules["datasets"], "__egonn_ref__"):
        m = types.ModuleType("datasets")                           # HuggingFace `datasets` shadows the reference's package
        m.__path__ = [os.path.join(ref_root, "datasets")]
        m.__egonn_ref__ = True
        sys.modules["datasets"] = m
    for name in [n for n in sys.modules if n.split(".")[0] in ("models", "layers", "misc")]:
        del sys.modules[name]
    from misc.utils import ModelParams                                  # noqa: E402  (reference)
    from models.model_factory import model_factory                      # noqa: E402  (reference, unmodified)
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("[MODEL]\nmodel = egonn\ncoordinates = cartesian\nquantization_step = 0.4\n")
        cfg = f.name
    ref_model = model_factory(ModelParams(cfg))
    ref_model.load_state_dict(weights)
    ref_model = ref_model.eval().to(cuda)
    g = load_golden("mini3_cartesian")
    coords = torch.from_numpy(g["coords"]).to(cuda)
    batch = {"coords": coords, "features": torch.ones((coords.shape[0], 1), device=cuda)}
    with torch.no_grad():
        y = ref_model(batch)
    fused = _engine_model(weights, GOLDEN_CASES["mini3_cartesian"], cuda)
    z = fused(batch)
    assert_close_rel(y["global"], z["global"], 1e-4, "global")
    for k in ("descriptors", "keypoints", "sigma"):
        for a, b in zip(y[k], z[k]):
            assert_close_rel(a, b, 1e-4, k)

"""Golden vectors of ONE TRAINING STEP (train-mode BatchNorm + backward, training/trainer.py:141-195) from the UNMODIFIED
reference graph: ``models.model_factory.model_factory`` of /root/reference on the ME-semantics CPU shim (oracle/me_shim),
shipped checkpoint, ``model.train()``, two forwards + one backward of the loss in tests/train_case.py; gradients by
torch's own autograd through the shim's gather -> mm -> index_add convolutions.

    python tests/golden/make_golden_train.py        # here, where /root/reference exists -> tests/golden/train_mini3.npz

Stored: the loss, every parameter's gradient (whole when <= 8192 elements, else 4096 strided samples) and its L2 norm, and
the BatchNorm running statistics after the two forwards."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, HERE)

from oracle import ref_import  # noqa: E402
import train_case  # noqa: E402


def main(real_me: bool = False, reference_root: str = None, suffix: str = ""):
    """``real_me`` (tools/verify_against_me.py --write-golden): the same step on the REAL MinkowskiEngine ->
    ``train_mini3_me.npz``, which the training tests then check as well."""
    if real_me:
        import make_golden
        make_golden.enable_real_me(reference_root or ref_import.REFERENCE_ROOT)
    else:
        ref_import.enable()
    from models.model_factory import model_factory
    from misc.utils import ModelParams
    sd = torch.load(os.path.join(HERE, "egonn_weights.pth"), map_location="cpu", weights_only=True)
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(f"[MODEL]\nmodel = egonn\ncoordinates = {train_case.QUANT['coordinates']}\nquantization_step = {train_case.QUANT['step']}\n")
    mp = ModelParams(f.name)
    os.unlink(f.name)
    model = model_factory(mp)
    model.load_state_dict(sd)
    coords = torch.from_numpy(np.load(os.path.join(HERE, "mini3_cartesian.npz"))["coords"])
    torch.manual_seed(0)
    loss = train_case.step(model, coords)
    rec = train_case.record(model, loss)
    path = os.path.join(HERE, "train_mini3" + suffix + ".npz")
    np.savez_compressed(path, **rec)
    print("loss", loss, "tensors", len(rec), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()

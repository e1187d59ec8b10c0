"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/egonn_b200.h declares, struct layouts match, the model mirrors the reference state_dict."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from egonn_b200 import lib as L
    header = open(os.path.join(REPO, "include", "egonn_b200.h")).read()
    declared = set(re.findall(r"\b(egn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    assert L.load().egn_version() >= 100


def test_struct_layout_matches_c(tmp_path):
    from egonn_b200 import lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "egonn_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu", sizeof(egn_net), sizeof(egn_layer), sizeof(egn_head),'
                   'sizeof(egn_coords_info), offsetof(egn_net, global_head), offsetof(egn_net, quant_step));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == [ctypes.sizeof(L.Net), ctypes.sizeof(L.Layer), ctypes.sizeof(L.Head), ctypes.sizeof(L.CoordsInfo),
                   L.Net.global_head.offset, L.Net.quant_step.offset]


def test_model_factory_accepts_reference_checkpoint(weights):
    import egonn_b200 as E
    from egonn_b200 import weights as W
    mp = E.ModelParams.from_dict(model="egonn", coordinates="polar", quantization_step=[1., 0.3, 0.2])
    m = E.model_factory(mp)
    own = m.state_dict()
    assert list(own.keys()) == list(weights.keys())                  # same names, same order as Appendix B
    for k in own:
        assert own[k].shape == weights[k].shape, k
    m.load_state_dict(weights)
    blob, net = W.pack_egonn(m.state_dict(), mp.quantizer.describe())
    assert net.n_levels == 7 and net.conv0_ksize == 5 and net.polar == 1
    assert [net.eca_k[i] for i in range(1, 8)] == [3, 3, 3, 5, 5, 5, 5]
    assert abs(net.gem_p - float(weights["global_pooling.pooling.p"])) < 1e-7
    assert blob.numel() % 4 == 0
    with pytest.raises(NotImplementedError):
        E.model_factory(E.ModelParams.from_dict(model="nope", coordinates="cartesian", quantization_step=0.1))


def test_no_cpu_path():
    import egonn_b200 as E
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from egonn_b200.lib import EgnError
    with pytest.raises(EgnError):
        E.Engine()
    m = E.model_factory(E.ModelParams.from_dict(model="egonn", coordinates="cartesian", quantization_step=0.3)).eval()
    with pytest.raises(RuntimeError):
        m({"coords": torch.zeros((4, 4), dtype=torch.int32), "features": torch.ones((4, 1))})


def test_product_never_imports_oracle():
    """③: nothing under egonn_b200/ may import, link or execute anything under oracle/."""
    for root, _, files in os.walk(os.path.join(REPO, "egonn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "oracle/" not in text and "oracle." not in text.replace("oracle (", ""), f

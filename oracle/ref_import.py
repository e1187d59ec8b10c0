"""Make the UNMODIFIED reference importable in this container on top of the oracle ME shim.

TEST INFRASTRUCTURE - used only by tests/golden/make_golden.py and by CPU tests that are skipped when
/root/reference is absent (it is absent on the GPU box).  Works around SURVEY.md Appendix C.1:
  1. HuggingFace ``datasets`` shadows the reference's namespace package ``datasets/``;
  2. the reference has no __init__.py files (its root must be on sys.path);
  3. ``import MinkowskiEngine`` must resolve to ``oracle/me_shim/MinkowskiEngine``.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EGONN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def enable():
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    for p in (repo, os.path.join(here, "me_shim"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "datasets" not in sys.modules or not hasattr(sys.modules["datasets"], "__egonn_ref__"):
        m = types.ModuleType("datasets")
        m.__path__ = [os.path.join(REFERENCE_ROOT, "datasets")]
        m.__egonn_ref__ = True
        sys.modules["datasets"] = m
    import MinkowskiEngine  # noqa: F401
    assert "oracle-shim" in MinkowskiEngine.__version__

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_built():
    """a fresh checkout has no built artefacts (they are git-ignored): build them once, in-tree"""
    import glob
    have = (os.path.exists(os.path.join(ROOT, "eva_b200", "lib", "libevab200.so")) and glob.glob(os.path.join(ROOT, "eva_b200", "_eva_b200*.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "libckks_oracle.so")))
    if not have:
        import __graft_entry__
        __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    if not hasattr(config, "workerinput"):      # xdist workers inherit the controller's build
        _ensure_built()
    # the test fixtures (programs compiled by the reference compiler) are found by name next to the shipped workloads
    from eva_b200 import program_io
    fixtures = os.path.join(ROOT, "tests", "golden", "programs")
    if fixtures not in program_io.SEARCH:
        program_io.SEARCH.append(fixtures)


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)

"""The C-ABI library loads here (no GPU) and exports every symbol declared in
include/evab200.h; no compute calls."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from eva_b200 import build, cabi
    build.build()
    lib = cabi.load()
    hdr = open(os.path.join(ROOT, "include", "evab200.h")).read()
    declared = sorted(set(re.findall(r"\b(evab_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(cabi.exported_symbols())
    assert lib.evab_version() >= 1
    assert lib.evab_galois_elt_from_step(16384, 1) == 3
    assert lib.evab_galois_elt_from_step(16384, 0) == 32767


def test_no_device_fails_loudly():
    import ctypes as C
    import numpy as np
    import torch
    if torch.cuda.is_available():
        return
    from eva_b200 import cabi
    lib = cabi.load()
    h = C.c_void_p()
    p = np.array([0xffffffffffc0001], dtype=np.uint64)
    rc = lib.evab_ctx_create(C.c_uint64(16384), p.ctypes.data_as(cabi.u64p), 1, 0, C.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.evab_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "eva_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
            if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert '#include "../../oracle' not in src and "ckks_oracle" not in src, f

"""The C-ABI library loads and exports every symbol include/ldic.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    import ldic_b200
    if not os.path.exists(ldic_b200.lib_path()):
        ge.build()
    hdr = open(os.path.join(ROOT, "include", "ldic.h")).read()
    declared = set(re.findall(r"LDIC_API\s+[\w\s\*]+?\b(ldic_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(ldic_b200.lib_path())
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(ldic_b200.EXPORTED_SYMBOLS), declared ^ set(ldic_b200.EXPORTED_SYMBOLS)
    lib.ldic_version.restype = ctypes.c_int
    assert lib.ldic_version() >= 100
    lib.ldic_likelihood_workspace_bytes.restype = ctypes.c_size_t
    assert lib.ldic_likelihood_workspace_bytes() > 0


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "learning-driven-image-compression-algorithm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle\.|ref_path|/root/reference", src, re.M), f


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    import ldic_b200
    with pytest.raises(ldic_b200.LdicError):
        ldic_b200.ops.mse_sum(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 4, 4))
    with pytest.raises(ldic_b200.LdicError):
        ldic_b200.Net((1, 64, 64, 3), (1, 64, 64, 3), False, False)(torch.zeros(1, 3, 64, 64), "test")

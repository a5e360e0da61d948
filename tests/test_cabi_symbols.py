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


def test_torch_library_ops_registered_with_fake_impls():
    """the ops are registered with the PyTorch dispatcher (namespace ldic::).  Without a GPU the fake
    (meta) implementations still propagate shapes / dtypes on fake CUDA tensors, and CPU tensors are refused."""
    import pytest
    import torch
    import ldic_b200
    from ldic_b200 import torch_ops
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in torch_ops.REGISTERED:
        assert hasattr(torch.ops.ldic, name), name
    K = ldic_b200._lib
    with FakeTensorMode():
        x = torch.empty(2, 192, 8, 12, device="cuda")
        assert torch.ops.ldic.gdn(x, torch.empty(192, device="cuda"), torch.empty(192, 192, device="cuda"), False, True).shape == x.shape
        vh, lik, s = torch.ops.ldic.round_likelihood_bpp(x, x, x, 1, 0, 1e-8, 0.11)
        assert vh.shape == x.shape and lik.shape == x.shape and s.shape == (1,)
        xi = torch.empty(2, 16, 24, 192, device="cuda", dtype=torch.bfloat16)
        w = torch.empty(16, device="cuda", dtype=torch.bfloat16)
        b = torch.empty(192, device="cuda")
        y = torch.ops.ldic.conv_forward(xi, w, b, None, None, K.LDIC_CONV_S2_5x5_P12, 192, 192, 192, 192, K.ACT_NONE, True, 0, 0)
        assert tuple(y.shape) == (2, 8, 12, 192) and y.dtype == torch.float32
        y = torch.ops.ldic.conv_forward(xi, w, b, None, None, K.LDIC_DECONV_GS_5x5, 192, 384, 192, 384, K.ACT_NONE, False, 0, 0)
        assert tuple(y.shape) == (2, 32, 48, 384) and y.dtype == torch.bfloat16        # 384 output channels: the wide kernel
        img = torch.empty(2, 3, 64, 128, device="cuda", dtype=torch.uint8)
        y = torch.ops.ldic.conv_forward(img, w, b, None, None, K.LDIC_CONV_FIRST_5x5S2, 3, 192, 128, 192, K.ACT_NONE, False, 0, 0)
        assert tuple(y.shape) == (2, 32, 64, 192)
        assert torch.ops.ldic.mse_sum(torch.empty(3, 3, 8, 8, device="cuda"), torch.empty(3, 3, 8, 8, device="cuda"), False).shape == (3,)
    with pytest.raises(NotImplementedError):
        torch.ops.ldic.lower_bound(torch.zeros(4), 0.1)

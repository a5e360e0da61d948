"""Loads the UNMODIFIED reference classes from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container (``/root/reference`` does not exist on the
GPU box).  Used by ``tests/golden/make_golden.py`` to generate the committed
fixtures, and by ``tests/test_oracle_vs_reference_live.py`` (auto-skipped when
the reference tree is absent).  Call it in a dedicated process: it mutates
``sys.modules``, ``torch.Tensor.cuda`` and ``PIL.Image.Image.save``.

Recipe: SURVEY.md Appendix A (stubs for compressai / timm / model.Haar, the
``.cuda()`` identity patch on CPU, PNG-save no-op, CUDA_VISIBLE_DEVICES restore).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("LDIC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "net.py"))


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def load_leaf(relpath: str, name: str):
    """Import one reference file by path (no package side effects)."""
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_net_module():
    """Returns the reference ``model.net`` module, importable on CPU."""
    import torch
    import torch.nn as nn

    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")

    class _NA(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("stub: not used by model/net.py")

    _mod("compressai")
    _mod("compressai.entropy_models", EntropyBottleneck=_NA, GaussianConditional=_NA)
    _mod("compressai.layers", AttentionBlock=_NA, ResidualBlock=_NA, ResidualBlockUpsample=_NA,
         ResidualBlockWithStride=_NA,
         conv3x3=lambda i, o, stride=1: nn.Conv2d(i, o, 3, stride, 1),
         subpel_conv3x3=lambda i, o, r=1: nn.Sequential(nn.Conv2d(i, o * r * r, 3, padding=1), nn.PixelShuffle(r)))
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=nn.Identity, to_2tuple=lambda x: (x, x),
         trunc_normal_=nn.init.trunc_normal_)
    import model  # noqa: F401  (reference package; __init__ only imports numpy)
    _mod("model.Haar", define_G=lambda *a, **k: None)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    import PIL.Image
    PIL.Image.Image.save = lambda self, *a, **k: None
    import model.net as net_mod
    if cvd is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    return net_mod

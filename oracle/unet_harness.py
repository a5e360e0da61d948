"""Loads the reference's U-Net-family model file (model/net_unet_ha_hs.py, UNMODIFIED) on CPU with its missing
dependencies RESTATED (TEST INFRASTRUCTURE, build container only; used by tests/golden/make_golden_unet.py).

PARITY CAVEAT ("restated deps"): the reference imports `compressai` and `timm` (not installed, not vendored, not
version-pinned: SURVEY 8c) and `model/DepthwiseSeparableConv.py`, `model/Haar.py` (absent from the reference tree).
Everything below the `--- restated ---` line follows SURVEY Appendix B, i.e. the published upstream semantics; the
`DepthwiseSeparableConv` definition is a GUESS from its call shape (model/net_unet_ha_hs.py:536-542).  Fixtures made
through this harness pin our U-Net assembly to "the reference's own code + these restatements", not to an
installation of the reference.  Run in a dedicated process: it mutates sys.modules / sys.argv / torch.Tensor.to.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ref_harness
from .ref_harness import REF_ROOT, _mod


# --- restated ------------------------------------------------------------------------------------------------
def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


def _make_layers(GDN):
    class ResidualBlock(nn.Module):
        """conv3x3 -> LeakyReLU -> conv3x3 -> LeakyReLU, + skip (1x1 conv iff in != out)."""

        def __init__(self, in_ch, out_ch):
            super().__init__()
            self.conv1 = conv3x3(in_ch, out_ch)
            self.leaky_relu = nn.LeakyReLU(inplace=True)
            self.conv2 = conv3x3(out_ch, out_ch)
            self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

        def forward(self, x):
            identity = x
            out = self.leaky_relu(self.conv1(x))
            out = self.leaky_relu(self.conv2(out))
            if self.skip is not None:
                identity = self.skip(x)
            return out + identity

    class ResidualBlockWithStride(nn.Module):
        """conv3x3(stride) -> LeakyReLU -> conv3x3 -> GDN, + conv1x1(stride) skip."""

        def __init__(self, in_ch, out_ch, stride=2):
            super().__init__()
            self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
            self.leaky_relu = nn.LeakyReLU(inplace=True)
            self.conv2 = conv3x3(out_ch, out_ch)
            self.gdn = GDN(out_ch)
            self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

        def forward(self, x):
            identity = x
            out = self.leaky_relu(self.conv1(x))
            out = self.gdn(self.conv2(out))
            if self.skip is not None:
                identity = self.skip(x)
            return out + identity

    class ResidualBlockUpsample(nn.Module):
        def __init__(self, in_ch, out_ch, upsample=2):
            super().__init__()
            self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
            self.leaky_relu = nn.LeakyReLU(inplace=True)
            self.conv = conv3x3(out_ch, out_ch)
            self.igdn = GDN(out_ch, inverse=True)
            self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

        def forward(self, x):
            out = self.leaky_relu(self.subpel_conv(x))
            out = self.igdn(self.conv(out))
            return out + self.upsample(x)

    class AttentionBlock(nn.Module):
        """a = 3 x ResidualUnit(x); b = conv1x1(3 x ResidualUnit(x)); out = a * sigmoid(b) + x."""

        def __init__(self, N):
            super().__init__()

            class ResidualUnit(nn.Module):
                def __init__(self):
                    super().__init__()
                    self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                              nn.ReLU(inplace=True), conv1x1(N // 2, N))
                    self.relu = nn.ReLU(inplace=True)

                def forward(self, x):
                    return self.relu(self.conv(x) + x)

            self.conv_a = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit())
            self.conv_b = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit(), conv1x1(N, N))

        def forward(self, x):
            return self.conv_a(x) * torch.sigmoid(self.conv_b(x)) + x

    return ResidualBlock, ResidualBlockWithStride, ResidualBlockUpsample, AttentionBlock


class EntropyBottleneck(nn.Module):
    """Medians-only restatement: the reference discards the likelihoods (model/net_unet_ha_hs.py:882) and uses only
    `_get_medians()` (:885), which is 0 at init (`quantiles` = [-10, 0, 10] per channel)."""

    def __init__(self, channels, *a, **k):
        super().__init__()
        q = torch.tensor([-10.0, 0.0, 10.0]).repeat(channels, 1, 1)
        self.quantiles = nn.Parameter(q)

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()

    def forward(self, x):
        m = self._get_medians().reshape(1, -1, 1, 1)
        return torch.round(x - m) + m, torch.ones_like(x)


def _make_gaussian_conditional(LowerBound):
    class GaussianConditional(nn.Module):
        """eval forward of CompressAI's GaussianConditional(None)."""

        def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9, likelihood_bound=1e-9):
            super().__init__()
            self.lower_bound_scale = LowerBound(scale_bound)
            self.likelihood_lower_bound = LowerBound(likelihood_bound)

        @staticmethod
        def _standardized_cumulative(inputs):
            return 0.5 * torch.erfc(-(2 ** -0.5) * inputs)

        def forward(self, inputs, scales, means=None):
            if self.training:
                outputs = inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
            else:
                outputs = torch.round(inputs - means) + means if means is not None else torch.round(inputs)
            values = outputs - means if means is not None else outputs
            scales = self.lower_bound_scale(scales)
            values = torch.abs(values)
            upper = self._standardized_cumulative((0.5 - values) / scales)
            lower = self._standardized_cumulative((-0.5 - values) / scales)
            return outputs, self.likelihood_lower_bound(upper - lower)

    return GaussianConditional


class DepthwiseSeparableConv(nn.Module):
    """GUESSED (module absent from the reference): depthwise 3x3 (padding 1) followed by pointwise 1x1, the textbook
    block of that name; only the call shape `DepthwiseSeparableConv(in_channels=c, out_channels=c)` is known."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.depthwise = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding, groups=in_channels)
        self.pointwise = nn.Conv2d(in_channels, out_channels, 1)

    def forward(self, x):
        return self.pointwise(self.depthwise(x))


# --- harness -------------------------------------------------------------------------------------------------
def load_unet_module(name: str = "net_unet_ha_hs"):
    """Returns the reference `model.<name>` module, importable on CPU."""
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    sys.argv = sys.argv[:1]                                    # argparse inside Net.__init__
    import ops as ref_ops                                      # the reference's vendored CompressAI ops (LowerBound, ...)
    ref_gdn = ref_harness.load_leaf("layers/gdn.py", "ref_layers_gdn_for_unet")
    RB, RBS, RBU, AB = _make_layers(ref_gdn.GDN)
    _mod("compressai")
    _mod("compressai.entropy_models", EntropyBottleneck=EntropyBottleneck,
         GaussianConditional=_make_gaussian_conditional(ref_ops.LowerBound))
    _mod("compressai.layers", AttentionBlock=AB, ResidualBlock=RB, ResidualBlockUpsample=RBU, ResidualBlockWithStride=RBS,
         conv3x3=conv3x3, subpel_conv3x3=subpel_conv3x3, GDN=ref_gdn.GDN)
    _mod("compressai.ops", LowerBound=ref_ops.LowerBound, ste_round=ref_ops.ste_round)
    _mod("timm")
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    _mod("timm.models")
    _mod("timm.models.registry", register_model=lambda f: f)
    _mod("timm.models.layers", DropPath=lambda *a, **k: nn.Identity(), to_2tuple=lambda x: (x, x),
         trunc_normal_=nn.init.trunc_normal_)
    plt = _mod("matplotlib.pyplot", rcParams={}, rc=lambda *a, **k: None,
               style=types.SimpleNamespace(use=lambda *a, **k: None))
    _mod("matplotlib", pyplot=plt)
    _mod("seaborn", set_style=lambda *a, **k: None)
    import model  # noqa: F401
    _mod("model.Haar", define_G=lambda *a, **k: None)
    _mod("model.DepthwiseSeparableConv", DepthwiseSeparableConv=DepthwiseSeparableConv)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
        _to = torch.Tensor.to

        def to_cpu(self, *a, **k):                              # NoiseQuant: torch.tensor(..).to(torch.device("cuda"))
            a = tuple(torch.device("cpu") if (isinstance(x, torch.device) and x.type == "cuda") or x == "cuda" else x for x in a)
            return _to(self, *a, **k)
        torch.Tensor.to = to_cpu
    import PIL.Image
    PIL.Image.Image.save = lambda self, *a, **k: None
    import importlib
    m = importlib.import_module(f"model.{name}")
    if cvd is None:
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = cvd
    return m

"""U-Net-family ``Net`` of the reference (model/net_unet_ha_hs.py:658-1032; BASELINE configs[2] / [3]) assembled on
the B200 kernels.

What runs where (the hot path on libldic_b200, the family's other blocks as stock torch modules):

  libldic_b200   * the 5x5 stride-2 convs of g_a with their GDN fused in the epilogue, the stand-alone GDN layers
                   (model/gdn.py GDN after ResidualBlockWithStride; the CompressAI-style GDN inside it)
                 * WinBasedAttention at dim 192 (the four attention blocks of every Win_noShift_Attention)
                 * GaussianConditional: round(y - mu) + mu, erfc-form likelihood, sigma >= 0.11, L >= 1e-9, sum ln L
                 * g_s: the four 5x5 stride-2 transposed convs with IGDN fused; batch_conv + tanh + clamp + the 8-bit
                   level squared error fused into the last one's epilogue
                 * ste_round / bypass_round, the bpp / PSNR scalar tail
  stock torch    ResidualBottleneck, ResidualBlock(WithStride) 3x3 / 7x7 convs, SWAtten (Swin blocks of the slice
                 entropy model), the U-Net hyperprior Unet_ha_new / Unet_hs_new, the slice transforms, the syntax
                 model.  These are cuDNN / cuBLAS calls exactly as in the reference; they are not kernel targets of
                 this tier and bench.py reports their share of the step.

Same constructor, forward contract (test mode: ``(bpp, v_mse[B], v_psnr)``) and state-dict keys as the reference
(1189 keys besides the one-hot sampler buffers and the HAN head, which are accepted and dropped like in ``Net``).

PARITY NOTE ("restated deps"): the reference imports third-party ``compressai`` / ``timm`` classes and two modules
that are absent from its own tree.  The classes ResidualBlock, ResidualBlockWithStride,
AttentionBlock (base of SWAtten), EntropyBottleneck (medians only), GaussianConditional follow SURVEY Appendix B;
``DepthwiseSeparableConv`` is a guess from its call shape.  Fixtures (tests/golden/make_golden_unet.py) are produced
by the reference's own file running on the same restatements (oracle/unet_harness.py).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .layers import (GDN as CAGDN, GaussianConditional, GaussianModel, LowerBound, ModelGDN, ModelIGDN, WinBasedAttention)
from .net import PredictionModel_Context, _DISCARDED_PREFIXES, conv_generator
from .transforms import _PlannedTransform


# The 3x3 / 1x1 convolutions of ResidualBlock, ResidualBlockWithStride and Win_noShift_Attention (13 + 1 per attention
# module at dim 192: 70 % of the family's FLOPs) run on the tcgen05 conv kernel (bf16 operands, fp32 accumulation, the
# LeakyReLU / GDN in the epilogue) when the channel count is a multiple of 64.  False: stock torch convs (cuDNN), which
# the tests use as the cross-check of the kernel path.
KERNEL_CONVS = True


def _versions(*mods):
    return tuple((p._version, p.data_ptr()) for m in mods for p in m.parameters())


def _kernel_ok(x, *channels):
    return KERNEL_CONVS and x.is_cuda and all(c % 64 == 0 for c in channels)


def conv1x1(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=1, stride=stride)


def conv3x3(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=3, stride=stride, padding=1)


def conv5x5(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=5, stride=stride, padding=2)


def conv7x7(i, o, stride=1):
    return nn.Conv2d(i, o, kernel_size=7, stride=stride, padding=3)


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """model/net_unet_ha_hs.py:626-633."""
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


# ---- stock-torch blocks (restated third-party / reference classes; parameter containers + torch forward) ----------
class ResidualBottleneck(nn.Module):
    """model/net_unet_ha_hs.py:90-104, model/Block_unet.py:401-416."""

    def __init__(self, N=192, act=nn.GELU):
        super().__init__()
        self.branch = nn.Sequential(conv1x1(N, N // 2), act(), nn.Conv2d(N // 2, N // 2, 3, 1, 1), act(), conv1x1(N // 2, N))

    def forward(self, x):
        return x + self.branch(x)


class ResidualBlock(nn.Module):
    """CompressAI ResidualBlock: conv3x3 -> LeakyReLU -> conv3x3 -> LeakyReLU, + skip."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    _plan = _plan_key = None

    def _ldic(self):
        key = _versions(self.conv1, self.conv2)
        if self._plan is None or self._plan_key != key:
            mk = lambda c, f32: ops.ConvTC(_lib.LDIC_CONV_S1_3x3_P1, c.weight.detach(), c.bias.detach(),
                                           act=_lib.ACT_LEAKY001, out_f32=f32)
            self._plan, self._plan_key = (mk(self.conv1, False), mk(self.conv2, True), mk(self.conv2, False)), key
        return self._plan

    def forward_nhwc(self, t):
        """NHWC bf16 -> NHWC bf16; the residual is added in conv2's epilogue (fp32, before the bf16 rounding)."""
        c1, c2, c2b = self._ldic()
        return c2b(c1(t), residual=t)

    def forward(self, x):
        if self.skip is None and _kernel_ok(x, self.conv1.in_channels, self.conv1.out_channels):
            c1, c2, _ = self._ldic()
            o = c2(c1(ops.nchw_to_nhwc_bf16(x, c1.cin_pad)))                 # NHWC fp32
            return ops.residual_nhwc_to_nchw(o, x)                           # + identity, back to NCHW
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        return out + (x if self.skip is None else self.skip(x))


class ResidualBlockWithStride(nn.Module):
    """CompressAI ResidualBlockWithStride: conv3x3(s) -> LeakyReLU -> conv3x3 -> GDN, + conv1x1(s) skip.
    The GDN is the CompressAI-style class (layers/gdn.py) and runs on ldic_gdn_nchw_f32."""

    def __init__(self, in_ch, out_ch, stride=2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = CAGDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

    _plan = _plan_key = None

    def forward(self, x):
        h = self.leaky_relu(self.conv1(x))
        identity = x if self.skip is None else self.skip(x)
        if _kernel_ok(x, self.conv2.in_channels, self.conv2.out_channels):
            # conv2 + GDN in one tensor-core kernel (the GDN as the epilogue's gamma contraction)
            key = _versions(self.conv2, self.gdn)
            if self._plan is None or self._plan_key != key:
                g = self.gdn
                self._plan = ops.ConvTC(_lib.LDIC_CONV_S1_3x3_P1, self.conv2.weight.detach(), self.conv2.bias.detach(),
                                        act=_lib.ACT_GDN, out_f32=True, gdn=(g.beta.detach(), g.gamma.detach()) + g.constants())
                self._plan_key = key
            o = self._plan(ops.nchw_to_nhwc_bf16(h, self._plan.cin_pad))
            return ops.residual_nhwc_to_nchw(o, identity)
        return self.gdn(self.conv2(h)) + identity


class _ResidualUnit(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2), nn.ReLU(inplace=True),
                                  conv1x1(N // 2, N))
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.conv(x) + x)


class Win_noShift_Attention(nn.Module):
    """layers/layers.py:56-111: a = 3 x ResidualBlock(x); b = [WinBasedAttention, conv1x1, WinBasedAttention, ResidualBlock,
    conv3x3, WinBasedAttention, ResidualBlock, conv7x7, WinBasedAttention, ResidualBlock](x); out = a * sigmoid(b) + x.
    The four WinBasedAttention blocks run on the window-attention kernel when the shape is one it supports."""

    def __init__(self, dim, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        N = dim
        wa = lambda: WinBasedAttention(dim=dim, num_heads=num_heads, window_size=window_size, shift_size=shift_size)
        self.conv_a = nn.Sequential(ResidualBlock(N, N), ResidualBlock(N, N), ResidualBlock(N, N))
        self.conv_b = nn.Sequential(wa(), conv1x1(N, N), wa(), ResidualBlock(N, N), conv3x3(N, N), wa(), ResidualBlock(N, N),
                                    conv7x7(N, N), wa(), ResidualBlock(N, N))

    _plan = _plan_key = None

    def _singles(self):
        """conv_b[1] / conv_b[4] (a lone 1x1 / 3x3 conv, no activation) as kernel layers with bf16 output."""
        key = _versions(self.conv_b[1], self.conv_b[4])
        if self._plan is None or self._plan_key != key:
            mk = lambda c, k: ops.ConvTC(k, c.weight.detach(), c.bias.detach(), out_f32=False)
            self._plan = {1: mk(self.conv_b[1], _lib.LDIC_CONV_1x1), 4: mk(self.conv_b[4], _lib.LDIC_CONV_S1_3x3_P1)}
            self._plan_key = key
        return self._plan

    def forward(self, x):
        C = x.shape[1]
        if not (_kernel_ok(x, C) and all(m.kernel_shape() for m in self.conv_b if isinstance(m, WinBasedAttention))):
            return self.conv_a(x) * torch.sigmoid(self.conv_b(x)) + x
        # Kernel path: the whole module runs on NHWC bf16 tensors -- one layout change in, one (fused with the gate and
        # the outer residual) out; every inner residual is added in a conv epilogue.  Only the 7x7 conv goes through
        # cuDNN (its 49 taps exceed the tap table of the conv kernel's producer warps).
        xb = ops.nchw_to_nhwc_bf16(x, C)
        a = xb
        for rb in self.conv_a:
            a = rb.forward_nhwc(a)
        singles = self._singles()
        b = xb
        for i, m in enumerate(self.conv_b):
            if i in singles:
                b = singles[i](b)
            elif isinstance(m, (WinBasedAttention, ResidualBlock)):
                b = m.forward_nhwc(b)
            else:                               # conv7x7
                b = ops.nchw_to_nhwc_bf16(m(ops.nhwc_to_nchw_f32(b, C)), C)
        return ops.gate_residual_nhwc_to_nchw(a, b, x)


class WMSA(nn.Module):
    """model/Block_unet.py:170-252 (window multi-head self-attention of the TCM Swin block), torch ops."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, type):
        super().__init__()
        self.input_dim, self.output_dim, self.head_dim = input_dim, output_dim, head_dim
        self.scale = head_dim ** -0.5
        self.n_heads = input_dim // head_dim
        self.window_size, self.type = window_size, type
        self.embedding_layer = nn.Linear(input_dim, 3 * input_dim, bias=True)
        self.relative_position_params = nn.Parameter(torch.zeros(self.n_heads, 2 * window_size - 1, 2 * window_size - 1))
        nn.init.trunc_normal_(self.relative_position_params, std=.02)
        self.linear = nn.Linear(input_dim, output_dim)
        cord = torch.tensor([[i, j] for i in range(window_size) for j in range(window_size)])
        self.register_buffer("_relation", cord[:, None, :] - cord[None, :, :] + window_size - 1, persistent=False)

    def _mask(self, hw, ww, p, shift, device):
        m = torch.zeros(hw, ww, p, p, p, p, dtype=torch.bool, device=device)
        s = p - shift
        m[-1, :, :s, :, s:, :] = True
        m[-1, :, s:, :, :s, :] = True
        m[:, -1, :, :s, :, s:] = True
        m[:, -1, :, s:, :, :s] = True
        return m.reshape(1, 1, hw * ww, p * p, p * p)

    def forward(self, x):                       # x: (b, h, w, c)
        p = self.window_size
        if self.type != 'W':
            x = torch.roll(x, shifts=(-(p // 2), -(p // 2)), dims=(1, 2))
        b, H, W, c = x.shape
        hw, ww = H // p, W // p
        x = x.reshape(b, hw, p, ww, p, c).permute(0, 1, 3, 2, 4, 5).reshape(b, hw * ww, p * p, c)
        qkv = self.embedding_layer(x).reshape(b, hw * ww, p * p, 3 * self.n_heads, self.head_dim).permute(3, 0, 1, 2, 4)
        q, k, v = qkv.chunk(3, dim=0)
        sim = torch.einsum('hbwpc,hbwqc->hbwpq', q, k) * self.scale
        rel = self.relative_position_params[:, self._relation[:, :, 0].long(), self._relation[:, :, 1].long()]
        sim = sim + rel[:, None, None]
        if self.type != 'W':
            sim = sim.masked_fill_(self._mask(hw, ww, p, p // 2, x.device), float("-inf"))
        out = torch.einsum('hbwij,hbwjc->hbwic', F.softmax(sim, dim=-1), v)
        out = out.permute(1, 2, 3, 0, 4).reshape(b, hw * ww, p * p, self.n_heads * self.head_dim)
        out = self.linear(out)
        out = out.reshape(b, hw, ww, p, p, -1).permute(0, 1, 3, 2, 4, 5).reshape(b, H, W, -1)
        if self.type != 'W':
            out = torch.roll(out, shifts=(p // 2, p // 2), dims=(1, 2))
        return out


class Block_1(nn.Module):
    """model/net_unet_ha_hs.py:107-129."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path, type='W', input_resolution=None):
        super().__init__()
        self.ln1 = nn.LayerNorm(input_dim)
        self.msa = WMSA(input_dim, input_dim, head_dim, window_size, type)
        self.ln2 = nn.LayerNorm(input_dim)
        self.mlp = nn.Sequential(nn.Linear(input_dim, 4 * input_dim), nn.GELU(), nn.Linear(4 * input_dim, output_dim))

    def forward(self, x):
        x = x + self.msa(self.ln1(x))
        return x + self.mlp(self.ln2(x))


class SwinBlock(nn.Module):
    """model/net_unet_ha_hs.py:132-151."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path):
        super().__init__()
        self.block_1 = Block_1(input_dim, output_dim, head_dim, window_size, drop_path, type='W')
        self.block_2 = Block_1(input_dim, output_dim, head_dim, window_size, drop_path, type='SW')
        self.window_size = window_size

    def forward(self, x):
        if x.size(-1) <= self.window_size or x.size(-2) <= self.window_size:
            # the reference pads here and then multiplies tensors of different sizes (:139-150 with SWAtten.forward): it
            # cannot run on latents this small, and neither do we
            raise ops.LdicError("SwinBlock: latent must be larger than the 8x8 window in both dimensions (image >= 144 px)")
        t = x.permute(0, 2, 3, 1)
        t = self.block_2(self.block_1(t))
        return t.permute(0, 3, 1, 2)


class SWAtten(nn.Module):
    """model/net_unet_ha_hs.py:154-175 on the restated CompressAI AttentionBlock (conv_a / conv_b of ResidualUnits)."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path, inter_dim=192):
        super().__init__()
        N = inter_dim if inter_dim is not None else input_dim
        self.conv_a = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N))
        self.conv_b = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N), conv1x1(N, N))
        self.non_local_block = SwinBlock(N, N, head_dim, window_size, drop_path)
        if inter_dim is not None:
            self.in_conv = conv1x1(input_dim, inter_dim)
            self.out_conv = conv1x1(inter_dim, output_dim)

    def forward(self, x):
        x = self.in_conv(x)
        identity = x
        z = self.non_local_block(x)
        out = self.conv_a(x) * torch.sigmoid(self.conv_b(z)) + identity
        return self.out_conv(out)


class ResidualBlock3_5(nn.Module):
    """model/Block_unet.py:295-332."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv5x5(out_ch, out_ch)
        self.conv3 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        out = self.leaky_relu(self.conv3(self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))))
        return out + (x if self.skip is None else self.skip(x))


class ResidualBlock5x5(nn.Module):
    """model/Block_unet.py:335-364 (only conv2 is used by its forward; conv1 / conv3 are parameters of the checkpoint)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv5x5(out_ch, out_ch)
        self.conv3 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(x))
        return out + (x if self.skip is None else self.skip(x))


class ResidualBlock3x3(nn.Module):
    """model/Block_unet.py:367-398."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv3 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x):
        out = self.leaky_relu(self.conv3(self.leaky_relu(self.conv1(x))))
        return out + (x if self.skip is None else self.skip(x))


class Unet_ha_new(nn.Module):
    """model/Block_unet.py:774-838 (hyper analysis of the U-Net family; its small WinBasedAttention blocks -- dims 96 /
    128 / 512, windows 4 / 2 -- take the torch path of WinBasedAttention)."""

    def __init__(self, inchannels, num_heads, depth):
        super().__init__()
        self.inchannels, self.num_heads, self.depth = inchannels, num_heads, depth
        self.SpatialTransformer1 = WinBasedAttention(inchannels // 2, num_heads, window_size=4, shift_size=2)
        self.ResBlock1 = ResidualBottleneck(96)
        self.SpatialTransformer2 = WinBasedAttention(128, num_heads, window_size=4, shift_size=2)
        self.ResBlock2 = ResidualBottleneck(128)
        self.ResBlock3 = ResidualBottleneck(256)
        self.conv1 = ResidualBlock3_5(inchannels // 2, inchannels // 2)
        self.conv2 = ResidualBlock5x5(128, 128)
        self.down0 = nn.Conv2d(inchannels, inchannels, 1, 1, 0, bias=True)
        self.down1 = nn.Conv2d(inchannels, 256, kernel_size=3, stride=2, padding=1)
        self.down2 = nn.Conv2d(256, 512, kernel_size=3, stride=2, padding=1)
        self.down3 = nn.Conv2d(256, 256, 1, 1, 0, bias=True)
        self.middle = nn.Sequential(ResidualBottleneck(512), WinBasedAttention(512, num_heads, window_size=2, shift_size=1),
                                    ResidualBottleneck(512))
        self.relu = nn.GELU()

    def forward(self, x):
        trans_down_x, conv_down_x = torch.split(x, (x.shape[1] // 2, x.shape[1] // 2), dim=1)
        down_x1 = self.down0(torch.cat((self.conv1(conv_down_x), self.SpatialTransformer1(trans_down_x)), dim=1)) + x
        down_x1 = self.relu(self.down1(down_x1))
        conv_down_y, trans_down_y = torch.split(down_x1, (down_x1.shape[1] // 2, down_x1.shape[1] // 2), dim=1)
        down_x2 = self.down3(torch.cat((self.conv2(conv_down_y), self.SpatialTransformer2(trans_down_y)), dim=1)) + down_x1
        down_x2 = self.relu(self.down2(down_x2))
        middle_x = self.middle(down_x2)
        return middle_x, middle_x, down_x1, x


class Unet_hs_new(nn.Module):
    """model/Block_unet.py:841-890.  NB: the reference's forward never reads its first argument (the quantised z)."""

    def __init__(self, out_channels, num_heads, depth):
        super().__init__()
        self.out_channels, self.num_heads, self.depth = out_channels, num_heads, depth
        self.SpatialTransformer2 = WinBasedAttention(128, num_heads, window_size=2, shift_size=1)
        self.SpatialTransformer3 = WinBasedAttention(256, num_heads, window_size=2, shift_size=1)
        self.up0 = nn.Conv2d(512, 512, 1, 1, 0, bias=True)
        self.up1 = nn.ConvTranspose2d(512, 256, 5, 2, 2, output_padding=1, bias=True)
        self.up2 = nn.ConvTranspose2d(256, 192, 5, 2, 2, output_padding=1, bias=True)
        self.up3 = nn.ConvTranspose2d(512, 256, 1, 1, 0, bias=True)
        self.up4 = nn.ConvTranspose2d(384, out_channels, 1, 1, 0, bias=True)
        self.up5 = nn.Conv2d(256, 256, 1, 1, 0, bias=True)
        self.conv3 = ResidualBlock3x3(256, 256)
        self.conv4 = ResidualBlock3x3(128, 128)
        self.relu = nn.GELU()

    def forward(self, x, middle_x, down_x1, input):
        trans_up_x, conv_up_x = torch.split(middle_x, (middle_x.shape[1] // 2, middle_x.shape[1] // 2), dim=1)
        up_x1 = self.up0(torch.cat((self.conv3(conv_up_x), self.SpatialTransformer3(trans_up_x)), dim=1)) + middle_x
        up_x1 = self.relu(self.up1(up_x1))
        up_x1 = self.relu(self.up3(torch.cat((up_x1, down_x1), dim=1)))
        conv_up_y, trans_up_y = torch.split(up_x1, (up_x1.shape[1] // 2, up_x1.shape[1] // 2), dim=1)
        up_x2 = self.up5(torch.cat((self.conv4(conv_up_y), self.SpatialTransformer2(trans_up_y)), dim=1)) + up_x1
        up_x2 = self.relu(self.up2(up_x2))
        return self.up4(torch.cat((up_x2, input), dim=1))


class DepthwiseSeparableConv(nn.Module):
    """GUESSED: the reference imports it from a file that is not in its tree (model/net_unet_ha_hs.py:45); depthwise 3x3
    + pointwise 1x1 is the textbook block of that name and fits the call shape (:536-542)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.depthwise = nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding, groups=in_channels)
        self.pointwise = nn.Conv2d(in_channels, out_channels, 1)

    def forward(self, x):
        return self.pointwise(self.depthwise(x))


class Syntax_Model(nn.Module):
    """model/net_unet_ha_hs.py:533-570."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.Depth_down0 = DepthwiseSeparableConv(in_channels=16, out_channels=16)
        self.down0 = nn.Conv2d(in_dim, 32, 3, 2, 1)
        self.Depth_down1 = DepthwiseSeparableConv(in_channels=32, out_channels=32)
        self.down1 = nn.Conv2d(32, 64, 3, 2, 1)
        self.Depth_down2 = DepthwiseSeparableConv(in_channels=64, out_channels=64)
        self.down2 = nn.Conv2d(64, 128, 3, 2, 1)
        self.WAM = Win_noShift_Attention(dim=64, num_heads=8, window_size=4, shift_size=2)
        self.conv = nn.Conv2d(in_dim + 32 + 64 + 128, out_dim, 1, 1, 0)
        self.pooling = nn.AdaptiveAvgPool2d(1)

    def forward(self, syntax):
        out1 = self.pooling(syntax)
        ds1 = F.relu(self.down0(self.Depth_down0(syntax)))
        out2 = self.pooling(ds1)
        ds2 = self.WAM(F.relu(self.down1(self.Depth_down1(ds1))))
        out3 = self.pooling(ds2)
        ds3 = F.relu(self.down2(self.Depth_down2(ds2)))
        out4 = self.pooling(ds3)
        return self.conv(torch.cat((out1, out2, out3, out4), 1))


class PredictionModel_Syntax(nn.Module):
    """model/net_unet_ha_hs.py:573-610 (held for checkpoint compatibility: the U-Net forward never calls it)."""

    def __init__(self, in_dim, dim=192, trainable=True, outdim=None):
        super().__init__()
        outdim = dim if outdim is None else outdim
        self.down0 = nn.Conv2d(in_dim, dim, 3, 2, 1)
        self.down1 = nn.Conv2d(dim, dim, 3, 2, 1)
        self.pooling = nn.AdaptiveAvgPool2d(1)
        self.WAM = Win_noShift_Attention(dim=dim, num_heads=8, window_size=4, shift_size=2)
        self.fc = nn.Linear(dim * 2 + in_dim, outdim)
        self.flatten = nn.Flatten()


class EntropyBottleneck(nn.Module):
    """CompressAI EntropyBottleneck reduced to what the reference observes: `_get_medians()` (model/net_unet_ha_hs.py:885);
    its likelihoods are computed and discarded (:882).  Only the `quantiles` parameter is kept."""

    def __init__(self, channels):
        super().__init__()
        self.quantiles = nn.Parameter(torch.tensor([-10.0, 0.0, 10.0]).repeat(channels, 1, 1))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()


# ---- transforms on the kernels ------------------------------------------------------------------------------------
class analysisTransformModel(_PlannedTransform):
    """model/net_unet_ha_hs.py:197-232.  Indices of `transform` as in the reference; the two 5x5 stride-2 convs
    (6, 15) run on the tcgen05 kernel, 6 with its GDN (7) fused."""

    def __init__(self, in_dim, num_filters, conv_trainable=True):
        super().__init__()
        f = num_filters
        self.transform = nn.Sequential(
            ResidualBottleneck(in_dim), ResidualBottleneck(in_dim), ResidualBottleneck(in_dim),
            ResidualBlockWithStride(in_dim, f[0], stride=2), ModelGDN(f[0]),
            nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(f[0], f[1], 5, 2, 0), ModelGDN(f[1]),
            Win_noShift_Attention(dim=f[1], num_heads=8, window_size=8, shift_size=4),
            ResidualBottleneck(f[1]), ResidualBottleneck(f[1]), ResidualBottleneck(f[1]),
            ResidualBlockWithStride(f[1], f[2], 2), ModelGDN(f[2]),
            nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(f[2], f[3], 5, 2, 0),
            Win_noShift_Attention(dim=f[3], num_heads=8, window_size=4, shift_size=2))

    def _build_plan(self):
        t = self.transform
        dev = t[6].weight.device

        def gdn_only(g):
            """A stand-alone GDN as a 1x1 conv with the identity matrix + the fused GDN epilogue: the channel contraction
            runs on the tensor cores and the result leaves as the NHWC bf16 tensor the next conv reads (the fp32
            CUDA-core kernel behind ModelGDN.forward costs ~4 ms per call at 1/2 resolution and batch 16)."""
            C = g.beta.numel()
            return ops.ConvTC(_lib.LDIC_CONV_1x1, torch.eye(C, device=dev), torch.zeros(C, device=dev), act=_lib.ACT_GDN,
                              out_f32=False, gdn=self._gdn_args(g))
        return [ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[6].weight.detach(), t[6].bias.detach(), act=_lib.ACT_GDN,
                           out_f32=True, gdn=self._gdn_args(t[7])),
                ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[15].weight.detach(), t[15].bias.detach(), out_f32=True),
                gdn_only(t[4]), gdn_only(t[13])]

    def forward(self, x):
        t, L = self.transform, self.plan()
        x = t[3](t[2](t[1](t[0](x))))
        if _kernel_ok(x, x.shape[1]):
            g = L[2](ops.nchw_to_nhwc_bf16(x, L[2].cin_pad))                                         # 4 (GDN) -> NHWC bf16
        else:
            g = ops.nchw_to_nhwc_bf16(t[4](x), L[0].cin_pad)
        x = ops.nhwc_to_nchw_f32(L[0](g), t[6].out_channels)                                         # 5 + 6 + 7
        x = t[12](t[11](t[10](t[9](t[8](x)))))
        if _kernel_ok(x, x.shape[1]):
            g = L[3](ops.nchw_to_nhwc_bf16(x, L[3].cin_pad))                                         # 13 (GDN)
        else:
            g = ops.nchw_to_nhwc_bf16(t[13](x), L[1].cin_pad)
        x = ops.nhwc_to_nchw_f32(L[1](g), t[15].out_channels)                                        # 14 + 15
        return t[16](x)


class synthesisTransformModel(_PlannedTransform):
    """model/net_unet_ha_hs.py:287-326: Win_noShift_Attention, 2 x [pad + ConvT5 s2 + IGDN], Win_noShift_Attention,
    2 x [pad + ConvT5 s2 + IGDN].  All four deconvs + IGDN on the tcgen05 kernel; the last one carries the fused tail."""

    def __init__(self, in_dim, num_filters, conv_trainable=True):
        super().__init__()
        f = num_filters
        dc = lambda i, o: nn.ConvTranspose2d(i, o, 5, 2, 3, output_padding=1)
        self.transform = nn.Sequential(
            Win_noShift_Attention(dim=in_dim, num_heads=8, window_size=4, shift_size=2),
            nn.ZeroPad2d((1, 0, 1, 0)), dc(in_dim, f[0]), ModelIGDN(f[0], inverse=True),
            nn.ZeroPad2d((1, 0, 1, 0)), dc(f[0], f[1]), ModelIGDN(f[1], inverse=True),
            Win_noShift_Attention(dim=f[1], num_heads=8, window_size=8, shift_size=2),
            nn.ZeroPad2d((1, 0, 1, 0)), dc(f[1], f[2]), ModelIGDN(f[2], inverse=True),
            nn.ZeroPad2d((1, 0, 1, 0)), dc(f[2], f[3]), ModelIGDN(f[3], inverse=True))

    def _build_plan(self):
        t = self.transform
        mk = lambda ci, gi, kind, f32: ops.ConvTC(kind, t[ci].weight.detach(), t[ci].bias.detach(), act=_lib.ACT_IGDN,
                                                 out_f32=f32, gdn=self._gdn_args(t[gi]))
        last_small = t[12].out_channels * 4 <= 256 and t[12].out_channels % 16 == 0
        return [mk(2, 3, _lib.LDIC_DECONV_GS_5x5, False), mk(5, 6, _lib.LDIC_DECONV_GS_5x5, True),
                mk(9, 10, _lib.LDIC_DECONV_GS_5x5, False),
                mk(12, 13, _lib.LDIC_DECONV_GS_5x5_MERGED if last_small else _lib.LDIC_DECONV_GS_5x5, True)]

    def body(self, y_hat):
        """Everything up to the input of the last deconv: NHWC bf16."""
        t, L = self.transform, self.plan()
        x = t[0](y_hat)
        x = L[1](L[0](ops.nchw_to_nhwc_bf16(x, L[0].cin_pad)))                     # 1..6 -> NHWC fp32
        x = t[7](ops.nhwc_to_nchw_f32(x, t[5].out_channels))
        return L[2](ops.nchw_to_nhwc_bf16(x, L[2].cin_pad))                         # 8..10

    def forward(self, y_hat):
        return ops.nhwc_to_nchw_f32(self.plan()[3](self.body(y_hat)), self.transform[12].out_channels)


class Net(nn.Module):
    """model/net_unet_ha_hs.py:658-1032."""

    def __init__(self, train_size, test_size, is_high, post_processing):
        super().__init__()
        if post_processing:
            raise NotImplementedError("HAN post-processing (model/han.py) is outside the rate-distortion forward path")
        self.num_slices, self.max_support_slices = 4, 4
        self.torch_tf32_matmul = True     # see rd_forward
        self.gaussian_conditional = GaussianConditional(None)
        self.gaussian_conditional.lower_bound_scale = LowerBound(0.11)          # state-dict keys of the CompressAI class
        self.gaussian_conditional.likelihood_lower_bound = LowerBound(1e-9)
        self.train_size, self.test_size = train_size, test_size
        self.post_processing, self.is_high = post_processing, is_high
        N, M = (384, 32) if is_high else (192, 16)
        if is_high:
            raise NotImplementedError("the reference hard-codes 192 channels in the U-Net family's hyperprior and slice "
                                      "transforms (model/net_unet_ha_hs.py:723-790): is_high cannot be constructed there either")
        self.M, self.N = M, N
        self.conv_1 = conv1x1(192, 4)
        self.conv_2 = conv1x1(4, 192)
        self.a_model = analysisTransformModel(3, [N, N, N, N])
        self.s_model = synthesisTransformModel(N, [N, N, N, M])
        self.syntax_model = Syntax_Model(M, M)
        self.conv_weights_gen = conv_generator(in_dim=M, out_dim=M)
        self.h_a = Unet_ha_new(192, 8, 3)
        self.h_s = Unet_hs_new(192, 8, 3)
        self.entropy_bottleneck_z2 = GaussianModel()
        self.entropy_bottleneck_z3 = GaussianModel()
        self.entropy_bottleneck = EntropyBottleneck(512)
        self.entropy_bottleneck_z3_syntax = GaussianModel()
        self.window_size = 8
        S = self.num_slices
        cin = lambda i: 192 + (192 // S) * min(i, 4)
        self.atten_mean = nn.ModuleList(nn.Sequential(SWAtten(cin(i), cin(i), 16, self.window_size, 0, inter_dim=128))
                                        for i in range(S))
        tr = lambda c0: nn.Sequential(conv(c0, 224, stride=1, kernel_size=3), nn.GELU(), conv(224, 128, stride=1, kernel_size=3),
                                      nn.GELU(), conv(128, 192 // S, stride=1, kernel_size=3))
        self.cc_mean_transforms = nn.ModuleList(tr(cin(i)) for i in range(S))
        self.atten_scale = nn.ModuleList(nn.Sequential(SWAtten(cin(i), cin(i), 16, self.window_size, 0, inter_dim=128))
                                         for i in range(S))
        self.cc_scale_transforms = nn.ModuleList(tr(cin(i)) for i in range(S))
        self.lrp_transforms = nn.ModuleList(tr(192 + (192 // S) * min(i + 1, 5)) for i in range(S))
        self.v_z2_sigma = nn.Parameter(torch.ones((1, N, 1, 1), dtype=torch.float32, requires_grad=True))
        self.register_parameter('z2_sigma', self.v_z2_sigma)
        self.prediction_model = PredictionModel_Context(in_dim=2 * N - M, dim=N, outdim=(N - M) * 2)
        self.prediction_model_syntax = PredictionModel_Syntax(in_dim=N, dim=M, outdim=M * 2)
        self.tail_fused = True

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        kept = {k: v for k, v in state_dict.items()
                if not (k.endswith("_sampler.sample_filter") or k.startswith(_DISCARDED_PREFIXES))}
        return super().load_state_dict(kept, strict=strict, assign=assign)

    @torch.no_grad()
    def rd_forward(self, inputs: torch.Tensor, want_x_hat: bool = False, want_bitstreams: bool = False) -> Dict[str, torch.Tensor]:
        """`want_bitstreams`: also rANS-code every slice's symbols round(y - mu) under N(0, max(scale, 0.11)) -- the very
        quantiser and model of the GaussianConditional call (:937) -- into one bitstream per image and slice
        (out["streams"][i], out["slice_params"][i] = (mu, scale) for ops.rans_decode).  The reference only estimates this
        rate; its z never reaches a likelihood (:882-885) and its h_s consumes encoder-side skip tensors (:892), so the
        family has no self-contained decoder and only the slice streams are coded."""
        if not inputs.is_cuda:
            raise ops.LdicError("Net runs on CUDA only (no CPU fallback)")
        # the stock torch blocks: cuDNN convolutions already run TF32 by default; `torch_tf32_matmul` puts the fp32
        # nn.Linear layers of SWAtten (cuBLAS SIMT sgemm, 5.6 of 59 ms at batch 16) on the same footing for this call
        prev = torch.backends.cuda.matmul.allow_tf32
        if self.torch_tf32_matmul:
            torch.backends.cuda.matmul.allow_tf32 = True
        try:
            with torch.cuda.device(inputs.device):
                return self._rd_forward(inputs.contiguous().float(), want_x_hat, want_bitstreams)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def _rd_forward(self, x, want_x_hat, want_bitstreams=False):
        B, _, H, W = x.shape
        if H % 256 or W % 256:
            # 16x down to y, 4x more inside the syntax model whose Win_noShift_Attention uses 4x4 windows
            # (model/net_unet_ha_hs.py:541-561): the reference itself fails on other sizes
            raise ops.LdicError("U-Net family: H and W must be multiples of 256")
        M, S = self.M, self.num_slices
        z3 = self.a_model(x)                                                        # :875
        y_shape = z3.shape[2:]
        z, middle_x, down_x1, inp = self.h_a(z3)                                    # :880
        z_offset = self.entropy_bottleneck._get_medians()                           # :885 (likelihoods of :882 are never used)
        z_hat = torch.ops.ldic.round_ste(z - z_offset) + z_offset                   # :889
        latent_scales = self.h_s(z_hat, middle_x, down_x1, inp)                     # :892
        latent_means = latent_scales                                                # :895 calls the same module on the same inputs
        z3_syntax = self.syntax_model(z3[:, :M])                                    # :904
        z3_syntax_rounded = torch.ops.ldic.round_ste(z3_syntax.contiguous())        # :905
        y_slices = z3.chunk(S, 1)
        y_hat_slices = []
        bits = torch.zeros(S, dtype=torch.float32, device=x.device)
        streams, slice_params, dequantised = [], [], []
        for i, y_slice in enumerate(y_slices):                                      # :916-950
            support = y_hat_slices[:self.max_support_slices]
            mean_support = self.atten_mean[i](torch.cat([latent_means] + support, dim=1))
            mu = self.cc_mean_transforms[i](mean_support)[:, :, :y_shape[0], :y_shape[1]]
            scale_support = self.atten_scale[i](torch.cat([latent_scales] + support, dim=1))
            scale = self.cc_scale_transforms[i](scale_support)[:, :, :y_shape[0], :y_shape[1]]
            # :937 gaussian_conditional + :941 ste_round(y - mu) + mu: one kernel (quant "dequantize", erfc form, sum ln L)
            y_hat_slice, _, _ = ops.gaussian_likelihood(y_slice.contiguous(), scale.contiguous(), mu.contiguous(),
                                                        quant=ops.QUANT_DEQUANT, form=ops.FORM_GAUSSIAN_CONDITIONAL,
                                                        lik_bound=self.gaussian_conditional.likelihood_bound,
                                                        scale_bound=self.gaussian_conditional.scale_bound,
                                                        want_lik=False, want_vhat=True, sum_out=bits[i:i + 1])
            if want_bitstreams:
                mu_c, sc_c = mu.contiguous(), scale.contiguous()
                streams.append(ops.rans_encode(y_slice.contiguous(), sc_c, mu_c, quant=ops.QUANT_DEQUANT,
                                               scale_bound=self.gaussian_conditional.scale_bound))
                slice_params.append((mu_c, sc_c))
                dequantised.append(y_hat_slice)
            lrp = self.lrp_transforms[i](torch.cat([mean_support, y_hat_slice], dim=1))
            y_hat_slices.append(y_hat_slice + 0.5 * torch.tanh(lrp))                # :947-948
        y_hat = torch.cat(y_hat_slices, dim=1)
        conv_w = self.conv_weights_gen(z3_syntax_rounded).reshape(B, 3, M).contiguous()    # :970
        body = self.s_model.body(y_hat)                                             # :966
        last = self.s_model.plan()[3]
        if self.tail_fused and last.kind == _lib.LDIC_DECONV_GS_5x5_MERGED:
            # :971-980 batch_conv + tanh, :1006 clamp (no-op after tanh), :1025-1028 level error: in the deconv's epilogue
            sq_err, x_hat, _ = last.fused_tail(body, x, conv_w, want_x_tilde=want_x_hat, tanh_out=True)
        else:
            sq_err, x_hat = ops.syntax_conv_mse(x, last(body), conv_w, want_x_tilde=want_x_hat, tanh_out=True)
        out = {"bits": bits, "sq_err": sq_err, "latents": {"y": z3, "z": z, "y_hat": y_hat, "z3_syntax": z3_syntax}}
        if want_x_hat:
            out["x_hat"] = x_hat
        if want_bitstreams:
            out["streams"], out["slice_params"], out["slice_symbols"] = streams, slice_params, dequantised
        return out

    def metrics(self, out, H: int, W: int):
        """:992-1030: bpp over the TRAIN size (h, w) even in test mode (reference quirk H2), y likelihoods only."""
        _, h, w, _ = self.train_size
        bits3 = torch.stack([out["bits"].sum(), out["bits"].new_zeros(()), out["bits"].new_zeros(())])
        packed, v_mse = ops.rd_pack_metrics(bits3.contiguous(), out["sq_err"], 3 * H * W)
        r = ops.rd_finish_metrics(packed, float(h * w))
        return r[0], v_mse, r[1]

    def forward(self, inputs, mode='train', num=1):
        if mode != 'test':
            raise NotImplementedError("only the rate-distortion forward (mode='test') is implemented; training is out of scope")
        out = self.rd_forward(inputs)
        return self.metrics(out, inputs.shape[2], inputs.shape[3])

"""g_a / g_s / h_a / h_s of the reference (model/net.py:91-216) on the tcgen05 conv kernels.

The classes keep the reference's names, constructor arguments and state-dict
keys (``transform.<idx>.{weight,bias,beta,gamma,...}``): the nn.Conv2d /
nn.ConvTranspose2d / GDN children are parameter containers only; ``forward``
runs a plan of ``ops.ConvTC`` layers (weights pre-packed to bf16 once per
weight version) on NHWC bf16 activations.

``forward(x)`` accepts and returns NCHW fp32 like the reference modules;
``forward_nhwc`` is the copy-free entry used by ``Net``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .layers import ModelGDN, ModelIGDN


def _versions(module: nn.Module):
    return tuple((p._version, p.data_ptr()) for p in module.parameters())


class _PlannedTransform(nn.Module):
    _plan = None
    _plan_key = None
    precision = "bf16"      # "tf32": TF32 parity mode of the analysis-side transforms (g_a, h_a): fp32 activations, kind::tf32

    def plan(self):
        key = _versions(self) + (self.precision,)
        if self._plan is None or self._plan_key != key:
            with torch.no_grad():
                self._plan = self._build_plan()
            self._plan_key = key
        return self._plan

    def _gdn_args(self, g):
        return (g.beta.detach(), g.gamma.detach()) + g.constants()


class analysisTransformModel(_PlannedTransform):
    """model/net.py:91-118.  4x [ZeroPad2d((1,2,1,2)) -> Conv2d(k5,s2)] with GDN after convs 1-3.

    First layer (Cin = 3): up to 192 output channels run on conv_first_kernel (patches built from the NCHW image in
    shared memory, GDN fused, weights and gamma resident); wider models (N = 384) build a bf16 patch matrix with
    ldic_im2col_5x5s2 and run the 1x1 GEMM + GDN on conv_wide_kernel.  `first_fused = False` forces the patch-matrix
    form (tests cross-check the two)."""
    first_fused = True

    def __init__(self, in_dim, num_filters, conv_trainable=True):
        super().__init__()
        f = num_filters
        self.in_dim = in_dim
        self.transform = nn.Sequential(
            nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(in_dim, f[0], 5, 2, 0), ModelGDN(f[0]),
            nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(f[0], f[1], 5, 2, 0), ModelGDN(f[1]),
            nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(f[1], f[2], 5, 2, 0), ModelGDN(f[2]),
            nn.ZeroPad2d((1, 2, 1, 2)), nn.Conv2d(f[2], f[3], 5, 2, 0),
        )

    def _build_plan(self):
        t = self.transform
        layers = []
        c1 = t[1]
        self._kp1 = ops._pad64(25 * c1.in_channels)
        self._first_fused = False
        if self.precision == "tf32":
            if c1.in_channels * 25 > 128:
                raise ops.LdicError("TF32 mode: the first layer goes through the 128-column patch matrix (Cin <= 5)")
            pr = dict(precision="tf32")
            w1 = c1.weight.detach().permute(0, 2, 3, 1).reshape(c1.out_channels, -1).contiguous()
            layers.append(ops.ConvTC(_lib.LDIC_CONV_1x1, w1, c1.bias.detach(), act=_lib.ACT_GDN, cin_pad=self._kp1,
                                     gdn=self._gdn_args(t[2]), **pr))
            self._first_im2col = True
            for ci, gi in ((4, 5), (7, 8)):
                layers.append(ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[ci].weight.detach(), t[ci].bias.detach(),
                                         act=_lib.ACT_GDN, gdn=self._gdn_args(t[gi]), **pr))
            layers.append(ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[10].weight.detach(), t[10].bias.detach(),
                                     act=_lib.ACT_NONE, precision="tf32_last"))
            return layers
        if c1.in_channels == 3 and c1.out_channels in (64, 128, 192) and self.first_fused:
            # first layer + GDN straight from the NCHW fp32 image (no patch matrix in HBM)
            layers.append(ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, c1.weight.detach(), c1.bias.detach(),
                                     act=_lib.ACT_GDN, gdn=self._gdn_args(t[2])))
            self._first_im2col = False
            self._first_fused = True
        elif c1.in_channels * 25 <= 128:
            # first layer: patch matrix (im2col kernel) + 1x1 GEMM; k = (ky*5+kx)*Cin + ci
            w1 = c1.weight.detach().permute(0, 2, 3, 1).reshape(c1.out_channels, -1).contiguous()
            layers.append(ops.ConvTC(_lib.LDIC_CONV_1x1, w1, c1.bias.detach(), act=_lib.ACT_GDN,
                                     cin_pad=self._kp1, gdn=self._gdn_args(t[2])))
            self._first_im2col = True
        else:
            layers.append(ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, c1.weight.detach(), c1.bias.detach(),
                                     act=_lib.ACT_GDN, gdn=self._gdn_args(t[2])))
            self._first_im2col = False
        layers.append(ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[4].weight.detach(), t[4].bias.detach(),
                                 act=_lib.ACT_GDN, gdn=self._gdn_args(t[5])))
        layers.append(ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[7].weight.detach(), t[7].bias.detach(),
                                 act=_lib.ACT_GDN, gdn=self._gdn_args(t[8])))
        layers.append(ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, t[10].weight.detach(), t[10].bias.detach(),
                                 act=_lib.ACT_NONE, out_f32=True))
        return layers

    def forward_nhwc(self, x_nchw: torch.Tensor) -> torch.Tensor:
        """x (B,Cin,H,W) NCHW, fp32 in [-1,1] or uint8 levels (mapped like eval_net.py:84) -> y (B,H/16,W/16,C) fp32 NHWC."""
        L = self.plan()
        B, _, H, W = x_nchw.shape
        if H % 16 or W % 16:
            raise ops.LdicError("analysis transform needs H and W to be multiples of 16")
        if x_nchw.dtype == torch.uint8 and not (self._first_fused and W % 16 == 0):   # (also the TF32 mode)
            x_nchw = ops.u8_to_f32_pm1(x_nchw)                               # only the fused first layer reads uint8 itself
        if self._first_fused:
            t = L[0](x_nchw.contiguous())                                    # (B,H/2,W/2,C) bf16 NHWC
        elif self._first_im2col:
            a = ops.im2col_5x5s2(x_nchw, Kp=self._kp1, out_f32=self.precision == "tf32")   # (B,H/2,W/2,Kp)
            t = L[0](a.view(1, 1, B * (H // 2) * (W // 2), self._kp1)).view(B, H // 2, W // 2, -1)
        else:
            t = L[0](ops.nchw_to_nhwc_bf16(x_nchw, ops._pad64(x_nchw.shape[1])))
        t = L[1](t)
        t = L[2](t)
        return L[3](t)

    def forward(self, inputs):
        return self.forward_nhwc(inputs).permute(0, 3, 1, 2)


class synthesisTransformModel(_PlannedTransform):
    """model/net.py:122-148.  4x [ZeroPad2d((1,0,1,0)) -> ConvTranspose2d(k5,s2,p3,op1) -> IGDN]."""

    def __init__(self, in_dim, num_filters, conv_trainable=True):
        super().__init__()
        f = num_filters
        self.in_dim = in_dim
        self.transform = nn.Sequential(
            nn.ZeroPad2d((1, 0, 1, 0)), nn.ConvTranspose2d(in_dim, f[0], 5, 2, 3, output_padding=1), ModelIGDN(f[0], inverse=True),
            nn.ZeroPad2d((1, 0, 1, 0)), nn.ConvTranspose2d(f[0], f[1], 5, 2, 3, output_padding=1), ModelIGDN(f[1], inverse=True),
            nn.ZeroPad2d((1, 0, 1, 0)), nn.ConvTranspose2d(f[1], f[2], 5, 2, 3, output_padding=1), ModelIGDN(f[2], inverse=True),
            nn.ZeroPad2d((1, 0, 1, 0)), nn.ConvTranspose2d(f[2], f[3], 5, 2, 3, output_padding=1), ModelIGDN(f[3], inverse=True),
        )
        self.cin_offset = 0      # Net places the content channels at offset M inside the full latent
        self.cin_pad = None

    def _build_plan(self):
        t = self.transform
        layers = []
        for ci, gi in ((1, 2), (4, 5), (7, 8), (10, 11)):
            conv = t[ci]
            last = ci == 10
            small = conv.out_channels * 4 <= 256 and conv.out_channels % 16 == 0 and (conv.out_channels * 4) % 64 == 0
            kind = _lib.LDIC_DECONV_GS_5x5_MERGED if (last and small) else _lib.LDIC_DECONV_GS_5x5
            kw = {}
            if ci == 1:
                kw = dict(cin_offset=self.cin_offset, cin_pad=self.cin_pad)
            layers.append(ops.ConvTC(kind, conv.weight.detach(), conv.bias.detach(), act=_lib.ACT_IGDN,
                                     out_f32=last, gdn=self._gdn_args(t[gi]), **kw))
        return layers

    def forward_nhwc(self, y_hat_nhwc_bf16: torch.Tensor) -> torch.Tensor:
        """(B,h,w,Cin_pad) bf16 NHWC -> (B,16h,16w,M) fp32 NHWC."""
        t = y_hat_nhwc_bf16
        for layer in self.plan():
            t = layer(t)
        return t

    def forward_nhwc_fused_tail(self, y_hat_nhwc_bf16, image_nchw, conv_w, want_x_tilde=False, want_out=False):
        """g_s with batch_conv (model/net.py:811) and the squared level error (:864-868) fused into the last
        deconv's epilogue.  Returns (sq_err int64[B], x_tilde or None, 16-channel g_s output or None); None when
        the last layer is not the merged small-Cout kind (caller falls back to forward_nhwc + syntax_conv_mse)."""
        if not self.has_fused_tail():
            return None
        t = self.forward_nhwc_body(y_hat_nhwc_bf16)
        return self.fused_tail(t, image_nchw, conv_w, want_x_tilde=want_x_tilde, want_out=want_out)

    def has_fused_tail(self) -> bool:
        return self.plan()[-1].kind == _lib.LDIC_DECONV_GS_5x5_MERGED

    def forward_nhwc_body(self, y_hat_nhwc_bf16, sm_limit: int = 0):
        """The first three deconv + IGDN layers (they depend on the rounded latent only)."""
        t = y_hat_nhwc_bf16
        for layer in self.plan()[:-1]:
            t = layer(t, sm_limit=sm_limit)
        return t

    def fused_tail(self, t, image_nchw, conv_w, want_x_tilde=False, want_out=False, tanh_out=False):
        """Last deconv + IGDN + batch_conv + squared level error on the output of forward_nhwc_body."""
        return self.plan()[-1].fused_tail(t, image_nchw, conv_w, want_x_tilde=want_x_tilde, want_out=want_out,
                                          tanh_out=tanh_out)

    def forward(self, inputs):
        L = self.plan()
        x = ops.nchw_to_nhwc_bf16(inputs, L[0].cin_pad) if self.cin_offset == 0 else None
        if x is None:
            raise ops.LdicError("synthesis transform with a channel offset is driven through Net")
        return self.forward_nhwc(x).permute(0, 3, 1, 2)


class h_analysisTransformModel(_PlannedTransform):
    """model/net.py:185-199: abs -> Conv3x3 s1 -> ReLU -> Conv5x5 s2 p2 -> ReLU -> Conv5x5 s2 p2."""

    def __init__(self, in_dim, num_filters, strides_list, conv_trainable=True):
        super().__init__()
        f, s = num_filters, strides_list
        if list(s) != [1, 2, 2]:
            raise ops.LdicError("h_a: only the reference's strides [1,2,2] are supported")
        self.transform = nn.Sequential(
            nn.Conv2d(in_dim, f[0], 3, s[0], 1), nn.ReLU(),
            nn.Conv2d(f[0], f[1], 5, s[1], 2), nn.ReLU(),
            nn.Conv2d(f[1], f[2], 5, s[2], 2))

    def _build_plan(self):
        t = self.transform
        pr = dict(precision=self.precision)
        return [ops.ConvTC(_lib.LDIC_CONV_S1_3x3_P1, t[0].weight.detach(), t[0].bias.detach(), act=_lib.ACT_RELU, **pr),
                ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P2, t[2].weight.detach(), t[2].bias.detach(), act=_lib.ACT_RELU, **pr),
                ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P2, t[4].weight.detach(), t[4].bias.detach(), act=_lib.ACT_NONE,
                           out_f32=True, precision="tf32_last" if self.precision == "tf32" else "bf16")]

    def forward_nhwc(self, y_abs_nhwc_bf16, sm_limit: int = 0):
        t = y_abs_nhwc_bf16
        if t.shape[1] % 4 or t.shape[2] % 4:
            raise ops.LdicError("hyper analysis needs latent H and W to be multiples of 4")
        for layer in self.plan():
            t = layer(t, sm_limit=sm_limit)
        return t                                       # (B,h/4,w/4,C) fp32 NHWC

    def forward(self, inputs):
        L = self.plan()
        x = ops.nchw_to_nhwc_bf16(inputs, L[0].cin_pad, apply_abs=True)       # torch.abs, model/net.py:197
        return self.forward_nhwc(x).permute(0, 3, 1, 2)


class h_synthesisTransformModel(_PlannedTransform):
    """model/net.py:203-216: ConvT5 s2 p2 op1 -> ReLU -> ConvT5 s2 p2 op1 -> ReLU -> ConvT3 s1 p1."""

    def __init__(self, in_dim, num_filters, strides_list, conv_trainable=True):
        super().__init__()
        f, s = num_filters, strides_list
        if list(s) != [2, 2, 1]:
            raise ops.LdicError("h_s: only the reference's strides [2,2,1] are supported")
        self.transform = nn.Sequential(
            nn.ConvTranspose2d(in_dim, f[0], 5, s[0], 2, output_padding=1), nn.ReLU(),
            nn.ConvTranspose2d(f[0], f[1], 5, s[1], 2, output_padding=1), nn.ReLU(),
            nn.ConvTranspose2d(f[1], f[2], 3, s[2], 1))

    def _build_plan(self):
        t = self.transform
        return [ops.ConvTC(_lib.LDIC_DECONV_HS_5x5, t[0].weight.detach(), t[0].bias.detach(), act=_lib.ACT_RELU),
                ops.ConvTC(_lib.LDIC_DECONV_HS_5x5, t[2].weight.detach(), t[2].bias.detach(), act=_lib.ACT_RELU),
                ops.ConvTC(_lib.LDIC_DECONV_S1_3x3, t[4].weight.detach(), t[4].bias.detach(), act=_lib.ACT_NONE,
                           out_f32=True)]

    def forward_nhwc(self, z_hat_nhwc_bf16, sm_limit: int = 0):
        t = z_hat_nhwc_bf16
        for layer in self.plan():
            t = layer(t, sm_limit=sm_limit)
        return t                                       # (B,4h,4w,C) fp32 NHWC

    def forward(self, inputs):
        L = self.plan()
        return self.forward_nhwc(ops.nchw_to_nhwc_bf16(inputs, L[0].cin_pad)).permute(0, 3, 1, 2)

"""Drop-in nn.Module surface of the reference's L1/L2 pieces on the hot path.

Same class names, constructor signatures, forward contracts and state-dict keys
as the reference; forward runs on libldic_b200 kernels (CUDA only, no fallback).

  LowerBound, NonNegativeParametrizer, ste_round   <- ops/bound_ops.py, ops/parametrizers.py, ops/ops.py
  GDN (CompressAI style, `inverse` flag)           <- layers/gdn.py:26-75
  ModelGDN / ModelIGDN                             <- model/gdn.py:29-156 (the classes model/net.py uses)
  GaussianModel, bypass_round                      <- model/net.py:266-286, 416-426
  GaussianConditional (eval forward)               <- CompressAI semantics used at model/Net_unet.py:1057
  WindowAttention, WinBasedAttention               <- layers/win_attention.py:38-209
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib, ops


class LowerBound(nn.Module):
    """ops/bound_ops.py:44-65.  Buffer `bound` kept for state-dict compatibility."""

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))
        self._bound = float(torch.tensor([float(bound)], dtype=torch.float32)[0])

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        super()._load_from_state_dict(state_dict, prefix, *a, **k)
        self._bound = float(self.bound.detach().cpu()[0])

    def forward(self, x):
        return torch.ops.ldic.lower_bound(x.contiguous(), self._bound)      # dispatcher op with the reference's backward


class NonNegativeParametrizer(nn.Module):
    """ops/parametrizers.py:23-49."""

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        bound = (self.minimum + self.reparam_offset ** 2) ** 0.5
        self.lower_bound = LowerBound(bound)
        self._pedestal = float(torch.tensor([pedestal], dtype=torch.float32)[0])

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            out = self.lower_bound(x)
            return out ** 2 - self.pedestal
        return torch.ops.ldic.nonneg_reparam(x.contiguous(), self.lower_bound._bound, self._pedestal)


def bypass_round(x):
    """model/net.py:416-426: round on our kernel (quant mode 1), identity gradient (torch.ops.ldic.round_ste)."""
    return torch.ops.ldic.round_ste(x.contiguous())


def ste_round(x):
    """ops/ops.py:20-34: round(x) - x.detach() + x (evaluated in that order)."""
    return torch.ops.ldic.round_ste(x.contiguous()) - x.detach() + x


class _GDNBase(nn.Module):
    _cache_key = None

    def _effective(self, bb, gb, ped):
        key = (self.beta._version, self.gamma._version, self.beta.data_ptr(), self.gamma.data_ptr())
        if self._cache_key != key:
            be, ge, _, _ = ops.gdn_prepare(self.beta.detach(), self.gamma.detach(), bb, gb, ped)
            self._eff = (be, ge)
            self._cache_key = key
        return self._eff


class GDN(_GDNBase):
    """layers/gdn.py:26-75 (CompressAI-style; x * rsqrt(norm) or x * sqrt(norm))."""

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

    def constants(self):
        return (self.beta_reparam.lower_bound._bound, self.gamma_reparam.lower_bound._bound, self.beta_reparam._pedestal)

    def forward(self, x):
        be, ge = self._effective(*self.constants())
        return torch.ops.ldic.gdn(x.contiguous(), be, ge, self.inverse, True)


class _ModelGDN(_GDNBase):
    """model/gdn.py:29-67: parameters + the `reparam_offset` / `pedestal` buffers."""
    _inverse = False

    def __init__(self, ch, inverse=False, beta_min=1e-6, gamma_init=.1, reparam_offset=2 ** -18):
        super().__init__()
        self.inverse = inverse          # stored and ignored, like the reference
        self.beta_min = beta_min
        self.gamma_init = gamma_init
        self.register_buffer("reparam_offset", torch.FloatTensor([reparam_offset]))
        self.register_buffer("pedestal", self.reparam_offset ** 2)
        # fp32 tensor arithmetic, exactly model/gdn.py:52-53
        self.beta_bound = float(((self.beta_min + self.reparam_offset ** 2) ** .5)[0])
        self.gamma_bound = float(self.reparam_offset[0])
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + self.pedestal))
        self.gamma = nn.Parameter(torch.sqrt(self.gamma_init * torch.eye(ch) + self.pedestal))

    def constants(self):
        return (self.beta_bound, self.gamma_bound, float(self.pedestal.detach().cpu()[0]))

    def forward(self, inputs):
        be, ge = self._effective(*self.constants())
        return torch.ops.ldic.gdn(inputs.contiguous(), be, ge, self._inverse, False)


class ModelGDN(_ModelGDN):
    """model/gdn.py:29-92  (x / sqrt(beta + gamma x^2))."""
    _inverse = False


class ModelIGDN(_ModelGDN):
    """model/gdn.py:94-156 (x * sqrt(beta + gamma x^2))."""
    _inverse = True


class GaussianModel(nn.Module):
    """model/net.py:266-286.  forward(inputs, hyper_sigma, hyper_mu) -> likelihood.
    `likelihood_bound` is 1e-8 there and 1e-12 in the U-Net family (model/Net_unet.py:602)."""

    def __init__(self, likelihood_bound: float = 1e-8):
        super().__init__()
        self.likelihood_bound = float(likelihood_bound)

    def forward(self, inputs, hyper_sigma, hyper_mu):
        _, lik, _ = ops.gaussian_likelihood(inputs, hyper_sigma, hyper_mu, quant=ops.QUANT_NONE,
                                            form=ops.FORM_GAUSSIAN_MODEL, lik_bound=self.likelihood_bound)
        return lik


class GaussianConditional(nn.Module):
    """Eval-mode forward of CompressAI's GaussianConditional(None) as the U-Net family calls it
    (model/Net_unet.py:1057, model/net_unet_ha_hs.py:937): returns (y_hat, likelihoods)."""

    def __init__(self, scale_table=None, scale_bound: float = 0.11, likelihood_bound: float = 1e-9):
        super().__init__()
        self.scale_bound = float(scale_bound)
        self.likelihood_bound = float(likelihood_bound)

    def forward(self, inputs, scales, means=None):
        vh, lik, _ = ops.gaussian_likelihood(inputs, scales, means, quant=ops.QUANT_DEQUANT if means is not None else ops.QUANT_ROUND,
                                             form=ops.FORM_GAUSSIAN_CONDITIONAL, lik_bound=self.likelihood_bound,
                                             scale_bound=self.scale_bound, want_vhat=True)
        return vh, lik


class WindowAttention(nn.Module):
    """layers/win_attention.py:38-126.  Holds the reference's parameters under the reference's names
    (relative_position_bias_table, relative_position_index, qkv, proj); the arithmetic runs in
    WinBasedAttention.forward on libldic_b200 (three 1x1 tcgen05 convs, the window kernel, one 1x1 conv)."""

    def __init__(self, dim=192, window_size=(8, 8), num_heads=8, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if window_size[0] != window_size[1]:
            raise ops.LdicError("WindowAttention: square windows only")
        if attn_drop or proj_drop:
            raise NotImplementedError("WindowAttention: dropout is a training feature (out of scope)")
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        ws = window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
        cf = torch.flatten(coords, 1)
        rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self._key = None

    def packed(self):
        """(q, k, v, proj ConvTC layers, bias[heads, N, N]) for the current parameter versions."""
        ps = [self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias, self.relative_position_bias_table]
        key = tuple((p._version, p.data_ptr()) for p in ps if p is not None)
        if self._key != key:
            Cd = self.dim
            w = self.qkv.weight.detach().float()
            b = self.qkv.bias.detach().float() if self.qkv.bias is not None else torch.zeros(3 * Cd, device=w.device)
            mk = lambda wt, bs, f32: ops.ConvTC(_lib.LDIC_CONV_1x1, wt.contiguous(), bs.contiguous(), out_f32=f32)
            q = mk(w[:Cd] * self.scale, b[:Cd] * self.scale, False)        # q = q * scale (:108) folded into the weights
            k = mk(w[Cd:2 * Cd], b[Cd:2 * Cd], False)
            v = mk(w[2 * Cd:], b[2 * Cd:], False)
            pr = mk(self.proj.weight.detach().float(), self.proj.bias.detach().float(), True)
            # the kernel indexes the (2ws-1)^2 x heads table itself (relative_position_index is the standard Swin index,
            # layers/win_attention.py:64-78; ops.window_attention_bias gathers the [heads, N, N] form for any other index)
            ws = self.window_size[0]
            cf = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
            rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0) + (ws - 1)
            standard = rel[:, :, 0] * (2 * ws - 1) + rel[:, :, 1]
            if torch.equal(self.relative_position_index.detach().cpu().long(), standard):
                bias = self.relative_position_bias_table.detach().float().contiguous()
            else:                # a checkpoint with a different index buffer: gather the [heads, N, N] form with it
                bias = ops.window_attention_bias(self.relative_position_bias_table.detach(), self.relative_position_index,
                                                 self.num_heads, ws)
            self._packed = (q, k, v, pr, bias)
            self._key = key
        return self._packed

    def packed_proj_bf16(self):
        """The proj Linear as a conv layer with bf16 output (NHWC pipelines that keep the residual stream in bf16)."""
        key = (self.proj.weight._version, self.proj.weight.data_ptr(), self.proj.bias._version)
        if getattr(self, "_pkey16", None) != key:
            self._p16 = ops.ConvTC(_lib.LDIC_CONV_1x1, self.proj.weight.detach().float().contiguous(),
                                   self.proj.bias.detach().float().contiguous(), out_f32=False)
            self._pkey16 = key
        return self._p16

    def forward(self, x, mask=None):
        """x: (num_windows*B, N, C) window tokens as in the reference; mask must be None here (the shifted-window
        mask is derived inside the kernel when called through WinBasedAttention)."""
        if mask is not None:
            raise NotImplementedError("WindowAttention.forward(mask=...): call WinBasedAttention, which derives the mask")
        Bn, N, Cd = x.shape
        ws = self.window_size[0]
        img = x.reshape(Bn, ws, ws, Cd).permute(0, 3, 1, 2).contiguous()
        o = _window_block(self, img, ws, 0, residual=False)
        return o.permute(0, 2, 3, 1).reshape(Bn, N, Cd)


def _window_block(attn: "WindowAttention", x: torch.Tensor, ws: int, shift: int, residual: bool) -> torch.Tensor:
    B, Cd, H, W = x.shape
    q, k, v, pr, bias = attn.packed()
    t = ops.nchw_to_nhwc_bf16(x, q.cin_pad)
    o = ops.window_attention_core(q(t), k(t), v(t), bias, attn.num_heads, ws, shift)
    if o.shape[-1] != pr.cin_pad:
        raise ops.LdicError("window attention: dim must be a multiple of 64")
    o = pr(o)
    if residual:
        return ops.residual_nhwc_to_nchw(o, x)
    return ops.nhwc_to_nchw_f32(o, Cd)


def _window_block_nhwc(attn: "WindowAttention", t: torch.Tensor, ws: int, shift: int) -> torch.Tensor:
    """NHWC bf16 in, NHWC bf16 out, x + (S)W-MSA(x) with the residual added in the proj conv's epilogue."""
    q, k, v, pr, bias = attn.packed()
    if t.shape[-1] != q.cin_pad:
        raise ops.LdicError("window attention: dim must be a multiple of 64")
    o = ops.window_attention_core(q(t), k(t), v(t), bias, attn.num_heads, ws, shift)
    prb = attn.packed_proj_bf16()
    return prb(o, residual=t)


class WinBasedAttention(nn.Module):
    """layers/win_attention.py:129-209: x (B,C,H,W) -> x + (S)W-MSA(x).  Same constructor and state-dict keys
    (attn.qkv.*, attn.proj.*, attn.relative_position_bias_table, attn.relative_position_index)."""

    def __init__(self, dim=192, num_heads=8, window_size=8, shift_size=0, qkv_bias=True, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0.):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        if drop_path:
            raise NotImplementedError("WinBasedAttention: drop_path is a training feature (out of scope)")
        self.attn = WindowAttention(dim, window_size=(window_size, window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()

    def kernel_shape(self) -> bool:
        """True when (dim, heads, window) is a shape ldic_window_attention_core is built for: the hot blocks of the
        reference (dim 192 / 128 / 64, 8 heads, windows 8 and 4).  The U-Net hyperprior also instantiates this class at
        dims 96 / 256 / 512 with 2x2 and 4x4 windows on 1/64-resolution maps (model/Block_unet.py:783-806,849-851);
        those few small blocks run the same arithmetic as torch ops (`forward_torch`) -- an explicit choice by shape,
        reported by `backend`, not a fallback for a missing library."""
        hd = self.dim // self.num_heads
        return self.dim % 64 == 0 and self.dim % self.num_heads == 0 and hd % 2 == 0 and hd <= 32 and \
            self.window_size in (4, 8) and self.num_heads <= 16

    @property
    def backend(self) -> str:
        return "ldic" if self.kernel_shape() else "torch"

    def forward_torch(self, x):
        """layers/win_attention.py:150-209 with torch ops (window partition, cyclic shift, 0 / -100 mask, softmax)."""
        B, Cd, H, W = x.shape
        ws, sh, nh = self.window_size, self.shift_size, self.num_heads
        a = self.attn
        t = x.permute(0, 2, 3, 1)
        mask = None
        if sh > 0:
            img = torch.zeros((1, H, W, 1), device=x.device)
            cnt = 0
            for hs in (slice(0, -ws), slice(-ws, -sh), slice(-sh, None)):
                for wsl in (slice(0, -ws), slice(-ws, -sh), slice(-sh, None)):
                    img[:, hs, wsl, :] = cnt
                    cnt += 1
            mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
            mask = mw.unsqueeze(1) - mw.unsqueeze(2)
            mask = mask.masked_fill(mask != 0, float(-100.0)).masked_fill(mask == 0, float(0.0))
            t = torch.roll(t, shifts=(-sh, -sh), dims=(1, 2))
        win = t.reshape(B, H // ws, ws, W // ws, ws, Cd).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, Cd)
        Bn, Nt, _ = win.shape
        qkv = a.qkv(win).reshape(Bn, Nt, 3, nh, Cd // nh).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * a.scale, qkv[1], qkv[2]
        attn = q @ k.transpose(-2, -1)
        bias = a.relative_position_bias_table[a.relative_position_index.view(-1)].view(Nt, Nt, -1).permute(2, 0, 1)
        attn = attn + bias.unsqueeze(0)
        if mask is not None:
            nW = mask.shape[0]
            attn = (attn.view(Bn // nW, nW, nh, Nt, Nt) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nh, Nt, Nt)
        o = (torch.softmax(attn, dim=-1) @ v).transpose(1, 2).reshape(Bn, Nt, Cd)
        o = a.proj(o).view(B, H // ws, W // ws, ws, ws, Cd).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, Cd)
        if sh > 0:
            o = torch.roll(o, shifts=(sh, sh), dims=(1, 2))
        return x + o.permute(0, 3, 1, 2)

    def forward_nhwc(self, t):
        """(B,H,W,C) bf16 NHWC -> same: the block inside an NHWC bf16 pipeline (Win_noShift_Attention on the kernels)."""
        return _window_block_nhwc(self.attn, t, self.window_size, self.shift_size)

    def forward(self, x):
        if not self.kernel_shape():
            if not x.is_cuda:
                raise ops.LdicError("WinBasedAttention runs on CUDA only")
            return self.forward_torch(x)
        return _window_block(self.attn, x, self.window_size, self.shift_size, residual=True)


def psnr_from_sq_err(sq_err: torch.Tensor, chw: int):
    """model/net.py:868-869 from the exact integer sums: v_mse (B,), v_psnr scalar."""
    v_mse = sq_err.to(torch.float64) / float(chw)
    v_psnr = torch.mean(20.0 * torch.log10(255.0 / torch.sqrt(v_mse)), 0)
    return v_mse.to(torch.float32), v_psnr.to(torch.float32)

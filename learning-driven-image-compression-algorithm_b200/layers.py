"""Drop-in nn.Module surface of the reference's L1/L2 pieces on the hot path.

Same class names, constructor signatures, forward contracts and state-dict keys
as the reference; forward runs on libldic_b200 kernels (CUDA only, no fallback).

  LowerBound, NonNegativeParametrizer, ste_round   <- ops/bound_ops.py, ops/parametrizers.py, ops/ops.py
  GDN (CompressAI style, `inverse` flag)           <- layers/gdn.py:26-75
  ModelGDN / ModelIGDN                             <- model/gdn.py:29-156 (the classes model/net.py uses)
  GaussianModel, bypass_round                      <- model/net.py:266-286, 416-426
  GaussianConditional (eval forward)               <- CompressAI semantics used at model/Net_unet.py:1057
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops


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
        return ops.lower_bound(x, self._bound)


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
        return ops.nonneg_reparam(x, self.lower_bound._bound, self._pedestal)


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        # round on our kernel (quant mode 1), identity gradient: model/net.py:416-426
        _, _, yf = ops.latent_prep(x, want_round_bf16=False, want_abs_bf16=False, want_round_f32=True)
        return yf

    @staticmethod
    def backward(ctx, g):
        return g


def bypass_round(x):
    """model/net.py:426."""
    return _RoundSTE.apply(x)


def ste_round(x):
    """ops/ops.py:20-34: round(x) - x.detach() + x (evaluated in that order)."""
    return _RoundSTE.apply(x) - x.detach() + x


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
        return ops.gdn_nchw(x, be, ge, self.inverse, use_rsqrt=True)


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
        return ops.gdn_nchw(inputs, be, ge, self._inverse, use_rsqrt=False)


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


def psnr_from_sq_err(sq_err: torch.Tensor, chw: int):
    """model/net.py:868-869 from the exact integer sums: v_mse (B,), v_psnr scalar."""
    v_mse = sq_err.to(torch.float64) / float(chw)
    v_psnr = torch.mean(20.0 * torch.log10(255.0 / torch.sqrt(v_mse)), 0)
    return v_mse.to(torch.float32), v_psnr.to(torch.float32)

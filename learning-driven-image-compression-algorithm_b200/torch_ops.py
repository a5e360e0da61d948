"""``torch.library`` registration of the hot-path operators (namespace ``ldic::``).

SURVEY 8(b) / BASELINE north_star: the kernels sit behind a thin C-ABI torch custom-op layer.  The C ABI
(include/ldic.h, loaded with ctypes by ``_lib``) is the boundary a non-Python host would bind; this module
registers the same entry points with the PyTorch dispatcher so they are callable as ``torch.ops.ldic.<op>``,
carry fake (meta) implementations for shape propagation / ``torch.compile`` tracing / FakeTensorMode, and --
for ``lower_bound`` and ``round_ste`` -- the reference's custom backward formulas
(ops/bound_ops.py:30-41, model/net.py:416-426).  Every CUDA implementation goes straight to libldic_b200;
no op has a CPU or composite fallback (a CPU tensor raises ``LdicError`` / NotImplementedError).

Registered operators
    ldic::lower_bound(x, bound) -> y                                            a3   ops/bound_ops.py:21-65
    ldic::nonneg_reparam(p, bound, pedestal) -> out                             a3   ops/parametrizers.py:23-49
    ldic::round_ste(x) -> y                                                     a6   model/net.py:416-426
    ldic::gdn(x, beta_eff, gamma_eff, inverse, use_rsqrt) -> y                  a2   model/gdn.py:69-92, layers/gdn.py:62-75
    ldic::round_likelihood_bpp(v, sigma, mu?, quant, form, lik_bound, scale_bound) -> (v_hat, lik, sum_ln)
                                                                                a6-a9 model/net.py:266-286,856-861
    ldic::mse_sum(x, x_tilde, clamp_pm1) -> int64[B]                            a11  model/net.py:864-868
    ldic::rd_metrics(bits3, sq_err, chw, pixels_per_image) -> (bpp_psnr[2], v_mse[B])   a9/a11 model/net.py:856-869
    ldic::conv_forward(x, w_packed, bias_packed, gamma_bf16?, beta_tiled?, kind, cin, cout, cin_pad, cout_pad,
                       act, out_f32, aux0, aux1) -> y                           a1/a4/a5/a10 model/net.py:91-216
    ldic::window_attention(q, k, v, bias, heads, ws, shift) -> out              f2   layers/win_attention.py:38-127
    ldic::rans_encode(v, sigma, mu?, quant, scale_bound, streams) -> (bytes[B, cap], sizes[B], status[B])
    ldic::rans_decode(bytes, sizes, sigma, mu?, shape, quant, scale_bound, streams) -> (v_hat, status[B])
                                                                                f4   (no reference coder: csrc/rans.cu)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops

NS = "ldic"
_dev = "cuda"


def _register():
    lib = torch.library

    # ---- a3 -------------------------------------------------------------------------------------
    @lib.custom_op(f"{NS}::lower_bound", mutates_args=(), device_types=_dev)
    def lower_bound(x: Tensor, bound: float) -> Tensor:
        return ops.lower_bound_fwd(x, bound)

    @lower_bound.register_fake
    def _(x, bound):
        return torch.empty_like(x, memory_format=torch.contiguous_format)

    @lib.custom_op(f"{NS}::lower_bound_bwd", mutates_args=(), device_types=_dev)
    def lower_bound_bwd(x: Tensor, bound: float, grad_out: Tensor) -> Tensor:
        return ops.lower_bound_bwd(x, bound, grad_out)

    @lower_bound_bwd.register_fake
    def _(x, bound, grad_out):
        return torch.empty_like(grad_out, memory_format=torch.contiguous_format)

    def _lb_setup(ctx, inputs, output):
        x, bound = inputs
        ctx.save_for_backward(x)
        ctx.bound = float(bound)

    def _lb_backward(ctx, grad):
        (x,) = ctx.saved_tensors
        return torch.ops.ldic.lower_bound_bwd(x, ctx.bound, grad.contiguous()), None

    lower_bound.register_autograd(_lb_backward, setup_context=_lb_setup)

    @lib.custom_op(f"{NS}::nonneg_reparam", mutates_args=(), device_types=_dev)
    def nonneg_reparam(p: Tensor, bound: float, pedestal: float) -> Tensor:
        return ops.nonneg_reparam(p, bound, pedestal)

    @nonneg_reparam.register_fake
    def _(p, bound, pedestal):
        return torch.empty_like(p, memory_format=torch.contiguous_format)

    # ---- a6 -------------------------------------------------------------------------------------
    @lib.custom_op(f"{NS}::round_ste", mutates_args=(), device_types=_dev)
    def round_ste(x: Tensor) -> Tensor:
        return ops.latent_prep(x, want_round_bf16=False, want_abs_bf16=False, want_round_f32=True)[2]

    @round_ste.register_fake
    def _(x):
        return torch.empty_like(x, memory_format=torch.contiguous_format)

    round_ste.register_autograd(lambda ctx, g: g)              # identity gradient (BypassRound, model/net.py:421-423)

    # ---- a2 -------------------------------------------------------------------------------------
    @lib.custom_op(f"{NS}::gdn", mutates_args=(), device_types=_dev)
    def gdn(x: Tensor, beta_eff: Tensor, gamma_eff: Tensor, inverse: bool, use_rsqrt: bool) -> Tensor:
        return ops.gdn_nchw(x, beta_eff, gamma_eff, inverse, use_rsqrt)

    @gdn.register_fake
    def _(x, beta_eff, gamma_eff, inverse, use_rsqrt):
        return torch.empty_like(x, memory_format=torch.contiguous_format)

    # ---- a6-a9 ----------------------------------------------------------------------------------
    @lib.custom_op(f"{NS}::round_likelihood_bpp", mutates_args=(), device_types=_dev)
    def round_likelihood_bpp(v: Tensor, sigma: Tensor, mu: Optional[Tensor], quant: int, form: int, lik_bound: float,
                             scale_bound: float) -> Tuple[Tensor, Tensor, Tensor]:
        shape = v.shape
        if v.dim() != 4:                       # any shape with elementwise sigma / mu: one row of v.numel() elements
            if sigma.shape != shape or (mu is not None and mu.shape != shape):
                raise ops.LdicError("round_likelihood_bpp: non-4D inputs need sigma / mu of v's shape")
            v, sigma = v.reshape(1, 1, 1, -1), sigma.reshape(1, 1, 1, -1)
            mu = None if mu is None else mu.reshape(1, 1, 1, -1)
        vh, lik, s = ops.gaussian_likelihood(v, sigma, mu, quant=quant, form=form, lik_bound=lik_bound,
                                             scale_bound=scale_bound, want_lik=True, want_vhat=True)
        return vh.reshape(shape), lik.reshape(shape), s

    @round_likelihood_bpp.register_fake
    def _(v, sigma, mu, quant, form, lik_bound, scale_bound):
        return torch.empty_like(v), torch.empty_like(v), v.new_empty((1,))

    # ---- a11 / scalar tail ------------------------------------------------------------------------
    @lib.custom_op(f"{NS}::mse_sum", mutates_args=(), device_types=_dev)
    def mse_sum(x: Tensor, x_tilde: Tensor, clamp_pm1: bool) -> Tensor:
        return ops.mse_sum(x, x_tilde, clamp_pm1)

    @mse_sum.register_fake
    def _(x, x_tilde, clamp_pm1):
        return x.new_empty((x.shape[0],), dtype=torch.int64)

    @lib.custom_op(f"{NS}::rd_metrics", mutates_args=(), device_types=_dev)
    def rd_metrics(bits3: Tensor, sq_err: Tensor, chw: int, pixels_per_image: float) -> Tuple[Tensor, Tensor]:
        packed, v_mse = ops.rd_pack_metrics(bits3, sq_err, chw)
        return ops.rd_finish_metrics(packed, pixels_per_image), v_mse

    @rd_metrics.register_fake
    def _(bits3, sq_err, chw, pixels_per_image):
        return bits3.new_empty((2,)), bits3.new_empty((sq_err.shape[0],))

    # ---- a1 / a4 / a5 / a10 -----------------------------------------------------------------------
    def _conv_desc(kind, B, H, W, cin, cout, cin_pad, cout_pad, act, out_f32, aux0, aux1):
        return _lib.ConvDesc(kind, B, H, W, cin, cout, cin_pad, cout_pad, act, int(out_f32), aux0, aux1, 0, 0)

    def _conv_geometry(x, kind):
        if kind == _lib.LDIC_CONV_FIRST_5x5S2:
            B, _, H, W = x.shape
        else:
            B, H, W, _ = x.shape
        return B, H, W

    @lib.custom_op(f"{NS}::conv_forward", mutates_args=(), device_types=_dev)
    def conv_forward(x: Tensor, w_packed: Tensor, bias_packed: Tensor, gamma_bf16: Optional[Tensor],
                     beta_tiled: Optional[Tensor], kind: int, cin: int, cout: int, cin_pad: int, cout_pad: int, act: int,
                     out_f32: bool, aux0: int, aux1: int) -> Tensor:
        ops._req(x, None, "x")
        if not x.is_contiguous():
            raise ops.LdicError("conv_forward: x must be contiguous")
        B, H, W = _conv_geometry(x, kind)
        d = _conv_desc(kind, B, H, W, cin, cout, cin_pad, cout_pad, act, out_f32, aux0, aux1)
        if kind == _lib.LDIC_CONV_FIRST_5x5S2 and x.dtype == torch.uint8:
            d.aux0 = 1
        dims = (C.c_int * 4)()
        _lib.check(_lib.load().ldic_conv_out_dims(C.byref(d), dims), "ldic_conv_out_dims")
        y = torch.empty(tuple(dims), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
        _lib.check(_lib.load().ldic_conv_forward(C.byref(d), ops._ptr(x), ops._ptr(w_packed), ops._ptr(bias_packed),
                                                 ops._ptr(gamma_bf16), ops._ptr(beta_tiled), ops._ptr(y), ops._stream()),
                   "ldic_conv_forward")
        return y

    @conv_forward.register_fake
    def _(x, w_packed, bias_packed, gamma_bf16, beta_tiled, kind, cin, cout, cin_pad, cout_pad, act, out_f32, aux0, aux1):
        B, H, W = _conv_geometry(x, kind)
        d = _conv_desc(kind, int(B), int(H), int(W), cin, cout, cin_pad, cout_pad, act, out_f32, aux0, aux1)
        dims = (C.c_int * 4)()
        _lib.check(_lib.load().ldic_conv_out_dims(C.byref(d), dims), "ldic_conv_out_dims")   # host-only shape query
        return x.new_empty(tuple(dims), dtype=torch.float32 if out_f32 else torch.bfloat16)

    # ---- f2 ---------------------------------------------------------------------------------------
    @lib.custom_op(f"{NS}::window_attention", mutates_args=(), device_types=_dev)
    def window_attention(q: Tensor, k: Tensor, v: Tensor, bias: Tensor, heads: int, ws: int, shift: int) -> Tensor:
        return ops.window_attention_core(q, k, v, bias, heads, ws, shift)

    @window_attention.register_fake
    def _(q, k, v, bias, heads, ws, shift):
        return torch.empty_like(q)

    # ---- f4 ---------------------------------------------------------------------------------------
    def _rans_capacity(shape, streams):
        Bb, Cc, Hh, Ww = (int(d) for d in shape)
        n = Cc * Hh * Ww
        S = int(streams) if streams > 0 else ops.rans_streams_for(n)
        return S, (int(_lib.load().ldic_rans_max_bytes(n, S)) + 3) & ~3

    @lib.custom_op(f"{NS}::rans_encode", mutates_args=(), device_types=_dev)
    def rans_encode(v: Tensor, sigma: Tensor, mu: Optional[Tensor], quant: int, scale_bound: float,
                    streams: int) -> Tuple[Tensor, Tensor, Tensor]:
        e = ops.rans_encode(v, sigma, mu, quant=quant, scale_bound=scale_bound, streams=streams if streams > 0 else None)
        return e.buf, e.sizes, e.status

    @rans_encode.register_fake
    def _(v, sigma, mu, quant, scale_bound, streams):
        S, cap = _rans_capacity(v.shape, streams)
        B = int(v.shape[0])
        return (v.new_empty((B, cap), dtype=torch.uint8), v.new_empty((B,), dtype=torch.int32),
                v.new_empty((B,), dtype=torch.int32))

    @lib.custom_op(f"{NS}::rans_decode", mutates_args=(), device_types=_dev)
    def rans_decode(buf: Tensor, sizes: Tensor, sigma: Tensor, mu: Optional[Tensor], shape: List[int], quant: int,
                    scale_bound: float, streams: int) -> Tuple[Tensor, Tensor]:
        n = int(shape[1]) * int(shape[2]) * int(shape[3])
        S = int(streams) if streams > 0 else ops.rans_streams_for(n)
        data = ops.RansStreams(buf, sizes, None, S, n, quant)
        rows, cols, rps, kw = ops._rans_surface(tuple(int(d) for d in shape), sigma, mu)
        out = torch.empty(tuple(int(d) for d in shape), dtype=torch.float32, device=sigma.device)
        st = ops.rans_decode_rows(data, rows, cols, rps, out, v_hat_rs=cols, quant=quant, scale_bound=scale_bound,
                                  check_status=False, **kw)
        return out, st.clone()

    @rans_decode.register_fake
    def _(buf, sizes, sigma, mu, shape, quant, scale_bound, streams):
        return (sigma.new_empty(tuple(int(d) for d in shape), dtype=torch.float32),
                sigma.new_empty((int(shape[0]),), dtype=torch.int32))

    return ("lower_bound", "lower_bound_bwd", "nonneg_reparam", "round_ste", "gdn", "round_likelihood_bpp", "mse_sum",
            "rd_metrics", "conv_forward", "window_attention", "rans_encode", "rans_decode")


REGISTERED = _register()

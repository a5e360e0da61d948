"""Drop-in ``Net`` of the reference's model/net.py (429-871) on the B200 kernels.

Same constructor, same ``forward(inputs, mode, num)`` contract in test mode
(returns ``(bpp, v_mse[B], v_psnr)``), same parameter / buffer names, so a
reference checkpoint loads with ``strict=True`` (the four one-hot
``*_sampler.sample_filter`` buffers and the HAN post-processing head are
accepted and discarded: the sampler is a gather here, post-processing is out of
scope of this path).

Kernel coverage: g_a + GDN (a1,a2,a3), h_a (a4), h_s (a5), rounding
(a6), GaussianModel likelihoods (a7), bpp reduction (a9), g_s + IGDN (a10),
batch_conv + MSE/PSNR (a11) and the context model PredictionModel_Context
(TMA patch gather + 3 convs + fc on the tcgen05 kernel) run on
libldic_b200, and so does the small syntax branch (Syntax_Model, PredictionModel_Syntax, conv_generator:
``ldic_syntax_branch``).  Both widths of the reference are supported: N=192/M=16 and the
``is_high`` model N=384/M=32 (model/net.py:446-451; 384 accumulator columns run on ``conv_wide_kernel``).

No torch-op implementation of any stage lives in this module; tests that cross-check a stage against stock
torch ops pass their own implementation through ``rd_forward(..., overrides=...)``.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layers import GaussianModel, psnr_from_sq_err
from .transforms import (analysisTransformModel, h_analysisTransformModel, h_synthesisTransformModel,
                         synthesisTransformModel)

_DISCARDED_PREFIXES = ("HAN.", "conv_weights_gen_HAN.", "add_mean.")


class conv_generator(nn.Module):
    """model/net.py:322-343."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.transform = nn.Sequential(nn.Linear(in_dim, 128), nn.LeakyReLU(0.2), nn.Linear(128, 256),
                                       nn.LeakyReLU(0.2), nn.Linear(256, out_dim * 3))

    def forward(self, x):
        b = x.shape[0]
        return self.transform(x.reshape(b, -1)).view(b, 3, self.out_dim, 1, 1)


class Syntax_Model(nn.Module):
    """model/net.py:349-375."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.down0 = nn.Conv2d(in_dim, 32, 3, 2, 1)
        self.down1 = nn.Conv2d(32, 64, 3, 2, 1)
        self.conv = nn.Conv2d(in_dim + 32 + 64, out_dim, 1, 1, 0)
        self.pooling = nn.AdaptiveAvgPool2d(1)

    def forward(self, syntax):
        out1 = self.pooling(syntax)
        ds1 = F.relu(self.down0(syntax))
        out2 = self.pooling(ds1)
        ds2 = F.relu(self.down1(ds1))
        out3 = self.pooling(ds2)
        return self.conv(torch.cat((out1, out2, out3), 1))


class PredictionModel_Syntax(nn.Module):
    """model/net.py:378-413.  Returns (mu, sigma); Net binds them swapped like the reference (:789)."""

    def __init__(self, in_dim, dim=192, trainable=True, outdim=None):
        super().__init__()
        outdim = dim if outdim is None else outdim
        self.down0 = nn.Conv2d(in_dim, dim, 3, 2, 1)
        self.down1 = nn.Conv2d(dim, dim, 3, 2, 1)
        self.pooling = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(dim * 2 + in_dim, outdim)
        self.flatten = nn.Flatten()

    def forward(self, y_rounded, h_tilde, h_sampler=None):
        b, c, h, w = y_rounded.size()
        ds0 = F.relu(self.down0(h_tilde))
        ds1 = F.relu(self.down1(ds0))
        ctx = self.flatten(torch.cat((self.pooling(h_tilde), self.pooling(ds0), self.pooling(ds1)), 1))
        t = self.fc(ctx)
        mu = t[:, :c].view(b, h, w, c).permute(0, 3, 1, 2)
        sigma = torch.exp(t[:, c:]).contiguous().view(b, h, w, c).permute(0, 3, 1, 2)
        return mu, sigma


class PredictionModel_Context(nn.Module):
    """model/net.py:289-319.  The reference's BlockSample (one-hot 7x7 conv2d, :219-242) is
    replaced by the equivalent pad + unfold gather: patch cell (i,j) of latent position (y,x) is
    input[y+i-3, x+j-2]; the y sampler zeroes cells (3,2) and (3,3) (causal mask)."""

    def __init__(self, in_dim, dim=192, trainable=True, outdim=None):
        super().__init__()
        outdim = dim if outdim is None else outdim
        self.transform = nn.Sequential(nn.Conv2d(in_dim, dim, 3, 1, 1), nn.LeakyReLU(0.2),
                                       nn.Conv2d(dim, dim, 3, 2, 1), nn.LeakyReLU(0.2),
                                       nn.Conv2d(dim, dim, 3, 1, 1), nn.LeakyReLU(0.2))
        self.fc = nn.Linear(dim * 2 * 2, outdim)
        self.flatten = nn.Flatten()

    @staticmethod
    def sample(x, masked):
        b, c, h, w = x.shape
        t = F.unfold(F.pad(x, (2, 1, 3, 0)), kernel_size=4)             # (b, c*16, h*w)
        t = t.view(b, c, 4, 4, h * w).permute(0, 4, 1, 2, 3).reshape(b * h * w, c, 4, 4)
        if masked:
            t[:, :, 3, 2:] = 0
        return t

    # -- tensor-core path ----------------------------------------------------------
    _plan = None
    _plan_key = None

    def plan(self, N: int, M: int):
        from . import _lib
        key = tuple((p._version, p.data_ptr()) for p in self.parameters()) + (N, M)
        if self._plan is None or self._plan_key != key:
            t = self.transform
            with torch.no_grad():
                self._plan = [
                    ops.ConvTC(_lib.LDIC_CTX_CONV1, t[0].weight.detach(), t[0].bias.detach(), act=_lib.ACT_LEAKY02, aux=(N, M)),
                    ops.ConvTC(_lib.LDIC_CTX_CONV2, t[2].weight.detach(), t[2].bias.detach(), act=_lib.ACT_LEAKY02),
                    ops.ConvTC(_lib.LDIC_CTX_CONV3, t[4].weight.detach(), t[4].bias.detach(), act=_lib.ACT_LEAKY02),
                    ops.ConvTC(_lib.LDIC_CTX_FC, self.fc.weight.detach(), self.fc.bias.detach(), out_f32=True),
                ]
            self._plan_key = key
        return self._plan

    _sheared = None

    def sheared_conv1(self, N: int, M: int, shear: int):
        """The first context conv for an input image stored sheared by `shear` columns per row (Net.decompress)."""
        from . import _lib
        key = (self._plan_key, shear)
        if self._sheared is None or self._sheared[0] != key:
            t = self.transform
            with torch.no_grad():
                layer = ops.ConvTC(_lib.LDIC_CTX_CONV1, t[0].weight.detach(), t[0].bias.detach(), act=_lib.ACT_LEAKY02,
                                   aux=(N, M + 256 * shear))
            self._sheared = (key, layer)
        return self._sheared[1]

    def raw_tc(self, y_round_bf16: torch.Tensor, h2: torch.Tensor, M: int) -> torch.Tensor:
        """(B,h,w,N) bf16 rounded latent (all N channels; the first M are the syntax channels and
        get zero weights) + (B,h,w,N) fp32 h_s output -> (B*h*w, 1, 2, Cp) fp32: [..,0,:c] = mu,
        [..,1,:c] = log sigma.  The 4x4 patches of BlockSample (model/net.py:219-242) are gathered by
        TMA inside the first conv instead of being materialised."""
        N = h2.shape[-1]
        L = self.plan(N, M)
        x = ops.ctx_pack_input(y_round_bf16, h2)
        t = L[0](x)            # (P,4,4,N)
        t = L[1](t)            # (P,2,2,N)
        t = L[2](t)            # (P,2,2,N)
        return L[3](t)         # (P,1,2,Cp) fp32

    def forward(self, y_rounded, h_tilde, y_sampler=None, h_sampler=None):
        """Module surface of the reference: (B,c,h,w) rounded content latent + (B,N,h,w) h_s output
        -> (mu, sigma) as NCHW views of (b,h,w,c) storage (model/net.py:313-319)."""
        b, c, h, w = y_rounded.shape
        N = h_tilde.shape[1]
        M = N - c
        y_full = torch.zeros(b, h, w, N, dtype=torch.bfloat16, device=y_rounded.device)
        y_full[..., M:] = y_rounded.permute(0, 2, 3, 1)
        h2 = h_tilde.permute(0, 2, 3, 1).contiguous().float()
        t = self.raw_tc(y_full, h2, M)                       # (P,1,2,Cp)
        mu = t[:, 0, 0, :c].view(b, h, w, c).permute(0, 3, 1, 2)
        sigma = torch.exp(t[:, 0, 1, :c]).contiguous().view(b, h, w, c).permute(0, 3, 1, 2)
        return mu, sigma


class Net(nn.Module):
    def __init__(self, train_size, test_size, is_high, post_processing):
        super().__init__()
        if post_processing:
            raise NotImplementedError("HAN post-processing (model/han.py) is outside the rate-distortion forward path")
        self.train_size = train_size
        self.test_size = test_size
        self.post_processing = post_processing
        self.is_high = is_high
        N, M = (384, 32) if is_high else (192, 16)                 # model/net.py:446-451
        self.M, self.N = M, N
        self.a_model = analysisTransformModel(3, [N, N, N, N])
        self.s_model = synthesisTransformModel(N - M, [N, N, N, M])
        self.s_model.cin_offset = M          # g_s reads the full N-channel rounded latent; syntax rows get zero weights
        self.s_model.cin_pad = ops._pad64(N)
        self.syntax_model = Syntax_Model(M, M)
        self.conv_weights_gen = conv_generator(in_dim=M, out_dim=M)
        self.ha_model = h_analysisTransformModel(N, [N, N, N], [1, 2, 2])
        self.hs_model = h_synthesisTransformModel(N, [N, N, N], [2, 2, 1])
        self.entropy_bottleneck_z2 = GaussianModel()
        self.entropy_bottleneck_z3 = GaussianModel()
        self.entropy_bottleneck_z3_syntax = GaussianModel()
        self.v_z2_sigma = nn.Parameter(torch.ones((1, N, 1, 1), dtype=torch.float32, requires_grad=True))
        self.register_parameter('z2_sigma', self.v_z2_sigma)       # same tensor under two names (model/net.py:482-488)
        self.prediction_model = PredictionModel_Context(in_dim=2 * N - M, dim=N, outdim=(N - M) * 2)
        self.prediction_model_syntax = PredictionModel_Syntax(in_dim=N, dim=M, outdim=M * 2)
        self.tail_fused = True            # batch_conv + MSE inside the last deconv's epilogue (False: separate kernels)
        # SM partition: the hyperprior / syntax chain (h_a, likelihood z, h_s, syntax branch: small grids of 12..48 CTAs)
        # runs on a side stream on `side_sms` SMs while the first three g_s deconvs, which depend on round(y) only,
        # run on the remaining SMs of the main stream (persistent kernels with their grids capped accordingly).
        self.side_sms = 12                # SMs of the side stream; 0: single stream (same-box interleaved A/B: 3.31 -> 3.19 ms per step)
        self.side_eager = False           # also partition when launching eagerly (tests)
        self._side_streams = {}
        # forward(x, 'test'): transparent CUDA-graph capture / replay per input shape (see forward)
        self.auto_graph = True
        self._graphs = {}
        self._decode_graphs = {}
        self._discarded_logged = False

    # -- TF32 parity mode ---------------------------------------------------------
    @property
    def parity_tf32(self) -> bool:
        return self.a_model.precision == "tf32"

    @parity_tf32.setter
    def parity_tf32(self, on: bool):
        """True: the analysis side (g_a, h_a -- the layers that decide the symbols) runs with fp32 activations and
        kind::tf32 MMAs (10-bit mantissa operands instead of bf16's 7), at about half the rate of those layers.  It is
        the precision comparison SURVEY H4 asks for: how many symbols flip against the fp32 reference because of the
        operand type.  The rate-optimised product path is bf16 (default)."""
        p = "tf32" if on else "bf16"
        self.a_model.precision = p
        self.ha_model.precision = p
        self._graphs.clear()

    # -- checkpoint compatibility ---------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        kept, dropped = {}, []
        for k, v in state_dict.items():
            if k.endswith("_sampler.sample_filter") or k.startswith(_DISCARDED_PREFIXES):
                dropped.append(k)
            else:
                kept[k] = v
        if dropped and not self._discarded_logged:
            import logging
            logging.getLogger("ldic_b200").info(
                "Net.load_state_dict: %d checkpoint entries outside the rate-distortion path are not used "
                "(one-hot sampler filters, HAN head, add_mean), e.g. %s", len(dropped), dropped[0])
            self._discarded_logged = True
        self._graphs.clear()
        return super().load_state_dict(kept, strict=strict, assign=assign)

    def base_params(self):
        mods = [self.a_model, self.s_model, self.ha_model, self.hs_model, self.syntax_model, self.conv_weights_gen,
                self.prediction_model, self.prediction_model_syntax]
        params = [p for m in mods for p in m.parameters()]
        params.append(self.v_z2_sigma)
        return params

    def post_processing_params(self):
        return []

    def batch_conv(self, weights, inputs):
        """model/net.py:527-537 (kept for API parity; the forward uses the fused kernel)."""
        b, ch, _, _ = inputs.shape
        _, ch_out, _, k, _ = weights.shape
        weights = weights.reshape(b * ch_out, ch, k, k)
        inputs = torch.cat(torch.split(inputs, 1, dim=0), dim=1)
        out = F.conv2d(inputs, weights, stride=1, padding=0, groups=b)
        return torch.cat(torch.split(out, ch_out, dim=1), dim=0)

    # -- the hot path -----------------------------------------------------------------------
    @torch.no_grad()
    def rd_forward(self, inputs: torch.Tensor, want_x_hat: bool = False, want_likelihoods: bool = False,
                   want_xt16: bool = False, per_image_bits: bool = False,
                   overrides: Optional[Dict[str, Callable]] = None) -> Dict[str, torch.Tensor]:
        """Rate-distortion forward of Net.forward(mode='test') (model/net.py:539-871).
        Returns the per-stream sum(ln L) (`bits` = [z, y, syntax]), the exact per-image
        squared-error sums and, on request, x_hat / likelihood tensors.

        `inputs`: (B,3,H,W) fp32 in [-1,1], or uint8 = the 8-bit levels of the image; then the first layer applies
        the reference's map x = (u/255)*2-1 (ToTensor + eval_net.py:84) while it builds its patches and the fused
        tail compares against the levels, so only 1 byte per sample crosses PCIe.
        `overrides` (tests only): {"syntax": f(net, y_nchw, h2_nchw) -> (z3, z3_round, first, second, conv_w),
        "context": f(net, y_nchw, h2_nchw) -> (ctx, row_stride, sigma_offset)} replace a stage of the kernel path."""
        if not inputs.is_cuda:
            raise ops.LdicError("Net runs on CUDA only (no CPU fallback)")
        with torch.cuda.device(inputs.device):
            return self._rd_forward(inputs, want_x_hat, want_likelihoods, want_xt16, per_image_bits, overrides or {})

    def _rd_forward(self, inputs, want_x_hat, want_likelihoods, want_xt16, per_image_bits, overrides):
        want_likelihoods = want_likelihoods or per_image_bits
        x = inputs.contiguous() if inputs.dtype == torch.uint8 else inputs.contiguous().float()
        B, _, H, W = x.shape
        N, M = self.N, self.M
        if H % 64 or W % 64:
            raise ops.LdicError("H and W must be multiples of 64 (eval_net.py:68-81 pads to 64)")
        h, w = H // 16, W // 16
        P = B * h * w
        dev = x.device

        y = self.a_model.forward_nhwc(x)                                            # :627   (B,h,w,N) fp32 NHWC
        y_round_bf16, y_abs_bf16, _ = ops.latent_prep(y)                            # :197 abs, :741 round
        if self.ha_model.precision == "tf32":
            y_abs_bf16 = torch.abs(y)                                               # TF32 parity mode: h_a reads fp32 |y|
        fused_ok = self.tail_fused and self.s_model.has_fused_tail()
        # the SM partition pays off where launches cost no host time, i.e. inside a graph capture (forward() and
        # GraphedEvaluator replay graphs); issued eagerly, the fork / join and the side stream's allocator pool cost more
        # host time than the overlap returns (4.3 vs 2.9 ms per step from Python), unless `side_eager` asks for it
        split = (self.side_sms > 0 and fused_ok and not overrides
                 and (self.side_eager or torch.cuda.is_current_stream_capturing()))
        bits = torch.empty(3, dtype=torch.float32, device=dev)                      # sum(ln L) of z, y, syntax
        lik_z = None

        def hyper_chain(sm_limit: int):
            nonlocal lik_z
            z = self.ha_model.forward_nhwc(y_abs_bf16, sm_limit=sm_limit)           # :666   (B,h/4,w/4,N) fp32
            Pz = z.shape[0] * z.shape[1] * z.shape[2]
            z_hat_bf16 = torch.empty(z.shape, dtype=torch.bfloat16, device=dev)
            sigma_z = self.z2_sigma.detach().reshape(N).contiguous()
            lik_z = torch.empty_like(z) if want_likelihoods else None
            ops.likelihood_rows(z, Pz, N, v_rs=N, sigma=sigma_z, sigma_mode=1, quant=ops.QUANT_ROUND,         # :676,:781
                                lik_bound=self.entropy_bottleneck_z2.likelihood_bound,
                                v_hat_bf16=z_hat_bf16, vb_rs=N, lik=lik_z, sum_out=bits[0:1])
            h2 = self.hs_model.forward_nhwc(z_hat_bf16, sm_limit=sm_limit)          # :681   (B,h,w,N) fp32 NHWC
            if "syntax" in overrides:
                syn = overrides["syntax"](self, y.permute(0, 3, 1, 2), h2.permute(0, 3, 1, 2))
            else:
                syn = ops.syntax_branch(y, h2, M, self.syntax_model, self.prediction_model_syntax,
                                        self.conv_weights_gen)                      # :712-719,:753,:789,:805
            return z, h2, syn

        gs_body = None
        if split:
            main = torch.cuda.current_stream(dev)
            side = self._side_streams.get(dev.index)
            if side is None:
                side = self._side_streams[dev.index] = torch.cuda.Stream(device=dev)
            side.wait_stream(main)                                                  # fork: round(y), |y| are ready
            with torch.cuda.stream(side):
                z, h2, syn = hyper_chain(self.side_sms)
            gs_body = self.s_model.forward_nhwc_body(y_round_bf16, sm_limit=-self.side_sms)   # :800 first three deconv + IGDN
            main.wait_stream(side)                                                  # join
            if not torch.cuda.is_current_stream_capturing():
                for t in (z, h2, lik_z) + tuple(syn):
                    if t is not None:
                        t.record_stream(main)
        else:
            z, h2, syn = hyper_chain(0)
        z3_syntax, z3_syntax_rounded, syn_first, syn_second, conv_w = syn

        Cc = N - M
        if "context" in overrides:
            ctx, ctx_rs, ctx_sig_off = overrides["context"](self, y.permute(0, 3, 1, 2), h2.permute(0, 3, 1, 2))
        else:
            ctx = self.prediction_model.raw_tc(y_round_bf16, h2, M)                 # :784  (P,1,2,Cp): mu | log sigma
            ctx_rs, ctx_sig_off = 2 * ctx.shape[-1], ctx.shape[-1]
        lik_y = torch.empty(B, h, w, Cc, dtype=torch.float32, device=dev) if want_likelihoods else None
        ops.likelihood_rows(y, P, Cc, v_rs=N, v_off=M, mu=ctx, mu_mode=2, mu_rs=ctx_rs, mu_off=0,              # :786
                            sigma=ctx, sigma_mode=2, sigma_rs=ctx_rs, sigma_off=ctx_sig_off, sigma_is_log=True,
                            quant=ops.QUANT_ROUND, lik_bound=self.entropy_bottleneck_z3.likelihood_bound,
                            lik=lik_y, sum_out=bits[1:2])
        # syntax stream: "sigma" <- first return (mu), "mu" <- second (sigma): reference quirk H2
        _, lik_syn, _ = ops.gaussian_likelihood(z3_syntax_rounded.contiguous(), syn_first.contiguous(),
                                                syn_second.contiguous(), want_lik=want_likelihoods,
                                                lik_bound=self.entropy_bottleneck_z3_syntax.likelihood_bound,
                                                sum_out=bits[2:3])

        if fused_ok:             # g_s with batch_conv + squared level error in the last deconv's epilogue
            if gs_body is None:
                gs_body = self.s_model.forward_nhwc_body(y_round_bf16)              # :800
            sq_err, x_hat, xt16 = self.s_model.fused_tail(gs_body, x, conv_w.reshape(B, 3, M),                # :800,:811,:864-868
                                                          want_x_tilde=want_x_hat, want_out=want_xt16)
        else:
            xt16 = self.s_model.forward_nhwc(y_round_bf16)                          # :800   (B,H,W,M) fp32 NHWC
            xf = ops.u8_to_f32_pm1(x) if x.dtype == torch.uint8 else x
            sq_err, x_hat = ops.syntax_conv_mse(xf, xt16, conv_w.reshape(B, 3, M), want_x_tilde=want_x_hat)   # :811,:864-868

        out = {"bits": bits, "sq_err": sq_err}
        if want_x_hat:
            out["x_hat"] = x_hat
        if want_likelihoods:
            out["likelihoods"] = {"y": lik_y.permute(0, 3, 1, 2), "z": lik_z.permute(0, 3, 1, 2), "syntax": lik_syn}
        if per_image_bits:       # sum(ln L) per image and stream (the reference only forms the batch total, :857)
            out["bits_per_image"] = torch.stack([torch.log(l.reshape(B, -1).double()).sum(1)
                                                 for l in (lik_z, lik_y, lik_syn)], 1)
        out["latents"] = {"y": y, "z": z, "h2": h2, "ctx": ctx, "z3_syntax": z3_syntax, "conv_w": conv_w, "xt16": xt16,
                          "ctx_rs": ctx_rs, "ctx_sig_off": ctx_sig_off, "syn_first": syn_first, "syn_second": syn_second}
        return out

    # ---- f4: actual bitstreams (the reference only estimates the rate, model/net.py:856-861) ----------------------------
    def y_streams(self, h: int, w: int, symbols_per_stream: int = 2048) -> int:
        """Streams per image of the content latent: for every column group (y_groups) every row of the latent is cut into
        k equal runs (k the smallest divisor of w that brings a run to about `symbols_per_stream` symbols), so no stream
        crosses a row -- the decoder walks wavefronts (pixel (r, c) at step c + 2r, see `decompress`) and a stream must
        be consumed in order."""
        G = self.y_groups()
        cg = (self.N - self.M) // G
        limit = min(65535, max(cg, symbols_per_stream + symbols_per_stream // 4))
        k = next((d for d in range(1, w + 1) if w % d == 0 and (w // d) * cg <= limit), w)
        return G * h * k

    def y_groups(self) -> int:
        """Column groups of the content streams (LdicRansArgs.col_groups): the channels of a pixel are spread over four
        streams, which the wavefront decoder advances in parallel (176 = 4 x 44 symbols per pixel at N-M = 176)."""
        return 4 if (self.N - self.M) % 4 == 0 else 1

    def entropy_encode(self, out: Dict[str, torch.Tensor], symbols_per_stream: int = 2048) -> Dict[str, "ops.RansStreams"]:
        """rANS-codes the three symbol streams of one rd_forward result with the very (mu, sigma) its likelihoods were
        evaluated with: z = round(z) under the per-channel N(0, sigma_z) (:676,:781), y = the N-M content channels of
        round(y) under the context model's (mu, sigma) (:784-786), syntax = round(z3_syntax) under the syntax prior
        (:789).  One independent bitstream per image and stream; nothing is synchronised here."""
        lat = out["latents"]
        y, z = lat["y"], lat["z"]
        B, h, w, N = y.shape
        M, Cc = self.M, self.N - self.M
        hz, wz = z.shape[1], z.shape[2]
        sps = symbols_per_stream
        sigma_z = self.z2_sigma.detach().reshape(N).contiguous()
        enc = {}
        enc["z"] = ops.rans_encode_rows(z, B * hz * wz, N, hz * wz, v_rs=N, sigma=sigma_z, sigma_mode=1,
                                        streams=ops.rans_streams_for(hz * wz * N, sps))
        enc["y"] = ops.rans_encode_rows(y, B * h * w, Cc, h * w, v_rs=N, v_off=M, mu=lat["ctx"], mu_mode=2, mu_rs=lat["ctx_rs"],
                                        sigma=lat["ctx"], sigma_mode=2, sigma_rs=lat["ctx_rs"], sigma_off=lat["ctx_sig_off"],
                                        sigma_is_log=True, streams=self.y_streams(h, w, sps), col_groups=self.y_groups())
        enc["syntax"] = ops.rans_encode(lat["z3_syntax"].reshape(B, -1, 1, 1), lat["syn_first"].reshape(B, -1, 1, 1),
                                        lat["syn_second"].reshape(B, -1, 1, 1), streams=1)
        return enc

    def compress(self, inputs: torch.Tensor, symbols_per_stream: int = 2048):
        """x -> per-image bitstreams: a list of {"z": bytes, "y": bytes, "syntax": bytes} and the dict
        {"bpp_coded", "bpp_estimated"} (coded = 8 * bytes over the batch's pixels; estimated = the reference's
        sum(-log2 L) figure of the same forward, model/net.py:856-861)."""
        out = self.rd_forward(inputs)
        with torch.cuda.device(inputs.device):
            enc = self.entropy_encode(out, symbols_per_stream)
            parts = dict(zip(enc.keys(), ops.rans_tobytes(enc.values())))
        B, _, H, W = inputs.shape
        streams = [{k: parts[k][b] for k in parts} for b in range(B)]
        coded = 8.0 * sum(len(v) for s in streams for v in s.values()) / (B * H * W)
        est = float(out["bits"].double().sum().item()) / (-math.log(2.0) * B * H * W)
        return streams, {"bpp_coded": coded, "bpp_estimated": est}

    def decode_z(self, z_streams, batch: int, H: int, W: int, symbols_per_stream: int = 2048) -> torch.Tensor:
        """The hyper-latent symbols (B, H/64, W/64, N) fp32 NHWC from their bitstreams alone (the prior is a model
        parameter): the first step of a decoder; h_s of the result reproduces the encoder's h2 bit for bit."""
        N = self.N
        hz, wz = H // 64, W // 64
        dev = self.z2_sigma.device
        with torch.cuda.device(dev):
            z_hat = torch.empty(batch, hz, wz, N, dtype=torch.float32, device=dev)
            sigma_z = self.z2_sigma.detach().reshape(N).contiguous()
            ops.rans_decode_rows(z_streams, batch * hz * wz, N, hz * wz, z_hat, v_hat_rs=N, sigma=sigma_z, sigma_mode=1,
                                 streams=ops.rans_streams_for(hz * wz * N, symbols_per_stream))
        return z_hat

    def decode_y(self, y_streams, ctx: torch.Tensor, ctx_rs: int, ctx_sig_off: int, batch: int, H: int, W: int,
                 symbols_per_stream: int = 2048) -> torch.Tensor:
        """The content symbols (B, H/16, W/16, N-M) fp32 NHWC given the context model's (mu | log sigma) tensor.  The
        reference's context model is causal over y_hat (BlockSample with masked=True, model/net.py:219-242), so a real
        decoder alternates it with this call over wavefronts; here the whole tensor is supplied at once."""
        Cc = self.N - self.M
        h, w = H // 16, W // 16
        with torch.cuda.device(ctx.device):
            y_hat = torch.empty(batch, h, w, Cc, dtype=torch.float32, device=ctx.device)
            ops.rans_decode_rows(y_streams, batch * h * w, Cc, h * w, y_hat, v_hat_rs=Cc, mu=ctx, mu_mode=2, mu_rs=ctx_rs,
                                 sigma=ctx, sigma_mode=2, sigma_rs=ctx_rs, sigma_off=ctx_sig_off, sigma_is_log=True,
                                 streams=self.y_streams(h, w, symbols_per_stream), col_groups=self.y_groups())
        return y_hat

    def _wavefront_table(self, h: int, w: int, dev) -> torch.Tensor:
        """[T, h, G, 2] int32 (first symbol, count) per column group of the pixel (r, t - 2r) each row r decodes at step t
        (count 0: none)."""
        G = self.y_groups()
        cg = (self.N - self.M) // G
        T = w + 2 * (h - 1)
        table = torch.zeros(T, h, G, 2, dtype=torch.int32)
        for t in range(T):
            for r in range(max(0, (t - w + 2) // 2), min(h - 1, t // 2) + 1):
                for g in range(G):                      # group g of the segment starts at symbol g * (h * w * cg)
                    table[t, r, g, 0] = g * h * w * cg + (r * w + (t - 2 * r)) * cg
                    table[t, r, g, 1] = cg
        return table.to(dev)

    def _decode_y_steps(self, data, h2, table, B, h, w, symbols_per_stream, schedule):
        """The wavefront loop on `data` (RansStreams or list of bytes) -> (y_hat fp32 [B,h,w,Cc], y_hat bf16 [B,h,w,N],
        RansDecoder).  No synchronisation: capturable in a CUDA graph when `data` is device resident."""
        N, M, Cc = self.N, self.M, self.N - self.M
        dev = h2.device
        T = w + 2 * (h - 1)
        y_hat = torch.zeros(B, h, w, Cc, dtype=torch.float32, device=dev)
        y_hat_bf16 = torch.zeros(B, h, w, N, dtype=torch.bfloat16, device=dev)
        G = self.y_groups()
        dec = ops.RansDecoder(data, B * h * w, Cc, h * w, streams=self.y_streams(h, w, symbols_per_stream), device=dev,
                              col_groups=G)
        if schedule == "full":
            for t in range(T):
                r0, r1 = max(0, (t - w + 2) // 2), min(h - 1, t // 2)
                ctx = self.prediction_model.raw_tc(y_hat_bf16, h2, M)
                rs, so = 2 * ctx.shape[-1], ctx.shape[-1]
                dec.decode(table[t, r0:r1 + 1], (r1 - r0 + 1) * G, y_hat, v_hat_rs=Cc, v_hat_bf16=y_hat_bf16, vb_rs=N, vb_off=M,
                           mu=ctx, mu_mode=2, mu_rs=rs, sigma=ctx, sigma_mode=2, sigma_rs=rs, sigma_off=so, sigma_is_log=True)
        elif schedule == "band":
            # sheared image: pixel (r, c) lives in column c + 2r + 8 (8 = the reach of the taps to the left), so the
            # wavefront of step t is column t + 8 and its patches lie in columns [t, t + 10)
            SH, LEFT, BAND = 2, 8, 10
            w2 = w + SH * (h - 1) + LEFT + 1
            flat = ops.ctx_pack_input(y_hat_bf16, h2)                                        # [B,h,w,2N]: zeros | bf16(h2)
            x_s = torch.zeros(B, h, w2, 2 * N, dtype=torch.bfloat16, device=dev)
            for r in range(h):
                x_s[:, r, LEFT + SH * r:LEFT + SH * r + w] = flat[:, r]
            pix = torch.arange(B * h * w, device=dev, dtype=torch.int64)
            row_of = pix // w                                                                # b * h + r
            prow = row_of.to(torch.int32)
            vb_map = (row_of * w2 + (pix % w) + SH * (row_of % h) + LEFT).to(torch.int32)
            L = self.prediction_model.plan(N, M)
            L0s = self.prediction_model.sheared_conv1(N, M, SH)
            for t in range(T):
                r0, r1 = max(0, (t - w + 2) // 2), min(h - 1, t // 2)
                o1 = L0s.column_of_band(x_s[:, :, t:t + BAND], LEFT)                         # (B*h, 4, 4, N): the wavefront's pixels
                ctx = L[3](L[2](L[1](o1)))                                                   # (B*h, 1, 2, Cp)
                rs, so = 2 * ctx.shape[-1], ctx.shape[-1]
                dec.decode(table[t, r0:r1 + 1], (r1 - r0 + 1) * G, y_hat, v_hat_rs=Cc, v_hat_bf16=x_s, vb_rs=2 * N, vb_off=M,
                           param_row_map=prow, bf16_row_map=vb_map, mu=ctx, mu_mode=2, mu_rs=rs, sigma=ctx, sigma_mode=2,
                           sigma_rs=rs, sigma_off=so, sigma_is_log=True)
            y_hat_bf16[..., M:] = y_hat.to(torch.bfloat16)
        else:
            raise ops.LdicError("decompress: schedule must be 'band' or 'full'")
        return y_hat, y_hat_bf16, dec

    def _decode_y_wavefronts(self, y_blobs, h2, B, h, w, symbols_per_stream, schedule, use_graph):
        """Eager the first time a shape is seen; from the second time on the ~800 launches of the loop are one CUDA graph
        over static buffers (bitstreams, h2 in; y_hat out), keyed by shape and parameter versions like Net.forward's."""
        dev = h2.device
        Cc = self.N - self.M
        key = (dev.index, B, h, w, symbols_per_stream, schedule, tuple((p._version, p.data_ptr()) for p in self.parameters()))
        ent = self._decode_graphs.get(key) if use_graph else None
        if not use_graph or ent is None:
            table = self._wavefront_table(h, w, dev)
            y_hat, y_hat_bf16, dec = self._decode_y_steps(y_blobs, h2, table, B, h, w, symbols_per_stream, schedule)
            dec.finish()
            if use_graph:
                if len(self._decode_graphs) >= 4:
                    self._decode_graphs.clear()
                self._decode_graphs[key] = False
            return y_hat, y_hat_bf16
        if ent is False:                        # second sight: capture
            S = self.y_streams(h, w, symbols_per_stream)
            n = h * w * Cc
            cap = (int(ops._L().ldic_rans_max_bytes(n, S)) + 3) & ~3
            st = {"buf": torch.zeros(B, cap, dtype=torch.uint8, device=dev),
                  "sizes": torch.zeros(B, dtype=torch.int32, device=dev), "h2": torch.empty_like(h2),
                  "table": self._wavefront_table(h, w, dev), "cap": cap}
            data = ops.RansStreams(st["buf"], st["sizes"], None, S, n, ops.QUANT_ROUND)
            st["h2"].copy_(h2)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):       # warm-up on the static buffers (empty streams: every image reports a bad header)
                self._decode_y_steps(data, st["h2"], st["table"], B, h, w, symbols_per_stream, schedule)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                st["y_hat"], st["y_hat_bf16"], st["dec"] = self._decode_y_steps(data, st["h2"], st["table"], B, h, w,
                                                                                symbols_per_stream, schedule)
            st["graph"] = g
            ent = self._decode_graphs[key] = st
        sizes = [len(b) for b in y_blobs]
        if max(sizes) > ent["cap"]:
            raise ops.LdicError("decompress: a content bitstream is larger than any encoder output for this shape")
        host = torch.zeros(B, (max(sizes) + 3) & ~3, dtype=torch.uint8)
        for i, b in enumerate(y_blobs):
            if len(b):
                host[i, :len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
        ent["buf"][:, :host.shape[1]].copy_(host)
        ent["sizes"].copy_(torch.tensor(sizes, dtype=torch.int32))
        ent["h2"].copy_(h2)
        ent["graph"].replay()
        ent["dec"].finish()
        return ent["y_hat"].clone(), ent["y_hat_bf16"].clone()

    @torch.no_grad()
    def decompress(self, streams, H: int, W: int, symbols_per_stream: int = 2048, want_latents: bool = False,
                   schedule: str = "band", use_graph: bool = True):
        """The decoder: per-image {"z","y","syntax"} byte strings (Net.compress) -> x_hat (B,3,H,W), bit-identical to the
        encoder's reconstruction (rd_forward(want_x_hat=True)["x_hat"]), from the bytes and the model alone.
          z       from its stream under the factorised prior; h2 = h_s(z^)                       (model/net.py:676-681)
          syntax  prior from h2 (PredictionModel_Syntax, :789), symbols -> conv_generator (:805)
          y       the context model is causal over y^ (BlockSample masked=True, :219-242: pixel (r, c) sees rows r-3..r-1
                  at columns c-2..c+1 and (r, c-2), (r, c-1)), so pixels are decoded along wavefronts t = c + 2r: at every
                  step the context model runs on the partially decoded latent (the very kernels of the encoder, hence
                  the very (mu, sigma)) and the rANS decoder advances the streams of the pixels on the wavefront.
          x_hat   g_s + IGDN + batch_conv (:800-811)
        schedule="full" runs the context model over the whole latent at each of the w + 2(h-1) steps (110 for 768x512);
        schedule="band" (default) keeps [round(y) | h2] in an image SHEARED by two columns per row, where a wavefront is
        one column: per step the four context layers run on the wavefront's pixels only, the first one reading the
        10-column band its taps reach (LDIC_CTX_CONV1 with sheared tap offsets and a band origin).  Both give the encoder's
        (mu, sigma) bit for bit: an output pixel of these kernels depends on its own inputs in a fixed order."""
        B = len(streams)
        N, M, Cc = self.N, self.M, self.N - self.M
        if H % 64 or W % 64:
            raise ops.LdicError("H and W must be multiples of 64")
        h, w = H // 16, W // 16
        dev = self.z2_sigma.device
        with torch.cuda.device(dev):
            z_hat = self.decode_z([s["z"] for s in streams], B, H, W, symbols_per_stream)
            h2 = self.hs_model.forward_nhwc(z_hat.to(torch.bfloat16))
            no_y = torch.zeros(B, h, w, N, dtype=torch.float32, device=dev)
            mods = (self.syntax_model, self.prediction_model_syntax, self.conv_weights_gen)
            _, _, syn_first, syn_second, _ = ops.syntax_branch(no_y, h2, M, *mods)          # the prior needs h2 only
            z3r = ops.rans_decode([s["syntax"] for s in streams], (B, M, 1, 1), syn_first.reshape(B, -1, 1, 1),
                                  syn_second.reshape(B, -1, 1, 1), streams=1)
            conv_w = ops.syntax_branch(no_y, h2, M, *mods, z3_round_in=z3r)[4]
            y_hat, y_hat_bf16 = self._decode_y_wavefronts([s["y"] for s in streams], h2, B, h, w, symbols_per_stream,
                                                          schedule, use_graph)
            body = self.s_model.forward_nhwc_body(y_hat_bf16)
            blank = torch.zeros(B, 3, H, W, dtype=torch.uint8, device=dev)
            if self.s_model.has_fused_tail():
                _, x_hat, _ = self.s_model.fused_tail(body, blank, conv_w.reshape(B, 3, M), want_x_tilde=True)
            else:
                _, x_hat = ops.syntax_conv_mse(ops.u8_to_f32_pm1(blank), self.s_model.plan()[-1](body), conv_w.reshape(B, 3, M),
                                               want_x_tilde=True)
        if want_latents:
            return x_hat, {"z_hat": z_hat, "h2": h2, "y_hat": y_hat, "z3_round": z3r, "conv_w": conv_w}
        return x_hat

    def metrics(self, out: Dict[str, torch.Tensor], batch: int, H: int, W: int):
        """bpp / v_mse / v_psnr exactly as model/net.py:856-869 forms them."""
        tb, th, tw, tc = self.test_size
        packed, v_mse = ops.rd_pack_metrics(out["bits"], out["sq_err"], 3 * H * W)
        r = ops.rd_finish_metrics(packed, float(th * tw))
        return r[0], v_mse, r[1]

    def forward(self, inputs, mode='train', num=1):
        """model/net.py:539-871 in test mode: (bpp, v_mse[B], v_psnr).

        With `auto_graph` (default) the ~35 launches of a step are captured into a CUDA graph the second time an input
        shape is seen and replayed afterwards (the input is copied into the graph's static buffer first): the host
        cost of a step drops from ~35 ctypes calls to one copy and one graph launch.  The returned tensors of a
        replayed step are clones, so they stay valid across calls like the eager ones."""
        if mode != 'test':
            raise NotImplementedError("only the rate-distortion forward (mode='test') is implemented; training is out of scope")
        if not inputs.is_cuda:
            raise ops.LdicError("Net runs on CUDA only (no CPU fallback)")
        if self.auto_graph and not torch.cuda.is_current_stream_capturing() and ops.PROFILE is None:
            r = self._graph_forward(inputs)
            if r is not None:
                return r
        out = self.rd_forward(inputs)
        return self.metrics(out, inputs.shape[0], inputs.shape[2], inputs.shape[3])

    def _graph_key(self, inputs):
        return (inputs.device.index, inputs.dtype, tuple(inputs.shape), tuple(self.test_size), self.tail_fused, self.side_sms,
                self.a_model.precision,
                tuple((p._version, p.data_ptr()) for p in self.parameters()))

    def _graph_forward(self, inputs):
        from .graph import GraphedEvaluator
        key = self._graph_key(inputs)
        ent = self._graphs.get(key)
        if ent is None:                      # first sight of this shape: run eagerly, remember it
            if len(self._graphs) >= 8:
                self._graphs.clear()
            self._graphs[key] = False
            return None
        with torch.cuda.device(inputs.device):
            if ent is False:                 # second sight: capture
                static = inputs.contiguous().clone()
                ent = self._graphs[key] = GraphedEvaluator(self, [static])
            ent.inputs[0].copy_(inputs)
            bpp, psnr, out = ent(0)
            return bpp.clone(), out["v_mse"].clone(), psnr.clone()

    def forward_dict(self, inputs):
        """CompressAI-style view ({'x_hat', 'likelihoods': {'y','z'}}) that RateDistortionLoss
        (train_net_unet.py:62-87) consumes."""
        out = self.rd_forward(inputs, want_x_hat=True, want_likelihoods=True)
        return {"x_hat": out["x_hat"], "likelihoods": {"y": out["likelihoods"]["y"], "z": out["likelihoods"]["z"]}}

"""CPU oracle for the rate-distortion forward path (TEST INFRASTRUCTURE ONLY).

This module is a functional, state-dict driven restatement in plain torch-CPU
fp32 of the reference's hot path.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under the product package
``ldic_b200`` imports from ``oracle/``.

Parity status: PINNED.  Every function below is compared in
``tests/test_oracle_golden.py`` against fixtures under ``tests/golden/`` that
were produced by executing the unmodified reference classes
(``tests/golden/make_golden.py``, run in the build container where
``/root/reference`` exists), plus the reference's single known-answer block
(``ops/parametrizers.py:52-58``).  The one exception is
``gaussian_conditional`` (CompressAI ``GaussianConditional``): that class is a
third-party dependency which the reference neither vendors nor pins, so that
function is a restatement of the upstream published semantics anchored on the
reference call sites (``model/Net_unet.py:805,1057``) -- "parity unpinned" for
that single function.

All ``file:line`` citations are relative to the reference repository root.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------
# L1 ops
# ----------------------------------------------------------------------------


def lower_bound(x: Tensor, bound) -> Tensor:
    """max(x, bound).  ops/bound_ops.py:21-22, model/gdn.py:13-17."""
    b = bound if isinstance(bound, Tensor) else torch.tensor(float(bound), dtype=x.dtype)
    return torch.max(x, b.to(x.dtype))


def lower_bound_grad(x: Tensor, bound: float, grad_out: Tensor) -> Tensor:
    """Pass-through gradient iff x>=bound or grad<0.  ops/bound_ops.py:25-27."""
    keep = (x >= bound) | (grad_out < 0)
    return keep.to(grad_out.dtype) * grad_out


def parametrizer_constants(minimum: float = 0.0, reparam_offset: float = 2 ** -18) -> Tuple[float, float]:
    """(bound, pedestal) the way ops/parametrizers.py:32-41 builds them:
    python-double arithmetic, then stored in fp32 buffers."""
    pedestal = float(reparam_offset) ** 2
    bound = (float(minimum) + float(reparam_offset) ** 2) ** 0.5
    f32 = lambda v: float(torch.tensor([v], dtype=torch.float32)[0])
    return f32(bound), f32(pedestal)


def nonneg_init(x: Tensor, reparam_offset: float = 2 ** -18) -> Tensor:
    """NonNegativeParametrizer.init  ops/parametrizers.py:43-44."""
    pedestal = torch.tensor([float(reparam_offset) ** 2], dtype=torch.float32)
    return torch.sqrt(torch.max(x + pedestal, pedestal))


def nonneg_forward(p: Tensor, minimum: float = 0.0, reparam_offset: float = 2 ** -18) -> Tensor:
    """NonNegativeParametrizer.forward  ops/parametrizers.py:46-49."""
    bound, pedestal = parametrizer_constants(minimum, reparam_offset)
    out = lower_bound(p, bound)
    return out ** 2 - torch.tensor(pedestal, dtype=torch.float32)


def model_gdn_constants(beta_min: float = 1e-6, reparam_offset: float = 2 ** -18) -> Tuple[float, float, float]:
    """(beta_bound, gamma_bound, pedestal) the way model/gdn.py:43-53 builds
    them: fp32 *tensor* arithmetic (differs in the last ulp from the python
    double route of ops/parametrizers.py)."""
    off = torch.tensor([reparam_offset], dtype=torch.float32)
    pedestal = off ** 2
    beta_bound = (beta_min + off ** 2) ** 0.5
    return float(beta_bound[0]), float(off[0]), float(pedestal[0])


def ste_round(x: Tensor) -> Tensor:
    """round(x) - x + x, evaluated left to right in fp32.  ops/ops.py:34,
    model/Net_unet.py:778.  (Not always bit-identical to round(x).)"""
    return torch.round(x) - x + x


def bypass_round(x: Tensor) -> Tensor:
    """torch.round, half-to-even.  model/net.py:416-426."""
    return torch.round(x)


# ----------------------------------------------------------------------------
# GDN / IGDN
# ----------------------------------------------------------------------------


def gdn_effective_params_model(beta_p: Tensor, gamma_p: Tensor, beta_min: float = 1e-6,
                               reparam_offset: float = 2 ** -18) -> Tuple[Tensor, Tensor]:
    """beta/gamma reparametrisation of model/gdn.py:75-81 (and 140-146)."""
    bb, gb, ped = model_gdn_constants(beta_min, reparam_offset)
    ped_t = torch.tensor(ped, dtype=torch.float32)
    beta = lower_bound(beta_p, bb) ** 2 - ped_t
    gamma = lower_bound(gamma_p, gb) ** 2 - ped_t
    return beta, gamma


def gdn_effective_params_layers(beta_p: Tensor, gamma_p: Tensor, beta_min: float = 1e-6) -> Tuple[Tensor, Tensor]:
    """layers/gdn.py:65-66 via ops/parametrizers.py:46-49."""
    return nonneg_forward(beta_p, minimum=beta_min), nonneg_forward(gamma_p, minimum=0.0)


def gdn_model(x: Tensor, beta_p: Tensor, gamma_p: Tensor, inverse: bool = False) -> Tensor:
    """model.gdn.GDN.forward (x / sqrt(norm)) model/gdn.py:69-92 and
    model.gdn.IGDN.forward (x * sqrt(norm)) model/gdn.py:134-156."""
    C = x.shape[1]
    beta, gamma = gdn_effective_params_model(beta_p, gamma_p)
    norm = F.conv2d(x ** 2, gamma.view(C, C, 1, 1), beta)
    norm = torch.sqrt(norm)
    return x * norm if inverse else x / norm


def gdn_layers(x: Tensor, beta_p: Tensor, gamma_p: Tensor, inverse: bool = False,
               beta_min: float = 1e-6) -> Tensor:
    """layers.gdn.GDN.forward (x * rsqrt / x * sqrt)  layers/gdn.py:62-75."""
    C = x.shape[1]
    beta, gamma = gdn_effective_params_layers(beta_p, gamma_p, beta_min)
    norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


# ----------------------------------------------------------------------------
# Likelihood models, bpp, PSNR
# ----------------------------------------------------------------------------


def _std_normal_cdf(t: Tensor) -> Tensor:
    """torch.distributions.Normal(0,1).cdf as the reference calls it
    (model/net.py:277-278): 0.5 * (1 + erf(t / sqrt(2)))."""
    return 0.5 * (1 + torch.erf((t - 0.0) * 1.0 / math.sqrt(2)))


def gaussian_model_likelihood(v: Tensor, sigma: Tensor, mu: Tensor, bound: float = 1e-8) -> Tensor:
    """GaussianModel.forward(inputs, hyper_sigma, hyper_mu)  model/net.py:272-286
    (bound 1e-8); U-Net-family copy model/Net_unet.py:588-604 (bound 1e-12).
    No sigma bound, no abs: negative/zero sigma propagate exactly as in the
    reference."""
    half = 0.5
    upper = (v - mu + half) / sigma
    lower = (v - mu - half) / sigma
    res = _std_normal_cdf(upper) - _std_normal_cdf(lower)
    return torch.clamp(res, min=bound)


def gaussian_conditional(y: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                         scale_bound: float = 0.11, likelihood_bound: float = 1e-9,
                         ) -> Tuple[Tensor, Tensor]:
    """CompressAI ``GaussianConditional(None).forward(y, scales, means)`` in eval
    mode, restated from the upstream published semantics (third-party, NOT in
    /root/reference, version unpinned; call sites model/Net_unet.py:805,1057,
    model/net_unet_ha_hs.py:669,937).  PARITY UNPINNED for this function.

    y_hat = round(y - mu) + mu; v = |y_hat - mu|; s = max(scales, 0.11);
    L = max(0.5 erfc(-(0.5 - v)/(s sqrt2)) - 0.5 erfc(-(-0.5 - v)/(s sqrt2)), 1e-9)
    LowerBound itself is in-tree: ops/bound_ops.py:21-65.
    """
    if means is None:
        means = torch.zeros_like(y)
    y_hat = torch.round(y - means) + means
    values = torch.abs(y_hat - means)
    s = lower_bound(scales, scale_bound)
    const = float(-(2 ** -0.5))

    def std_cum(t):
        return 0.5 * torch.erfc(const * t)

    upper = std_cum((0.5 - values) / s)
    lower = std_cum((-0.5 - values) / s)
    lik = lower_bound(upper - lower, likelihood_bound)
    return y_hat, lik


def bits_from_likelihoods(*liks: Tensor) -> Tensor:
    """Per-stream sum(ln L) as model/net.py:857 does it (torch.sum over all dims)."""
    return torch.stack([torch.sum(torch.log(l), [0, 1, 2, 3]) for l in liks])


def bpp_from_likelihoods(liks, num_pixels: int) -> Tensor:
    """model/net.py:856-859: sum_l sum(log l)/(-ln2 * B*th*tw), one scalar."""
    parts = [torch.sum(torch.log(l), [0, 1, 2, 3]) / (-np.log(2) * num_pixels) for l in liks]
    out = parts[0]
    for p in parts[1:]:
        out = out + p
    return out


def mse_psnr(x: Tensor, x_tilde: Tensor, clamp_pm1: bool = False) -> Tuple[Tensor, Tensor]:
    """model/net.py:864-869 (U-Net family adds clamp(x_tilde,-1,1) first,
    model/net_unet_ha_hs.py:1006)."""
    if clamp_pm1:
        x_tilde = torch.clamp(x_tilde, -1, 1)
    gt = torch.round((x + 1) * 127.5)
    x_hat = torch.clamp((x_tilde + 1) * 127.5, 0, 255)
    x_hat = torch.round(x_hat).float()
    v_mse = torch.mean((x_hat - gt) ** 2, [1, 2, 3])
    v_psnr = torch.mean(20 * torch.log10(255 / torch.sqrt(v_mse)), 0)
    return v_mse, v_psnr


# ----------------------------------------------------------------------------
# Transforms (state-dict driven; key names are the reference's, Appendix C)
# ----------------------------------------------------------------------------


def analysis_transform(sd: Dict[str, Tensor], x: Tensor, prefix: str = "a_model.transform.") -> Tensor:
    """g_a: 4x [ZeroPad2d((1,2,1,2)) -> Conv2d(k5,s2,p0)], GDN after convs 1-3.
    model/net.py:96-114."""
    for conv_i, gdn_i in ((1, 2), (4, 5), (7, 8), (10, None)):
        x = F.pad(x, (1, 2, 1, 2))
        x = F.conv2d(x, sd[f"{prefix}{conv_i}.weight"], sd[f"{prefix}{conv_i}.bias"], stride=2)
        if gdn_i is not None:
            x = gdn_model(x, sd[f"{prefix}{gdn_i}.beta"], sd[f"{prefix}{gdn_i}.gamma"], inverse=False)
    return x


def synthesis_transform(sd: Dict[str, Tensor], y_hat: Tensor, prefix: str = "s_model.transform.") -> Tensor:
    """g_s: 4x [ZeroPad2d((1,0,1,0)) -> ConvTranspose2d(k5,s2,p3,op1) -> IGDN].
    model/net.py:126-144."""
    x = y_hat
    for conv_i, gdn_i in ((1, 2), (4, 5), (7, 8), (10, 11)):
        x = F.pad(x, (1, 0, 1, 0))
        x = F.conv_transpose2d(x, sd[f"{prefix}{conv_i}.weight"], sd[f"{prefix}{conv_i}.bias"],
                               stride=2, padding=3, output_padding=1)
        x = gdn_model(x, sd[f"{prefix}{gdn_i}.beta"], sd[f"{prefix}{gdn_i}.gamma"], inverse=True)
    return x


def h_analysis_transform(sd: Dict[str, Tensor], y: Tensor, prefix: str = "ha_model.transform.") -> Tensor:
    """h_a: abs -> Conv3x3 s1 p1 -> ReLU -> Conv5x5 s2 p2 -> ReLU -> Conv5x5 s2 p2.
    model/net.py:188-199."""
    x = torch.abs(y)
    x = F.relu(F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], stride=1, padding=1))
    x = F.relu(F.conv2d(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"], stride=2, padding=2))
    return F.conv2d(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"], stride=2, padding=2)


def h_synthesis_transform(sd: Dict[str, Tensor], z_hat: Tensor, prefix: str = "hs_model.transform.") -> Tensor:
    """h_s: ConvT5 s2 p2 op1 -> ReLU -> ConvT5 s2 p2 op1 -> ReLU -> ConvT3 s1 p1.
    model/net.py:206-216."""
    x = F.relu(F.conv_transpose2d(z_hat, sd[prefix + "0.weight"], sd[prefix + "0.bias"],
                                  stride=2, padding=2, output_padding=1))
    x = F.relu(F.conv_transpose2d(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"],
                                  stride=2, padding=2, output_padding=1))
    return F.conv_transpose2d(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"], stride=1, padding=1)


# ----------------------------------------------------------------------------
# Context / syntax branches (on the forward, "next" rows for kernels: SURVEY 8 f1)
# ----------------------------------------------------------------------------


def block_sample_filter(dim: int, masked: bool) -> Tensor:
    """One-hot 7x7 sampling filter of BlockSample.__init__  model/net.py:223-235."""
    flt = np.zeros((dim * 16, dim, 7, 7), dtype=np.float32)
    for i in range(4):
        for j in range(4):
            if masked and i == 3 and j >= 2:
                break
            for k in range(dim):
                flt[k * 16 + i * 4 + j, k, i, j + 1] = 1
    return torch.from_numpy(flt)


def block_sample_onehot(x: Tensor, masked: bool, flt: Optional[Tensor] = None) -> Tensor:
    """BlockSample.forward exactly as written (one-hot conv2d)  model/net.py:237-242."""
    b, c, h, w = x.shape
    if flt is None:
        flt = block_sample_filter(c, masked)
    t = F.conv2d(x, flt, padding=3)
    t = t.contiguous().view(b, c, 4, 4, h, w).permute(0, 4, 5, 1, 2, 3)
    return t.contiguous().view(b * h * w, c, 4, 4)


def block_sample_gather(x: Tensor, masked: bool) -> Tensor:
    """Same result as block_sample_onehot via pad + slice: patch cell (i,j) of
    position (y,x) is input[y+i-3, x+j-2]; the masked (y) sampler zeroes
    (i=3, j in {2,3}).  Bit-identical to the one-hot conv (tested)."""
    b, c, h, w = x.shape
    xp = F.pad(x, (2, 1, 3, 0))
    out = torch.zeros(b, h, w, c, 4, 4, dtype=x.dtype)
    for i in range(4):
        for j in range(4):
            if masked and i == 3 and j >= 2:
                continue
            out[:, :, :, :, i, j] = xp[:, :, i:i + h, j:j + w].permute(0, 2, 3, 1)
    return out.view(b * h * w, c, 4, 4)


def prediction_context(sd: Dict[str, Tensor], y_rounded: Tensor, h_tilde: Tensor,
                       faithful_sampler: bool = False,
                       prefix: str = "prediction_model.") -> Tuple[Tensor, Tensor]:
    """PredictionModel_Context.forward  model/net.py:305-319.  Returns (mu, sigma)
    as NCHW *views* of (b,h,w,c) storage, like the reference."""
    b, c, h, w = y_rounded.shape
    sample = block_sample_onehot if faithful_sampler else block_sample_gather
    merged = torch.cat([sample(y_rounded, True), sample(h_tilde, False)], 1)
    t = F.leaky_relu(F.conv2d(merged, sd[prefix + "transform.0.weight"], sd[prefix + "transform.0.bias"], 1, 1), 0.2)
    t = F.leaky_relu(F.conv2d(t, sd[prefix + "transform.2.weight"], sd[prefix + "transform.2.bias"], 2, 1), 0.2)
    t = F.leaky_relu(F.conv2d(t, sd[prefix + "transform.4.weight"], sd[prefix + "transform.4.bias"], 1, 1), 0.2)
    t = F.linear(t.flatten(1), sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])
    mu = t[:, :c].view(b, h, w, c).permute(0, 3, 1, 2)
    sigma = torch.exp(t[:, c:]).contiguous().view(b, h, w, c).permute(0, 3, 1, 2)
    return mu, sigma


def syntax_model(sd: Dict[str, Tensor], syntax: Tensor, prefix: str = "syntax_model.") -> Tensor:
    """Syntax_Model.forward  model/net.py:359-375."""
    pool = lambda t: F.adaptive_avg_pool2d(t, 1)
    out1 = pool(syntax)
    ds1 = F.relu(F.conv2d(syntax, sd[prefix + "down0.weight"], sd[prefix + "down0.bias"], 2, 1))
    out2 = pool(ds1)
    ds2 = F.relu(F.conv2d(ds1, sd[prefix + "down1.weight"], sd[prefix + "down1.bias"], 2, 1))
    out3 = pool(ds2)
    out = torch.cat((out1, out2, out3), 1)
    return F.conv2d(out, sd[prefix + "conv.weight"], sd[prefix + "conv.bias"])


def prediction_syntax(sd: Dict[str, Tensor], y_rounded: Tensor, h_tilde: Tensor,
                      prefix: str = "prediction_model_syntax.") -> Tuple[Tensor, Tensor]:
    """PredictionModel_Syntax.forward  model/net.py:391-413.  Returns (mu, sigma)
    in that order -- the caller at model/net.py:789 binds them swapped."""
    b, c, h, w = y_rounded.shape
    pool = lambda t: F.adaptive_avg_pool2d(t, 1)
    ds0 = F.relu(F.conv2d(h_tilde, sd[prefix + "down0.weight"], sd[prefix + "down0.bias"], 2, 1))
    ds1 = F.relu(F.conv2d(ds0, sd[prefix + "down1.weight"], sd[prefix + "down1.bias"], 2, 1))
    ctx = torch.cat((pool(h_tilde), pool(ds0), pool(ds1)), 1).flatten(1)
    t = F.linear(ctx, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])
    mu = t[:, :c].view(b, h, w, c).permute(0, 3, 1, 2)
    sigma = torch.exp(t[:, c:]).contiguous().view(b, h, w, c).permute(0, 3, 1, 2)
    return mu, sigma


def conv_generator(sd: Dict[str, Tensor], x: Tensor, out_dim: int, prefix: str = "conv_weights_gen.") -> Tensor:
    """conv_generator.forward  model/net.py:336-343."""
    b = x.shape[0]
    t = x.view(b, -1)
    t = F.leaky_relu(F.linear(t, sd[prefix + "transform.0.weight"], sd[prefix + "transform.0.bias"]), 0.2)
    t = F.leaky_relu(F.linear(t, sd[prefix + "transform.2.weight"], sd[prefix + "transform.2.bias"]), 0.2)
    t = F.linear(t, sd[prefix + "transform.4.weight"], sd[prefix + "transform.4.bias"])
    return t.view(b, 3, out_dim, 1, 1)


def batch_conv(weights: Tensor, inputs: Tensor) -> Tensor:
    """Per-image 1x1 conv via grouped conv2d.  model/net.py:527-537."""
    b, ch = inputs.shape[:2]
    _, ch_out, _, k, _ = weights.shape
    weights = weights.reshape(b * ch_out, ch, k, k)
    inputs = torch.cat(torch.split(inputs, 1, dim=0), dim=1)
    out = F.conv2d(inputs, weights, stride=1, padding=0, groups=b)
    return torch.cat(torch.split(out, ch_out, dim=1), dim=0)


# ----------------------------------------------------------------------------
# Net.forward(mode='test')  model/net.py:539-871
# ----------------------------------------------------------------------------


def net_forward_test(sd: Dict[str, Tensor], x: Tensor, test_size: Tuple[int, int, int, int],
                     M: int = 16, faithful_sampler: bool = False,
                     return_intermediates: bool = True) -> Dict[str, Tensor]:
    """``Net.forward(inputs, 'test')`` without post-processing (model/net.py:539-871).

    ``test_size`` is the constructor's (tb, th, tw, tc); bpp is normalised by
    ``inputs.size(0) * th * tw`` (model/net.py:856), NOT by the tensor's H, W.
    """
    tb, th, tw, tc = test_size
    z3 = analysis_transform(sd, x)                                   # :627
    z2 = h_analysis_transform(sd, z3)                                # :666
    z2_rounded = bypass_round(z2)                                    # :676
    h2 = h_synthesis_transform(sd, z2_rounded)                       # :681
    z2_sigma = sd["z2_sigma"]                                        # :706
    z2_mu = torch.zeros_like(z2_sigma)                               # :708
    z3_syntax = syntax_model(sd, z3[:, :M])                          # :712-719
    z3_content = z3[:, M:]                                           # :726
    z3_content_rounded = bypass_round(z3_content)                    # :741
    z3_syntax_rounded = bypass_round(z3_syntax)                      # :753
    z2_lik = gaussian_model_likelihood(z2_rounded, z2_sigma, z2_mu)  # :781
    mu, sigma = prediction_context(sd, z3_content_rounded, h2, faithful_sampler)     # :784
    y_lik = gaussian_model_likelihood(z3_content_rounded, sigma, mu)                 # :786
    # names swapped at the call site (:789): "sigma" receives mu and vice versa
    syn_sigma, syn_mu = prediction_syntax(sd, z3_syntax_rounded, h2)
    syn_lik = gaussian_model_likelihood(z3_syntax_rounded, syn_sigma, syn_mu)        # :790
    x_tilde16 = synthesis_transform(sd, z3_content_rounded)                          # :800
    w = conv_generator(sd, z3_syntax_rounded, M)                                     # :805
    x_tilde = batch_conv(w, x_tilde16)                                               # :811
    num_pixels = x.shape[0] * th * tw                                                # :856
    bpp = bpp_from_likelihoods([z2_lik, y_lik, syn_lik], num_pixels)                 # :857-859
    v_mse, v_psnr = mse_psnr(x, x_tilde)                                             # :864-869
    out = {"bpp": bpp, "v_mse": v_mse, "v_psnr": v_psnr}
    if return_intermediates:
        out.update(z3=z3, z2=z2, z2_rounded=z2_rounded, h2=h2, z3_syntax=z3_syntax,
                   z3_content_rounded=z3_content_rounded, z3_syntax_rounded=z3_syntax_rounded,
                   mu=mu, sigma=sigma, z2_lik=z2_lik, y_lik=y_lik, syn_lik=syn_lik,
                   syn_sigma=syn_sigma, syn_mu=syn_mu, x_tilde16=x_tilde16,
                   conv_weights=w, x_tilde=x_tilde,
                   bits=bits_from_likelihoods(z2_lik, y_lik, syn_lik))
    return out


def eval_pad(img: Tensor, multiple: int = 64) -> Tensor:
    """eval_net.py:68-84: pad (3,h,w) in [0,1] with ONES at bottom/right up to a
    multiple of 64, add batch dim, map to [-1,1]."""
    _, h, w = img.shape
    hp = h if h % multiple == 0 else (h // multiple) * multiple + multiple
    wp = w if w % multiple == 0 else (w // multiple) * multiple + multiple
    img = torch.cat((img, torch.ones(3, hp - h, w)), 1)
    img = torch.cat((img, torch.ones(3, hp, wp - w)), 2)
    return img.unsqueeze(0) * 2.0 - 1.0


# ---------------------------------------------------------------------------------------------
# window attention block (layers/win_attention.py:38-209), functional restatement.
# Pinned by tests/golden/win_attention.npz (outputs of the unmodified reference class).
# ---------------------------------------------------------------------------------------------
def win_rel_position_index(ws: int) -> Tensor:
    """layers/win_attention.py:66-76."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def win_shift_mask(H: int, W: int, ws: int, shift: int) -> Tensor:
    """layers/win_attention.py:160-177: (nW, ws*ws, ws*ws) with 0 / -100."""
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def win_based_attention(sd: Dict[str, Tensor], x: Tensor, num_heads: int, ws: int, shift: int, prefix: str = "") -> Tensor:
    """WinBasedAttention.forward (layers/win_attention.py:150-209): x (B,C,H,W) -> x + W-MSA(x)."""
    B, C, H, W = x.shape
    g = lambda k: sd[prefix + k]
    t = x.permute(0, 2, 3, 1)
    if shift > 0:
        t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
    win = t.reshape(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)
    Bn, N, _ = win.shape
    hd = C // num_heads
    qkv = F.linear(win, g("attn.qkv.weight"), g("attn.qkv.bias")).reshape(Bn, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = g("attn.relative_position_index").reshape(-1).long()
    bias = g("attn.relative_position_bias_table")[idx].reshape(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift > 0:
        m = win_shift_mask(H, W, ws, shift)
        nW = m.shape[0]
        attn = (attn.view(Bn // nW, nW, num_heads, N, N) + m.unsqueeze(1).unsqueeze(0)).view(-1, num_heads, N, N)
    attn = torch.softmax(attn, -1)
    o = (attn @ v).transpose(1, 2).reshape(Bn, N, C)
    o = F.linear(o, g("attn.proj.weight"), g("attn.proj.bias"))
    o = o.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    return x + o.permute(0, 3, 1, 2)

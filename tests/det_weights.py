"""Deterministic synthetic weights / inputs shared by the golden generator, the
parity tests, smoke() and bench.py.  Independent of the reference and of torch's
RNG (numpy PCG64), so the same tensors are rebuilt on the GPU box.

Key names and shapes are the reference ``model/net.py`` ``Net`` state-dict
 minus the HAN head and the one-hot sampler buffers (the
reference loads it with strict=False; our Net accepts and discards those keys).

``boost=True`` applies the gain recipe of SURVEY.md H8 so that symbols are not
all zero (default-scale init rounds every latent to 0 and parity would pass with
a broken conv).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch


def _rng(seed: int, tag: str) -> np.random.Generator:
    h = np.frombuffer(tag.encode(), dtype=np.uint8).astype(np.uint64)
    mix = int((h * np.arange(1, len(h) + 1, dtype=np.uint64)).sum() % (2 ** 31))
    return np.random.Generator(np.random.PCG64([seed, mix]))


def _uniform(seed, tag, shape, bound):
    return torch.from_numpy(_rng(seed, tag).uniform(-bound, bound, size=shape).astype(np.float32))


def _conv(sd, seed, name, cout, cin, k, gain=1.0, transposed=False):
    # fan_in as torch computes it: size(1) * receptive field (for ConvTranspose2d
    # the weight is (cin, cout, k, k), so fan_in uses cout -- same as torch's default)
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    fan_in = shape[1] * k * k
    b = 1.0 / math.sqrt(fan_in)
    sd[name + ".weight"] = _uniform(seed, name + ".w", shape, b) * gain
    sd[name + ".bias"] = _uniform(seed, name + ".b", (cout,), b) * gain


def _linear(sd, seed, name, cout, cin, gain=1.0):
    b = 1.0 / math.sqrt(cin)
    sd[name + ".weight"] = _uniform(seed, name + ".w", (cout, cin), b) * gain
    sd[name + ".bias"] = _uniform(seed, name + ".b", (cout,), b) * gain


def _gdn(sd, seed, name, ch):
    off = torch.tensor([2.0 ** -18], dtype=torch.float32)
    ped = off ** 2
    r = _rng(seed, name)
    beta = 1.0 + 0.5 * r.uniform(-1, 1, size=(ch,))
    gamma = 0.1 * np.eye(ch) + 0.02 * np.abs(r.standard_normal(size=(ch, ch))) / math.sqrt(ch / 16.0)
    sd[name + ".beta"] = torch.sqrt(torch.from_numpy(beta.astype(np.float32)) + ped)
    sd[name + ".gamma"] = torch.sqrt(torch.from_numpy(gamma.astype(np.float32)) + ped)
    sd[name + ".reparam_offset"] = off.clone()
    sd[name + ".pedestal"] = ped.clone()


def make_state_dict(seed: int = 0, N: int = 192, M: int = 16, boost: bool = True) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    g = (lambda v: v) if boost else (lambda v: 1.0)
    r = _rng(seed, "z2_sigma")
    sig = torch.from_numpy((0.6 + 1.2 * r.uniform(size=(1, N, 1, 1))).astype(np.float32))
    sd["v_z2_sigma"] = sig
    sd["z2_sigma"] = sig.clone()
    # g_a
    cins = [3, N, N, N]
    for li, (ci, idx) in enumerate(zip(cins, (1, 4, 7, 10))):
        _conv(sd, seed, f"a_model.transform.{idx}", N, ci, 5, gain=g(40.0) if idx == 10 else 1.0)
    for idx in (2, 5, 8):
        _gdn(sd, seed, f"a_model.transform.{idx}", N)
    # g_s (ConvTranspose2d)
    chans = [(N - M, N), (N, N), (N, N), (N, M)]
    for (ci, co), idx in zip(chans, (1, 4, 7, 10)):
        _conv(sd, seed, f"s_model.transform.{idx}", co, ci, 5, transposed=True,
              gain=g(5.0) if idx == 10 else (g(1.5) if idx == 1 else 1.0))
    for (ci, co), idx in zip(chans, (2, 5, 8, 11)):
        _gdn(sd, seed, f"s_model.transform.{idx}", co)
    # h_a / h_s
    _conv(sd, seed, "ha_model.transform.0", N, N, 3)
    _conv(sd, seed, "ha_model.transform.2", N, N, 5)
    _conv(sd, seed, "ha_model.transform.4", N, N, 5, gain=g(10.0))
    _conv(sd, seed, "hs_model.transform.0", N, N, 5, transposed=True)
    _conv(sd, seed, "hs_model.transform.2", N, N, 5, transposed=True)
    _conv(sd, seed, "hs_model.transform.4", N, N, 3, transposed=True, gain=g(30.0))
    # syntax branch
    _conv(sd, seed, "syntax_model.down0", 32, M, 3)
    _conv(sd, seed, "syntax_model.down1", 64, 32, 3)
    _conv(sd, seed, "syntax_model.conv", M, M + 32 + 64, 1, gain=g(12.0))
    _linear(sd, seed, "conv_weights_gen.transform.0", 128, M)
    _linear(sd, seed, "conv_weights_gen.transform.2", 256, 128)
    _linear(sd, seed, "conv_weights_gen.transform.4", 3 * M, 256, gain=g(2.0))
    # context model
    _conv(sd, seed, "prediction_model.transform.0", N, 2 * N - M, 3)
    _conv(sd, seed, "prediction_model.transform.2", N, N, 3)
    _conv(sd, seed, "prediction_model.transform.4", N, N, 3)
    _linear(sd, seed, "prediction_model.fc", 2 * (N - M), 4 * N, gain=g(8.0))
    _conv(sd, seed, "prediction_model_syntax.down0", M, N, 3)
    _conv(sd, seed, "prediction_model_syntax.down1", M, M, 3)
    _linear(sd, seed, "prediction_model_syntax.fc", 2 * M, 2 * M + N)
    return sd


def widths_state_dict(seed: int = 11):
    """Deterministic weights of the four transform classes at hidden width 128 / latent width 192."""
    sd = {}
    H, Lw, M = 128, 192, 16
    for (ci, co), idx in zip([(3, H), (H, H), (H, H), (H, Lw)], (1, 4, 7, 10)):
        _conv(sd, seed, f"a.transform.{idx}", co, ci, 5, gain=30.0 if idx == 10 else 1.0)
    for idx in (2, 5, 8):
        _gdn(sd, seed, f"a.transform.{idx}", H)
    for (ci, co), idx in zip([(Lw, H), (H, H), (H, H), (H, M)], (1, 4, 7, 10)):
        _conv(sd, seed, f"s.transform.{idx}", co, ci, 5, transposed=True, gain=4.0 if idx == 10 else 1.0)
    for co, idx in zip((H, H, H, M), (2, 5, 8, 11)):
        _gdn(sd, seed, f"s.transform.{idx}", co)
    _conv(sd, seed, "ha.transform.0", H, Lw, 3)
    _conv(sd, seed, "ha.transform.2", H, H, 5)
    _conv(sd, seed, "ha.transform.4", H, H, 5, gain=8.0)
    _conv(sd, seed, "hs.transform.0", H, H, 5, transposed=True)
    _conv(sd, seed, "hs.transform.2", H, H, 5, transposed=True)
    _conv(sd, seed, "hs.transform.4", Lw, H, 3, transposed=True, gain=20.0)
    return sd


def sub_state_dict(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def make_input(seed: int, B: int, H: int, W: int) -> torch.Tensor:
    """x in [-1,1) like eval_net.py:84; smooth-ish + noise so layers see structure."""
    r = _rng(seed, f"x{B}x{H}x{W}")
    base = r.uniform(-1, 1, size=(B, 3, H // 8 + 1, W // 8 + 1)).astype(np.float32)
    up = torch.nn.functional.interpolate(torch.from_numpy(base), size=(H, W), mode="bilinear", align_corners=True)
    noise = torch.from_numpy(r.uniform(-1, 1, size=(B, 3, H, W)).astype(np.float32))
    return (0.7 * up + 0.3 * noise).clamp(-1, 1).contiguous()


def likelihood_synthetic(seed: int, n: int):
    """Kernel-level synthetic of SURVEY 8(d): v~N(0,4^2), mu~N(0,1),
    sigma=exp(N(0,1)).clamp(0.05,20) + the edge set."""
    r = _rng(seed, f"lik{n}")
    v = (4.0 * r.standard_normal(n)).astype(np.float32)
    mu = r.standard_normal(n).astype(np.float32)
    sigma = np.clip(np.exp(r.standard_normal(n)), 0.05, 20).astype(np.float32)
    edge_v = np.array([0.5, -0.5, 1.5, -1.5, 2.5, -2.5, -0.0, 0.0, 3.0, 1.0, 0.49999997, 7.5], dtype=np.float32)
    k = min(n, len(edge_v))
    v[:k] = edge_v[:k]
    if n >= 24:
        sigma[12:16] = np.array([-1.0, 0.11, 0.0, 1e-3], dtype=np.float32)
        mu[16:20] = np.array([0.25, -0.25, 0.5, -0.5], dtype=np.float32)
    return torch.from_numpy(v), torch.from_numpy(mu), torch.from_numpy(sigma)


def unet_param_fill(named_shapes, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic values for the PARAMETERS of the U-Net-family Net (model/net_unet_ha_hs.py), by name and shape,
    so that the reference (build container) and our assembly (GPU box) can be given identical weights without shipping a
    200 MB state-dict.  `named_shapes`: iterable of (name, shape) over named_parameters().  Buffers keep their
    constructor values on both sides.  Gains are chosen so that latents quantise to non-trivial symbols."""
    out: Dict[str, torch.Tensor] = {}
    for name, shape in named_shapes:
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        if name in ("v_z2_sigma", "z2_sigma") or leaf == "quantiles":
            continue                                            # constructor values
        if leaf == "beta":                                      # GDN parameters are stored reparametrised (sqrt of the value + pedestal)
            r = _rng(seed, name)
            v = 1.0 + 0.5 * r.uniform(-1, 1, size=shape)
            out[name] = torch.sqrt(torch.from_numpy(v.astype(np.float32)) + 2.0 ** -36)
        elif leaf == "gamma":
            r = _rng(seed, name)
            ch = shape[0]
            v = 0.1 * np.eye(ch) + 0.02 * np.abs(r.standard_normal(size=shape)) / math.sqrt(ch / 16.0)
            out[name] = torch.sqrt(torch.from_numpy(v.astype(np.float32)) + 2.0 ** -36)
        elif leaf in ("relative_position_bias_table", "relative_position_params"):
            out[name] = _uniform(seed, name, shape, 0.3)
        elif len(shape) == 1:
            if leaf == "weight":                                # LayerNorm scale
                out[name] = 1.0 + _uniform(seed, name, shape, 0.1)
            elif name.startswith("cc_scale_transforms.") and name.split(".")[2] == "4":
                out[name] = 0.8 + _uniform(seed, name, shape, 0.3)   # scale head bias: most sigmas above the 0.11 bound
            else:
                out[name] = _uniform(seed, name, shape, 0.05)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 1.0
            if name.startswith("a_model.transform.15."):
                gain = 15.0                                     # latent y: std of a few quantisation steps
            elif name.startswith("cc_scale_transforms.") and name.split(".")[2] == "4":
                gain = 30.0
            elif name.startswith("cc_mean_transforms.") and name.split(".")[2] == "4":
                gain = 25.0
            elif name.startswith("s_model.transform."):         # transposed convs: hold the activation scale through g_s
                gain = {"2": 0.45, "5": 2.2, "9": 1.3, "12": 1.2}.get(name.split(".")[2], 1.0)   # (IGDN grows like x^2 above 1:
                # keep |x| below ~1 so that a flipped symbol is not amplified four times on its way to the image)
            elif name.startswith("syntax_model.conv."):
                gain = 12.0
            out[name] = _uniform(seed, name, shape, gain / math.sqrt(fan_in))
    return out

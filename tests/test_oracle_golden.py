"""Pins the CPU oracle (oracle/ref_path.py) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py) and the reference's one KAT."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_path as rp
import det_weights as dw

G = os.path.join(os.path.dirname(__file__), "golden")


def L(name):
    d = np.load(os.path.join(G, name))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


def test_kat_parametrizer():
    # ops/parametrizers.py:52-58 -- values the reference prints
    k = json.load(open(os.path.join(G, "kat_parametrizer.json")))
    gi = rp.nonneg_init(0.1 * torch.eye(5))
    gf = rp.nonneg_forward(gi)
    assert torch.equal(gi, torch.tensor(k["init"]))
    assert torch.equal(gf, torch.tensor(k["forward"]))
    assert abs(gi[0, 0].item() - 3.1623e-01) < 5e-6 and abs(gi[0, 1].item() - 3.8147e-06) < 5e-11
    assert abs(gf[0, 0].item() - 0.1) < 1e-7 and gf[0, 1].item() == 0.0


def test_leaf_ops_bit_exact():
    d = L("leaf_ops.npz")
    assert torch.equal(rp.lower_bound(d["lb_x"], 0.11), d["lb_out"])
    assert torch.equal(rp.lower_bound_grad(d["lb_x"], torch.tensor(0.11), d["lb_gout"]), d["lb_grad"])
    assert torch.equal(rp.lower_bound(d["lb_x"], 0.05), d["mlb_out"])
    assert torch.equal(rp.lower_bound_grad(d["lb_x"], torch.tensor(0.05), d["lb_gout"]), d["mlb_grad"])
    assert torch.equal(rp.ste_round(d["rnd_in"]), d["ste_round"])
    assert torch.equal(rp.bypass_round(d["rnd_in"]), d["bypass_round"])
    assert torch.equal(rp.nonneg_forward(d["nn_p"]), d["nn_fwd_min0"])
    assert torch.equal(rp.nonneg_forward(d["nn_p"], minimum=1e-6), d["nn_fwd_beta"])
    assert torch.equal(rp.nonneg_init(d["nn_p"]), d["nn_init"])
    bb, gb, ped = rp.model_gdn_constants()
    assert [bb, gb, ped] == d["model_gdn_consts"].tolist()
    b2, p2 = rp.parametrizer_constants(1e-6)
    g2, _ = rp.parametrizer_constants(0.0)
    assert [b2, g2, p2] == d["layers_gdn_consts"].tolist()


def test_gdn_variants_bit_exact():
    d = L("leaf_ops.npz")
    x, bp, gp = d["x"], d["beta_p"], d["gamma_p"]
    assert torch.equal(rp.gdn_model(x, bp, gp, False), d["model_gdn"])
    assert torch.equal(rp.gdn_model(x, bp, gp, True), d["model_igdn"])
    assert torch.equal(rp.gdn_layers(x, bp, gp, False), d["layers_gdn_inv0"])
    assert torch.equal(rp.gdn_layers(x, bp, gp, True), d["layers_gdn_inv1"])
    # model/ops.py:106-136 is the same arithmetic as model/gdn.py
    assert torch.equal(rp.gdn_model(x, bp, gp, False), d["model_ops_gdn_inv0"])
    assert torch.equal(rp.gdn_model(x, bp, gp, True), d["model_ops_gdn_inv1"])


def test_gaussian_model_bit_exact():
    d = L("gaussian_model.npz")
    assert torch.equal(rp.bypass_round(d["v"]), d["v_rounded"])
    lik = rp.gaussian_model_likelihood(d["v_rounded"], d["sigma"], d["mu"])
    # NaN where sigma == 0 (reference quirk H2) must match position-wise
    assert torch.equal(torch.isnan(lik), torch.isnan(d["lik"]))
    m = ~torch.isnan(lik)
    assert torch.equal(lik[m], d["lik"][m])
    zl = rp.gaussian_model_likelihood(d["z_rounded"], d["z_sigma"], torch.zeros_like(d["z_sigma"]))
    assert torch.equal(zl, d["z_lik"])


def test_block_sample_gather_equals_onehot():
    r = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 6, 7, generator=r)
    for masked in (True, False):
        assert torch.equal(rp.block_sample_gather(x, masked), rp.block_sample_onehot(x, masked))


@pytest.mark.parametrize("name", ["net_64x64_b1.npz", "net_64x128_b2.npz", "net_evalpad_60x50.npz"])
def test_net_forward_full(name):
    d = L(name)
    B, H, W, th, tw = (int(d[k]) for k in ("B", "H", "W", "th", "tw"))
    sd = dw.make_state_dict(int(d["seed"]), boost=bool(int(d["boost"])))
    x = rp.eval_pad(d["img"]) if "img" in d else dw.make_input(int(d["seed"]), B, H, W)
    with torch.no_grad():
        o = rp.net_forward_test(sd, x, (B, th, tw, 3))
    for k in ("z3", "z2", "h2", "mu", "sigma", "z2_lik", "y_lik", "syn_lik", "x_tilde16", "z3_syntax", "conv_weights"):
        assert torch.equal(o[k], d[k]), k
    # reference binds prediction_model_syntax's (mu, sigma) as (sigma, mu): model/net.py:789 vs :413
    assert torch.equal(o["syn_sigma"], d["syn_first"]) and torch.equal(o["syn_mu"], d["syn_second"])
    for k in ("bpp", "v_mse", "v_psnr"):
        assert torch.equal(o[k], d[k]), k


def test_net_forward_faithful_sampler_matches():
    d = L("net_64x64_b1.npz")
    sd = dw.make_state_dict(0)
    x = dw.make_input(0, 1, 64, 64)
    with torch.no_grad():
        o = rp.net_forward_test(sd, x, (1, 64, 64, 3), faithful_sampler=True)
    assert torch.equal(o["bpp"], d["bpp"]) and torch.equal(o["y_lik"], d["y_lik"])


def test_net_default_gain_and_config1():
    d = L("net_64x64_default_gain.npz")
    with torch.no_grad():
        o = rp.net_forward_test(dw.make_state_dict(0, boost=False), dw.make_input(0, 1, 64, 64), (1, 64, 64, 3))
    for k in ("bpp", "v_mse", "v_psnr", "z3", "z2"):
        assert torch.equal(o[k], d[k]), k
    d = L("net_256x256_b1.npz")
    with torch.no_grad():
        o = rp.net_forward_test(dw.make_state_dict(0), dw.make_input(0, 1, 256, 256), (1, 256, 256, 3))
    assert torch.equal(o["z3"], d["z3"]) and torch.equal(o["z2"], d["z2"])
    assert torch.equal(o["x_tilde16"][:, :, ::8, ::8], d["x_tilde16_sub"])
    # thread-count dependent summation order in torch.sum: compare scalars to 1e-6
    for k in ("bpp", "v_mse", "v_psnr", "bits"):
        assert torch.allclose(o[k], d[k], rtol=2e-6, atol=0), k


def test_eval_pad_matches_driver_padding():
    """ldic_b200.evaluation.pad_to_multiple == eval_net.py:68-84 (CPU-only host logic, no kernels involved)."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # import the module file directly: the package __init__ is importable on CPU, only kernel calls need a GPU
    sys.path.insert(0, root)
    import ldic_b200
    for h, w in ((60, 50), (64, 64), (65, 128), (1, 1)):
        img = torch.rand(3, h, w)
        a = ldic_b200.evaluation.pad_to_multiple(img)
        b = rp.eval_pad(img)
        assert a.shape == b.shape and torch.equal(a, b)
        assert a.shape[2] % 64 == 0 and a.shape[3] % 64 == 0


@pytest.mark.parametrize("tag", ["noshift", "shift", "small"])
def test_win_attention_oracle_vs_reference_golden(tag):
    """the oracle's window attention against outputs of the unmodified layers/win_attention.py."""
    d = np.load(os.path.join(G, "win_attention.npz"))
    dim, heads, ws, shift, B, H, W = [int(v) for v in d[f"{tag}_cfg"]]
    sd = {k[len(tag) + 4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith(f"{tag}_sd_")}
    y = rp.win_based_attention(sd, torch.from_numpy(d[f"{tag}_x"]), heads, ws, shift)
    assert torch.equal(sd["attn.relative_position_index"].long(), rp.win_rel_position_index(ws))
    torch.testing.assert_close(y, torch.from_numpy(d[f"{tag}_y"]), rtol=1e-5, atol=1e-5)


def test_net_high_forward_full():
    """Round 2: the `--high` model (N=384, M=32; model/net.py:446-451).  The oracle reproduces the unmodified reference
    bit for bit at this width too (tests/golden/make_golden_r2.py)."""
    d = L("net_high_64x64_b1.npz")
    sd = dw.make_state_dict(int(d["seed"]), N=384, M=32, boost=True)
    x = dw.make_input(int(d["seed"]), 1, 64, 64)
    with torch.no_grad():
        o = rp.net_forward_test(sd, x, (1, 64, 64, 3), M=32)
    for k in ("z3", "z2", "h2", "mu", "sigma", "z2_lik", "y_lik", "syn_lik", "x_tilde16", "z3_syntax", "conv_weights"):
        assert torch.equal(o[k], d[k]), k
    for k in ("bpp", "v_mse", "v_psnr"):
        assert torch.equal(o[k], d[k]), k


def test_transforms_at_128_192_widths():
    """BASELINE config 1 widths: oracle transforms vs the unmodified reference classes at [128,128,128,192]."""
    sys_path = os.path.join(os.path.dirname(__file__), "golden")
    import sys
    if sys_path not in sys.path:
        sys.path.insert(0, sys_path)
    from make_golden_r2 import widths_state_dict
    d = L("widths_128_192.npz")
    sd = widths_state_dict(int(d["seed"]))
    x = dw.make_input(int(d["xseed"]), int(d["B"]), int(d["H"]), int(d["W"]))
    with torch.no_grad():
        y = rp.analysis_transform(sd, x, prefix="a.transform.")
        assert torch.equal(y, d["y"])
        assert torch.equal(rp.synthesis_transform(sd, torch.round(y), prefix="s.transform."), d["xt16"])
        z = rp.h_analysis_transform(sd, y, prefix="ha.transform.")
        assert torch.equal(z, d["z"])
        assert torch.equal(rp.h_synthesis_transform(sd, torch.round(z), prefix="hs.transform."), d["h2"])


def test_uint8_input_map_identities():
    """The two facts the uint8 input path of the kernels relies on, checked for all 256 levels:
    (i) the bf16 operand of the first layer: bf16((u/255)*2-1) == bf16(fma(u, fl(2/255), -1));
    (ii) the a11 ground truth: round(((u/255)*2-1 + 1) * 127.5) == u  (model/net.py:864 on eval_net.py:84 inputs)."""
    u = np.arange(256, dtype=np.float32)
    x = (u / np.float32(255.0)) * np.float32(2.0) - np.float32(1.0)
    c = np.float32(2.0) / np.float32(255.0)
    fma = np.array([np.float32(np.float64(a) * np.float64(c) - 1.0) for a in u], dtype=np.float32)   # one rounding
    assert torch.equal(torch.from_numpy(fma).bfloat16(), torch.from_numpy(x).bfloat16())
    gt = np.rint((x + np.float32(1.0)) * np.float32(127.5))
    assert (gt == u).all()
    xt = (torch.arange(256, dtype=torch.uint8).float() / 255.0) * 2.0 - 1.0                           # torch's ToTensor path
    assert torch.equal(xt, torch.from_numpy(x))

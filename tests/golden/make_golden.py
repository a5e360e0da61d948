"""Generates the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture records inputs (or the seed that regenerates them through
tests/det_weights.py) and the outputs of the reference's own classes:
``ops/*``, ``model/gdn.py``, ``layers/gdn.py``, ``model/ops.py``,
``model/net.py`` (GaussianModel, BypassRound, the transform classes, Net).
The oracle (oracle/ref_path.py) and the CUDA path are both tested against these
files; nothing at test time reads /root/reference.
"""
from __future__ import annotations

import io
import json
import os
import sys
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness  # noqa: E402
import det_weights as dw  # noqa: E402


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().contiguous().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def main():
    assert ref_harness.available(), "reference tree not found"
    torch.set_num_threads(8)
    net_mod = ref_harness.load_net_module()
    import ops as ref_ops                      # /root/reference/ops
    from model import gdn as ref_model_gdn     # /root/reference/model/gdn.py
    from model import ops as ref_model_ops     # /root/reference/model/ops.py
    ref_layers_gdn = ref_harness.load_leaf("layers/gdn.py", "ref_layers_gdn")

    # ---- 1. the reference's only KAT: ops/parametrizers.py:52-58 -----------------
    nonn = ref_ops.NonNegativeParametrizer()
    g0 = 0.1 * torch.eye(5)
    gi = nonn.init(g0)
    gf = nonn(gi)
    with open(os.path.join(HERE, "kat_parametrizer.json"), "w") as f:
        json.dump({"source": "ops/parametrizers.py:52-58 executed",
                   "init": gi.tolist(), "forward": gf.tolist(),
                   "survey_printed": {"init_diag": 3.1623e-01, "init_offdiag": 3.8147e-06,
                                      "forward_diag": 0.1000, "forward_offdiag": 0.0}}, f, indent=1)

    # ---- 2. leaf ops ---------------------------------------------------------------
    r = np.random.Generator(np.random.PCG64(1234))
    C, Hh, Ww = 24, 5, 7
    x = torch.from_numpy(r.standard_normal((2, C, Hh, Ww)).astype(np.float32)) * 2
    xb = torch.from_numpy(r.uniform(-0.3, 0.3, 64).astype(np.float32))
    gout = torch.from_numpy(r.standard_normal(64).astype(np.float32))
    lb = ref_ops.LowerBound(0.11)
    xb_ = xb.clone().requires_grad_(True)
    lb_out = lb(xb_)
    lb_out.backward(gout)
    mlb_x = xb.clone().requires_grad_(True)
    mlb_out = ref_model_gdn.lower_bound(mlb_x, 0.05)
    mlb_out.backward(gout)
    rnd_in = torch.from_numpy(np.concatenate([r.standard_normal(40) * 3,
                                              [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, -0.0, 0.0, 2.3, 1e-8, -1e-8, 8388607.5]]
                                             ).astype(np.float32))
    nn_p = torch.from_numpy(r.uniform(-0.01, 1.2, (C, C)).astype(np.float32))
    nn_beta = ref_ops.NonNegativeParametrizer(minimum=1e-6)
    # GDN variants with non-trivial params (shared raw parameters)
    beta_p = torch.sqrt(torch.from_numpy((1 + 0.5 * r.uniform(-1, 1, C)).astype(np.float32)) + 2.0 ** -36)
    gamma_p = torch.sqrt(torch.from_numpy((0.1 * np.eye(C) + 0.03 * np.abs(r.standard_normal((C, C)))).astype(np.float32)) + 2.0 ** -36)
    gamma_p[0, 1] = 1e-7   # below the gamma bound -> exercises LowerBound
    beta_p[3] = 1e-4       # below the beta bound
    outs = {}
    with torch.no_grad():
        for nm, cls in (("model_gdn", ref_model_gdn.GDN), ("model_igdn", ref_model_gdn.IGDN)):
            m = cls(C)
            m.beta.copy_(beta_p); m.gamma.copy_(gamma_p)
            outs[nm] = m(x)
        for inv in (False, True):
            m = ref_layers_gdn.GDN(C, inverse=inv)
            m.beta.copy_(beta_p); m.gamma.copy_(gamma_p)
            outs[f"layers_gdn_inv{int(inv)}"] = m(x)
            m2 = ref_model_ops.GDN(C, inverse=inv)
            m2.beta.copy_(beta_p); m2.gamma.copy_(gamma_p)
            outs[f"model_ops_gdn_inv{int(inv)}"] = m2(x)
        mg = ref_model_gdn.GDN(C)
        lg = ref_layers_gdn.GDN(C)
        save("leaf_ops.npz",
             x=x, beta_p=beta_p, gamma_p=gamma_p, **outs,
             lb_x=xb, lb_gout=gout, lb_out=lb_out, lb_grad=xb_.grad,
             mlb_out=mlb_out, mlb_grad=mlb_x.grad,
             rnd_in=rnd_in, ste_round=ref_ops.ste_round(rnd_in), bypass_round=net_mod.bypass_round(rnd_in),
             nn_p=nn_p, nn_fwd_min0=ref_ops.NonNegativeParametrizer()(nn_p), nn_fwd_beta=nn_beta(nn_p),
             nn_init=ref_ops.NonNegativeParametrizer().init(nn_p),
             model_gdn_init_beta=mg.beta, model_gdn_init_gamma=mg.gamma,
             model_gdn_consts=np.array([float(mg.beta_bound), float(mg.gamma_bound), float(mg.pedestal)], dtype=np.float64),
             layers_gdn_init_beta=lg.beta, layers_gdn_init_gamma=lg.gamma,
             layers_gdn_consts=np.array([float(lg.beta_reparam.lower_bound.bound), float(lg.gamma_reparam.lower_bound.bound),
                                         float(lg.beta_reparam.pedestal)], dtype=np.float64))

    # ---- 3. GaussianModel on the kernel-level synthetic -------------------------------
    gm = net_mod.GaussianModel()
    v, mu, sigma = dw.likelihood_synthetic(0, 8192)
    vr = torch.round(v)
    with torch.no_grad():
        lik = gm(vr, sigma, mu)
        # factorised form: (B,C,H,W) symbols with (1,C,1,1) sigma, mu=0  (model/net.py:781)
        zc = 12
        zr = torch.round(v[: 2 * zc * 4 * 5].view(2, zc, 4, 5))
        zs = sigma[100:100 + zc].view(1, zc, 1, 1)
        zlik = gm(zr, zs, torch.zeros_like(zs))
    save("gaussian_model.npz", v=v, v_rounded=vr, mu=mu, sigma=sigma, lik=lik, z_rounded=zr, z_sigma=zs, z_lik=zlik)

    # ---- 4. Net forward, full intermediates -------------------------------------------
    def run_net(B, H, W, seed, boost, test_hw=None, x=None):
        th, tw = test_hw if test_hw else (H, W)
        sd = dw.make_state_dict(seed, boost=boost)
        with contextlib.redirect_stdout(io.StringIO()):
            net = net_mod.Net((B, th, tw, 3), (B, th, tw, 3), False, False).eval()
        missing, unexpected = net.load_state_dict(sd, strict=False)
        assert not unexpected
        assert all(k.startswith(("HAN", "conv_weights_gen_HAN", "add_mean")) or "sampler" in k for k in missing), missing
        if x is None:
            x = dw.make_input(seed, B, H, W)
        cap = {}

        def hook(name):
            def f(m, i, o):
                cap[name] = o
            return f
        for n in ["a_model", "ha_model", "hs_model", "prediction_model", "prediction_model_syntax", "s_model",
                  "entropy_bottleneck_z2", "entropy_bottleneck_z3", "entropy_bottleneck_z3_syntax",
                  "syntax_model", "conv_weights_gen"]:
            getattr(net, n).register_forward_hook(hook(n))
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            bpp, v_mse, v_psnr = net(x, "test", 1)
        del net
        r_ = dict(bpp=bpp, v_mse=v_mse, v_psnr=v_psnr, z3=cap["a_model"], z2=cap["ha_model"], h2=cap["hs_model"],
                  mu=cap["prediction_model"][0].contiguous(), sigma=cap["prediction_model"][1].contiguous(),
                  syn_first=cap["prediction_model_syntax"][0].contiguous(), syn_second=cap["prediction_model_syntax"][1].contiguous(),
                  z2_lik=cap["entropy_bottleneck_z2"], y_lik=cap["entropy_bottleneck_z3"],
                  syn_lik=cap["entropy_bottleneck_z3_syntax"], x_tilde16=cap["s_model"],
                  z3_syntax=cap["syntax_model"], conv_weights=cap["conv_weights_gen"])
        r_["bits"] = torch.stack([torch.log(r_[k]).sum() for k in ("z2_lik", "y_lik", "syn_lik")])
        return x, r_

    x, r64 = run_net(1, 64, 64, 0, True)
    save("net_64x64_b1.npz", seed=0, boost=1, B=1, H=64, W=64, th=64, tw=64, **r64)

    x, r2 = run_net(2, 64, 128, 3, True)
    save("net_64x128_b2.npz", seed=3, boost=1, B=2, H=64, W=128, th=64, tw=128, **r2)

    # eval_net.py semantics: 60x50 image padded with ONES to 64x64, bpp normalised by (60,50)
    rr = np.random.Generator(np.random.PCG64(77))
    img = torch.from_numpy(rr.uniform(0, 1, (3, 60, 50)).astype(np.float32))
    from oracle.ref_path import eval_pad
    xpad = eval_pad(img)
    _, rp_ = run_net(1, 64, 64, 5, True, test_hw=(60, 50), x=xpad)
    save("net_evalpad_60x50.npz", seed=5, boost=1, B=1, H=64, W=64, th=60, tw=50, img=img, **rp_)

    _, rd = run_net(1, 64, 64, 0, False)
    save("net_64x64_default_gain.npz", seed=0, boost=0, B=1, H=64, W=64, th=64, tw=64,
         **{k: rd[k] for k in ("bpp", "v_mse", "v_psnr", "bits", "z3", "z2")})

    # config 1 shape (1x256x256): summaries + the latents
    _, r256 = run_net(1, 256, 256, 0, True)
    save("net_256x256_b1.npz", seed=0, boost=1, B=1, H=256, W=256, th=256, tw=256,
         bpp=r256["bpp"], v_mse=r256["v_mse"], v_psnr=r256["v_psnr"], bits=r256["bits"],
         z3=r256["z3"], z2=r256["z2"], h2_sub=r256["h2"][:, ::8], mu_sub=r256["mu"][:, ::8], sigma_sub=r256["sigma"][:, ::8],
         x_tilde16_sub=r256["x_tilde16"][:, :, ::8, ::8])

    # config 2 shape, one image (768x512): scalars only
    _, rk = run_net(1, 512, 768, 0, True)
    with open(os.path.join(HERE, "net_512x768_b1.json"), "w") as f:
        json.dump({"seed": 0, "boost": 1, "B": 1, "H": 512, "W": 768,
                   "bpp": float(rk["bpp"]), "v_mse": rk["v_mse"].tolist(), "v_psnr": float(rk["v_psnr"]),
                   "bits": rk["bits"].tolist(),
                   "z3_abs_sum": float(rk["z3"].abs().double().sum()), "z3_std": float(rk["z3"].std()),
                   "nonzero_y": float((torch.round(rk["z3"][:, 16:]) != 0).float().mean()),
                   "nonzero_z": float((torch.round(rk["z2"]) != 0).float().mean())}, f, indent=1)
    print("done")


if __name__ == "__main__":
    main()

"""Golden fixtures of the U-Net-family Net (model/net_unet_ha_hs.py, BASELINE configs[2] / [3]), produced by running the
reference's own, unmodified model file on CPU -- with its missing third-party / absent dependencies RESTATED
(oracle/unet_harness.py: "restated deps").  Build container only:

    python tests/golden/make_golden_unet.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import unet_harness  # noqa: E402
import det_weights as dw  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    torch.set_num_threads(8)
    m = unet_harness.load_unet_module("net_unet_ha_hs")

    def run(B, H, W, seed, name):
        with contextlib.redirect_stdout(io.StringIO()):
            net = m.Net((B, H, W, 3), (B, H, W, 3), False, False).eval()
        sd = net.state_dict()
        keys = {k: list(v.shape) for k, v in sd.items()
                if not (k.endswith("sample_filter") or k.startswith(("HAN.", "conv_weights_gen_HAN.", "add_mean.")))}
        with open(os.path.join(HERE, "unet_state_keys.json"), "w") as f:
            json.dump(keys, f, indent=0)
        fill = dw.unet_param_fill([(n, tuple(p.shape)) for n, p in net.named_parameters()
                                   if not n.startswith(("HAN.", "conv_weights_gen_HAN.", "add_mean."))], seed)
        missing, unexpected = net.load_state_dict(fill, strict=False)
        assert not unexpected, unexpected
        x = dw.make_input(seed, B, H, W)
        cap = {}
        liks = []
        net.a_model.register_forward_hook(lambda mod, i, o: cap.__setitem__("z3", o))
        net.s_model.register_forward_hook(lambda mod, i, o: cap.__setitem__("x_tilde16", o))
        net.s_model.register_forward_hook(lambda mod, i, o: cap.__setitem__("y_hat", i[0]))
        net.h_s.register_forward_hook(lambda mod, i, o: cap.__setitem__("latent", o))
        net.gaussian_conditional.register_forward_hook(lambda mod, i, o: liks.append((i[1], i[2], o[1])))
        net.syntax_model.register_forward_hook(lambda mod, i, o: cap.__setitem__("z3_syntax", o))
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            bpp, v_mse, v_psnr = net(x, "test", 1)
        bits = torch.stack([torch.log(l[2]).sum() for l in liks])
        scales = torch.cat([l[0] for l in liks], 1)
        means = torch.cat([l[1] for l in liks], 1)
        yl = torch.cat([l[2] for l in liks], 1)
        y = cap["z3"]
        sym = torch.round(y - means)
        print(f"{name}: bpp {float(bpp):.4f} psnr {float(v_psnr):.3f} | y std {float(y.std()):.3f} nonzero symbols "
              f"{float((sym != 0).float().mean()):.2f} | scale<0.11 {float((scales < 0.11).float().mean()):.2f} | "
              f"lik at clamp {float((yl <= 1e-9).float().mean()):.3f} | x~16 range {float(cap['x_tilde16'].min()):.2f}.."
              f"{float(cap['x_tilde16'].max()):.2f} | latent std {float(cap['latent'].std()):.3f}")
        save(name, seed=seed, B=B, H=H, W=W, bpp=bpp, v_mse=v_mse, v_psnr=v_psnr, bits=bits, z3=y, means=means, scales=scales,
             y_hat=cap["y_hat"], z3_syntax=cap["z3_syntax"], latent_sub=cap["latent"][:, ::4], x_tilde16_sub=cap["x_tilde16"][:, :, ::4, ::4])

    run(1, 256, 256, 0, "unet_256x256_b1.npz")
    run(2, 256, 512, 1, "unet_256x512_b2.npz")
    print("done")


if __name__ == "__main__":
    main()

"""Round-2 golden fixtures, again produced by EXECUTING THE UNMODIFIED REFERENCE (build container only):

    python tests/golden/make_golden_r2.py

* net_high_64x64_b1.npz   -- model/net.py ``Net(is_high=True)`` (N=384, M=32; model/net.py:446-451) forward in test
                             mode on a 64x64 image with the deterministic weights of tests/det_weights.py.
* widths_128_192.npz      -- the transform classes at the widths of BASELINE config 1 ("N=128, M=192"):
                             analysisTransformModel(3,[128,128,128,192]), synthesisTransformModel(192,[128,128,128,16]),
                             h_analysisTransformModel(192,[128,128,128],[1,2,2]),
                             h_synthesisTransformModel(128,[128,128,192],[2,2,1]) (model/net.py:91-216).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness  # noqa: E402
import det_weights as dw  # noqa: E402
from make_golden import save  # noqa: E402


from det_weights import widths_state_dict, sub_state_dict as sub  # noqa: E402,F401


def main():
    assert ref_harness.available(), "reference tree not found"
    torch.set_num_threads(8)
    net_mod = ref_harness.load_net_module()

    # ---- transform classes at 128 / 192 ------------------------------------------------------------
    sd = widths_state_dict()
    B, H, W = 2, 64, 128                     # latent 4 x 8: h_a's stride-2 layers need even sizes on our side
    x = dw.make_input(21, B, H, W)
    with torch.no_grad():
        ga = net_mod.analysisTransformModel(3, [128, 128, 128, 192]).eval()
        ga.load_state_dict(sub(sd, "a."), strict=True)
        y = ga(x)
        gs = net_mod.synthesisTransformModel(192, [128, 128, 128, 16]).eval()
        gs.load_state_dict(sub(sd, "s."), strict=True)
        y_hat = torch.round(y)
        xt16 = gs(y_hat)
        ha = net_mod.h_analysisTransformModel(192, [128, 128, 128], [1, 2, 2]).eval()
        ha.load_state_dict(sub(sd, "ha."), strict=True)
        z = ha(y)
        hs = net_mod.h_synthesisTransformModel(128, [128, 128, 192], [2, 2, 1]).eval()
        hs.load_state_dict(sub(sd, "hs."), strict=True)
        h2 = hs(torch.round(z))
    print("widths: y std %.3f nonzero %.2f | z std %.3f nonzero %.2f | xt16 range %.2f..%.2f" % (
        y.std(), (y_hat != 0).float().mean(), z.std(), (torch.round(z) != 0).float().mean(), xt16.min(), xt16.max()))
    save("widths_128_192.npz", seed=11, xseed=21, B=B, H=H, W=W, y=y, xt16=xt16, z=z, h2=h2)

    # ---- Net(is_high=True): N = 384, M = 32 --------------------------------------------------------
    def run_high(B, H, W, seed):
        sdh = dw.make_state_dict(seed, N=384, M=32, boost=True)
        with contextlib.redirect_stdout(io.StringIO()):
            net = net_mod.Net((B, H, W, 3), (B, H, W, 3), True, False).eval()
        missing, unexpected = net.load_state_dict(sdh, strict=False)
        assert not unexpected, unexpected
        assert all(k.startswith(("HAN", "conv_weights_gen_HAN", "add_mean")) or "sampler" in k for k in missing), missing
        x = dw.make_input(seed, B, H, W)
        cap = {}

        def hook(name):
            def f(m, i, o):
                cap[name] = o
            return f
        for n in ["a_model", "ha_model", "hs_model", "prediction_model", "prediction_model_syntax", "s_model",
                  "entropy_bottleneck_z2", "entropy_bottleneck_z3", "entropy_bottleneck_z3_syntax", "syntax_model",
                  "conv_weights_gen"]:
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
        print("high %dx%dx%d: bpp %.4f psnr %.3f | y std %.3f nonzero %.2f | z nonzero %.2f | x~16 range %.2f..%.2f" % (
            B, H, W, bpp, v_psnr, r_["z3"].std(), (torch.round(r_["z3"][:, 32:]) != 0).float().mean(),
            (torch.round(r_["z2"]) != 0).float().mean(), r_["x_tilde16"].min(), r_["x_tilde16"].max()))
        return r_

    r_ = run_high(1, 64, 64, 2)
    save("net_high_64x64_b1.npz", seed=2, boost=1, B=1, H=64, W=64, th=64, tw=64, N=384, M=32, **r_)
    r_ = run_high(2, 128, 192, 4)
    save("net_high_128x192_b2.npz", seed=4, boost=1, B=2, H=128, W=192, th=128, tw=192, N=384, M=32,
         **{k: r_[k] for k in ("bpp", "v_mse", "v_psnr", "bits", "z3", "z2", "h2", "z3_syntax", "conv_weights")},
         x_tilde16_sub=r_["x_tilde16"][:, :, ::4, ::4])
    r_ = run_high(1, 256, 256, 6)
    save("net_high_256x256_b1.npz", seed=6, boost=1, B=1, H=256, W=256, th=256, tw=256, N=384, M=32,
         **{k: r_[k] for k in ("bpp", "v_mse", "v_psnr", "bits", "z3", "z2")})
    print("done")


if __name__ == "__main__":
    main()

"""End-to-end parity of the drop-in Net against golden fixtures produced by the unmodified
reference (gates of BASELINE.json: bpp within 0.5 %, PSNR within 0.01 dB)."""
import json
import os

import numpy as np
import pytest
import torch

import det_weights as dw
from oracle import ref_path as rp

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
BPP_RTOL, PSNR_ATOL_DB = 5e-3, 1e-2


def L(name):
    d = np.load(os.path.join(G, name))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


@pytest.fixture(scope="module")
def ldic():
    import ldic_b200
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(0), "device")
    return ldic_b200


def build(ldic, d):
    B, th, tw = int(d["B"]), int(d["th"]), int(d["tw"])
    net = ldic.Net((B, th, tw, 3), (B, th, tw, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(int(d["seed"]), boost=bool(int(d["boost"]))), strict=True)
    return net


@pytest.mark.parametrize("name", ["net_64x64_b1.npz", "net_64x128_b2.npz", "net_evalpad_60x50.npz",
                                  "net_64x64_default_gain.npz", "net_256x256_b1.npz"])
def test_net_forward_vs_reference_golden(ldic, name):
    d = L(name)
    B, H, W = int(d["B"]), int(d["H"]), int(d["W"])
    net = build(ldic, d)
    x = rp.eval_pad(d["img"]) if "img" in d else dw.make_input(int(d["seed"]), B, H, W)
    bpp, v_mse, v_psnr = net(x.cuda(), "test", 1)
    assert v_mse.shape == (B,)
    assert abs(bpp.item() / d["bpp"].item() - 1) < BPP_RTOL, (bpp.item(), d["bpp"].item())
    assert abs(v_psnr.item() - d["v_psnr"].item()) < PSNR_ATOL_DB, (v_psnr.item(), d["v_psnr"].item())
    # intermediates: bf16-operand budget on the latents
    out = net.rd_forward(x.cuda())
    y = out["latents"]["y"].permute(0, 3, 1, 2).cpu()
    rel = ((y - d["z3"]).pow(2).mean().sqrt() / d["z3"].pow(2).mean().sqrt()).item()
    assert rel < 1e-2, rel
    flips = (torch.round(y) != torch.round(d["z3"])).float().mean().item()
    assert flips < 0.02, flips


def test_net_512x768_scalars(ldic):
    k = json.load(open(os.path.join(G, "net_512x768_b1.json")))
    net = ldic.Net((1, 512, 768, 3), (1, 512, 768, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(k["seed"]), strict=True)
    x = dw.make_input(k["seed"], 1, 512, 768)
    bpp, v_mse, v_psnr = net(x.cuda(), "test", 1)
    assert abs(bpp.item() / k["bpp"] - 1) < BPP_RTOL
    assert abs(v_psnr.item() - k["v_psnr"]) < PSNR_ATOL_DB
    # batch independence at full size: image 0 of a batch of 2 gives the same per-image numbers
    x2 = torch.cat([x, dw.make_input(1, 1, 512, 768)], 0).cuda()
    out2 = net.rd_forward(x2, want_xt16=True)
    out1 = net.rd_forward(x.cuda(), want_xt16=True)
    # our kernels are batch-invariant bit for bit (g_a latent, g_s output); the torch-op context /
    # syntax branches may pick batch-dependent cuDNN algorithms, so the scalars get a tolerance
    assert torch.equal(out2["latents"]["y"][0], out1["latents"]["y"][0])
    assert torch.equal(out2["latents"]["z"][0], out1["latents"]["z"][0])
    assert torch.equal(out2["latents"]["xt16"][0], out1["latents"]["xt16"][0])
    assert abs(out2["sq_err"][0].item() / out1["sq_err"][0].item() - 1) < 1e-3


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 512, 768), (3, 192, 64)])
def test_syntax_branch_kernels_vs_torch_modules(ldic, B, H, W):
    """Syntax_Model / PredictionModel_Syntax / conv_generator (model/net.py:322-413) on ldic_syntax_branch
    against the same nn.Modules run as stock fp32 torch ops on identical inputs."""
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    g = torch.Generator().manual_seed(5)
    h, w, N, M = H // 16, W // 16, net.N, net.M
    y = (torch.randn(B, h, w, N, generator=g) * 1.5).cuda()
    h2 = (torch.randn(B, h, w, N, generator=g) * 0.8).cuda()
    z3, z3r, mu, sg, cw = ldic.ops.syntax_branch(y, h2, M, net.syntax_model, net.prediction_model_syntax, net.conv_weights_gen)
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            z3_t = net.syntax_model(y.permute(0, 3, 1, 2)[:, :M])
            mu_t, sg_t = net.prediction_model_syntax(torch.round(z3_t), h2.permute(0, 3, 1, 2))
            cw_t = net.conv_weights_gen(z3r)          # same rounded symbols on both sides
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    tol = dict(rtol=2e-5, atol=2e-5)
    assert z3.shape == z3_t.shape and torch.allclose(z3, z3_t, **tol), (z3 - z3_t).abs().max()
    assert torch.equal(z3r, torch.round(z3))
    assert torch.allclose(mu, mu_t.reshape(mu.shape), **tol) and torch.allclose(sg, sg_t.reshape(sg.shape), **tol)
    assert cw.shape == cw_t.shape and torch.allclose(cw, cw_t, **tol), (cw - cw_t).abs().max()
    # whole forward: torch-op syntax branch and the kernel branch agree
    x = dw.make_input(3, B, H, W).cuda()
    o1 = net.rd_forward(x)

    def syntax_torch(net_, y_nchw, h2_nchw):        # the stage as stock fp32 torch ops (model/net.py:712-719,:753,:789,:805)
        prev_ = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            z3_ = net_.syntax_model(y_nchw[:, :net_.M])
            z3r_ = torch.round(z3_)
            first, second = net_.prediction_model_syntax(z3r_, h2_nchw)
            return z3_, z3r_, first, second, net_.conv_weights_gen(z3r_)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev_
    o2 = net.rd_forward(x, overrides={"syntax": syntax_torch})
    assert torch.allclose(o1["bits"], o2["bits"], rtol=1e-5)
    assert (o1["sq_err"] - o2["sq_err"]).abs().max().item() <= 1e-4 * o2["sq_err"].max().item()


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 512, 768)])
def test_fused_tail_matches_separate_kernels(ldic, B, H, W):
    """batch_conv + squared level error inside the last deconv's epilogue (ldic_conv_forward_fused_tail) against the
    unfused path (16-channel g_s output in HBM + ldic_syntax_conv_mse): same arithmetic, exact integer sums."""
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    x = dw.make_input(7, B, H, W).cuda()
    net.tail_fused = True
    o1 = net.rd_forward(x, want_x_hat=True, want_xt16=True)
    net.tail_fused = False
    o2 = net.rd_forward(x, want_x_hat=True)
    assert torch.equal(o1["sq_err"], o2["sq_err"])
    assert torch.equal(o1["x_hat"], o2["x_hat"])
    assert torch.equal(o1["latents"]["xt16"], o2["latents"]["xt16"])
    net.tail_fused = True
    o3 = net.rd_forward(x)                     # product configuration: nothing but the sums leaves the kernel
    assert torch.equal(o3["sq_err"], o2["sq_err"]) and o3["latents"]["xt16"] is None


def test_eval_driver_matches_reference_eval_net(ldic):
    """eval_net.py:68-96 (pad with ones to x64, bpp over the unpadded size) through ldic_b200.evaluation, against the
    golden fixture the unmodified reference produced for a 60x50 image; batching does not change per-image numbers."""
    d = L("net_evalpad_60x50.npz")
    net = build(ldic, d)
    img = d["img"]
    x = ldic.evaluation.pad_to_multiple(img)
    assert torch.equal(x, rp.eval_pad(img))
    other = torch.rand(3, 64, 40, generator=torch.Generator().manual_seed(1))
    res = ldic.evaluation.evaluate_images(net, [img, other, img], batch_size=8)
    for r in (res[0], res[2]):
        assert abs(r["bpp"] / d["bpp"].item() - 1) < BPP_RTOL
        assert abs(r["psnr"] - d["v_psnr"].item()) < PSNR_ATOL_DB
        assert (r["h"], r["w"]) == (60, 50)
    assert res[0] == res[2] and res[1]["w"] == 40


def test_state_dict_contract(ldic):
    net = ldic.Net((1, 64, 64, 3), (1, 64, 64, 3), False, False)
    sd = dw.make_state_dict(0)
    # a reference checkpoint also carries sampler buffers and the HAN head: accepted under strict=True
    sd["test_y_sampler.sample_filter"] = torch.zeros(1)
    sd["HAN.head.0.weight"] = torch.zeros(1)
    net.load_state_dict(sd, strict=True)
    keys = set(net.state_dict().keys())
    assert {"v_z2_sigma", "z2_sigma", "a_model.transform.1.weight", "a_model.transform.2.beta", "a_model.transform.2.gamma",
            "a_model.transform.2.reparam_offset", "a_model.transform.2.pedestal", "s_model.transform.11.gamma",
            "ha_model.transform.4.bias", "hs_model.transform.4.weight", "prediction_model.fc.weight",
            "prediction_model_syntax.fc.bias", "syntax_model.conv.weight", "conv_weights_gen.transform.4.weight"} <= keys
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 3, 64, 64), "train")


def test_forward_dict_view(ldic):
    d = L("net_64x64_b1.npz")
    net = build(ldic, d)
    x = dw.make_input(0, 1, 64, 64).cuda()
    o = net.forward_dict(x)
    assert o["x_hat"].shape == (1, 3, 64, 64)
    assert o["likelihoods"]["y"].shape == (1, 176, 4, 4) and o["likelihoods"]["z"].shape == (1, 192, 1, 1)
    # RateDistortionLoss-style bpp (train_net_unet.py:76-79) agrees with the forward's y+z bits
    bits = sum(torch.log(l).sum() for l in o["likelihoods"].values())
    out = net.rd_forward(x)
    assert abs(bits.item() - out["bits"][:2].sum().item()) < 1e-3 * abs(bits.item())


def test_metric_kernels_match_reference_arithmetic(ldic):
    """ldic_rd_pack_metrics / ldic_rd_finish_metrics against model/net.py:856-869 evaluated with torch on the host."""
    import math
    bits = torch.tensor([-12345.678, -987654.3, -321.5])
    sq = torch.tensor([3 * 64 * 64 * 900, 3 * 64 * 64 * 17 + 5, 123456789], dtype=torch.int64)
    chw, th, tw = 3 * 64 * 64, 64, 64
    packed, v_mse = ldic.ops.rd_pack_metrics(bits.cuda(), sq.cuda(), chw)
    r = ldic.ops.rd_finish_metrics(packed, float(th * tw)).cpu()
    mse = sq.double() / chw
    assert torch.equal(v_mse.cpu(), mse.float())
    psnr = (20 * torch.log10(255.0 / torch.sqrt(mse))).mean()
    bpp = bits.double().sum() / (-math.log(2) * 3 * th * tw)
    assert packed.cpu()[4].item() == 3.0 and torch.equal(packed.cpu()[:3], bits.double())
    assert abs(r[0].item() / bpp.item() - 1) < 1e-6 and abs(r[1].item() - psnr.item()) < 1e-5


def test_graph_replay_is_bit_identical_to_eager(ldic):
    """GraphedEvaluator replays exactly the eager launch sequence: same sums, same squared errors, new data per replay."""
    B, H, W = 2, 64, 128
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    xs = [dw.make_input(s, B, H, W).cuda() for s in (0, 1)]
    eager = []
    for x in xs:
        out = net.rd_forward(x)
        bpp, v_mse, v_psnr = net.metrics(out, B, H, W)
        eager.append((out["bits"].clone(), out["sq_err"].clone(), bpp.clone(), v_psnr.clone(), v_mse.clone()))
    bufs = [torch.empty_like(xs[0]) for _ in range(2)]
    gev = ldic.GraphedEvaluator(net, bufs)
    assert gev.launches_per_replay > 20
    for rep in range(2):
        for k in (0, 1):
            src = xs[(k + rep) % 2]
            bufs[k].copy_(src)
            n0 = ldic.ops.launch_count()
            bpp, psnr, out = gev(k)
            assert ldic.ops.launch_count() - n0 == 1          # only the finish kernel is launched from the host
            e = eager[(k + rep) % 2]
            assert torch.equal(out["bits"], e[0]) and torch.equal(out["sq_err"], e[1])
            assert torch.equal(bpp, e[2]) and torch.equal(psnr, e[3]) and torch.equal(out["v_mse"], e[4])


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (4, 256, 384)])
def test_two_stream_forward_is_bit_identical_to_single_stream(ldic, B, H, W):
    """The SM partition (Net.side_sms: hyperprior / syntax chain on a side stream next to the first three g_s deconvs,
    grids capped through LdicConvDesc.sm_limit) changes the schedule, not one bit of the results."""
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    x = dw.make_input(3, B, H, W).cuda()
    res = []
    net.side_eager = True                                  # eager launches take the partition only on request
    for ov in (28, 0, 12):
        net.side_sms = ov
        out = net.rd_forward(x, want_x_hat=True)
        torch.cuda.synchronize()
        res.append((out["bits"].clone(), out["sq_err"].clone(), out["x_hat"].clone()))
    for r in res[1:]:
        assert torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1]) and torch.equal(r[2], res[0][2])


def test_full_size_batch_properties(ldic):
    """BASELINE configs[1] shape (768x512) through size-independent properties: images are independent units, so a
    batch must give exactly the per-image squared errors of its images run alone (integer sums) and the sum of their
    log-likelihood sums; a second run of the same batch is bit-identical (fixed summation order everywhere except the
    exact integer atomics)."""
    B, H, W = 4, 512, 768
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    x = dw.make_input(11, B, H, W).cuda()
    out = net.rd_forward(x)
    bits, sq = out["bits"].clone(), out["sq_err"].clone()
    out2 = net.rd_forward(x)
    assert torch.equal(out2["bits"], bits) and torch.equal(out2["sq_err"], sq)
    single_bits = torch.zeros(3, dtype=torch.float64)
    for i in range(B):
        oi = net.rd_forward(x[i:i + 1].contiguous())
        assert oi["sq_err"].item() == sq[i].item()
        single_bits += oi["bits"].double().cpu()
    assert torch.allclose(single_bits, bits.double().cpu(), rtol=2e-6, atol=0)
    # symbols: the rounded latent the kernels used is torch.round of the latent they produced, bit for bit
    y = out["latents"]["y"]
    z = out["latents"]["z"]
    _, _, yr = ldic.ops.latent_prep(y, want_round_bf16=False, want_abs_bf16=False, want_round_f32=True)
    assert torch.equal(yr, torch.round(y)) and torch.isfinite(z).all()
    bpp, v_mse, v_psnr = net.metrics(out, B, H, W)
    assert v_mse.shape == (B,) and torch.isfinite(bpp) and torch.isfinite(v_psnr)


@pytest.mark.parametrize("B,H,W", [(3, 64, 128), (5, 128, 64), (1, 192, 320), (7, 64, 64)])
def test_batch_independence_odd_shapes(ldic, B, H, W):
    """Odd batch sizes / tile counts (the CTA-pair kernels pad the last pair, the wide tail overlaps tiles in x):
    every image of a batch gives exactly the squared error it gives alone, and the batch's log-likelihood sums are the
    sums over its images."""
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    x = dw.make_input(5, B, H, W).cuda()
    out = net.rd_forward(x)
    bits, sq = out["bits"].double().cpu(), out["sq_err"].cpu()
    acc = torch.zeros(3, dtype=torch.float64)
    for i in range(B):
        oi = net.rd_forward(x[i:i + 1].contiguous())
        assert oi["sq_err"].item() == sq[i].item(), (i, oi["sq_err"].item(), sq[i].item())
        acc += oi["bits"].double().cpu()
    assert torch.allclose(acc, bits, rtol=5e-6, atol=0), (acc, bits)

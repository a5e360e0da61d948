"""GPU parity of the tcgen05 conv / deconv kernels (+ fused GDN / IGDN epilogue) against the CPU
oracle.  The oracle is evaluated on the SAME bf16-rounded operands the tensor cores see (fp32
accumulation on both sides), so the tolerances below only cover summation order and the bf16
rounding of x^2 / gamma inside the GDN epilogue; the end-to-end effect of bf16 operands is gated
separately in test_gpu_net.py (bpp 0.5 %, PSNR 0.01 dB)."""
import pytest
import torch
import torch.nn.functional as F

import det_weights as dw
from oracle import ref_path as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ldic():
    import ldic_b200
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(0), "device")
    return ldic_b200


def bf(t):
    return t.to(torch.bfloat16).float()


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def gdn_params(C, seed):
    sd = {}
    dw._gdn(sd, seed, "g", C)
    return sd["g.beta"], sd["g.gamma"]


def gdn_oracle_bf16(x, beta_p, gamma_p, inverse):
    """model/gdn.py arithmetic with the two tensor-core operands (x^2, gamma) rounded to bf16."""
    C = x.shape[1]
    beta, gamma = rp.gdn_effective_params_model(beta_p, gamma_p)
    norm = F.conv2d(bf(x * x), bf(gamma).view(C, C, 1, 1), beta)
    return x * torch.sqrt(norm) if inverse else x * torch.rsqrt(norm)


def to_nhwc_bf16(x, Cp=None):
    B, C, H, W = x.shape
    Cp = Cp or C
    y = torch.zeros(B, H, W, Cp, dtype=torch.bfloat16)
    y[..., :C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return y.cuda()


def close(a, b, rtol, atol):
    err = (a - b).abs()
    bad = err > rtol * b.abs() + atol
    assert not bad.any(), f"{int(bad.sum())}/{a.numel()} off; max abs err {err.max():.3e}, ref max {b.abs().max():.3e}"


def test_gemm_1x1_ragged_rows(ldic):
    L = ldic._lib
    P, K, Cout = 300, 75, 192            # 300 rows: last tile is partial (TMA zero fill + masked stores)
    a, w, b = bf(rnd((P, K), 1)), rnd((Cout, K), 2, 0.1), rnd((Cout,), 3)
    layer = ldic.ops.ConvTC(L.LDIC_CONV_1x1, w.cuda(), b.cuda(), out_f32=True, cin_pad=128)
    x = torch.zeros(1, 1, P, 128, dtype=torch.bfloat16)
    x[0, 0, :, :K] = a.to(torch.bfloat16)
    y = layer(x.cuda()).cpu().view(P, Cout)
    close(y, a @ bf(w).t() + b, 1e-4, 1e-4)


@pytest.mark.parametrize("B,H,W,act", [(2, 16, 24, "none"), (1, 64, 96, "gdn"), (3, 8, 8, "relu")])
def test_conv_s2_p12(ldic, B, H, W, act):
    L = ldic._lib
    C = 192
    x, w, b = bf(rnd((B, C, H, W), 4)), rnd((C, C, 5, 5), 5, 0.02), rnd((C,), 6, 0.1)
    ref = F.conv2d(F.pad(x, (1, 2, 1, 2)), bf(w), b, stride=2)
    kw = {}
    if act == "gdn":
        bp, gp = gdn_params(C, 7)
        ref = gdn_oracle_bf16(ref, bp, gp, False)
        kw = dict(act=L.ACT_GDN, gdn=(bp.cuda(), gp.cuda()) + rp.model_gdn_constants())
    elif act == "relu":
        ref = F.relu(ref)
        kw = dict(act=L.ACT_RELU)
    layer = ldic.ops.ConvTC(L.LDIC_CONV_S2_5x5_P12, w.cuda(), b.cuda(), out_f32=True, **kw)
    y = layer(to_nhwc_bf16(x)).cpu().permute(0, 3, 1, 2)
    close(y, ref, 2e-3 if act == "gdn" else 1e-4, 2e-4)
    # bf16 output path = the same values rounded once
    layer16 = ldic.ops.ConvTC(L.LDIC_CONV_S2_5x5_P12, w.cuda(), b.cuda(), out_f32=False, **kw)
    y16 = layer16(to_nhwc_bf16(x)).float().cpu().permute(0, 3, 1, 2)
    close(y16, ref, 1e-2, 1e-3)


@pytest.mark.parametrize("B,H,W,C,act", [(2, 16, 24, 192, "gdn"), (1, 64, 200, 192, "gdn"), (3, 6, 8, 128, "none"),
                                         (1, 130, 132, 64, "gdn")])
def test_first_layer_fused_from_nchw_image(ldic, B, H, W, C, act):
    """model/net.py:97-99: ZeroPad2d((1,2,1,2)) + Conv2d(3,C,5,2) + GDN read straight from the NCHW fp32 image
    (ragged widths: W/2 not a multiple of the 64-pixel tile; H/2 odd)."""
    L = ldic._lib
    x, w, b = rnd((B, 3, H, W), 40), rnd((C, 3, 5, 5), 41, 0.2), rnd((C,), 42, 0.1)
    ref = F.conv2d(F.pad(bf(x), (1, 2, 1, 2)), bf(w), b, stride=2)
    kw = {}
    if act == "gdn":
        bp, gp = gdn_params(C, 43)
        ref = gdn_oracle_bf16(ref, bp, gp, False)
        kw = dict(act=L.ACT_GDN, gdn=(bp.cuda(), gp.cuda()) + rp.model_gdn_constants())
    layer = ldic.ops.ConvTC(L.LDIC_CONV_FIRST_5x5S2, w.cuda(), b.cuda(), out_f32=True, **kw)
    y = layer(x.cuda()).cpu().permute(0, 3, 1, 2)
    assert y.shape == ref.shape
    close(y, ref, 2e-3 if act == "gdn" else 1e-4, 2e-4)
    layer16 = ldic.ops.ConvTC(L.LDIC_CONV_FIRST_5x5S2, w.cuda(), b.cuda(), out_f32=False, **kw)
    y16 = layer16(x.cuda()).float().cpu().permute(0, 3, 1, 2)
    close(y16, ref, 1e-2, 1e-3)


def test_conv_s2_p2_and_s1_3x3(ldic):
    L = ldic._lib
    C = 192
    x = bf(rnd((2, C, 8, 12), 8))
    w5, b5 = rnd((C, C, 5, 5), 9, 0.02), rnd((C,), 10, 0.1)
    w3, b3 = rnd((C, C, 3, 3), 11, 0.03), rnd((C,), 12, 0.1)
    y = ldic.ops.ConvTC(L.LDIC_CONV_S2_5x5_P2, w5.cuda(), b5.cuda(), act=L.ACT_RELU, out_f32=True)(to_nhwc_bf16(x))
    close(y.cpu().permute(0, 3, 1, 2), F.relu(F.conv2d(x, bf(w5), b5, stride=2, padding=2)), 1e-4, 2e-4)
    y = ldic.ops.ConvTC(L.LDIC_CONV_S1_3x3_P1, w3.cuda(), b3.cuda(), act=L.ACT_LEAKY02, out_f32=True)(to_nhwc_bf16(x))
    close(y.cpu().permute(0, 3, 1, 2), F.leaky_relu(F.conv2d(x, bf(w3), b3, stride=1, padding=1), 0.2), 1e-4, 2e-4)


@pytest.mark.parametrize("B,H,W", [(1, 6, 8), (2, 32, 48)])
def test_deconv_gs_igdn_with_channel_offset(ldic, B, H, W):
    L = ldic._lib
    N, M = 192, 16
    x = torch.round(rnd((B, N, H, W), 13, 3.0))                       # rounded latent: exact in bf16
    w, b = rnd((N - M, N, 5, 5), 14, 0.02), rnd((N,), 15, 0.1)
    bp, gp = gdn_params(N, 16)
    ref = F.conv_transpose2d(F.pad(x[:, M:], (1, 0, 1, 0)), bf(w), b, stride=2, padding=3, output_padding=1)
    ref = gdn_oracle_bf16(ref, bp, gp, True)
    layer = ldic.ops.ConvTC(L.LDIC_DECONV_GS_5x5, w.cuda(), b.cuda(), act=L.ACT_IGDN, out_f32=True, cin_pad=N,
                            cin_offset=M, gdn=(bp.cuda(), gp.cuda()) + rp.model_gdn_constants())
    y = layer(to_nhwc_bf16(x)).cpu().permute(0, 3, 1, 2)
    assert y.shape == ref.shape
    close(y, ref, 2e-3, 5e-4)


@pytest.mark.parametrize("B,H,W", [(2, 12, 20), (1, 4, 30), (1, 5, 31), (3, 3, 1), (1, 9, 61), (1, 32, 48)])
def test_deconv_gs_merged_small_cout(ldic, B, H, W):
    """Merged last deconv (wide-N form: dx taps in N, 32 x 4 tiles overlapping by one halo pixel per side in x):
    widths around the 30-pixel tile stride, heights that are not multiples of 4, a one-pixel-wide image."""
    L = ldic._lib
    N, M = 192, 16
    x = bf(rnd((B, N, H, W), 17))
    w, b = rnd((N, M, 5, 5), 18, 0.02), rnd((M,), 19, 0.1)
    bp, gp = gdn_params(M, 20)
    ref = F.conv_transpose2d(F.pad(x, (1, 0, 1, 0)), bf(w), b, stride=2, padding=3, output_padding=1)
    ref = gdn_oracle_bf16(ref, bp, gp, True)
    layer = ldic.ops.ConvTC(L.LDIC_DECONV_GS_5x5_MERGED, w.cuda(), b.cuda(), act=L.ACT_IGDN, out_f32=True,
                            gdn=(bp.cuda(), gp.cuda()) + rp.model_gdn_constants())
    y = layer(to_nhwc_bf16(x)).cpu().permute(0, 3, 1, 2)
    assert y.shape == ref.shape == (B, M, 2 * H, 2 * W)
    close(y, ref, 2e-3, 5e-4)


def test_deconv_gs_merged_wide_vs_tap_accumulating_form(ldic):
    """The wide-N kernel against the generic formulation of the same layer (nine N=64 taps accumulated in TMEM,
    tuning switch tail_wide=0): same operands, fp32 accumulation on both sides, only the summation order of the three dx
    contributions differs -- which can flip the bf16 rounding of an x^2 operand of the IGDN contraction, hence the
    bf16-operand tolerance of the oracle comparisons rather than an fp32 one."""
    L = ldic._lib
    N, M = 192, 16
    x = to_nhwc_bf16(bf(rnd((2, N, 20, 45), 21)))
    w, b = rnd((N, M, 5, 5), 22, 0.02), rnd((M,), 23, 0.1)
    bp, gp = gdn_params(M, 24)
    layer = ldic.ops.ConvTC(L.LDIC_DECONV_GS_5x5_MERGED, w.cuda(), b.cuda(), act=L.ACT_IGDN, out_f32=True,
                            gdn=(bp.cuda(), gp.cuda()) + rp.model_gdn_constants())
    y_wide = layer(x).cpu()
    prev = ldic.ops.set_tuning("tail_wide", 0)
    try:
        y_taps = layer(x).cpu()
    finally:
        ldic.ops.set_tuning("tail_wide", prev)
    close(y_wide, y_taps, 2e-3, 5e-4)


def test_deconv_hs_and_s1(ldic):
    L = ldic._lib
    C = 192
    x = bf(rnd((2, C, 3, 5), 21))
    w5, b5 = rnd((C, C, 5, 5), 22, 0.02), rnd((C,), 23, 0.1)
    w3, b3 = rnd((C, C, 3, 3), 24, 0.03), rnd((C,), 25, 0.1)
    y = ldic.ops.ConvTC(L.LDIC_DECONV_HS_5x5, w5.cuda(), b5.cuda(), act=L.ACT_RELU, out_f32=True)(to_nhwc_bf16(x))
    ref = F.relu(F.conv_transpose2d(x, bf(w5), b5, stride=2, padding=2, output_padding=1))
    close(y.cpu().permute(0, 3, 1, 2), ref, 1e-4, 2e-4)
    y = ldic.ops.ConvTC(L.LDIC_DECONV_S1_3x3, w3.cuda(), b3.cuda(), out_f32=True)(to_nhwc_bf16(x))
    close(y.cpu().permute(0, 3, 1, 2), F.conv_transpose2d(x, bf(w3), b3, stride=1, padding=1), 1e-4, 2e-4)


def test_tc_conv_matches_cuda_core_reference_kernel(ldic):
    """Same layer on the fp32 CUDA-core validation kernel (covers sizes the CPU oracle is slow at)."""
    L = ldic._lib
    C = 192
    x, w, b = bf(rnd((2, C, 64, 96), 26)), rnd((C, C, 5, 5), 27, 0.02), rnd((C,), 28, 0.1)
    y = ldic.ops.ConvTC(L.LDIC_CONV_S2_5x5_P12, w.cuda(), b.cuda(), out_f32=True)(to_nhwc_bf16(x))
    yr = ldic.ops.conv_reference_f32(L.LDIC_CONV_S2_5x5_P12, x.permute(0, 2, 3, 1).contiguous().cuda(), bf(w).cuda(), b.cuda())
    close(y.cpu(), yr.cpu(), 1e-4, 2e-4)
    ref = F.conv2d(F.pad(x, (1, 2, 1, 2)), bf(w), b, stride=2)
    close(yr.cpu().permute(0, 3, 1, 2), ref, 1e-4, 2e-4)


def test_transform_modules_vs_oracle(ldic):
    """Module surface (NCHW fp32 in / out) of g_a, h_a, h_s against the fp32 oracle: bf16-operand
    budget, relative RMS error."""
    sd = dw.make_state_dict(0)
    net = ldic.Net((1, 64, 64, 3), (1, 64, 64, 3), False, False).cuda()
    net.load_state_dict(sd, strict=True)
    x = dw.make_input(0, 2, 64, 128)
    rel = lambda a, b: ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
    with torch.no_grad():
        y_ref = rp.analysis_transform(sd, x)
        y = net.a_model(x.cuda()).cpu()
        assert y.shape == y_ref.shape and rel(y, y_ref) < 1e-2
        z_ref = rp.h_analysis_transform(sd, y_ref)
        z = net.ha_model(y_ref.cuda()).cpu()
        assert z.shape == z_ref.shape and rel(z, z_ref) < 1e-2
        h_ref = rp.h_synthesis_transform(sd, torch.round(z_ref))
        hh = net.hs_model(torch.round(z_ref).cuda()).cpu()
        assert hh.shape == h_ref.shape and rel(hh, h_ref) < 1e-2


def test_context_model_tc_vs_oracle(ldic):
    """PredictionModel_Context on the tcgen05 kernels (TMA patch gather) against the
    oracle's BlockSample + convs + fc (model/net.py:219-242, 289-319)."""
    sd = dw.make_state_dict(0)
    net = ldic.Net((1, 64, 64, 3), (1, 64, 64, 3), False, False).cuda()
    net.load_state_dict(sd, strict=True)
    B, h, w, N, M = 2, 6, 9, 192, 16
    yr = torch.round(rnd((B, N - M, h, w), 31, 2.0))
    h2 = bf(rnd((B, N, h, w), 32, 0.8))                      # bf16-exact input isolates the kernel error
    with torch.no_grad():
        mu_ref, sg_ref = rp.prediction_context(sd, yr, h2)
        mu, sg = net.prediction_model(yr.cuda(), h2.cuda())
    assert mu.shape == mu_ref.shape and sg.shape == sg_ref.shape
    rel = lambda a, b: ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
    assert rel(mu.cpu(), mu_ref) < 1e-2, rel(mu.cpu(), mu_ref)
    assert rel(torch.log(sg.cpu()), torch.log(sg_ref)) < 1e-2
    # weights rounded to bf16 on the oracle side as well: only summation order / bf16 activations remain
    sdb = {k: (bf(v) if k.startswith("prediction_model.") and k.endswith("weight") else v) for k, v in sd.items()}
    with torch.no_grad():
        mu_b, _ = rp.prediction_context(sdb, yr, h2)
    assert rel(mu.cpu(), mu_b) < 6e-3


@pytest.mark.parametrize("C,act,f32", [(192, "leaky001", False), (64, "none", True), (128, "leaky001", False)])
def test_conv_fused_residual(ldic, C, act, f32):
    """ldic_conv_forward_residual: y = act(conv3x3(x) + b) + r, r NHWC bf16 (ResidualBlock of CompressAI as used by
    layers/layers.py:87-102: conv -> LeakyReLU(0.01) -> + identity), against torch on the same bf16 operands."""
    L = ldic._lib
    x = bf(rnd((2, C, 20, 28), 51))
    r = bf(rnd((2, C, 20, 28), 52))
    w, b = rnd((C, C, 3, 3), 53, 0.03), rnd((C,), 54, 0.1)
    ref = F.conv2d(x, bf(w), b, padding=1)
    a = L.ACT_NONE
    if act == "leaky001":
        ref = F.leaky_relu(ref, 0.01)
        a = L.ACT_LEAKY001
    ref = ref + r
    layer = ldic.ops.ConvTC(L.LDIC_CONV_S1_3x3_P1, w.cuda(), b.cuda(), act=a, out_f32=f32)
    y = layer(to_nhwc_bf16(x), residual=to_nhwc_bf16(r))
    assert y.dtype == (torch.float32 if f32 else torch.bfloat16)
    if f32:
        close(y.cpu().permute(0, 3, 1, 2), ref, 1e-4, 3e-4)
    else:
        close(y.float().cpu().permute(0, 3, 1, 2), ref, 1e-2, 4e-3)
    y0 = layer(to_nhwc_bf16(x))                                  # the same layer without a residual still works
    close(y0.float().cpu().permute(0, 3, 1, 2), ref - r, 1e-2, 4e-3)

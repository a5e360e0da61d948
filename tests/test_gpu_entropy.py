"""GPU parity of the memory-bound kernels (through the C ABI) against the CPU oracle and the
golden fixtures generated from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

import det_weights as dw
from oracle import ref_path as rp

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")

# likelihood gate of BASELINE.md section 4 / SURVEY H1
LIK_RTOL, LIK_ATOL = 1e-4, 5e-7


def L(name):
    d = np.load(os.path.join(G, name))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


@pytest.fixture(scope="module")
def ldic():
    import ldic_b200
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(0), "device")
    return ldic_b200


def lik_close(a, b):
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    assert torch.equal(nan_a, nan_b), "NaN positions differ"
    a, b = a[~nan_a], b[~nan_b]
    bad = (a - b).abs() > LIK_RTOL * b.abs() + LIK_ATOL
    assert not bad.any(), f"{int(bad.sum())} likelihoods outside rtol {LIK_RTOL} atol {LIK_ATOL}; worst {(a - b).abs().max()}"


def test_lower_bound_and_reparam_bit_exact(ldic):
    d = L("leaf_ops.npz")
    x = d["lb_x"].cuda().requires_grad_(True)
    y = ldic.LowerBound(0.11).cuda()(x)
    y.backward(d["lb_gout"].cuda())
    assert torch.equal(y.detach().cpu(), d["lb_out"]) and torch.equal(x.grad.cpu(), d["lb_grad"])
    p = ldic.NonNegativeParametrizer().cuda()
    pb = ldic.NonNegativeParametrizer(minimum=1e-6).cuda()
    assert torch.equal(p(d["nn_p"].cuda()).cpu(), d["nn_fwd_min0"])
    assert torch.equal(pb(d["nn_p"].cuda()).cpu(), d["nn_fwd_beta"])
    # init() is a torch-op helper (parameter initialisation, not a kernel): CUDA sqrt may differ by 1 ulp
    assert torch.allclose(p.init(d["nn_p"].cuda()).cpu(), d["nn_init"], rtol=3e-7, atol=0)
    assert torch.equal(ldic.NonNegativeParametrizer().init(d["nn_p"]), d["nn_init"])
    # reference KAT ops/parametrizers.py:52-58
    g = p(p.init(0.1 * torch.eye(5).cuda())).cpu()
    assert abs(g[0, 0].item() - 0.1) < 1e-7 and g[0, 1].item() == 0.0
    # empty input
    assert ldic.LowerBound(0.11).cuda()(torch.empty(0, device="cuda")).numel() == 0


def test_round_bit_exact(ldic):
    d = L("leaf_ops.npz")
    x = d["rnd_in"].cuda()
    assert torch.equal(ldic.bypass_round(x).cpu(), d["bypass_round"])
    assert torch.equal(ldic.ste_round(x).cpu(), d["ste_round"])
    # signed zeros survive (-0.5 -> -0.0)
    r = ldic.bypass_round(torch.tensor([-0.5, -0.2, 0.5, 1.5, 2.5], device="cuda")).cpu()
    assert torch.equal(torch.signbit(r), torch.tensor([True, True, False, False, False]))
    assert r.tolist() == [-0.0, -0.0, 0.0, 2.0, 2.0]


@pytest.mark.parametrize("cls,key,tol", [("ModelGDN", "model_gdn", 2e-6), ("ModelIGDN", "model_igdn", 2e-6)])
def test_model_gdn_standalone(ldic, cls, key, tol):
    d = L("leaf_ops.npz")
    C = d["x"].shape[1]
    m = getattr(ldic, cls)(C).cuda()
    with torch.no_grad():
        m.beta.copy_(d["beta_p"]); m.gamma.copy_(d["gamma_p"])
        y = m(d["x"].cuda()).cpu()
    # fp32 CUDA-core kernel: same arithmetic as the reference, different summation order
    assert torch.allclose(y, d[key], rtol=tol, atol=1e-7)
    assert sorted(m.state_dict().keys()) == ["beta", "gamma", "pedestal", "reparam_offset"]


@pytest.mark.parametrize("inv", [False, True])
def test_layers_gdn_standalone(ldic, inv):
    d = L("leaf_ops.npz")
    C = d["x"].shape[1]
    m = ldic.GDN(C, inverse=inv).cuda()
    with torch.no_grad():
        m.beta.copy_(d["beta_p"]); m.gamma.copy_(d["gamma_p"])
        y = m(d["x"].cuda()).cpu()
    assert torch.allclose(y, d[f"layers_gdn_inv{int(inv)}"], rtol=2e-6, atol=1e-7)
    assert sorted(m.state_dict().keys()) == ["beta", "beta_reparam.lower_bound.bound", "beta_reparam.pedestal", "gamma",
                                             "gamma_reparam.lower_bound.bound", "gamma_reparam.pedestal"]


def test_gdn_init_matches_reference(ldic):
    d = L("leaf_ops.npz")
    C = d["x"].shape[1]
    assert torch.equal(ldic.ModelGDN(C).beta.data, d["model_gdn_init_beta"])
    assert torch.equal(ldic.ModelGDN(C).gamma.data, d["model_gdn_init_gamma"])
    assert torch.equal(ldic.GDN(C).beta.data, d["layers_gdn_init_beta"])
    assert torch.equal(ldic.GDN(C).gamma.data, d["layers_gdn_init_gamma"])
    assert list(ldic.ModelGDN(C).constants()) == d["model_gdn_consts"].tolist()
    assert list(ldic.GDN(C).constants()) == d["layers_gdn_consts"].tolist()


def test_gaussian_model_vs_reference_golden(ldic):
    d = L("gaussian_model.npz")
    n = d["v"].numel()
    v, mu, sg = (d[k].cuda().view(1, 1, 1, n) for k in ("v", "mu", "sigma"))
    # symbols bit-exact when fed the reference's latents (quant mode 1 == torch.round)
    vh, lik, s = ldic.ops.gaussian_likelihood(v, sg, mu, quant=ldic.ops.QUANT_ROUND, want_vhat=True)
    assert torch.equal(vh.cpu().flatten(), d["v_rounded"])
    lik_close(lik.cpu().flatten(), d["lik"])
    # module surface: GaussianModel()(inputs, sigma, mu)
    lik2 = ldic.GaussianModel()(d["v_rounded"].cuda().view(1, 1, 1, n), sg, mu)
    lik_close(lik2.cpu().flatten(), d["lik"])
    # factorised (1,C,1,1) sigma, mu = 0
    zl = ldic.GaussianModel()(d["z_rounded"].cuda(), d["z_sigma"].cuda(), torch.zeros_like(d["z_sigma"]).cuda())
    lik_close(zl.cpu(), d["z_lik"])


@pytest.mark.parametrize("n", [0, 1, 3, 4, 1000, 4099, 1 << 20])
def test_likelihood_sizes_and_sum(ldic, n):
    v, mu, sg = dw.likelihood_synthetic(1, max(n, 1))
    v, mu, sg = v[:n], mu[:n], sg[:n]
    if n >= 24:
        sg = sg.clone(); sg[14] = 0.5      # keep the sum finite here (sigma == 0 -> NaN is covered by the golden test)
    ref = rp.gaussian_model_likelihood(torch.round(v), sg, mu)
    _, lik, s = ldic.ops.gaussian_likelihood(v.cuda().view(1, 1, 1, n), sg.cuda().view(1, 1, 1, n), mu.cuda().view(1, 1, 1, n),
                                             quant=ldic.ops.QUANT_ROUND)
    lik_close(lik.cpu().flatten(), ref)
    ref_sum = torch.log(ref.double()).sum().item()
    # in the tails (true mass < 1e-6) the fp32 erf difference is quantised to 2^-24 steps, so the
    # CPU and CUDA erf disagree by up to that quantum there; on this heavy-tailed synthetic the total bits
    # move by < 0.2 % (measured 5e-4), well inside the 0.5 % bpp gate.
    assert abs(s.item() - ref_sum) <= 2e-3 * abs(ref_sum) + 1e-6
    # run twice: the workspace ticket is left reusable and the reduction is deterministic
    _, _, s2 = ldic.ops.gaussian_likelihood(v.cuda().view(1, 1, 1, n), sg.cuda().view(1, 1, 1, n), mu.cuda().view(1, 1, 1, n),
                                            quant=ldic.ops.QUANT_ROUND)
    assert s.item() == s2.item()


def test_likelihood_nan_propagates_like_reference(ldic):
    v = torch.tensor([[-0.5, 1.0, 2.0, 3.0]]).view(1, 1, 1, 4)
    sg = torch.tensor([[0.0, 1.0, 1.0, 1.0]]).view(1, 1, 1, 4)    # (v - mu + .5)/0 = 0/0 -> NaN
    mu = torch.zeros(1, 1, 1, 4)
    ref = rp.gaussian_model_likelihood(v, sg, mu)
    _, lik, s = ldic.ops.gaussian_likelihood(v.cuda(), sg.cuda(), mu.cuda())
    assert torch.isnan(ref[0, 0, 0, 0]) and torch.isnan(lik[0, 0, 0, 0].cpu()) and torch.isnan(s.cpu()).all()


def test_likelihood_strided_rows_and_log_sigma(ldic):
    # the Net layout: v = channels [16:] of an NHWC latent, mu / log-sigma = halves of the fc output
    P, N, M = 37, 192, 16
    Cc = N - M
    g = torch.Generator().manual_seed(5)
    y = torch.randn(P, N, generator=g) * 3
    ctx = torch.randn(P, 2 * Cc, generator=g) * 0.5
    ref = rp.gaussian_model_likelihood(torch.round(y[:, M:]), torch.exp(ctx[:, Cc:]), ctx[:, :Cc])
    yc, cc = y.cuda(), ctx.cuda()
    lik = torch.empty(P, Cc, device="cuda")
    vh16 = torch.zeros(P, N, dtype=torch.bfloat16, device="cuda")
    s = ldic.ops.likelihood_rows(yc, P, Cc, v_rs=N, v_off=M, mu=cc, mu_mode=2, mu_rs=2 * Cc, sigma=cc, sigma_mode=2,
                                 sigma_rs=2 * Cc, sigma_off=Cc, sigma_is_log=True, quant=ldic.ops.QUANT_ROUND, lik=lik,
                                 v_hat_bf16=vh16, vb_rs=N, vb_off=M)
    lik_close(lik.cpu(), ref)
    assert torch.equal(vh16[:, M:].float().cpu(), torch.round(y[:, M:])) and (vh16[:, :M] == 0).all()
    assert abs(s.item() - torch.log(ref.double()).sum().item()) < 2e-3 * abs(s.item())


def test_gaussian_conditional_vs_oracle(ldic):
    g = torch.Generator().manual_seed(9)
    y = torch.randn(2, 8, 5, 6, generator=g) * 4
    mu = torch.randn(2, 8, 5, 6, generator=g)
    sc = torch.exp(torch.randn(2, 8, 5, 6, generator=g)) * 0.3     # some below the 0.11 bound
    yh_ref, lik_ref = rp.gaussian_conditional(y, sc, mu)
    yh, lik = ldic.GaussianConditional()(y.cuda(), sc.cuda(), mu.cuda())
    assert torch.equal(yh.cpu(), yh_ref)            # round(y - mu) + mu, two separately rounded fp32 ops
    lik_close(lik.cpu(), lik_ref)


def test_mse_psnr_exact(ldic):
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3, 3, 40, 52, generator=g) * 2 - 1
    xt = x + 0.3 * torch.randn(3, 3, 40, 52, generator=g)
    for clamp in (False, True):
        mse_ref, psnr_ref = rp.mse_psnr(x, xt, clamp_pm1=clamp)
        sq = ldic.ops.mse_sum(x.cuda(), xt.cuda(), clamp_pm1=clamp)
        gt = torch.round((x + 1) * 127.5)
        xh = torch.round(torch.clamp(((torch.clamp(xt, -1, 1) if clamp else xt) + 1) * 127.5, 0, 255))
        assert torch.equal(sq.cpu(), ((xh - gt).double() ** 2).sum((1, 2, 3)).long())      # exact integers
        v_mse, v_psnr = ldic.psnr_from_sq_err(sq, 3 * 40 * 52)
        assert torch.allclose(v_mse.cpu(), mse_ref, rtol=1e-6)
        assert abs(v_psnr.item() - psnr_ref.item()) < 1e-4                                  # gate: 0.01 dB


def test_syntax_conv_mse(ldic):
    g = torch.Generator().manual_seed(3)
    B, M, H, W = 2, 16, 24, 40
    x = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    xt16 = torch.randn(B, M, H, W, generator=g)
    wt = torch.randn(B, 3, M, 1, 1, generator=g) * 0.2
    xt = rp.batch_conv(wt, xt16)
    mse_ref, _ = rp.mse_psnr(x, xt)
    sq, xo = ldic.ops.syntax_conv_mse(x.cuda(), xt16.permute(0, 2, 3, 1).contiguous().cuda(), wt.view(B, 3, M).cuda(),
                                      want_x_tilde=True)
    assert torch.allclose(xo.cpu(), xt, rtol=1e-5, atol=1e-5)
    v_mse, _ = ldic.psnr_from_sq_err(sq, 3 * H * W)
    assert torch.allclose(v_mse.cpu(), mse_ref, rtol=2e-3)      # a few .5 ties may round differently


def test_layout_glue_roundtrip(ldic):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 19, 7, 9, generator=g)
    y = ldic.ops.nchw_to_nhwc_bf16(x.cuda(), 64)
    assert y.shape == (2, 7, 9, 64) and (y[..., 19:] == 0).all()
    assert torch.equal(y[..., :19].float().cpu(), x.to(torch.bfloat16).float().permute(0, 2, 3, 1))
    back = ldic.ops.nhwc_to_nchw_f32(y, 19)
    assert torch.equal(back.cpu(), x.to(torch.bfloat16).float())
    ya = ldic.ops.nchw_to_nhwc_bf16(x.cuda(), 64, apply_abs=True)
    assert torch.equal(ya[..., :19].float().cpu(), x.abs().to(torch.bfloat16).float().permute(0, 2, 3, 1))


def test_no_cpu_fallback(ldic):
    with pytest.raises(ldic.LdicError):
        ldic.ops.gaussian_likelihood(torch.zeros(1, 1, 1, 4), torch.ones(1, 1, 1, 4))
    with pytest.raises(ldic.LdicError):
        ldic.ModelGDN(4)(torch.zeros(1, 4, 2, 2))

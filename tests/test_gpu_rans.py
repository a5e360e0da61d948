"""rANS entropy coder (csrc/rans.cu) on the GPU, through the C ABI: byte-for-byte equality with the CPU
restatement oracle/rans_ref.py and the committed known-answer vectors, exact round trips at the bench size, coded size
against the rate the likelihood kernel estimates for the same symbols, and the error paths.  The reference holds no
entropy coder, so these properties (not a reference output) are what the coder is pinned to."""
import math
import os

import numpy as np
import pytest
import torch

import det_weights as dw
from oracle import rans_ref as rr

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ldic():
    import ldic_b200
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(0), "device")
    return ldic_b200


def _flat_encode(ops, v, mu, sigma, S, quant=1, **kw):
    n = v.numel()
    return ops.rans_encode_rows(v, 1, n, 1, v_rs=n, mu=mu, mu_mode=0 if mu is None else 2, mu_rs=n, sigma=sigma, sigma_mode=2,
                                sigma_rs=n, quant=quant, streams=S, **kw)


def test_known_answer_vectors_on_the_gpu(ldic):
    ops = ldic.ops
    d = np.load(os.path.join(G, "rans_kat.npz"))
    for i in range(int(d["cases"])):
        c = {k: d[f"c{i}_{k}"] for k in ("v", "mu", "sigma", "k", "S", "quant", "blob")}
        S, quant, n = int(c["S"]), int(c["quant"]), c["v"].size
        v, mu, sigma = (torch.from_numpy(c[k]).cuda() for k in ("v", "mu", "sigma"))
        enc = _flat_encode(ops, v, mu, sigma, S, quant)
        blob = enc.tobytes()[0]
        assert blob == c["blob"].tobytes(), i
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        ops.rans_decode_rows([c["blob"].tobytes()], 1, n, 1, out, v_hat_rs=n, mu=mu, mu_mode=2, mu_rs=n, sigma=sigma, sigma_mode=2,
                             sigma_rs=n, quant=quant, streams=S)
        want = c["k"].astype(np.float32) if quant == 1 else (c["k"].astype(np.float32) + c["mu"]).astype(np.float32)
        assert np.array_equal(out.cpu().numpy(), want), i


def test_strided_nhwc_slices_and_broadcast_modes_vs_oracle(ldic):
    """Three images, content channels [M, N) of an NHWC latent, mu from a wider context tensor, sigma per channel
    (mode 1), then per row (mode 3, NCHW per-channel prior): every segment equals the oracle's bytes."""
    ops = ldic.ops
    g = torch.Generator().manual_seed(5)
    B, h, w, N, M = 3, 6, 10, 24, 4
    Cc = N - M
    y = (torch.randn(B, h, w, N, generator=g) * 4).cuda()
    ctx = torch.randn(B * h * w, 2 * 32, generator=g).cuda()                    # mu at [0, Cc), sigma at [32, 32 + Cc)
    ctx[:, 32:] = torch.exp(ctx[:, 32:] * 0.7)
    sig_c = torch.exp(torch.randn(Cc, generator=g)).cuda()
    S = 5
    enc = ops.rans_encode_rows(y, B * h * w, Cc, h * w, v_rs=N, v_off=M, mu=ctx, mu_mode=2, mu_rs=64, sigma=ctx, sigma_mode=2,
                               sigma_rs=64, sigma_off=32, streams=S)
    blobs = enc.tobytes()
    for b in range(B):
        k = torch.round(y[b, :, :, M:]).reshape(-1).cpu().numpy().astype(np.int64)
        rows = slice(b * h * w, (b + 1) * h * w)
        mu = ctx[rows, :Cc].reshape(-1).cpu().numpy()
        sg = ctx[rows, 32:32 + Cc].reshape(-1).cpu().numpy()
        assert blobs[b] == rr.encode_segment(k, mu, sg, S), b
    out = torch.zeros(B, h, w, N, device="cuda")
    ops.rans_decode_rows(blobs, B * h * w, Cc, h * w, out, v_hat_rs=N, v_hat_off=M, mu=ctx, mu_mode=2, mu_rs=64, sigma=ctx,
                         sigma_mode=2, sigma_rs=64, sigma_off=32, streams=S)
    assert torch.equal(out[..., M:], torch.round(y[..., M:])) and out[..., :M].abs().sum().item() == 0
    # column groups: the same symbols coded group by group (two groups of 10 channels), against the oracle on the reordered arrays
    enc = ops.rans_encode_rows(y, B * h * w, Cc, h * w, v_rs=N, v_off=M, mu=ctx, mu_mode=2, mu_rs=64, sigma=ctx, sigma_mode=2,
                               sigma_rs=64, sigma_off=32, streams=6, col_groups=2)
    gblobs = enc.tobytes()
    order = rr.group_order(h * w, Cc, 2)
    for b in range(B):
        k = torch.round(y[b, :, :, M:]).reshape(-1).cpu().numpy().astype(np.int64)
        rows = slice(b * h * w, (b + 1) * h * w)
        mu = ctx[rows, :Cc].reshape(-1).cpu().numpy()
        sg = ctx[rows, 32:32 + Cc].reshape(-1).cpu().numpy()
        assert gblobs[b] == rr.encode_segment(k[order], mu[order], sg[order], 6, groups=2), b
    out = torch.zeros(B, h, w, N, device="cuda")
    ops.rans_decode_rows(gblobs, B * h * w, Cc, h * w, out, v_hat_rs=N, v_hat_off=M, mu=ctx, mu_mode=2, mu_rs=64, sigma=ctx,
                         sigma_mode=2, sigma_rs=64, sigma_off=32, streams=6, col_groups=2)
    assert torch.equal(out[..., M:], torch.round(y[..., M:]))
    with pytest.raises(ldic.LdicError, match="bad header"):                      # decoded with another grouping than coded with
        ops.rans_decode_rows(gblobs, B * h * w, Cc, h * w, out, v_hat_rs=N, v_hat_off=M, mu=ctx, mu_mode=2, mu_rs=64, sigma=ctx,
                             sigma_mode=2, sigma_rs=64, sigma_off=32, streams=6)
    # per-channel sigma, no mean (the z stream of model/net.py:676,:781)
    enc = ops.rans_encode_rows(y, B * h * w, Cc, h * w, v_rs=N, v_off=M, sigma=sig_c, sigma_mode=1, streams=S)
    blobs = enc.tobytes()
    for b in range(B):
        k = torch.round(y[b, :, :, M:]).reshape(-1).cpu().numpy().astype(np.int64)
        sg = sig_c.cpu().numpy()[None, :].repeat(h * w, 0).reshape(-1)
        assert blobs[b] == rr.encode_segment(k, np.zeros_like(sg), sg, S), b
    # module surface, NCHW with (1,C,1,1) sigma: rows = B*C, mode 3
    v = y.permute(0, 3, 1, 2).contiguous()
    sig_n = torch.exp(torch.randn(1, N, 1, 1, generator=g)).cuda()
    enc = ops.rans_encode(v, sig_n, streams=7)
    blobs = enc.tobytes()
    for b in range(B):
        k = torch.round(v[b]).reshape(-1).cpu().numpy().astype(np.int64)
        sg = sig_n.expand(1, N, h, w).reshape(-1).cpu().numpy()
        assert blobs[b] == rr.encode_segment(k, np.zeros_like(sg), sg, 7), b
    assert torch.equal(ops.rans_decode(blobs, v.shape, sig_n, streams=7), torch.round(v))
    assert torch.equal(ops.rans_decode(enc, v.shape, sig_n), torch.round(v))       # device-resident streams
    # batched readback (two synchronisations for any number of streams), optionally on a copy stream
    enc2 = ops.rans_encode(v, sig_n, streams=3)
    assert ops.rans_tobytes([enc, enc2]) == [blobs, enc2.tobytes()]
    assert ops.rans_tobytes([enc, enc2], copy_stream=torch.cuda.Stream()) == [blobs, enc2.tobytes()]


def test_dequantize_form_with_scale_bound_vs_oracle(ldic):
    """quant 2 (round(v - mu) + mu under N(0, max(sigma, 0.11)): GaussianConditional, model/net_unet_ha_hs.py:937)."""
    ops = ldic.ops
    g = torch.Generator().manual_seed(9)
    v = (torch.randn(2, 12, 8, 8, generator=g) * 3).cuda()
    mu = torch.randn(2, 12, 8, 8, generator=g).cuda()
    sigma = torch.exp(torch.randn(2, 12, 8, 8, generator=g) - 1.5).cuda()
    enc = ops.rans_encode(v, sigma, mu, quant=ops.QUANT_DEQUANT, scale_bound=0.11, streams=3)
    blobs = enc.tobytes()
    for b in range(2):
        k = torch.round(v[b] - mu[b]).reshape(-1).cpu().numpy().astype(np.int64)
        sg = torch.clamp(sigma[b], min=0.11).reshape(-1).cpu().numpy()
        assert blobs[b] == rr.encode_segment(k, np.zeros_like(sg), sg, 3, quant=2)
    v_hat, _, _ = ops.gaussian_likelihood(v, sigma, mu, quant=ops.QUANT_DEQUANT, form=ops.FORM_GAUSSIAN_CONDITIONAL, want_vhat=True)
    dec = ops.rans_decode(blobs, v.shape, sigma, mu, quant=ops.QUANT_DEQUANT, scale_bound=0.11, streams=3)
    assert torch.equal(dec, v_hat)                                   # the very tensor the synthesis transform consumes


def test_round_trip_and_rate_at_the_bench_size(ldic):
    """16 images' content latents (768x512: 32x48x176 each) with a (mu | log sigma) context tensor, as Net.forward holds
    them: exact round trip, and 8 * bytes within 1 % + headers of the sum(-log2 L) the likelihood kernel returns."""
    ops = ldic.ops
    g = torch.Generator(device="cuda").manual_seed(1)
    B, h, w, N, M, Cp = 16, 32, 48, 192, 16, 192
    Cc, P = N - M, B * h * w
    ctx = torch.empty(P, 2 * Cp, device="cuda")
    ctx[:, :Cp] = torch.randn(P, Cp, device="cuda", generator=g) * 3
    ctx[:, Cp:] = torch.randn(P, Cp, device="cuda", generator=g) * 0.8 + 0.2          # log sigma
    y = torch.zeros(B, h, w, N, device="cuda")
    y[..., M:] = ctx[:, :Cc].reshape(B, h, w, Cc) + torch.exp(ctx[:, Cp:Cp + Cc]).reshape(B, h, w, Cc) * torch.randn(
        B, h, w, Cc, device="cuda", generator=g)
    kw = dict(mu=ctx, mu_mode=2, mu_rs=2 * Cp, sigma=ctx, sigma_mode=2, sigma_rs=2 * Cp, sigma_off=Cp, sigma_is_log=True)
    enc = ops.rans_encode_rows(y, P, Cc, h * w, v_rs=N, v_off=M, **kw)
    sizes = enc.nbytes()
    S = enc.streams
    assert S == ops.rans_streams_for(h * w * Cc) == 132
    out = torch.empty(B, h, w, Cc, device="cuda")
    ops.rans_decode_rows(enc, P, Cc, h * w, out, v_hat_rs=Cc, **kw)
    assert torch.equal(out, torch.round(y[..., M:]))
    s = ops.likelihood_rows(y, P, Cc, v_rs=N, v_off=M, quant=ops.QUANT_ROUND, lik_bound=1e-8, **kw)
    est_bits = -s.item() / math.log(2.0)
    hdr = [np.frombuffer(b[:32], dtype="<u4") for b in enc.tobytes()]
    overhead = sum(8 * (32 + 6 * S + 8 * int(hh[3])) for hh in hdr)
    coded = 8 * sum(sizes)
    assert est_bits * 0.99 < coded - overhead < est_bits * 1.01, (coded, overhead, est_bits)
    assert overhead < 0.01 * coded
    # launches: 6 kernels for the 16 streams together, 3 to decode
    n0 = ops.launch_count()
    ops.rans_encode_rows(y, P, Cc, h * w, v_rs=N, v_off=M, **kw)
    assert ops.launch_count() - n0 == 6


def test_escapes_bad_symbols_capacity_and_corrupt_streams(ldic):
    ops = ldic.ops
    g = torch.Generator().manual_seed(2)
    n = 5000
    mu = (torch.randn(n, generator=g) * 2).cuda()
    sigma = torch.exp(torch.randn(n, generator=g) * 0.5).cuda()
    v = (mu + sigma * torch.randn(n, generator=g).cuda()).contiguous()
    v[::50] += 777777.0                                               # escapes: far outside every window
    v[7::300] -= 3.0e8
    sigma[11::170] = 0.0
    sigma[12::190] = float("nan")
    enc = _flat_encode(ops, v, mu, sigma, 4)
    blob = enc.tobytes()[0]
    k = torch.round(v).cpu().numpy().astype(np.int64)
    assert blob == rr.encode_segment(k, mu.cpu().numpy(), sigma.cpu().numpy(), 4)
    assert np.frombuffer(blob[:32], dtype="<u4")[3] >= 100
    out = torch.empty(n, device="cuda")
    kw = dict(mu=mu, mu_mode=2, mu_rs=n, sigma=sigma, sigma_mode=2, sigma_rs=n, streams=4)
    ops.rans_decode_rows([blob], 1, n, 1, out, v_hat_rs=n, **kw)
    assert torch.equal(out, torch.round(v))
    # the CPU oracle decodes the GPU's stream too
    assert np.array_equal(rr.decode_segment(blob, mu.cpu().numpy(), sigma.cpu().numpy(), 4), k)
    # NaN / huge symbols cannot be coded: flagged, not silently wrapped
    vb = v.clone(); vb[3] = float("nan")
    with pytest.raises(ldic.LdicError, match="NaN or beyond"):
        _flat_encode(ops, vb, mu, sigma, 4).tobytes()
    vb = v.clone(); vb[3] = 3e9
    with pytest.raises(ldic.LdicError, match="NaN or beyond"):
        _flat_encode(ops, vb, mu, sigma, 4).tobytes()
    # capacity
    with pytest.raises(ldic.LdicError, match="capacity"):
        _flat_encode(ops, v, mu, sigma, 4, capacity=len(blob) - 4).tobytes()
    assert _flat_encode(ops, v, mu, sigma, 4, capacity=(len(blob) + 3) & ~3).tobytes()[0] == blob
    # corrupt input: header, truncation, payload
    for bad in (b"\x00" + blob[1:], blob[:-2], blob[:40], b""):
        with pytest.raises(ldic.LdicError, match="bad header"):
            ops.rans_decode_rows([bad], 1, n, 1, out, v_hat_rs=n, **kw)
    with pytest.raises(ldic.LdicError, match="bad header"):                       # other stream count than coded with
        ops.rans_decode_rows([blob], 1, n, 1, out, v_hat_rs=n, **{**kw, "streams": 5})
    flipped = bytearray(blob); flipped[-3] ^= 0x40
    st = ops.rans_decode_rows([bytes(flipped)], 1, n, 1, out, v_hat_rs=n, check_status=False, **kw)
    assert st.item() == 8 or not torch.equal(out, torch.round(v))
    words = bytearray(blob); words[32 + 16] ^= 0xFF                               # a stream's word count
    st = ops.rans_decode_rows([bytes(words)], 1, n, 1, out, v_hat_rs=n, check_status=False, **kw)
    assert st.item() & 8


def test_empty_and_tiny_inputs(ldic):
    ops = ldic.ops
    sigma = torch.ones(3, device="cuda")
    v = torch.tensor([0.4, -2.6, 7.0], device="cuda")
    for S in (1, 2, 3, 8):
        enc = _flat_encode(ops, v, None, sigma, S)
        blob = enc.tobytes()[0]
        assert blob == rr.encode_segment(np.array([0, -3, 7]), np.zeros(3, np.float32), np.ones(3, np.float32), S)
        out = torch.empty(3, device="cuda")
        ops.rans_decode_rows([blob], 1, 3, 1, out, v_hat_rs=3, sigma=sigma, sigma_mode=2, sigma_rs=3, streams=S)
        assert out.tolist() == [0.0, -3.0, 7.0]
    e = torch.empty(0, 4, 2, 2, device="cuda")
    assert ops.rans_encode(e, torch.ones(1, 4, 1, 1, device="cuda")).tobytes() == []
    with pytest.raises(ldic.LdicError):
        ops.rans_encode_rows(v, 1, 3, 1, v_rs=3, sigma=sigma, sigma_mode=2, sigma_rs=3, quant=0)
    with pytest.raises(ldic.LdicError):                                           # > 65535 symbols per stream
        big = torch.zeros(70000, device="cuda")
        _flat_encode(ops, big, None, torch.ones(70000, device="cuda"), 1)


@pytest.mark.parametrize("B,H,W", [(2, 128, 192), (1, 256, 256)])
def test_net_compress_bitstreams_round_trip(ldic, B, H, W):
    """Net.compress: real bytes next to the estimated rate of the same forward.  The z stream decodes from the model
    parameters alone and h_s of it reproduces the encoder's h2 bit for bit; y and syntax decode given the (mu, sigma)
    the encoder used."""
    ops = ldic.ops
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(3), strict=True)
    x = dw.make_input(3, B, H, W).cuda()
    streams, info = net.compress(x)
    assert len(streams) == B and set(streams[0]) == {"z", "y", "syntax"}
    out = net.rd_forward(x)
    lat = out["latents"]
    bpp_ref = net.metrics(out, B, H, W)[0].item()
    assert abs(info["bpp_estimated"] / bpp_ref - 1) < 1e-5
    hz, wz, h, w = H // 64, W // 64, H // 16, W // 16
    Sz, Sy = ops.rans_streams_for(hz * wz * net.N), net.y_streams(h, w)
    # untrained context models put many symbols deep in the tails, where the estimate charges up to -log2(1e-8) = 26.6
    # bits and the coder 16 (window) or 80 (escape): only the upper bound is tight here; the two-sided 1 % check is
    # test_round_trip_and_rate_at_the_bench_size, where the symbols follow their model
    esc = sum(int(np.frombuffer(s[k][:32], dtype="<u4")[3]) for s in streams for k in s)
    fixed = 8.0 * (B * (3 * 32 + 6 * (Sz + Sy + 1)) + 10 * esc) / (B * H * W)
    assert info["bpp_estimated"] * 0.9 < info["bpp_coded"] - fixed < info["bpp_estimated"] * 1.01 + 1e-3, (info, esc)
    z_hat = net.decode_z([s["z"] for s in streams], B, H, W)
    assert torch.equal(z_hat, torch.round(lat["z"]))
    h2 = net.hs_model.forward_nhwc(z_hat.to(torch.bfloat16))
    assert torch.equal(h2, lat["h2"])
    y_hat = net.decode_y([s["y"] for s in streams], lat["ctx"], lat["ctx_rs"], lat["ctx_sig_off"], B, H, W)
    assert torch.equal(y_hat, torch.round(lat["y"][..., net.M:]))
    syn = ops.rans_decode([s["syntax"] for s in streams], (B, net.M, 1, 1), lat["syn_first"].reshape(B, -1, 1, 1),
                          lat["syn_second"].reshape(B, -1, 1, 1), streams=1)
    assert torch.equal(syn.reshape(-1), torch.round(lat["z3_syntax"]).reshape(-1))
    # deterministic: the same input gives the same bytes
    again, _ = net.compress(x)
    assert again == streams


def test_torch_library_ops_reach_the_coder(ldic):
    """torch.ops.ldic.rans_encode / rans_decode (the C ABI behind torch custom ops) give the same bytes."""
    ops = ldic.ops
    g = torch.Generator().manual_seed(4)
    v = (torch.randn(3, 8, 6, 6, generator=g) * 3).cuda()
    sigma = torch.exp(torch.randn(3, 8, 6, 6, generator=g) * 0.5).cuda()
    mu = torch.randn(3, 8, 6, 6, generator=g).cuda()
    buf, sizes, status = torch.ops.ldic.rans_encode(v, sigma, mu, 1, 0.0, 2)
    assert status.abs().sum().item() == 0
    ref = ops.rans_encode(v, sigma, mu, streams=2).tobytes()
    host = buf.cpu().numpy()
    assert [host[i, :n].tobytes() for i, n in enumerate(sizes.tolist())] == ref
    out, st = torch.ops.ldic.rans_decode(buf, sizes, sigma, mu, list(v.shape), 1, 0.0, 2)
    assert st.abs().sum().item() == 0 and torch.equal(out, torch.round(v))


def test_unet_family_slice_bitstreams(ldic):
    """The four slice streams of the U-Net family (round(y - mu) under N(0, max(scale, 0.11)), model/net_unet_ha_hs.py:937):
    decoding each with the (mu, scale) of its slice returns exactly the dequantised tensor the forward used, and the coded
    size follows the estimated rate of the same call."""
    from ldic_b200 import net_unet
    ops = ldic.ops
    B, H, W, seed = 1, 256, 256, 0
    net = net_unet.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    fill = dw.unet_param_fill([(n, tuple(p.shape)) for n, p in net.named_parameters()], seed)
    net.load_state_dict({**net.state_dict(), **{k: v.cuda() for k, v in fill.items()}}, strict=True)
    x = dw.make_input(seed, B, H, W).cuda()
    out = net.rd_forward(x, want_bitstreams=True)
    plain = net.rd_forward(x)
    assert torch.equal(out["bits"], plain["bits"]) and torch.equal(out["sq_err"], plain["sq_err"])
    coded = 0
    for i, (enc, (mu, scale), want) in enumerate(zip(out["streams"], out["slice_params"], out["slice_symbols"])):
        blobs = enc.tobytes()
        coded += 8 * sum(len(b) for b in blobs)
        dec = ops.rans_decode(blobs, want.shape, scale, mu, quant=ops.QUANT_DEQUANT, scale_bound=0.11)
        assert torch.equal(dec, want), i
    est = -out["bits"].double().sum().item() / math.log(2.0)
    esc = sum(int(np.frombuffer(b[:32], dtype="<u4")[3]) for enc in out["streams"] for b in enc.tobytes())
    # untrained slice transforms leave many symbols deep in the tails of their model, where the estimate charges up to
    # -log2(1e-9) = 29.9 bits, the coder 16 (inside the window) or 96 (escape): only the upper bound is tight here (the
    # two-sided 1 % check on symbols that follow their model is test_round_trip_and_rate_at_the_bench_size)
    assert 0.5 * est < coded - 96 * esc < 1.02 * est + 8 * 4 * B * 200, (coded, est, esc)


def test_eval_driver_writes_decodable_containers(ldic):
    """evaluation.evaluate_images(bitstreams=True): per image an LDIC container whose z stream decodes from the file and
    the model alone (padded size from the header) and whose size gives a coded bpp next to the estimated one."""
    ev = ldic.evaluation
    net = ldic.Net((1, 64, 64, 3), (1, 64, 64, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(5), strict=True)
    g = torch.Generator().manual_seed(8)
    imgs = [torch.rand(3, 100, 70, generator=g), torch.rand(3, 128, 128, generator=g), torch.rand(3, 100, 70, generator=g)]
    res = ev.evaluate_images(net, imgs, batch_size=4, bitstreams=True)
    plain = ev.evaluate_images(net, imgs, batch_size=4)
    for r, p, img in zip(res, plain, imgs):
        assert r["bpp"] == p["bpp"] and r["psnr"] == p["psnr"]
        h, w, hp, wp, streams = ev.unpack_container(r["container"])
        assert (h, w) == (img.shape[1], img.shape[2]) and hp % 64 == 0 and wp % 64 == 0
        assert r["bpp_coded"] == 8.0 * len(r["container"]) / (h * w)
        x = ev.pad_to_multiple(img).cuda()
        lat = net.rd_forward(x)["latents"]
        assert torch.equal(net.decode_z([streams["z"]], 1, hp, wp), torch.round(lat["z"]))
        y_hat = net.decode_y([streams["y"]], lat["ctx"], lat["ctx_rs"], lat["ctx_sig_off"], 1, hp, wp)
        assert torch.equal(y_hat, torch.round(lat["y"][..., net.M:]))
        assert 0.5 * r["bpp"] < r["bpp_coded"] < 1.1 * r["bpp"] + 8.0 * 700 / (h * w)
    # and back: the files alone give the reconstruction of the encoder, cropped to the unpadded size
    dec = ev.decode_containers(net, [r["container"] for r in res], batch_size=2)
    for img, d in zip(imgs, dec):
        x = ev.pad_to_multiple(img).cuda()
        ref = net.rd_forward(x, want_x_hat=True)["x_hat"][0, :, :img.shape[1], :img.shape[2]]
        assert d.shape == img.shape and torch.equal(d, ((ref + 1.0) * 0.5).clamp(0.0, 1.0))


@pytest.mark.parametrize("B,H,W", [(2, 128, 192), (1, 256, 256)])
def test_net_decompress_reproduces_the_encoder_reconstruction(ldic, B, H, W):
    """The whole codec: x -> Net.compress -> bytes -> Net.decompress -> x_hat, from the bytes and the model alone (the
    causal context model walked along wavefronts with the encoder's own kernels).  The reconstruction and every decoded
    latent are bit-identical to the encoder's."""
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(7), strict=True)
    x = dw.make_input(7, B, H, W).cuda()
    streams, info = net.compress(x)
    enc = net.rd_forward(x, want_x_hat=True)
    x_hat, lat = net.decompress(streams, H, W, want_latents=True)
    assert torch.equal(lat["z_hat"], torch.round(enc["latents"]["z"]))
    assert torch.equal(lat["h2"], enc["latents"]["h2"])
    assert torch.equal(lat["y_hat"], torch.round(enc["latents"]["y"][..., net.M:]))
    assert torch.equal(lat["conv_w"].reshape(-1), enc["latents"]["conv_w"].reshape(-1))
    assert torch.equal(x_hat, enc["x_hat"])
    assert torch.equal(net.decompress(streams, H, W, schedule="full"), x_hat)      # whole-latent context passes: same result
    # second call of a shape: the wavefront loop is captured into a CUDA graph; third: replayed on other bitstreams
    assert torch.equal(net.decompress(streams, H, W), x_hat)
    x2 = torch.roll(x, shifts=17, dims=3)
    streams2, _ = net.compress(x2)
    assert torch.equal(net.decompress(streams2, H, W), net.rd_forward(x2, want_x_hat=True)["x_hat"])
    assert len(net._decode_graphs) >= 1 and any(isinstance(v, dict) for v in net._decode_graphs.values())
    # a damaged content stream is detected, not silently decoded
    bad = [dict(s) for s in streams]
    yb = bytearray(bad[0]["y"]); yb[len(yb) // 2] ^= 0x5A; bad[0]["y"] = bytes(yb)
    try:
        x_bad = net.decompress(bad, H, W)
        assert not torch.equal(x_bad, enc["x_hat"])
    except ldic.LdicError:
        pass


def test_incremental_decode_order_is_enforced(ldic):
    ops = ldic.ops
    g = torch.Generator().manual_seed(6)
    n, S = 4000, 4
    mu = (torch.randn(n, generator=g) * 2).cuda()
    sigma = torch.exp(torch.randn(n, generator=g) * 0.5).cuda()
    v = (mu + sigma * torch.randn(n, generator=g).cuda()).contiguous()
    v[10] += 5000.0                                                    # one escape inside the first range
    blob = _flat_encode(ops, v, mu, sigma, S).tobytes()
    kw = dict(mu=mu, mu_mode=2, mu_rs=n, sigma=sigma, sigma_mode=2, sigma_rs=n)
    out = torch.zeros(n, device="cuda")
    outb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    dec = ops.RansDecoder(blob, 1, n, 1, streams=S)
    # four streams of 1000 symbols, decoded in interleaved pieces
    pieces = [(0, 300), (1000, 1000), (300, 700), (2000, 1), (2001, 999), (3000, 1000)]
    for first, count in pieces:
        r = torch.tensor([[first, count]], dtype=torch.int32, device="cuda")
        dec.decode(r, 1, out, v_hat_rs=n, v_hat_bf16=outb, vb_rs=n, **kw)
    dec.finish()
    assert torch.equal(out, torch.round(v)) and torch.equal(outb.float(), torch.round(v).to(torch.bfloat16).float())
    dec = ops.RansDecoder(blob, 1, n, 1, streams=S)
    dec.decode(torch.tensor([[300, 100]], dtype=torch.int32, device="cuda"), 1, out, v_hat_rs=n, **kw)   # skips [0, 300)
    with pytest.raises(ldic.LdicError, match="out of order"):
        dec.finish()
    dec = ops.RansDecoder(blob, 1, n, 1, streams=S)
    dec.decode(torch.tensor([[900, 200]], dtype=torch.int32, device="cuda"), 1, out, v_hat_rs=n, **kw)   # crosses streams
    with pytest.raises(ldic.LdicError, match="out of order"):
        dec.finish()


def test_high_model_codec_round_trip(ldic):
    """The N=384 / M=32 model (model/net.py:446-451): compress -> decompress reproduces the encoder's reconstruction
    (wide kernels in the context model, 352 content channels per pixel), incl. the graph-replayed second call."""
    B, H, W = 2, 128, 128
    net = ldic.Net((B, H, W, 3), (B, H, W, 3), True, False).cuda().eval()
    net.load_state_dict(dw.make_state_dict(2, N=384, M=32, boost=True), strict=True)
    x = dw.make_input(2, B, H, W).cuda()
    streams, info = net.compress(x)
    ref = net.rd_forward(x, want_x_hat=True)["x_hat"]
    assert torch.equal(net.decompress(streams, H, W), ref)
    assert torch.equal(net.decompress(streams, H, W), ref)
    assert 0.5 * info["bpp_estimated"] < info["bpp_coded"] < 1.05 * info["bpp_estimated"] + 0.2

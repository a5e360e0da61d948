"""GPU parity of the window attention block against outputs of the UNMODIFIED reference class
layers/win_attention.py::WinBasedAttention (tests/golden/win_attention.npz, made by make_golden_winattn.py) and
against the CPU oracle (oracle/ref_path.py::win_based_attention) on fresh seeded inputs.

The block computes q, k, v, the softmax weights and the attention output in bf16 on the tensor cores (fp32
accumulation, fp32 softmax, fp32 residual), so the comparison is made on the attention term o = y - x with
  max |o - o_ref| <= 4e-2 * max |o_ref|   and   rms(o - o_ref) <= 1e-2 * rms(o_ref)
(measured bf16 operand noise is about 3e-3 rms); a wrong window / shift / mask / bias mapping gives O(1) errors."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_path as rp

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ldic():
    import ldic_b200
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(0), "device")
    return ldic_b200


def check_attention_term(y, y_ref, x):
    o, o_ref = (y - x).double(), (y_ref - x).double()
    err = (o - o_ref).abs().max().item()
    rms = (o - o_ref).pow(2).mean().sqrt().item()
    ref_max, ref_rms = o_ref.abs().max().item(), o_ref.pow(2).mean().sqrt().item()
    assert err <= 4e-2 * ref_max and rms <= 1e-2 * ref_rms, f"max err {err:.3e} (ref max {ref_max:.3e}), rms {rms:.3e} (ref rms {ref_rms:.3e})"


@pytest.mark.parametrize("tag", ["noshift", "shift", "small"])
def test_block_vs_reference_golden(ldic, tag):
    d = np.load(os.path.join(G, "win_attention.npz"))
    dim, heads, ws, shift, B, H, W = [int(v) for v in d[f"{tag}_cfg"]]
    sd = {k[len(tag) + 4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith(f"{tag}_sd_")}
    blk = ldic.WinBasedAttention(dim=dim, num_heads=heads, window_size=ws, shift_size=shift).cuda().eval()
    blk.load_state_dict(sd, strict=True)              # the reference's own keys
    x = torch.from_numpy(d[f"{tag}_x"])
    n0 = ldic.ops.launch_count()
    with torch.no_grad():
        y = blk(x.cuda()).cpu()
    assert ldic.ops.launch_count() - n0 >= 7          # transpose, q, k, v, core, proj, residual
    check_attention_term(y, torch.from_numpy(d[f"{tag}_y"]), x)


@pytest.mark.parametrize("dim,heads,ws,shift,B,H,W", [
    (192, 8, 8, 4, 2, 32, 48),        # the U-Net family's 1/4-resolution block shape class (head_dim 24)
    (192, 8, 4, 2, 1, 16, 8),         # window 4 (second block, model/net_unet_ha_hs.py:228)
    (128, 8, 8, 3, 1, 24, 16),        # head_dim 16, odd shift
    (192, 8, 8, 0, 3, 8, 8),          # one window per image
])
def test_block_vs_oracle(ldic, dim, heads, ws, shift, B, H, W):
    torch.manual_seed(dim + ws + shift + H)
    blk = ldic.WinBasedAttention(dim=dim, num_heads=heads, window_size=ws, shift_size=shift).eval()
    with torch.no_grad():
        blk.attn.relative_position_bias_table.normal_(0, 0.7)
        blk.attn.qkv.weight.mul_(2.5)
    sd = {k: v.clone() for k, v in blk.state_dict().items()}
    x = torch.randn(B, dim, H, W)
    with torch.no_grad():
        y_ref = rp.win_based_attention(sd, x, heads, ws, shift)
        y = blk.cuda()(x.cuda()).cpu()
    check_attention_term(y, y_ref, x)


def test_window_tokens_surface(ldic):
    """WindowAttention.forward(x (nW*B, N, C), mask=None) as the reference exposes it (layers/win_attention.py:85)."""
    torch.manual_seed(5)
    wa = ldic.WindowAttention(dim=192, window_size=(8, 8), num_heads=8).eval()
    with torch.no_grad():
        wa.relative_position_bias_table.normal_(0, 0.5)
    sd = {"attn." + k: v.clone() for k, v in wa.state_dict().items()}
    t = torch.randn(5, 64, 192)
    img = t.reshape(5, 8, 8, 192).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        ref = (rp.win_based_attention(sd, img, 8, 8, 0) - img).permute(0, 2, 3, 1).reshape(5, 64, 192)
        out = wa.cuda()(t.cuda()).cpu()
    zero = torch.zeros_like(out)
    check_attention_term(out, ref, zero)
    with pytest.raises(NotImplementedError):
        wa(t.cuda(), mask=torch.zeros(1, 64, 64).cuda())


def test_rejects_bad_shapes(ldic):
    blk = ldic.WinBasedAttention(dim=192, num_heads=8, window_size=8, shift_size=0).cuda().eval()
    with pytest.raises(ldic.LdicError):
        blk(torch.randn(1, 192, 12, 16).cuda())       # H not a multiple of the window
    with pytest.raises(ldic.LdicError):
        blk(torch.randn(1, 192, 16, 16))              # CPU tensor: no fallback


@pytest.mark.parametrize("dim,heads,ws,shift,H,W", [(192, 8, 8, 4, 32, 48), (192, 8, 4, 2, 16, 24), (128, 8, 4, 0, 8, 8), (64, 8, 4, 2, 12, 8)])
def test_bias_table_path_equals_gathered_bias(ldic, dim, heads, ws, shift, H, W):
    """ldic_window_attention_core_table (bias indexed from the (2ws-1)^2 x heads table in shared memory) is bit-identical
    to the kernel fed the gathered [heads, N, N] bias (layers/win_attention.py:101-104)."""
    g = torch.Generator().manual_seed(ws + shift + H)
    q, k, v = (torch.randn(2, H, W, dim, generator=g).to(torch.bfloat16).cuda() for _ in range(3))
    blk = ldic.WinBasedAttention(dim=dim, num_heads=heads, window_size=ws, shift_size=shift)
    table = (torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5).cuda()
    gathered = ldic.ops.window_attention_bias(table, blk.attn.relative_position_index.cuda(), heads, ws)
    o_tab = ldic.ops.window_attention_core(q, k, v, table, heads, ws, shift)
    o_gat = ldic.ops.window_attention_core(q, k, v, gathered, heads, ws, shift)
    assert torch.equal(o_tab, o_gat)

"""Build-container-only check (auto-skipped where /root/reference is absent, e.g. on the GPU box): the stock-torch
blocks of ldic_b200.net_unet against the same blocks of the UNMODIFIED reference file (model/net_unet_ha_hs.py run
through oracle/unet_harness.py, "restated deps"), module by module on CPU with identical weights.  The kernel-backed
pieces are covered on the GPU by tests/test_gpu_unet.py against the committed fixtures."""
import contextlib
import io
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, io, contextlib
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import torch
from oracle import unet_harness
import det_weights as dw
import ldic_b200
from ldic_b200 import net_unet
from ldic_b200.layers import WinBasedAttention
WinBasedAttention.forward = WinBasedAttention.forward_torch   # CPU: every window block on its torch path
m = unet_harness.load_unet_module("net_unet_ha_hs")
with contextlib.redirect_stdout(io.StringIO()):
    ref = m.Net((1, 256, 256, 3), (1, 256, 256, 3), False, False).eval()
ours = net_unet.Net((1, 256, 256, 3), (1, 256, 256, 3), False, False).eval()
fill = dw.unet_param_fill([(n, tuple(p.shape)) for n, p in ours.named_parameters()], 0)
ref.load_state_dict(fill, strict=False)
ours.load_state_dict({**ours.state_dict(), **fill}, strict=True)
g = torch.Generator().manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g)
def same(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= 1e-5 * max(1.0, b.abs().max().item()), (what, err)
    print("ok", what, err)
with torch.no_grad():
    y = R(1, 192, 16, 16)
    ro, oo = ref.h_a(y), ours.h_a(y)
    for i in range(4):
        same(oo[i], ro[i], f"h_a[{i}]")
    same(ours.h_s(None, *oo[1:]), ref.h_s(None, *ro[1:]), "h_s")
    for i in range(4):
        c = 192 + 48 * i
        t = R(1, c, 16, 16)
        a_ref = ref.atten_mean[i](t)
        same(ours.atten_mean[i](t), a_ref, f"atten_mean[{i}]")
        same(ours.atten_scale[i](t), ref.atten_scale[i](t), f"atten_scale[{i}]")
        same(ours.cc_mean_transforms[i](a_ref), ref.cc_mean_transforms[i](a_ref), f"cc_mean[{i}]")
        same(ours.cc_scale_transforms[i](a_ref), ref.cc_scale_transforms[i](a_ref), f"cc_scale[{i}]")
        t2 = R(1, 192 + 48 * min(i + 1, 5), 16, 16)
        same(ours.lrp_transforms[i](t2), ref.lrp_transforms[i](t2), f"lrp[{i}]")
    s = R(1, 16, 16, 16)
    same(ours.syntax_model(s), ref.syntax_model(s), "syntax_model")
    w = R(1, 192, 16, 16)
    same(ours.s_model.transform[0](w), ref.s_model.transform[0](w), "s_model Win_noShift_Attention(4,2)")
    w8 = R(1, 192, 16, 24)
    same(ours.a_model.transform[8](w8), ref.a_model.transform[8](w8), "a_model Win_noShift_Attention(8,4)")
    for i in (0, 9):
        t = R(1, 3 if i == 0 else 192, 20, 20)
        same(ours.a_model.transform[i](t), ref.a_model.transform[i](t), f"ResidualBottleneck a_model[{i}]")
    zs = R(1, 16, 1, 1)
    same(ours.conv_weights_gen(zs), ref.conv_weights_gen(zs), "conv_weights_gen")
print("ALL OK")
'''


@pytest.mark.skipif(not os.path.isfile("/root/reference/model/net_unet_ha_hs.py"), reason="reference tree not present")
def test_unet_torch_blocks_match_reference_modules():
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]

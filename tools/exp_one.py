"""Runs one heavy layer a few times (for ncu): python tools/exp_one.py {conv2|deconv3|deconv4|ctx1|first} [n]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
which = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
if which == "conv2":
    layer = ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g); x = torch.randn(16, 256, 384, C, device=dev).to(torch.bfloat16)
elif which == "deconv3":
    layer = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g); x = torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)
elif which == "deconv4":
    wt = torch.randn(C, 16, 5, 5, device=dev) * 0.02; bt = torch.zeros(16, device=dev)
    g16 = (torch.ones(16, device=dev), torch.eye(16, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
    layer = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5_MERGED, wt, bt, act=_lib.ACT_IGDN, out_f32=True, gdn=g16); x = torch.randn(16, 256, 384, C, device=dev).to(torch.bfloat16)
elif which == "ctx1":
    wc = torch.randn(C, 2 * C - 16, 3, 3, device=dev) * 0.02
    layer = ops.ConvTC(_lib.LDIC_CTX_CONV1, wc, b, act=_lib.ACT_LEAKY02, aux=(C, 16)); x = torch.randn(16, 32, 48, 2 * C, device=dev).to(torch.bfloat16)
else:
    w1 = torch.randn(C, 3, 5, 5, device=dev) * 0.1
    layer = ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, w1, b, act=_lib.ACT_GDN, gdn=g); x = torch.randn(16, 3, 512, 768, device=dev)
for _ in range(n): layer(x)
torch.cuda.synchronize()
print("ok")

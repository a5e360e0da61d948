"""One heavy layer in isolation for an ncu capture: python tools/prof_layer.py {conv2|deconv3|deconv2|first|tail} [n]."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "deconv3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C = 192
torch.manual_seed(0)
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
if which == "conv2":
    layer, x = ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g), bf(16, 256, 384, C)
elif which == "deconv3":
    layer, x = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 128, 192, C)
elif which == "deconv2":
    layer, x = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 64, 96, C)
elif which == "first":
    layer = ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, torch.randn(C, 3, 5, 5, device=dev) * 0.1, b, act=_lib.ACT_GDN, gdn=g)
    x = torch.randn(16, 3, 512, 768, device=dev)
elif which == "tail":
    wt = torch.randn(C, 16, 5, 5, device=dev) * 0.02; bt = torch.zeros(16, device=dev)
    g16 = (torch.ones(16, device=dev), torch.eye(16, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
    layer, x = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5_MERGED, wt, bt, act=_lib.ACT_IGDN, out_f32=True, gdn=g16), bf(16, 256, 384, C)
for _ in range(n):
    layer(x)
torch.cuda.synchronize()
print("done", which)

import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
gs3 = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g)
x3 = torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)
w1 = torch.randn(C, 3, 5, 5, device=dev) * 0.1
for name, layer, x in (("deconv3", gs3, x3),):
    for _ in range(2): layer(x)
    torch.cuda.synchronize()
    os.environ["LDIC_DEBUG_TIMING"] = "1"
    print(name, flush=True)
    layer(x); layer(x)
    torch.cuda.synchronize()
    del os.environ["LDIC_DEBUG_TIMING"]

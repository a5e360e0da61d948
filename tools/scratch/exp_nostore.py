import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
layer = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g)
x = torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)
first = ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, torch.randn(C, 3, 5, 5, device=dev) * 0.1, b, act=_lib.ACT_GDN, gdn=g)
img = torch.randn(16, 3, 512, 768, device=dev)
def timeit(f, n=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for ns in (0, 1):
    if ns: os.environ["LDIC_DEBUG_NOSTORE"] = "1"
    print("nostore", ns, "deconv3", f"{timeit(lambda: layer(x)):.4f}", "first", f"{timeit(lambda: first(img)):.4f}", flush=True)
    os.environ["LDIC_DEBUG_TIMING"] = "1"
    layer(x); torch.cuda.synchronize(); first(img); torch.cuda.synchronize()
    del os.environ["LDIC_DEBUG_TIMING"]

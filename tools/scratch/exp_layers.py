"""Isolated per-layer timing (back-to-back launches, CUDA events) of the heavy layers of the 768x512 batch-16 step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
def timeit(layer, x, n=20):
    for _ in range(3): layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): layer(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
res = {}
conv2 = ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g)
x2 = torch.randn(16, 256, 384, C, device=dev).to(torch.bfloat16)
res["conv2"] = timeit(conv2, x2)
gs3 = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g)
x3 = torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)
res["deconv3"] = timeit(gs3, x3)
wt = torch.randn(C, 16, 5, 5, device=dev) * 0.02; bt = torch.zeros(16, device=dev)
g16 = (torch.ones(16, device=dev), torch.eye(16, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
gs4 = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5_MERGED, wt, bt, act=_lib.ACT_IGDN, out_f32=True, gdn=g16)
res["deconv4"] = timeit(gs4, x2)
wc = torch.randn(C, 2 * C - 16, 3, 3, device=dev) * 0.02
ctx1 = ops.ConvTC(_lib.LDIC_CTX_CONV1, wc, b, act=_lib.ACT_LEAKY02, aux=(C, 16))
xc = torch.randn(16, 32, 48, 2 * C, device=dev).to(torch.bfloat16)
res["ctx1"] = timeit(ctx1, xc)
if hasattr(_lib, "LDIC_CONV_FIRST_5x5S2") and not os.environ.get("LDIC_LIB_PATH"):
    w1 = torch.randn(C, 3, 5, 5, device=dev) * 0.1
    l1 = ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, w1, b, act=_lib.ACT_GDN, gdn=g)
    x1 = torch.randn(16, 3, 512, 768, device=dev)
    res["first"] = timeit(l1, x1)
print(os.environ.get("LDIC_LIB_PATH", "new"), " ".join(f"{k} {v:.4f}" for k, v in res.items()), flush=True)

"""Wait-time breakdown (LDIC_DEBUG_TIMING) of the fused tail layer and the first layer at the bench shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
C = 192
torch.manual_seed(0)
b = torch.zeros(C, device=dev)
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
wt = torch.randn(C, 16, 5, 5, device=dev) * 0.02; bt = torch.zeros(16, device=dev)
g16 = (torch.ones(16, device=dev), torch.eye(16, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
tail = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5_MERGED, wt, bt, act=_lib.ACT_IGDN, out_f32=True, gdn=g16)
x = torch.randn(16, 256, 384, C, device=dev).to(torch.bfloat16)
img = torch.rand(16, 3, 512, 768, device=dev) * 2 - 1
cw = torch.randn(16, 3, 16, device=dev) * 0.1
first = ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, torch.randn(C, 3, 5, 5, device=dev) * 0.1, b, act=_lib.ACT_GDN, gdn=g)
def timeit(f, n=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("tail fused", f"{timeit(lambda: tail.fused_tail(x, img, cw)):.4f} ms", flush=True)
print("tail plain (out f32)", f"{timeit(lambda: tail(x)):.4f} ms", flush=True)
print("first", f"{timeit(lambda: first(img)):.4f} ms", flush=True)
os.environ["LDIC_DEBUG_TIMING"] = "1"
print("--- tail fused", file=sys.stderr, flush=True); tail.fused_tail(x, img, cw); torch.cuda.synchronize()
print("--- tail plain", file=sys.stderr, flush=True); tail(x); torch.cuda.synchronize()
print("--- first", file=sys.stderr, flush=True); first(img); torch.cuda.synchronize()

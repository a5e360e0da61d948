import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
def timeit(layer, x, n=5):
    for _ in range(2): layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): layer(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
w = torch.randn(192, 192, 5, 5, device=dev) * 0.02; b = torch.zeros(192, device=dev)
conv2 = ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w, b, act=_lib.ACT_RELU)
x2 = torch.randn(16, 256, 384, 192, device=dev).to(torch.bfloat16)
wt = torch.randn(192, 16, 5, 5, device=dev) * 0.02; bt = torch.zeros(16, device=dev)
gs4 = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5_MERGED, wt, bt, act=_lib.ACT_RELU, out_f32=True)
wd = torch.randn(192, 192, 5, 5, device=dev) * 0.02
gs3 = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, wd, b, act=_lib.ACT_RELU)
x3 = torch.randn(16, 128, 192, 192, device=dev).to(torch.bfloat16)
for st in [8, 6, 5, 4, 3, 2]:
    os.environ["LDIC_STAGES"] = str(st)
    print(f"stages<={st}: conv2 {timeit(conv2, x2):.3f} ms  gs4 {timeit(gs4, x2):.3f} ms  gs3 {timeit(gs3, x3):.3f} ms", flush=True)

"""Small-grid hyperprior layers in isolation: time per launch and the wait breakdown."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
C = 192
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev); w3 = torch.randn(C, C, 3, 3, device=dev) * 0.02
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
layers = [
 ("h_a conv3x3 s1 32x48", ops.ConvTC(_lib.LDIC_CONV_S1_3x3_P1, w3, b, act=_lib.ACT_RELU), bf(16, 32, 48, C)),
 ("h_a conv5x5 s2 32x48", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P2, w5, b, act=_lib.ACT_RELU), bf(16, 32, 48, C)),
 ("h_a conv5x5 s2 16x24", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P2, w5, b, act=_lib.ACT_NONE, out_f32=True), bf(16, 16, 24, C)),
 ("h_s deconv 8x12", ops.ConvTC(_lib.LDIC_DECONV_HS_5x5, w5, b, act=_lib.ACT_RELU), bf(16, 8, 12, C)),
 ("h_s deconv 16x24", ops.ConvTC(_lib.LDIC_DECONV_HS_5x5, w5, b, act=_lib.ACT_RELU), bf(16, 16, 24, C)),
 ("h_s deconv3x3 s1 32x48", ops.ConvTC(_lib.LDIC_DECONV_S1_3x3, w3, b, act=_lib.ACT_NONE, out_f32=True), bf(16, 32, 48, C)),
]
def timeit(layer, x, n=50):
    for _ in range(3): layer(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        layer(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(n): layer(x)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, layer, x in layers:
    print(name, f"{timeit(layer, x) * 1e3:.1f} us per launch (graph of 50 back-to-back launches)", flush=True)
os.environ["LDIC_DEBUG_TIMING"] = "1"
for name, layer, x in layers:
    print("---", name, file=sys.stderr, flush=True)
    layer(x); torch.cuda.synchronize()

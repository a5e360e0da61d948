import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
w3 = torch.randn(C, C, 3, 3, device=dev) * 0.02
layers = [
 ("conv2 (streaming s2, GDN)", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g), torch.randn(16, 256, 384, C, device=dev).to(torch.bfloat16)),
 ("deconv3 (halo, IGDN)", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)),
 ("deconv3 (halo, relu)", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_RELU), torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)),
 ("conv3x3 s1 (halo, relu) 128x192", ops.ConvTC(_lib.LDIC_CONV_S1_3x3_P1, w3, b, act=_lib.ACT_RELU), torch.randn(16, 128, 192, C, device=dev).to(torch.bfloat16)),
]
for name, layer, x in layers:
    for _ in range(2): layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): layer(x)
    e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1)/5:.4f} ms", f"{layer.flops(*x.shape[:3])/ (e0.elapsed_time(e1)/5*1e-3)/1e12:.0f} TF/s", flush=True)
    os.environ["LDIC_DEBUG_TIMING"] = "1"
    layer(x)
    torch.cuda.synchronize()
    del os.environ["LDIC_DEBUG_TIMING"]
w1 = torch.randn(C, 3, 5, 5, device=dev) * 0.1
l1 = ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, w1, b, act=_lib.ACT_GDN, gdn=g)
x1 = torch.randn(16, 3, 512, 768, device=dev)
for _ in range(2): l1(x1)
torch.cuda.synchronize()
os.environ["LDIC_DEBUG_TIMING"] = "1"
l1(x1); torch.cuda.synchronize()
os.environ["LDIC_DEBUG_NOSTORE"] = "1"
l1(x1); torch.cuda.synchronize()

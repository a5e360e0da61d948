import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["LDIC_HALO"] = "0"
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
w = torch.randn(192, 192, 5, 5, device=dev) * 0.02; b = torch.zeros(192, device=dev)
conv2 = ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w, b, act=_lib.ACT_RELU)
x2 = torch.randn(16, 256, 384, 192, device=dev).to(torch.bfloat16)
wt = torch.randn(192, 16, 5, 5, device=dev) * 0.02; bt = torch.zeros(16, device=dev)
gs4 = ops.ConvTC(_lib.LDIC_DECONV_GS_5x5_MERGED, wt, bt, act=_lib.ACT_RELU, out_f32=True)
for name, layer in (("conv2", conv2), ("gs4", gs4)):
    for _ in range(2): layer(x2)
    torch.cuda.synchronize()
    os.environ["LDIC_DEBUG_TIMING"] = "1"
    print(name, flush=True)
    layer(x2); layer(x2)
    torch.cuda.synchronize()
    del os.environ["LDIC_DEBUG_TIMING"]

import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ldic_b200
from ldic_b200 import ops, _lib
L = _lib
B, H, W, Cc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
act = sys.argv[5] if len(sys.argv) > 5 else "none"
torch.manual_seed(0)
x = torch.randn(B, 3, H, W); w = torch.randn(Cc, 3, 5, 5) * 0.2; b = torch.randn(Cc) * 0.1
kw = {}
if act == "gdn":
    import det_weights as dw
    from oracle import ref_path as rp
    sd = {}; dw._gdn(sd, 43, "g", Cc)
    kw = dict(act=L.ACT_GDN, gdn=(sd["g.beta"].cuda(), sd["g.gamma"].cuda()) + rp.model_gdn_constants())
layer = ops.ConvTC(L.LDIC_CONV_FIRST_5x5S2, w.cuda(), b.cuda(), out_f32=True, **kw)
try:
    y = layer(x.cuda()); torch.cuda.synchronize()
    bf = lambda t: t.to(torch.bfloat16).float()
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(bf(x), (1, 2, 1, 2)), bf(w), b, stride=2)
    if act == "none":
        print("max err", (y.cpu().permute(0, 3, 1, 2) - ref).abs().max().item(), "ref max", ref.abs().max().item())
    else:
        print("ran; out abs mean", y.abs().mean().item())
except Exception as e:
    print("FAILED:", str(e).splitlines()[0])
out = (C.c_ulonglong * 5)()
print("timeout recorded:", _lib.load().ldic_debug_last_timeout(out), list(out))

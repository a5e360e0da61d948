"""Per-layer time of the GDN / IGDN layers against LDIC_GDN_INSERT (conv stages of tile t+1 issued before the
gamma stages of tile t), plus the wait-time breakdown (LDIC_DEBUG_TIMING) at the default and the best setting."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0)
C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
layers = [
 ("conv2", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g), bf(16, 256, 384, C)),
 ("conv3", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g), bf(16, 128, 192, C)),
 ("deconv1", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 32, 48, C)),
 ("deconv2", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 64, 96, C)),
 ("deconv3", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 128, 192, C)),
]
def timeit(layer, x, n=10):
    for _ in range(2): layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): layer(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
vals = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2,4,6,8,10,12,16").split(",")]
for name, layer, x in layers:
    row = []
    for v in vals:
        os.environ["LDIC_GDN_INSERT"] = str(v)
        row.append(f"{v}:{timeit(layer, x):.4f}")
    print(name, " ".join(row), flush=True)
wc = torch.randn(C, 2 * C - 16, 3, 3, device=dev) * 0.02
ctx1 = ops.ConvTC(_lib.LDIC_CTX_CONV1, wc, b, act=_lib.ACT_LEAKY02, aux=(C, 16))
print("ctx1", f"{timeit(ctx1, bf(16, 32, 48, 2 * C)):.4f}", flush=True)
for v in (4, 8):
    os.environ["LDIC_GDN_INSERT"] = str(v)
    os.environ["LDIC_DEBUG_TIMING"] = "1"
    for name, layer, x in (layers[0], layers[-1]):
        print(f"--- {name} insert {v}", file=sys.stderr, flush=True)
        layer(x); torch.cuda.synchronize()
    del os.environ["LDIC_DEBUG_TIMING"]

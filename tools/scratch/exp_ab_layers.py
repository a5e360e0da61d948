"""A/B of two builds of libldic_b200 on the same box: runs tools/exp_nostore.py-like timings in subprocesses alternating
LDIC_LIB_PATH (usage: exp_ab_layers.py <alt .so>)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
alt = os.path.abspath(sys.argv[1])
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
import ldic_b200
from ldic_b200 import ops, _lib
dev = torch.device("cuda", 0); C = 192
g = (torch.ones(C, device=dev), torch.eye(C, device=dev) * 0.1 + 0.001, 1e-3, 2.0 ** -18, 2.0 ** -36)
w5 = torch.randn(C, C, 5, 5, device=dev) * 0.02; b = torch.zeros(C, device=dev)
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
L = [("conv2", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g), bf(16, 256, 384, C)),
     ("conv3", ops.ConvTC(_lib.LDIC_CONV_S2_5x5_P12, w5, b, act=_lib.ACT_GDN, gdn=g), bf(16, 128, 192, C)),
     ("deconv2", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 64, 96, C)),
     ("deconv3", ops.ConvTC(_lib.LDIC_DECONV_GS_5x5, w5, b, act=_lib.ACT_IGDN, gdn=g), bf(16, 128, 192, C)),
     ("first", ops.ConvTC(_lib.LDIC_CONV_FIRST_5x5S2, torch.randn(C, 3, 5, 5, device=dev) * 0.1, b, act=_lib.ACT_GDN, gdn=g), torch.randn(16, 3, 512, 768, device=dev))]
out = []
for name, layer, x in L:
    for _ in range(3): layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): layer(x)
    e1.record(); torch.cuda.synchronize()
    out.append(f"{name} {e0.elapsed_time(e1) / 20:.4f}")
print(" ".join(out))
''' % ROOT
for rep in range(3):
    for tag, env in (("new", {}), ("alt", {"LDIC_LIB_PATH": alt})):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
        print(tag, r.stdout.strip() or r.stderr[-300:], flush=True)

"""Runs the C5-size trit-plane kernel a few times (for ncu / timing): 16 x 192 x 128 x 128 latents, 4 planes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ldic_b200 import ops
n = 16 * 192 * 128 * 128
dev = torch.device("cuda", 0)
v = torch.randn(n, device=dev) * 4; mu = torch.randn(n, device=dev); sg = torch.exp(torch.randn(n, device=dev)).clamp_(0.05, 20)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    e0.record()
    ops.tritplane_likelihood(v, sg, mu, planes=4)
    e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1):.4f} ms  {20.0 * n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")

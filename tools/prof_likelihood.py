"""Runs the C5-size likelihood kernel a few times (for ncu / timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ldic_b200 import ops
n = 16 * 192 * 128 * 128
dev = torch.device("cuda", 0)
v = torch.randn(n, device=dev) * 4; mu = torch.randn(n, device=dev); sg = torch.exp(torch.randn(n, device=dev)).clamp_(0.05, 20)
vh = torch.empty_like(v); lk = torch.empty_like(v)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    e0.record()
    ops.likelihood_rows(v, 1, n, v_rs=n, mu=mu, mu_mode=2, mu_rs=n, sigma=sg, sigma_mode=2, sigma_rs=n, quant=ops.QUANT_ROUND, v_hat=vh, v_hat_rs=n, lik=lk)
    e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1):.4f} ms  {20.0 * n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")

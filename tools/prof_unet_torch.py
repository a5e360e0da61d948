"""Kernel-time breakdown of one U-Net-family step with torch.profiler (CUPTI): python tools/prof_unet_torch.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ldic_b200
from ldic_b200 import net_unet
import det_weights as dw
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 512, 768
dev = torch.device("cuda", 0)
net = net_unet.Net((B, H, W, 3), (B, H, W, 3), False, False).to(dev).eval()
fill = dw.unet_param_fill([(n, tuple(p.shape)) for n, p in net.named_parameters()], 0)
net.load_state_dict({**net.state_dict(), **{k: v.to(dev) for k, v in fill.items()}}, strict=True)
x = dw.make_input(0, 2, H, W).repeat((B + 1) // 2, 1, 1, 1)[:B].to(dev)
for _ in range(2):
    net.rd_forward(x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    net.rd_forward(x)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 or getattr(e, "cuda_time_total", 0) > 0]
rows = sorted(((getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0), e.count, e.key) for e in prof.key_averages()), reverse=True)
tot = sum(r[0] for r in rows)
print("total self device time %.2f ms" % (tot / 1e3))
for t, c, k in rows[:45]:
    if t > 0:
        print("%9.3f ms %5.1f%% x%-4d %s" % (t / 1e3, 100 * t / tot, c, k[:110]))

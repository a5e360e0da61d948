"""LDIC_DEBUG_TIMING on the first analysis layer at the bench shape (prints the per-role wait / busy cycles)."""
import os, sys
os.environ["LDIC_DEBUG_TIMING"] = "1"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import ldic_b200
import det_weights as dw
import bench
torch.cuda.set_device(0)
B, H, W = 16, 512, 768
net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
net.load_state_dict(dw.make_state_dict(0), strict=True)
x = bench.make_u8_batches(0, B, 1)[0].cuda()
net.auto_graph = False
L = net.a_model.plan()
for _ in range(2):
    y = L[0](x)
torch.cuda.synchronize()
print("first layer out", tuple(y.shape), y.dtype)
xf = (x.float() / 255.0) * 2 - 1
for _ in range(2):
    y = L[0](xf)
torch.cuda.synchronize()

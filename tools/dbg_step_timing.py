"""LDIC_DEBUG_TIMING over one eager step at the bench shape: per-role wait cycles of every conv launch."""
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
os.environ["X"] = "1"
for i in range(2):
    if i == 1:
        print("=== second (warm) step ===", file=sys.stderr, flush=True)
    net.rd_forward(x)
    torch.cuda.synchronize()

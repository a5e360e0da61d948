"""First analysis layer: global stores per lane vs output tile through shared memory + TMA stores (first_tma_store)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import ldic_b200
from ldic_b200 import ops
import det_weights as dw
import bench
torch.cuda.set_device(0)
B, H, W = 16, 512, 768
net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
net.load_state_dict(dw.make_state_dict(0), strict=True)
xu = bench.make_u8_batches(0, B, 1)[0].cuda()
xf = (xu.float() / 255.0) * 2 - 1
L = net.a_model.plan()
def t(x, n=20):
    for _ in range(3): L[0](x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): L[0](x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = {}
for mode in (0, 1, 0, 1):
    ops.set_tuning("first_tma_store", mode)
    yu, yf = L[0](xu), L[0](xf)
    torch.cuda.synchronize()
    if not ref:
        ref = {"u": yu.clone(), "f": yf.clone()}
    same = torch.equal(yu, ref["u"]) and torch.equal(yf, ref["f"])
    print("first_tma_store", mode, "u8 %.4f ms  f32 %.4f ms  identical: %s" % (t(xu), t(xf), same), flush=True)
# small / ragged shapes
for (b, h, w) in [(1, 64, 64), (2, 64, 192), (3, 192, 64), (1, 256, 256)]:
    x = torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, device="cuda")
    ops.set_tuning("first_tma_store", 0); y0 = L[0](x).clone()
    ops.set_tuning("first_tma_store", 1); y1 = L[0](x).clone()
    torch.cuda.synchronize()
    print((b, h, w), "identical:", torch.equal(y0, y1), flush=True)

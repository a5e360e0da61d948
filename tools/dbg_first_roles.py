"""First analysis layer: time per launch against the MMA issue order (first_insert) and with the stores / the patch
assembly switched off (which role bounds the kernel?)."""
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
for ins in (2, 1, 0, 2, 0):
    ops.set_tuning("first_insert", ins)
    yu, yf = L[0](xu), L[0](xf)
    if not ref:
        ref = {"u": yu.clone(), "f": yf.clone()}
    same = torch.equal(yu, ref["u"]) and torch.equal(yf, ref["f"])
    print("first_insert", ins, "u8 %.4f ms  f32 %.4f ms  identical to insert=2: %s" % (t(xu), t(xf), same), flush=True)
for mode in (1, 2, 3):
    ops.set_tuning("debug_nostore", mode)
    print("first_insert 0, debug mode", mode, "(bit0 no stores, bit1 no patch build): u8 %.4f ms  f32 %.4f ms" % (t(xu), t(xf)), flush=True)
ops.set_tuning("debug_nostore", 0)

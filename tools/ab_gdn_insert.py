"""Per-layer time of the GDN / IGDN conv layers against gdn_insert (conv stages of tile t+1 issued before the GDN
contraction of tile t)."""
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
x = bench.make_u8_batches(0, B, 1)[0].cuda()
out = net.rd_forward(x)
A, S = net.a_model.plan(), net.s_model.plan()
a1 = A[0](x); a2 = A[1](a1); a3 = A[2](a2)
yb = torch.zeros(B, H // 16, W // 16, net.N, dtype=torch.bfloat16, device="cuda")
yb[..., net.M:] = torch.round(out["latents"]["y"][..., net.M:]).to(torch.bfloat16)
s1 = S[0](yb); s2 = S[1](s1); s3 = S[2](s2)
cw = out["latents"]["conv_w"].reshape(B, 3, net.M)
cases = {"conv2": lambda: A[1](a1), "conv3": lambda: A[2](a2), "deconv1": lambda: S[0](yb), "deconv2": lambda: S[1](s1),
         "deconv3": lambda: S[2](s2), "tail": lambda: S[3].fused_tail(s3, x, cw)}
def t(fn, n=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for v in (4, 1, 2, 6, 8, 10, 12, 16, 4):
    ops.set_tuning("gdn_insert", v)
    r = {k: round(t(f) * 1e3) for k, f in cases.items()}
    sq = S[3].fused_tail(s3, x, cw)[0]
    if ref is None: ref = sq.clone()
    print("gdn_insert", v, r, "sum", sum(r.values()), "tail identical", bool(torch.equal(sq, ref)), flush=True)

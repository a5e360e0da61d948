import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ldic_b200, det_weights as dw
B, H, W = 16, 512, 768
net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), False, False).cuda().eval()
net.load_state_dict(dw.make_state_dict(0), strict=True)
y = torch.randn(B, 32, 48, 192, device="cuda"); h2 = torch.randn(B, 32, 48, 192, device="cuda")
def run_k():
    return ldic_b200.ops.syntax_branch(y, h2, 16, net.syntax_model, net.prediction_model_syntax, net.conv_weights_gen)
def run_t():
    with torch.no_grad():
        z = net.syntax_model(y.permute(0, 3, 1, 2)[:, :16]); zr = torch.round(z)
        a, b = net.prediction_model_syntax(zr, h2.permute(0, 3, 1, 2)); c = net.conv_weights_gen(zr)
    return z, zr, a, b, c
for name, fn in (("kernels", run_k), ("torch", run_t)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")

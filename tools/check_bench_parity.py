"""Per-image bpp / PSNR of the bench's own images: bf16 path, TF32 parity mode and the fp32 CPU oracle."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import math
import torch
import ldic_b200
import det_weights as dw
import bench
from oracle import ref_path as rp
torch.cuda.set_device(0)
B, H, W = 4, 512, 768
sd = dw.make_state_dict(0)
net = ldic_b200.Net((1, H, W, 3), (1, H, W, 3), False, False).cuda().eval()
net.load_state_dict(sd, strict=True)
net.auto_graph = False
xu = bench.make_u8_batches(0, B, 1)[0]
xf = (xu.float() / 255.0) * 2.0 - 1.0
res = {}
for mode in ("bf16", "tf32"):
    net.parity_tf32 = mode == "tf32"
    rows = []
    for b in range(B):
        bpp, v_mse, v_psnr = net(xu[b:b + 1].cuda(), "test", 1)
        rows.append((bpp.item(), v_psnr.item()))
    res[mode] = rows
torch.set_num_threads(16)
rows = []
with torch.no_grad():
    for b in range(B):
        r = rp.net_forward_test(sd, xf[b:b + 1], (1, H, W, 3), return_intermediates=False)
        rows.append((r["bpp"].item(), r["v_psnr"].item()))
res["oracle"] = rows
for b in range(B):
    o = res["oracle"][b]
    print("image", b, "oracle bpp %.5f psnr %.4f |" % o,
          " ".join("%s: bpp %+.3f%% psnr %+.4f dB" % (m, 100 * (res[m][b][0] / o[0] - 1), res[m][b][1] - o[1]) for m in ("bf16", "tf32")), flush=True)

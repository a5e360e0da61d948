"""Times the rANS encoder / decoder at the bench size (16 x 768x512: 16 x 32x48x176 content symbols) with CUDA events."""
import json
import sys

import torch

sys.path.insert(0, ".")
import ldic_b200
from ldic_b200 import ops

torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(1)
B, h, w, N, M, Cp = 16, 32, 48, 192, 16, 192
Cc, P = N - M, B * h * w
ctx = torch.empty(P, 2 * Cp, device="cuda")
ctx[:, :Cp] = torch.randn(P, Cp, device="cuda", generator=g) * 3
ctx[:, Cp:] = torch.randn(P, Cp, device="cuda", generator=g) * 0.8 + 0.2
y = torch.zeros(B, h, w, N, device="cuda")
y[..., M:] = ctx[:, :Cc].reshape(B, h, w, Cc) + torch.exp(ctx[:, Cp:Cp + Cc]).reshape(B, h, w, Cc) * torch.randn(
    B, h, w, Cc, device="cuda", generator=g)
kw = dict(mu=ctx, mu_mode=2, mu_rs=2 * Cp, sigma=ctx, sigma_mode=2, sigma_rs=2 * Cp, sigma_off=Cp, sigma_is_log=True)
res = {}
for sps in (256, 512, 1024, 2048, 4096):
    S = ops.rans_streams_for(h * w * Cc, sps)
    enc = ops.rans_encode_rows(y, P, Cc, h * w, v_rs=N, v_off=M, streams=S, **kw)
    out = torch.empty(B, h, w, Cc, device="cuda")
    ops.rans_decode_rows(enc, P, Cc, h * w, out, v_hat_rs=Cc, **kw)
    assert torch.equal(out, torch.round(y[..., M:]))
    t = []
    for fn in (lambda: ops.rans_encode_rows(y, P, Cc, h * w, v_rs=N, v_off=M, streams=S, **kw),
               lambda: ops.rans_decode_rows(enc, P, Cc, h * w, out, v_hat_rs=Cc, check_status=False, **kw)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) / 10)
    res[sps] = {"streams": S, "encode_ms": round(t[0], 4), "decode_ms": round(t[1], 4), "bytes": sum(enc.nbytes()),
                "symbols": B * h * w * Cc}
    print(sps, res[sps], flush=True)
json.dump(res, open("gpurun_out/prof_rans.json", "w"))

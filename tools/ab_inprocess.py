"""Same-process, interleaved A/B of Net configurations (same box, same thermal state): alternates timed blocks of
graph replays and prints the per-configuration median ms/step.  Usage: python tools/ab_inprocess.py side_sms 0 12 16"""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ldic_b200
import det_weights as dw
from ldic_b200.graph import GraphedEvaluator
from bench import make_u8_batches

attr, vals = sys.argv[1], [int(v) for v in sys.argv[2:]]
B, H, W = 16, 512, 768
dev = torch.device("cuda", 0)
x = [b.to(dev) for b in make_u8_batches(0, B, 2)]
evs = []
for v in vals:
    net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), False, False).to(dev).eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    if attr == "first_epi":
        ldic_b200.ops.set_tuning("first_epi", v)
    else:
        setattr(net, attr, v)
    evs.append(GraphedEvaluator(net, x))
res = {v: [] for v in vals}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rnd in range(8):
    for v, ev in zip(vals, evs):
        for i in range(5):
            ev(i % 2)
        torch.cuda.synchronize()
        e0.record()
        for i in range(40):
            ev(i % 2)
        e1.record()
        torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / 40)
ref = None
for v in vals:
    print(attr, v, "median ms/step %.4f" % statistics.median(res[v]), "min %.4f" % min(res[v]), ["%.3f" % t for t in res[v]])
out0 = evs[0](0)[2]; outs = [ev(0)[2] for ev in evs]
print("bit-identical results:", all(torch.equal(o["bits"], outs[0]["bits"]) and torch.equal(o["sq_err"], outs[0]["sq_err"]) for o in outs))

"""Host-side timeline of the codec serving loop (bench.py --config codec, e2e leg): where do the milliseconds go?"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import ldic_b200
from ldic_b200 import ops
import det_weights as dw
import bench

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
B, H, W, NBUF = 16, 512, 768, 4
net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), False, False).to(dev).eval()
net.load_state_dict(dw.make_state_dict(0), strict=True)
host = [t.pin_memory() for t in bench.make_u8_batches(0, B, NBUF)]
xs = [t.to(dev) for t in host]
ev = ldic_b200.GraphedEvaluator(net, xs, entropy_code=True)
for k in range(NBUF):
    ev(k)
torch.cuda.synchronize()
T = {"enqueue": 0.0, "meta": 0.0, "bytes": 0.0, "slice": 0.0}
out_stream = torch.cuda.Stream()
done = [torch.cuda.Event() for _ in range(NBUF)]
steps = 30
t_all = time.perf_counter()
pending = None
for i in range(steps + 1):
    t0 = time.perf_counter()
    if i < steps:
        k = i % NBUF
        _, _, out = ev(k)
        done[k].record()
    t1 = time.perf_counter(); T["enqueue"] += t1 - t0
    if pending is not None:
        enc, kk = pending
        out_stream.wait_event(done[kk])
        with torch.cuda.stream(out_stream):
            streams = list(enc.values())
            meta = torch.cat([t.reshape(-1) for e in streams for t in (e.sizes, e.status)]).cpu().tolist()
            t2 = time.perf_counter(); T["meta"] += t2 - t1
            hosts, sizes, pos = [], [], 0
            for slot, e in enumerate(streams):
                kq = e.sizes.numel(); sz = meta[pos:pos + kq]; pos += 2 * kq; sizes.append(sz)
                m = max(sz)
                h = ops._pinned_bytes(slot, len(sz), m)
                h.copy_(e.buf[:, :m], non_blocking=True)
                hosts.append(h)
            torch.cuda.current_stream().synchronize()
            t3 = time.perf_counter(); T["bytes"] += t3 - t2
            blobs = [[h[j, :n].numpy().tobytes() for j, n in enumerate(sz)] for h, sz in zip(hosts, sizes)]
            T["slice"] += time.perf_counter() - t3
    pending = (out["streams"], k) if i < steps else None
torch.cuda.synchronize()
tot = time.perf_counter() - t_all
print("per step ms:", {k: round(1e3 * v / steps, 3) for k, v in T.items()}, "total", round(1e3 * tot / steps, 3))

#!/usr/bin/env python
"""Headline benchmark: 768x512 images/s through the rate-distortion forward
(g_a -> hyperprior -> round -> likelihood -> bpp -> g_s -> PSNR), BASELINE.json configs[1]
(model/net.py forward on batch 16 synthetic Kodak-size images per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1 is launched by torchrun (one rank per GPU, NCCL); every rank runs its own batch of 16
(weak scaling) and the ranks exchange ONE all-reduce of five scalars per step.
`--impl reference` times the CPU port of the reference's own path (oracle/ref_path.py; the
reference is pure Python and cannot travel to the GPU box) on all host cores, one image per step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "768x512 imgs/sec (g_a->hyperprior->bpp->g_s)"
UNIT = "images/s"
H, W = 512, 768


def measured_traffic():
    """dram__bytes_read+write of the roofline kernels from the committed ncu --set full capture (profiles/)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            return float(t["conv_kernels"]["dram_bytes_per_step"]), float(t["likelihood_c5"]["dram_bytes_per_launch"])
        except Exception:
            continue
    return None, None


def measured_traffic_of(key):
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return float(json.load(f)[key]["dram_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_u8_batches(rank: int, B: int, nbuf: int):
    """Synthetic 8-bit imagery (what the reference's eval driver reads from PNG files): `nbuf` distinct batches of B
    images, uint8 levels (B,3,H,W)."""
    import torch
    import det_weights as dw
    base = dw.make_input(rank, 4, H, W)
    out = []
    for i in range(nbuf):     # distinct batches: images rolled so no two buffers are equal
        xb = torch.cat([torch.roll(base, shifts=(i * 37 + j * 11), dims=3) for j in range((B + 3) // 4)], 0)[:B]
        out.append(torch.round((xb + 1.0) * 127.5).clamp_(0, 255).to(torch.uint8).contiguous())
    return out


def cpu_port_images_per_s(n_steps: int, n_warm: int, batch: int = 1, high: bool = False):
    """The reference's CPU path (torch-CPU port, one-hot sampler convs as written), all host cores."""
    import torch
    import det_weights as dw
    from oracle import ref_path as rp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N, M = (384, 32) if high else (192, 16)
    sd = dw.make_state_dict(0, N=N, M=M)
    x = (make_u8_batches(0, batch, 1)[0].float() / 255.0) * 2.0 - 1.0          # ToTensor + eval_net.py:84
    flt_y = rp.block_sample_filter(N - M, True)
    flt_h = rp.block_sample_filter(N, False)
    orig = rp.block_sample_onehot

    def cached(xx, masked, flt=None):
        return orig(xx, masked, flt_y if masked else flt_h)
    rp.block_sample_onehot = cached
    try:
        with torch.no_grad():
            for _ in range(n_warm):
                rp.net_forward_test(sd, x, (batch, H, W, 3), M=M, faithful_sampler=True, return_intermediates=False)
            t0 = time.perf_counter()
            for _ in range(n_steps):
                rp.net_forward_test(sd, x, (batch, H, W, 3), M=M, faithful_sampler=True, return_intermediates=False)
            dt = time.perf_counter() - t0
    finally:
        rp.block_sample_onehot = orig
    return batch * n_steps / dt, dt / n_steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    high = args.config == "high"
    ips, spi, cores = cpu_port_images_per_s(args.steps, args.warmup, batch=1, high=high)
    sample = f"1 image 768x512 per step, {args.steps} steps after {args.warmup} warm-up, torch CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": spi * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"model/net.py Net.forward(test) {'N=384 M=32' if high else 'N=192 M=16'}, 768x512, CPU port of the reference path "
                                   "(oracle/ref_path.py, one-hot BlockSample convs as written), 1 image per step"},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ldic_b200
    from ldic_b200 import ops
    from ldic_b200.dist import ShardedEvaluator
    import det_weights as dw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(local), "device")
    if world > 1:
        # keep stdout to the one JSON line (some boxes export NCCL_DEBUG=VERSION, which prints to stdout)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    hbm_peak, tc_peak_sus, tc_peak_burst, peak_src = peaks()

    high = args.config == "high"
    Nw, Mw = (384, 32) if high else (192, 16)
    net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), high, False).to(dev).eval()
    net.load_state_dict(dw.make_state_dict(0, N=Nw, M=Mw), strict=True)
    if args.side_sms >= 0:
        net.side_sms = args.side_sms
    net.auto_graph = not args.no_graph
    if args.parity_tf32:
        net.parity_tf32 = True          # g_a / h_a with kind::tf32 MMAs (SURVEY H4's precision comparison, ~half rate there)
    ev = ShardedEvaluator(net)
    NBUF = 4
    # 8-bit imagery: the host buffers hold uint8 levels; the first layer applies x = (u/255)*2-1 (eval_net.py:84) itself
    host = [h.pin_memory() for h in make_u8_batches(rank, B, NBUF)]
    if args.input == "f32":
        host = [((h.float() / 255.0) * 2.0 - 1.0).contiguous().pin_memory() for h in host]
    devbuf = [h.to(dev) for h in host]
    in_bytes = host[0].numel() * host[0].element_size()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    use_graph = not args.no_graph
    if use_graph:           # the whole launch sequence of a step captured once per static input buffer, replayed per step
        from ldic_b200.graph import GraphedEvaluator
        gev = GraphedEvaluator(net, devbuf)
        step = lambda i: gev(i % NBUF)
    else:
        step = lambda i: ev(devbuf[i % NBUF])
    for i in range(max(args.warmup, 3)):
        bpp, psnr, _ = step(i)
    sync_all()

    # ---------------- device-resident throughput (`value`) ----------------
    sampler = ClockSampler(local)
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        bpp, psnr, _ = step(i)
    e1.record()
    sync_all()
    n1 = ops.launch_count()
    launches = int(n1 - n0) + (args.steps * gev.launches_per_replay if use_graph else 0)
    t_ms = e0.elapsed_time(e1)
    tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    value = world * B * args.steps / (t_ms * 1e-3)

    # per-launch CUDA events around every conv launch: the same K steps issued eagerly (events cannot be read back
    # from inside a graph replay); the kernels and their durations are the ones of the timed region
    # (single stream: with the SM partition the per-launch times would include the concurrent kernels)
    side_was, net.side_sms = net.side_sms, 0
    ops.PROFILE = []
    sync_all()
    e0.record()
    for i in range(args.steps):
        ev(devbuf[i % NBUF])
    e1.record()
    sync_all()
    prof, ops.PROFILE = ops.PROFILE, None
    net.side_sms = side_was
    profiled_eager_ms = e0.elapsed_time(e1)
    # the nn.Module surface as a user calls it: net(x, 'test') (transparent graph replay unless --no-graph)
    for i in range(3):
        net(devbuf[i % NBUF], "test", 1)
    sync_all()
    e0.record()
    for i in range(args.steps):
        r_mod = net(devbuf[i % NBUF], "test", 1)
    e1.record()
    sync_all()
    eager_ms = e0.elapsed_time(e1)

    # conv kernel roofline from the events recorded around every conv_tc launch of the timed region
    # ---------------- multi-GPU correctness on hardware: the all-reduced (bpp, PSNR) of one step equals
    # the single-GPU result on the same global batch, recomputed serially by rank 0 outside the timed region
    multi_gpu_equal = None
    if world > 1:
        kb = 1 % NBUF
        bpp_d, psnr_d, _ = step(kb)
        dist_res = (float(bpp_d.item()), float(psnr_d.item()))
        sync_all()
        if rank == 0:
            packed = torch.zeros(5, dtype=torch.float64, device=dev)
            for r in range(world):
                xr = make_u8_batches(r, B, NBUF)[kb]
                if args.input == "f32":
                    xr = ((xr.float() / 255.0) * 2.0 - 1.0).contiguous()
                o = net.rd_forward(xr.to(dev))
                packed += ops.rd_pack_metrics(o["bits"], o["sq_err"], 3 * H * W, want_v_mse=False)[0]
            single = ops.rd_finish_metrics(packed, float(H * W)).cpu()
            err = (abs(dist_res[0] / float(single[0]) - 1.0), abs(dist_res[1] - float(single[1])))
            multi_gpu_equal = {"equal": bool(err[0] < 1e-6 and err[1] < 1e-5), "bpp_rel_err": err[0], "psnr_abs_err_db": err[1],
                               "bpp_all_reduced": dist_res[0], "bpp_single_gpu": float(single[0]),
                               "psnr_all_reduced": dist_res[1], "psnr_single_gpu": float(single[1]),
                               "global_batch": world * B}
            if not multi_gpu_equal["equal"]:
                raise SystemExit(f"multi-GPU result differs from the single-GPU result on the same global batch: {multi_gpu_equal}")
        sync_all()

    conv_ms = sum(a.elapsed_time(b) for (_, _, a, b) in prof)
    conv_flops = sum(layer.flops(*shp) for (layer, shp, _, _) in prof)
    per_layer = {}
    for (layer, shp, a, b) in prof:
        k = f"kind{layer.kind}_{shp[1]}x{shp[2]}_{layer.cin}->{layer.cout}"
        d = per_layer.setdefault(k, [0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += layer.flops(*shp); d[2] += 1
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    # ---------------- end to end through Net.forward with HOST buffers ----------------
    # Every step: H2D copy of that step's batch from pinned host memory (copy stream, double buffered so it
    # overlaps the previous step's kernels), the forward, and a D2H read of the step's result (bpp, PSNR,
    # per-image MSE) into pinned memory, consumed on the host one step later.
    sync_all()
    d2h_bytes = 4 + 4 + 4 * B
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(device=dev)
    res_host = [torch.empty(2 + B, dtype=torch.float32).pin_memory() for _ in range(2)]
    res_ev = [torch.cuda.Event() for _ in range(2)]
    h2d_ev = [torch.cuda.Event() for _ in range(2)]
    used_ev = [torch.cuda.Event() for _ in range(2)]         # forward of the step that read xin[k] has been enqueued
    xin = [torch.empty_like(devbuf[0]) for _ in range(2)]    # preallocated device input buffers (no allocator traffic)
    results = []
    if use_graph:
        gev2 = GraphedEvaluator(net, xin)                    # one graph per H2D landing buffer
        sync_all()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(used_ev[i % 2])        # the kernels that read this buffer two steps ago are done
            xin[i % 2].copy_(host[i % NBUF], non_blocking=True)               # H2D from pinned memory
            h2d_ev[i % 2].record(copy_stream)

    w0 = time.perf_counter()
    e0.record()
    prefetch(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            prefetch(i + 1)
        main.wait_event(h2d_ev[i % 2])
        x = xin[i % 2]
        if use_graph:
            bpp_i, psnr_i, out = gev2(i % 2)
            used_ev[i % 2].record(main)
            res_host[i % 2][0:2].copy_(gev2.result, non_blocking=True)                                     # D2H
            res_host[i % 2][2:].copy_(out["v_mse"], non_blocking=True)
        else:
            bpp_i, psnr_i, out = ev(x)
            used_ev[i % 2].record(main)
            v_mse = (out["sq_err"].to(torch.float64) / (3 * H * W)).to(torch.float32)
            res_host[i % 2].copy_(torch.cat([bpp_i.reshape(1), psnr_i.reshape(1), v_mse]), non_blocking=True)   # D2H
        res_ev[i % 2].record(main)
        if i > 0:                                   # host consumes step i-1's result while step i runs
            res_ev[(i - 1) % 2].synchronize()
            results.append(res_host[(i - 1) % 2].clone())
    res_ev[(args.steps - 1) % 2].synchronize()
    results.append(res_host[(args.steps - 1) % 2].clone())
    e1.record()
    sync_all()
    wall_ms = (time.perf_counter() - w0) * 1e3
    clocks = sampler.stop()
    e2e_ms = e0.elapsed_time(e1)
    tt = torch.tensor([max(e2e_ms, 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(tt.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- likelihood/bpp kernel against the HBM roofline (C5-size problem) ----------------
    n_el = 16 * 192 * 128 * 128
    v = torch.randn(n_el, device=dev) * 4
    mu = torch.randn(n_el, device=dev)
    sg = torch.exp(torch.randn(n_el, device=dev)).clamp_(0.05, 20)
    vh = torch.empty_like(v)
    lk = torch.empty_like(v)
    for _ in range(3):
        ops.likelihood_rows(v, 1, n_el, v_rs=n_el, mu=mu, mu_mode=2, mu_rs=n_el, sigma=sg, sigma_mode=2, sigma_rs=n_el,
                            quant=ops.QUANT_ROUND, v_hat=vh, v_hat_rs=n_el, lik=lk)
    torch.cuda.synchronize(dev)
    reps = 10
    e0.record()
    for _ in range(reps):
        ops.likelihood_rows(v, 1, n_el, v_rs=n_el, mu=mu, mu_mode=2, mu_rs=n_el, sigma=sg, sigma_mode=2, sigma_rs=n_el,
                            quant=ops.QUANT_ROUND, v_hat=vh, v_hat_rs=n_el, lik=lk)
    e1.record()
    torch.cuda.synchronize(dev)
    lik_ms = e0.elapsed_time(e1) / reps
    lik_gbs = 20.0 * n_el / (lik_ms * 1e-3) / 1e9
    del v, mu, sg, vh, lk

    # ---------------- window attention core against the HBM roofline: the 1/4-resolution block of
    # BASELINE configs[3] (1152x1920 padded crops, 4 images per GPU): 288 x 480 tokens x 192 channels, 8x8 windows, shift 4
    wa_B, wa_H, wa_W, wa_C, wa_heads, wa_ws = 4, 288, 480, 192, 8, 8
    qkv = [torch.randn(wa_B, wa_H, wa_W, wa_C, device=dev).to(torch.bfloat16) for _ in range(3)]
    wa_bias = torch.randn((2 * wa_ws - 1) ** 2, wa_heads, device=dev) * 0.1      # relative_position_bias_table, indexed in the kernel
    for _ in range(3):
        ops.window_attention_core(qkv[0], qkv[1], qkv[2], wa_bias, wa_heads, wa_ws, 4)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        ops.window_attention_core(qkv[0], qkv[1], qkv[2], wa_bias, wa_heads, wa_ws, 4)
    e1.record()
    torch.cuda.synchronize(dev)
    wa_ms = e0.elapsed_time(e1) / reps
    wa_tokens = wa_B * wa_H * wa_W
    wa_bytes = wa_tokens * wa_C * 2 * 4                      # q, k, v read + out written, bf16
    wa_gbs = wa_bytes / (wa_ms * 1e-3) / 1e9
    del qkv, wa_bias

    conv_traffic, lik_traffic = measured_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if not args.parity_tf32 else "tf32 (g_a, h_a) + bf16", "data": "synthetic",
        "config": {"workload": f"model/net.py Net.forward(test) N={Nw} M={Mw} on batch {B} synthetic 768x512 per GPU "
                               + ("(BASELINE configs[1])" if not high else "(the reference's --high width, model/net.py:446-451)"),
                   "global_batch": world * B, "height": H, "width": W,
                   "input": ("uint8 levels (8-bit imagery as eval_net.py reads it); x = (u/255)*2-1 applied inside the first layer"
                             if args.input == "u8" else "fp32 in [-1,1]"),
                   "weights": "random init (tests/det_weights.py seed 0, gain-boosted)",
                   "l2": f"{NBUF} rotating input batches ({NBUF * in_bytes >> 20} MiB) and ~3 GB of activations per step, both > 126 MB L2",
                   "context_model": "PredictionModel_Context on conv_tc_kernel (TMA patch gather, SURVEY 8 f1); syntax branch on ldic_syntax_branch (fp32 CUDA-core kernels)",
                   "parallelism": f"batch sharded over {world} GPU(s), 1 all-reduce of 5 scalars per step"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": float(tt.item()) / args.steps, "wall_ms_per_step": wall_ms / args.steps},
        "gpu_launches": launches,
        "launch_mode": ("CUDA graph replay: one graph per static input buffer holding every kernel of the step "
                        f"({gev.launches_per_replay} launches of libldic_b200), + 1 metric kernel per step" if use_graph
                        else "eager launches"),
        "eager_ms_per_step": eager_ms / args.steps,
        "eager_note": "net(x, 'test') on device-resident inputs: the nn.Module call a user makes"
                      + (" (transparent CUDA-graph replay per input shape)" if net.auto_graph else " (every kernel launched from Python)"),
        "profiled_eager_ms_per_step": profiled_eager_ms / args.steps,
        "streams": (f"SM partition: hyperprior / syntax chain on a side stream on {net.side_sms} SMs next to g_s deconv 1-3"
                    if net.side_sms else "single stream"),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_first_kernel / conv_tc2_kernel (g_a, g_s, h_a, h_s, context convs + fused GDN/IGDN)",
                     "achieved": achieved_tf, "peak": tc_peak_sus, "unit": "TFLOP/s",
                     "frac": achieved_tf / tc_peak_sus if tc_peak_sus else None,
                     "traffic": conv_traffic, "traffic_note": "DRAM bytes of all conv launches of one step (ncu --set full, profiles/r02_traffic.json)",
                     "peak_source": f"{peak_src} bf16_tflops_sustained",
                     "algorithmic_gflop_per_image": conv_flops / 1e9 / (B * args.steps),
                     "conv_ms_per_step": conv_ms / args.steps,
                     "share_of_step": conv_ms / profiled_eager_ms if profiled_eager_ms else None,
                     "share_note": "conv launch time / step time of the eager per-launch-event pass",
                     "per_layer_tflops": {k: round(d[1] / (d[0] * 1e-3) / 1e12, 1) for k, d in per_layer.items() if d[0] > 0},
                     "per_layer_ms_per_step": {k: round(d[0] / args.steps, 4) for k, d in per_layer.items()}},
        "roofline_likelihood": {"bound": "hbm", "kernel": "k_likelihood_fast<1,false> (round + Gaussian likelihood + sum ln L)",
                                "achieved": lik_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": lik_gbs / hbm_peak,
                                "traffic": lik_traffic, "bytes_per_elem": 20, "elems": n_el, "ms": lik_ms,
                                "peak_source": f"{peak_src} hbm_gbs", "workload": "C5-size 16x192x128x128, per-element mu/sigma"},
        "roofline_window_attention": {"bound": "hbm", "kernel": "k_window_attention<32,64> (softmax(q k^T + bias + shift mask) v per 8x8 window)",
                                      "achieved": wa_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": wa_gbs / hbm_peak,
                                      "traffic": measured_traffic_of("window_attention"), "bytes_per_token": wa_C * 2 * 4, "tokens": wa_tokens, "ms": wa_ms,
                                      "peak_source": f"{peak_src} hbm_gbs",
                                      "workload": "4 x 288x480 tokens x 192 ch (1/4 resolution of a 1152x1920 crop), 8 heads, window 8, shift 4"},
        "parity": {"bpp": float(bpp.item()), "psnr_db": float(psnr.item()), "multi_gpu_equal": multi_gpu_equal,
                   "note": "gates (bpp 0.5 %, PSNR 0.01 dB vs the unmodified reference) are enforced by tests/test_gpu_net.py at "
                           "B=1 768x512 and smaller; kernels are bit-exactly batch independent (test_full_size_batch_properties), "
                           "so B=16 itself is covered by B=1 + batch independence"},
    }
    if world == 1 and not args.no_cpu_baseline:
        ips, spi, cores = cpu_port_images_per_s(2, 1, batch=1, high=high)
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 image 768x512 per step (same net, same weights), 2 timed steps after 1 warm-up, "
                                          "torch CPU fp32 port of the reference path with its one-hot sampler convs"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_tritplane(args):
    """BASELINE configs[4]: progressive trit-plane quantisation + per-plane likelihood at 2048x2048 (latents
    192 x 128 x 128 per image).  A step = one pass of ldic_tritplane_likelihood over this rank's images; images shard
    over the ranks, the only exchange is one all-reduce of the L per-plane sums.  HBM bound: 12 B read +
    (L + 4) B written per element.  Parity for this row is UNPINNED (the reference file crashes); the
    bench checks the bit-exact symbol property q == clamp(round(v - mu)) and the exact reconstruction from the planes."""
    import torch
    import torch.distributed as dist
    import ldic_b200
    from ldic_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            import numpy as np
            from oracle import tritplane_ref
            n = 192 * 128 * 128 // 64                               # bounded sample: 1/64 of one image's latents
            r = np.random.Generator(np.random.PCG64(0))
            v = (4 * r.standard_normal(n)).astype(np.float32); mu = r.standard_normal(n).astype(np.float32)
            sg = np.clip(np.exp(r.standard_normal(n)), 0.05, 20).astype(np.float32)
            t0 = time.perf_counter()
            for _ in range(max(args.steps, 1)):
                tritplane_ref.tritplane(v, sg, mu, planes=4)
            dt = (time.perf_counter() - t0) / max(args.steps, 1)
            ips = (n / (192 * 128 * 128)) / dt
            emit({"impl": "reference", "metric": "2048x2048 imgs/sec (trit-plane symbols + per-plane likelihood)", "value": ips,
                  "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                  "config": {"workload": "numpy restatement of the trit-plane extension (oracle/tritplane_ref.py), 1/64 image per step"},
                  "cpu_baseline": {"value": ips, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{n} latent elements per step"},
                  "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(local), "device")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, _, _, peak_src = peaks()
    B, Lp = args.batch, 4
    per_img = 192 * 128 * 128
    n = B * per_img
    g = torch.Generator(device=dev).manual_seed(rank)
    NBUF = 3                                                        # rotating inputs: 3 x 12 B x n >> L2 at the default batch
    bufs = [(torch.randn(n, device=dev, generator=g) * 4, torch.randn(n, device=dev, generator=g),
             torch.exp(torch.randn(n, device=dev, generator=g)).clamp_(0.05, 20)) for _ in range(NBUF)]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step(i):
        v, mu, sg = bufs[i % NBUF]
        pl, q, sums = ops.tritplane_likelihood(v, sg, mu, planes=Lp)
        s64 = sums.double()
        if world > 1:
            dist.all_reduce(s64)
        return pl, q, s64
    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pl, q, s64 = step(i)
    e1.record()
    sync_all()
    launches = int(ops.launch_count() - n0)
    tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    # kernel alone (events around the launches only, no allocation / reduction in between)
    v, mu, sg = bufs[0]
    torch.cuda.synchronize(dev)
    k_ms = 0.0
    for i in range(args.steps):
        v, mu, sg = bufs[i % NBUF]
        e0.record()
        ops.tritplane_likelihood(v, sg, mu, planes=Lp)
        e1.record()
        torch.cuda.synchronize(dev)
        k_ms += e0.elapsed_time(e1)
    k_ms /= args.steps
    clocks = sampler.stop()
    # end to end: host buffers in, symbols + planes + sums out
    hv = [tuple(t.cpu().pin_memory() for t in b) for b in bufs[:2]]
    out_host = (torch.empty((Lp, n), dtype=torch.int8).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory())
    sync_all()
    e0.record()
    for i in range(args.steps):
        dv = [t.to(dev, non_blocking=True) for t in hv[i % 2]]
        pl_, q_, sums_ = ops.tritplane_likelihood(dv[0], dv[2], dv[1], planes=Lp)
        out_host[0].copy_(pl_.view(Lp, n), non_blocking=True); out_host[1].copy_(q_, non_blocking=True)
        sums_.cpu()
    e1.record()
    sync_all()
    tt2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
    # bit-exact symbol property on the last step's outputs
    v, mu, sg = bufs[(args.steps - 1) % NBUF]
    Hh = (3 ** Lp - 1) // 2
    q_ref = torch.clamp(torch.round(v - mu), -Hh, Hh).to(torch.int32)
    recon = sum(pl[l].to(torch.int32) * (3 ** l) for l in range(Lp)) - Hh
    sym_ok = bool(torch.equal(q, q_ref)) and bool(torch.equal(recon, q))
    if world > 1:
        ok_t = torch.tensor([int(sym_ok)], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        sym_ok = bool(ok_t.item())
    if rank == 0:
        bytes_per_elem = 12 + Lp + 4
        gbs = bytes_per_elem * n / (k_ms * 1e-3) / 1e9
        emit({"metric": "2048x2048 imgs/sec (trit-plane symbols + per-plane likelihood)", "value": world * B * args.steps / (t_ms * 1e-3),
              "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": f"BASELINE configs[4]: trit-plane quantisation + likelihood, {B} images of 2048x2048 "
                                     f"(latents 192x128x128) per GPU, {Lp} planes", "global_batch": world * B,
                         "l2": f"{NBUF} rotating input sets of {12 * n >> 20} MiB", "parity": "UNPINNED (no reference behaviour, SURVEY 8 a12)"},
              "e2e": {"value": world * B * args.steps / (float(tt2.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 12 * n,
                      "d2h_bytes_per_step": (Lp + 4) * n + 4 * Lp},
              "gpu_launches": launches, "clocks": clocks,
              "roofline": {"bound": "hbm", "kernel": "k_tritplane", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gbs / hbm_peak, "traffic": measured_traffic_of("tritplane_c5") if B == 16 else None,
                           "bytes_per_elem": bytes_per_elem, "elems": n, "ms": k_ms,
                           "note": "ALU bound, not HBM bound: 10 erfc + 5 log evaluations per element (L + 1 nested interval masses)",
                           "peak_source": f"{peak_src} hbm_gbs"},
              "parity": {"symbols_bit_exact": sym_ok, "sum_ln_per_plane": [float(x) for x in s64.tolist()]}})
    if world > 1:
        dist.destroy_process_group()


def run_codec(args):
    """SURVEY row f4 (the "what comes next" row): the whole encoder x -> bitstreams.  A step = Net.rd_forward on one batch
    of 768x512 images + rANS coding of its three symbol streams (z, y content, syntax) into one bitstream per image and
    stream.  Images shard over the ranks; nothing is exchanged.  The reference has no entropy coder (it estimates the
    rate), so the parity of this stage is UNPINNED against it; the bench checks the exact round trip of the coded
    symbols and reports coded vs estimated bpp."""
    import torch
    import torch.distributed as dist
    import ldic_b200
    from ldic_b200 import ops
    import det_weights as dw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "768x512 imgs/sec encoded to bitstreams (g_a->hyperprior->context model->rANS)"
    if args.impl == "reference":
        if rank == 0:
            import numpy as np
            from oracle import rans_ref
            # one image per step: the CPU port of the reference forward (all host cores) + the numpy restatement of the
            # coder on one image's worth of content symbols (synthetic symbols following their model; 1 core)
            n_fw = max(1, min(args.steps, 2))
            _, fw_s, cores = cpu_port_images_per_s(n_fw, 1)
            n, S = 32 * 48 * 176, 132
            r = np.random.Generator(np.random.PCG64(0))
            mu = (3 * r.standard_normal(n)).astype(np.float32)
            sg = np.exp(0.8 * r.standard_normal(n) + 0.2).astype(np.float32)
            k = np.rint(mu + sg * r.standard_normal(n)).astype(np.int64)
            t0 = time.perf_counter()
            for _ in range(n_fw):
                rans_ref.encode_segment(k, mu, sg, S)
            rans_s = (time.perf_counter() - t0) / n_fw
            dt = fw_s + rans_s
            ips = 1.0 / dt
            emit({"impl": "reference", "metric": metric, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                  "dtype": "f32 + int64", "data": "synthetic",
                  "config": {"workload": "CPU port of the reference forward (oracle/ref_path.py, 1 image 768x512 per step) + numpy "
                                         "restatement of the rANS coder on one image's content symbols (oracle/rans_ref.py); the "
                                         "reference itself has no coder", "forward_s": fw_s, "rans_s": rans_s},
                  "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{n_fw} image(s): forward on {cores} cores + {n} symbols coded on 1 core"},
                  "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(local), "device")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, _, _, peak_src = peaks()
    B = args.batch
    net = ldic_b200.Net((B, H, W, 3), (B, H, W, 3), False, False).to(dev).eval()
    net.load_state_dict(dw.make_state_dict(0), strict=True)
    NBUF = 4
    host = [t.pin_memory() for t in make_u8_batches(rank, B, NBUF)]
    xs = [t.to(dev) for t in host]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    ev = None if args.no_graph else ldic_b200.GraphedEvaluator(net, xs, entropy_code=True)

    def step(k):
        if ev is not None:                       # one CUDA graph per static input buffer: forward + coder
            _, _, out = ev(k)
            return out, out["streams"]
        out = net.rd_forward(xs[k])
        return out, net.entropy_encode(out)
    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step(i % NBUF)
        sync_all()
        sampler = ClockSampler(local)
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            out, enc = step(i % NBUF)
        e1.record()
        sync_all()
        launches = int(ops.launch_count() - n0) + (ev.launches_per_replay * args.steps if ev is not None else 0)
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        # the entropy-coding stage alone (events around the encoder / decoder calls of the last step's latents)
        lat = out["latents"]
        h, w, Cc = H // 16, W // 16, net.N - net.M
        ekw = dict(mu=lat["ctx"], mu_mode=2, mu_rs=lat["ctx_rs"], sigma=lat["ctx"], sigma_mode=2, sigma_rs=lat["ctx_rs"],
                   sigma_off=lat["ctx_sig_off"], sigma_is_log=True)
        y_hat = torch.empty(B, h, w, Cc, device=dev)
        enc_ms = dec_ms = 0.0
        Sy = net.y_streams(h, w)
        for i in range(3):                      # back-to-back launches, one event pair around all of them
            ency = ops.rans_encode_rows(lat["y"], B * h * w, Cc, h * w, v_rs=net.N, v_off=net.M, streams=Sy, **ekw)
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(args.steps):
            ency = ops.rans_encode_rows(lat["y"], B * h * w, Cc, h * w, v_rs=net.N, v_off=net.M, streams=Sy, **ekw)
        e1.record()
        torch.cuda.synchronize(dev); enc_ms = e0.elapsed_time(e1) / args.steps
        e0.record()
        for i in range(args.steps):
            ops.rans_decode_rows(ency, B * h * w, Cc, h * w, y_hat, v_hat_rs=Cc, check_status=False, **ekw)
        e1.record()
        torch.cuda.synchronize(dev); dec_ms = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop()
        round_trip = bool(torch.equal(y_hat, torch.round(lat["y"][..., net.M:])))
        z_hat = net.decode_z(enc["z"], B, H, W)
        round_trip = round_trip and bool(torch.equal(z_hat, torch.round(lat["z"])))
        sizes = {k: v.nbytes() for k, v in enc.items()}
        coded_bits = 8.0 * sum(sum(v) for v in sizes.values())
        # the whole decoder (bytes -> x_hat, context model walked along wavefronts) on this step's bitstreams
        parts = dict(zip(enc.keys(), ops.rans_tobytes(enc.values())))
        per_image = [{k: parts[k][b] for k in parts} for b in range(B)]
        x_ref = net.rd_forward(xs[(args.steps - 1) % NBUF], want_x_hat=True)["x_hat"]
        for _ in range(2):                                        # first call of a shape runs eagerly, the second captures
            net.decompress(per_image, H, W)                       # the wavefront loop into a CUDA graph
        torch.cuda.synchronize(dev)
        t_dec = time.perf_counter()
        x_dec = net.decompress(per_image, H, W)
        torch.cuda.synchronize(dev)
        t_dec = time.perf_counter() - t_dec
        decoder_exact = bool(torch.equal(x_dec, x_ref))
        est_bits = float(out["bits"].double().sum().item()) / -math.log(2.0)
        # end to end: uint8 host images in, bitstream bytes out (sizes first, then exactly the coded bytes)
        sync_all()
        d2h = 0
        e0.record()
        # a serving loop: the H2D copy of step i+1 runs on its own stream under step i's kernels, and step i's bitstreams
        # are read back on a third stream (sizes first, then exactly the coded bytes) while step i+1 runs
        main = torch.cuda.current_stream(dev)
        in_stream, out_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        h2d_ev = [torch.cuda.Event() for _ in range(NBUF)]
        used_ev = [torch.cuda.Event() for _ in range(NBUF)]
        done_ev = [torch.cuda.Event() for _ in range(NBUF)]

        def prefetch(i):
            k = i % NBUF
            with torch.cuda.stream(in_stream):
                if i >= NBUF:
                    in_stream.wait_event(used_ev[k])
                xs[k].copy_(host[k], non_blocking=True)
                h2d_ev[k].record(in_stream)
        prefetch(0)
        pending = None
        for i in range(args.steps + 1):
            if i < args.steps:
                k = i % NBUF
                if i + 1 < args.steps:
                    prefetch(i + 1)
                main.wait_event(h2d_ev[k])
                enc_k = step(k)[1]
                used_ev[k].record(main); done_ev[k].record(main)
            if pending is not None:
                out_stream.wait_event(done_ev[pending[1]])
                with torch.cuda.stream(out_stream):
                    d2h += sum(len(b) + 8 for v in ops.rans_tobytes(pending[0].values()) for b in v)
            pending = (enc_k, k) if i < args.steps else None
        e1.record()
        sync_all()
        tt2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
        ok_t = torch.tensor([int(round_trip)], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        round_trip = bool(ok_t.item())
    t_ms = float(tt.item())
    if rank == 0:
        nsym = B * h * w * Cc
        alg_bytes = 12.0 * nsym + sum(sizes["y"])                  # v, mu, log sigma read once; bitstream written once
        gbs = alg_bytes / (enc_ms * 1e-3) / 1e9
        emit({"metric": metric, "value": world * B * args.steps / (t_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
              "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "bf16 (transforms) + int32 (coder)", "data": "synthetic",
              "config": {"workload": f"SURVEY f4: model/net.py Net forward + rANS coding of z / y / syntax symbols, 768x512, batch {B} "
                                     "per GPU, one bitstream per image and stream", "global_batch": world * B, "height": H, "width": W,
                         "weights": "deterministic random init (tests/det_weights.py)",
                         "launch_mode": "eager" if ev is None else "CUDA graph replay (forward + coder in one graph per input buffer)",
                         "l2": f"{NBUF} rotating input batches; activations of one step exceed L2",
                         "parity": "UNPINNED against the reference (it has no entropy coder); exact round trip + CPU restatement (tests/test_gpu_rans.py)"},
              "e2e": {"value": world * B * args.steps / (float(tt2.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * 3 * H * W,
                      "d2h_bytes_per_step": d2h // max(args.steps, 1)},
              "gpu_launches": launches, "clocks": clocks,
              "roofline": {"bound": "hbm", "kernel": "rANS encoder of the content symbols (k_rans_ops / k_rans_enc_streams / k_rans_pack + 3 small)",
                           "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                           "bytes_per_symbol": alg_bytes / nsym, "symbols": nsym, "ms": enc_ms, "decode_ms": dec_ms,
                           "note": "latency bound, not HBM bound: the rANS state recurrence is sequential within a stream "
                                   "(about 2048 symbols per stream; ~175 cycles per symbol on one thread)",
                           "peak_source": f"{peak_src} hbm_gbs"},
              "decoder": {"images_per_s": B / t_dec, "ms_per_batch": t_dec * 1e3, "x_hat_bit_identical_to_encoder": decoder_exact,
                          "note": "Net.decompress: bytes -> x_hat from the bitstreams and the model alone; the causal context "
                                  "model is re-run on the partially decoded latent at each of the w + 2(h-1) wavefront steps "
                                  "(band schedule, the loop replayed as one CUDA graph); rank 0's batch, wall clock"},
              "parity": {"round_trip_exact": round_trip, "bpp_coded": coded_bits / (B * H * W), "bpp_estimated": est_bits / (B * H * W),
                         "bytes_per_image": {k: sum(v) / B for k, v in sizes.items()}}})
    if world > 1:
        dist.destroy_process_group()


def run_unet(args):
    """BASELINE configs[2] (768x512, global batch 64 sharded over the GPUs) and, with --crop 1280x2048, configs[3]
    (1920x1080 crops padded to the multiple of 256 the model needs, global batch 32): the U-Net-family Net (model/net_unet_ha_hs.py) with
    its hot path on libldic_b200 and its other blocks as stock torch modules.  Random-init weights by name
    (tests/det_weights.py::unet_param_fill), synthetic images."""
    import torch
    import torch.distributed as dist
    import ldic_b200
    from ldic_b200 import ops, net_unet
    from ldic_b200.layers import WinBasedAttention
    import det_weights as dw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            emit({"impl": "reference", "unavailable": "the U-Net family needs compressai / timm and two modules missing from the "
                  "reference tree; it only runs with restated dependencies in the build container (oracle/unet_harness.py)"})
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ldic_b200._lib.check(ldic_b200._lib.load().ldic_check_device(local), "device")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    Hh, Ww = (int(v) for v in args.crop.split("x"))
    gb = args.global_batch if args.global_batch else (64 if (Hh, Ww) == (512, 768) else 32)
    B = max(1, gb // world) if not args.batch_set else args.batch
    net = net_unet.Net((B, Hh, Ww, 3), (B, Hh, Ww, 3), False, False).to(dev).eval()
    fill = dw.unet_param_fill([(n, tuple(p.shape)) for n, p in net.named_parameters()], 0)
    net.load_state_dict({**net.state_dict(), **{k: v.to(dev) for k, v in fill.items()}}, strict=True)
    base = dw.make_input(rank, 2, Hh, Ww)
    NBUF = 2
    host = [torch.cat([torch.roll(base, shifts=i * 37 + j * 11, dims=3) for j in range((B + 1) // 2)], 0)[:B].contiguous().pin_memory()
            for i in range(NBUF)]
    devbuf = [h.to(dev) for h in host]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step(x):
        out = net.rd_forward(x)
        _, h, w, _ = net.train_size
        bits3 = torch.stack([out["bits"].sum(), out["bits"].new_zeros(()), out["bits"].new_zeros(())]).contiguous()
        packed, _ = ops.rd_pack_metrics(bits3, out["sq_err"], 3 * Hh * Ww, want_v_mse=False)
        if world > 1:
            dist.all_reduce(packed)
        return ops.rd_finish_metrics(packed, float(h * w))
    for i in range(max(args.warmup, 3)):
        r = step(devbuf[i % NBUF])
    sync_all()
    sampler = ClockSampler(local)
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        r = step(devbuf[i % NBUF])
    e1.record()
    sync_all()
    launches = int(ops.launch_count() - n0)
    tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms = float(tt.item())
    # share of the step inside the window-attention blocks / the Win_noShift_Attention modules (events around the modules)
    spans = {"WinBasedAttention": [], "Win_noShift_Attention": []}
    hooks = []
    for mod in net.modules():
        if isinstance(mod, net_unet.Win_noShift_Attention):
            def pre(m, i):
                e = torch.cuda.Event(enable_timing=True); e.record(); m._e0 = e
            def post(m, i, o):
                e = torch.cuda.Event(enable_timing=True); e.record(); spans["Win_noShift_Attention"].append((m._e0, e))
            hooks += [mod.register_forward_pre_hook(pre), mod.register_forward_hook(post)]
    # the attention blocks are entered through forward() (NCHW surface) or forward_nhwc() (inside the NHWC pipeline)
    orig = (WinBasedAttention.forward, WinBasedAttention.forward_nhwc)

    def timed(fn):
        def f(self, t):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(); r_ = fn(self, t); eb.record()
            spans["WinBasedAttention"].append((ea, eb))
            return r_
        return f
    WinBasedAttention.forward, WinBasedAttention.forward_nhwc = timed(orig[0]), timed(orig[1])
    try:
        e0.record()
        step(devbuf[0])
        e1.record()
        torch.cuda.synchronize(dev)
    finally:
        WinBasedAttention.forward, WinBasedAttention.forward_nhwc = orig
    hooked_ms = e0.elapsed_time(e1)
    share = {k: sum(a.elapsed_time(b) for a, b in v) / hooked_ms for k, v in spans.items()}
    for h in hooks:
        h.remove()
    # end to end
    xin = [torch.empty_like(devbuf[0]) for _ in range(2)]
    sync_all()
    e0.record()
    for i in range(args.steps):
        xin[i % 2].copy_(host[i % NBUF], non_blocking=True)
        rr = step(xin[i % 2]).cpu()
    e1.record()
    sync_all()
    tt2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
    clocks = sampler.stop()
    if rank == 0:
        cfgname = "configs[2]" if (Hh, Ww) == (512, 768) else "configs[3]"
        emit({"metric": f"{Ww}x{Hh} imgs/sec (U-Net family: g_a -> U-Net hyperprior -> slice entropy model -> bpp -> g_s)",
              "value": world * B * args.steps / (t_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
              "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True,
              "scaling": "strong" if not args.batch_set else "weak", "vs_baseline": None, "dtype": "bf16 (kernels) / fp32-TF32 (torch blocks)",
              "data": "synthetic",
              "config": {"workload": f"BASELINE {cfgname}: model/net_unet_ha_hs.py Net.forward(test), {Ww}x{Hh}, global batch {world * B} "
                                     f"({B} per GPU)", "global_batch": world * B, "height": Hh, "width": Ww,
                         "weights": "random init by name (tests/det_weights.py::unet_param_fill)",
                         "l2": "activations of one step are several GB (> 126 MB L2)", "launch_mode": "eager",
                         "parity": "restated deps (tests/test_gpu_unet.py)"},
              "e2e": {"value": world * B * args.steps / (float(tt2.item()) * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": host[0].numel() * 4, "d2h_bytes_per_step": 8},
              "gpu_launches": launches, "clocks": clocks,
              "share_of_step": {"WinBasedAttention_blocks": share["WinBasedAttention"],
                                "Win_noShift_Attention_modules": share["Win_noShift_Attention"],
                                "note": "CUDA-event spans around the modules in one eager step"},
              "roofline": {"bound": "tensor", "kernel": "mixed: libldic_b200 kernels + cuDNN/cuBLAS blocks", "achieved": None,
                           "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None},
              "parity": {"bpp": float(r[0].item()), "psnr_db": float(r[1].item())}})
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print to fd 1 (NCCL's version banner, torchrun notices)
    is sent to stderr; emit() writes the line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--config", default="net", choices=["net", "high", "tritplane", "unet", "codec"],
                    help="net: BASELINE configs[1] (headline); high: the N=384 model; tritplane: BASELINE configs[4]; "
                         "codec: forward + rANS bitstreams (SURVEY f4); unet: the U-Net family, configs[2] (768x512, global batch 64) or with --crop 1280x2048 configs[3]")
    ap.add_argument("--crop", default="512x768", help="unet: image size HxW, multiples of 256 (configs[3]: 1080x1920 crops pad to 1280x2048)")
    ap.add_argument("--global-batch", type=int, default=0, help="unet: global batch (default 64 at 768x512, else 32)")
    ap.add_argument("--input", default="u8", choices=["u8", "f32"], help="image type of the input buffers")
    ap.add_argument("--parity-tf32", action="store_true", help="run g_a / h_a in the TF32 parity mode (not the headline configuration)")
    ap.add_argument("--side-sms", type=int, default=-1, help="SM partition for the hyperprior / syntax side stream (0: single stream; default: Net's)")
    args = ap.parse_args()
    args.batch_set = any(a == "--batch" or a.startswith("--batch=") for a in sys.argv[1:])
    if args.config == "unet":
        run_unet(args)
    elif args.config == "tritplane":
        run_tritplane(args)
    elif args.config == "codec":
        run_codec(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""CUDA-graph replay of the rank-local rate-distortion forward.

One step of Net.rd_forward is ~30 kernel launches of 5 us .. 0.7 ms each; issued one by one from Python the
gaps between them add up to ~7 % of the step (serialised kernel time 3.43 ms vs 3.68 ms per step, measured with
`ncu --metrics gpu__time_duration.sum`, profiles/).  For fixed-shape batches the whole launch sequence -- every
kernel of libldic_b200 from the first conv to the packed metric sums -- is captured once per static input buffer
and replayed with one `cudaGraphLaunch`.  Nothing is skipped: the graph holds exactly the launches the eager
forward makes (TMA descriptors are encoded on the host at capture time and baked into the kernel parameters,
activations live in the graph's private memory pool, so their addresses are stable across replays).

The multi-GPU exchange stays outside the graph: replay -> ONE all-reduce of the five packed doubles (world > 1)
-> ldic_rd_finish_metrics.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops


class GraphedEvaluator:
    """`ev = GraphedEvaluator(net, [x0, x1, ...])`; write a batch into `x_k` (any stream-ordered way: H2D copy,
    `copy_`), then `bpp, psnr, out = ev(k)`.  The returned tensors are the graph's static outputs: they are
    overwritten by the next replay of the same graph."""

    def __init__(self, net, inputs: Sequence[torch.Tensor], group: Optional[dist.ProcessGroup] = None,
                 rd_kwargs: Optional[dict] = None, entropy_code: bool = False):
        if not inputs:
            raise ops.LdicError("GraphedEvaluator needs at least one static input buffer")
        for x in inputs:
            if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype in (torch.float32, torch.uint8) and x.is_contiguous()
                    and x.dim() == 4 and x.shape == inputs[0].shape and x.dtype == inputs[0].dtype):
                raise ops.LdicError("GraphedEvaluator: inputs must be contiguous CUDA fp32 / uint8 (B,3,H,W) tensors of one shape")
        self.net, self.group, self.inputs = net, group, list(inputs)
        self.rd_kwargs = dict(rd_kwargs or {})
        self.entropy_code = bool(entropy_code)    # also capture Net.entropy_encode: out["streams"] = the rANS bitstreams
        B, _, H, W = inputs[0].shape
        self.chw = 3 * H * W
        _, th, tw, _ = net.test_size
        self.pixels_per_image = float(th * tw)
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.outs: List[Dict[str, torch.Tensor]] = []
        self.packed: List[torch.Tensor] = []
        self.v_mse: List[torch.Tensor] = []
        self.result = torch.empty(2, dtype=torch.float32, device=inputs[0].device)
        if ops.PROFILE is not None:
            raise ops.LdicError("GraphedEvaluator: per-launch event profiling (ops.PROFILE) cannot be captured")
        # warm-up on a side stream: first-call work (function attributes, workspaces, weight packing) must not be captured
        side = torch.cuda.Stream(device=inputs[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                self._step(self.inputs[0])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(inputs[0].device)
        pool = torch.cuda.graph_pool_handle()
        self.launches_per_replay = 0          # kernels of libldic_b200 inside one graph (counted at capture)
        for x in self.inputs:
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"), torch.no_grad():
                out, packed, v_mse = self._step(x)
            self.launches_per_replay = ops.launch_count() - n0
            self.graphs.append(g)
            self.outs.append(out)
            self.packed.append(packed)
            self.v_mse.append(v_mse)

    def _step(self, x) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor]:
        out = self.net.rd_forward(x, **self.rd_kwargs)
        if self.entropy_code:
            out["streams"] = self.net.entropy_encode(out)
        packed, v_mse = ops.rd_pack_metrics(out["bits"], out["sq_err"], self.chw)
        return out, packed, v_mse

    @torch.no_grad()
    def __call__(self, k: int = 0):
        self.graphs[k].replay()
        packed = self.packed[k]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
        r = ops.rd_finish_metrics(packed, self.pixels_per_image, out=self.result)
        out = dict(self.outs[k])
        out["v_mse"] = self.v_mse[k]
        return r[0], r[1], out

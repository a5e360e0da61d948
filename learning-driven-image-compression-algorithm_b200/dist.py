"""Multi-GPU evaluation: the image batch shards across ranks (images are independent units,
weights replicated); the only exchange is ONE all-reduce of five scalars per batch
(SURVEY 8e).  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

The reference has no distributed eval (single GPU, eval_net.py:204); the reduction below
reproduces what its single-process forward would return on the concatenated global batch:
  bpp    = sum over ALL images of sum(ln L) / (-ln2 * B_global * th * tw)   (model/net.py:856-859)
  v_psnr = mean over ALL images of 20 log10(255 / sqrt(v_mse_i))             (model/net.py:869)
PSNR is summed AFTER the per-image log, so MSE itself is never all-reduced.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of the batch; the first (global_batch % world_size) ranks get one extra image."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_local(bits: torch.Tensor, sq_err: torch.Tensor, chw: int) -> torch.Tensor:
    """[sum ln L (z), (y), (syntax), sum_i psnr_i, n_images] in float64 on the tensors' device."""
    v_mse = sq_err.to(torch.float64) / float(chw)
    psnr_i = 20.0 * torch.log10(255.0 / torch.sqrt(v_mse))
    n = torch.tensor([float(sq_err.numel())], dtype=torch.float64, device=bits.device)
    return torch.cat([bits.to(torch.float64), psnr_i.sum().reshape(1), n])


def finish(packed: torch.Tensor, th: int, tw: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n = packed[4]
    bpp = packed[:3].sum() / (-math.log(2) * n * th * tw)
    return bpp.to(torch.float32), (packed[3] / n).to(torch.float32)


def reduce_metrics(bits: torch.Tensor, sq_err: torch.Tensor, chw: int, th: int, tw: int,
                   group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global (bpp, v_psnr) from each rank's local sums: one all-reduce(sum) of 5 doubles."""
    if bits.is_cuda:            # two launches on libldic_b200 instead of ~25 elementwise ones (same arithmetic, in double)
        from . import ops
        packed, _ = ops.rd_pack_metrics(bits, sq_err, chw, want_v_mse=False)
    else:                       # host-side logic (gloo tests)
        packed = pack_local(bits, sq_err, chw)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    if bits.is_cuda:
        r = ops.rd_finish_metrics(packed, float(th * tw))
        return r[0], r[1]
    return finish(packed, th, tw)


def gather_mse(sq_err: torch.Tensor, chw: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Optional: the reference's (B_global,) v_mse return, in rank order (equal shards)."""
    v = (sq_err.to(torch.float64) / float(chw)).to(torch.float32)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return v
    outs = [torch.empty_like(v) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, v, group=group)
    return torch.cat(outs)


class ShardedEvaluator:
    """Runs Net's rate-distortion forward on this rank's shard and reduces the metrics."""

    def __init__(self, net, group: Optional[dist.ProcessGroup] = None):
        self.net = net
        self.group = group

    @torch.no_grad()
    def __call__(self, x_local: torch.Tensor):
        out = self.net.rd_forward(x_local)
        _, th, tw, _ = self.net.test_size
        H, W = x_local.shape[2], x_local.shape[3]
        bpp, psnr = reduce_metrics(out["bits"], out["sq_err"], 3 * H * W, th, tw, self.group)
        return bpp, psnr, out

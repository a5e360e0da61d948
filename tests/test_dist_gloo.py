"""world_size-2 gloo test of the multi-GPU reduction (host logic; no GPU needed)."""
import math
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ldic_b200 import dist as ld
    g = torch.Generator().manual_seed(123)
    B, chw, th, tw = 6, 3 * 64 * 64, 60, 50
    bits_all = -torch.rand(B, 3, generator=g) * 1000          # per-image sum(ln L) of the 3 streams
    sq_all = (torch.rand(B, generator=g) * 1e7).long() + 1
    lo, hi = ld.shard_bounds(B, world, rank)
    bpp, psnr = ld.reduce_metrics(bits_all[lo:hi].sum(0), sq_all[lo:hi], chw, th, tw)
    mse = ld.gather_mse(sq_all[lo:hi], chw)
    q.put((rank, bpp.item(), psnr.item(), mse.tolist(), (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_reduction_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process expectation on the concatenated global batch (model/net.py:856-869)
    g = torch.Generator().manual_seed(123)
    B, chw, th, tw = 6, 3 * 64 * 64, 60, 50
    bits_all = -torch.rand(B, 3, generator=g) * 1000
    sq_all = (torch.rand(B, generator=g) * 1e7).long() + 1
    bpp_ref = bits_all.double().sum().item() / (-math.log(2) * B * th * tw)
    v_mse = sq_all.double() / chw
    psnr_ref = (20 * torch.log10(255 / torch.sqrt(v_mse))).mean().item()
    assert [r[4] for r in res] == [(0, 3), (3, 6)]
    for _, bpp, psnr, mse, _ in res:
        assert abs(bpp - bpp_ref) < 1e-6 * abs(bpp_ref)
        assert abs(psnr - psnr_ref) < 1e-5
        assert torch.allclose(torch.tensor(mse), v_mse.float(), rtol=1e-6)


def test_shard_bounds_ragged_and_empty():
    sys.path.insert(0, ROOT)
    from ldic_b200.dist import shard_bounds
    assert [shard_bounds(7, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 7)]
    assert [shard_bounds(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(64, 8, 7) == (56, 64)

"""world_size-2 gloo tests (CPU) of the multi-GPU batch-split plumbing (pfd_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pfd_b200.parallel import broadcast_conditioning, gather_images, shard_range, sharded_noise


def test_shard_ranges_partition_the_batch():
    for total in (1, 4, 5, 16, 32, 33):
        for world in (1, 2, 4, 8):
            rs = [shard_range(total, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = [total, 4, 8, 8]
        mine = sharded_noise(shape, seed=20, rank=rank, world=world)
        full = torch.randn(shape, generator=torch.Generator().manual_seed(20))
        a, b = shard_range(total, world, rank)
        ok_noise = torch.equal(mine, full[a:b])
        cond = torch.arange(148 * 8, dtype=torch.float32).reshape(1, 148, 8) if rank == 0 else None
        cond = broadcast_conditioning(cond, 0, shape=(1, 148, 8), dtype=torch.float32)
        ok_bcast = torch.equal(cond, torch.arange(148 * 8, dtype=torch.float32).reshape(1, 148, 8))
        imgs = (mine * 2.0).reshape(b - a, 4, 8, 8)                       # stand-in for the decode of this shard
        allim = gather_images(imgs, total)
        ok_gather = torch.equal(allim, full * 2.0)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                          # bench.py's max-over-ranks timing reduction
        q.put((rank, ok_noise, ok_bcast, ok_gather, t.item() == world))
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_split_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(all(r[1:]) for r in res), res

"""CPU, world_size 2 over gloo: the rank plumbing bench.py uses for N>1 (replica sharding by image, barrier,
max-over-ranks reduction of the per-rank time).  The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per_rank_batch, steps = 4, 3
    torch.manual_seed(1234 + rank)                        # each replica draws its own images (bench.py does the same)
    x = torch.rand(per_rank_batch, 3, 8, 8)
    dist.barrier()
    ms = torch.tensor([10.0 * (rank + 1)])                # pretend rank r needed 10*(r+1) ms
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_images = world * per_rank_batch * steps
    checks = torch.tensor([x.sum().item()])
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, checks)
    if rank == 0:
        out.put((ms.item(), total_images / (ms.item() / 1e3), [g.item() for g in gathered]))
    dist.destroy_process_group()


def test_replica_sharding_and_max_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, value, sums = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 20.0                                     # max over ranks, not the mean
    assert abs(value - 2 * 4 * 3 / 0.020) < 1e-6          # whole-job images / slowest rank's time
    assert sums[0] != sums[1]                             # ranks really processed different shards

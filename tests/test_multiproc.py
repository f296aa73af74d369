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


def _ddp_worker(rank, world, port, out):
    """Data-parallel gradient path of the training step (uformer_b200.training): FlatArena + GradReducer over gloo.
    The model is two LeWin blocks + samplers stated with uformer_b200.restated on CPU (the native forward needs a B200;
    the arena / bucket / hook / all-reduce logic under test is device agnostic)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn
    import uformer_b200 as U
    from uformer_b200 import restated as R
    from uformer_b200 import training as T
    from paramgen import randomize_state

    class Tiny(nn.Module):
        def __init__(self):
            super().__init__()
            self.b0 = U.LeWinTransformerBlock(16, (16, 16), 1, win_size=8, shift_size=0)
            self.down = U.Downsample(16, 32)
            self.b1 = U.LeWinTransformerBlock(32, (8, 8), 2, win_size=8, shift_size=0, modulator=True)
            self.unused = nn.Linear(4, 4)                        # never touched by forward: its bucket must still reduce

        def forward(self, x):
            return R.lewin_block(self.b1, R.downsample(self.down, R.lewin_block(self.b0, x)))

    net = Tiny()
    net.load_state_dict(randomize_state(net.state_dict(), 7))    # same weights on every rank
    torch.manual_seed(0)
    x_all = torch.randn(4, 256, 16)
    # single-process full-batch gradient (the thing data parallelism must reproduce)
    full = Tiny()
    full.load_state_dict(net.state_dict())
    full(x_all).pow(2).mean().backward()
    want = {k: p.grad.clone() for k, p in full.named_parameters() if p.grad is not None}

    arena = T.FlatArena(list(net.parameters())[::-1])
    red = T.GradReducer(arena, None, bucket_bytes=16 << 10)
    assert red.world == world and len(red.buckets) >= 3
    for it in range(2):                                          # two steps: begin() must re-arm the hooks
        arena.zero_grad()
        red.begin()
        shard = x_all[rank * 2:(rank + 1) * 2]
        net(shard).pow(2).mean().backward()
        fired_in_backward = list(red.launch_order)
        red.finish()
        arena.grad.div_(world)                                   # FlatAdamW folds this scale into its kernel
        err = max(((p.grad - want[k]).norm() / want[k].norm()).item() for k, p in net.named_parameters() if k in want)
        unused_zero = float(net.unused.weight.grad.abs().sum())
    if rank == 0:
        out.put((err, fired_in_backward, list(red.launch_order), len(red.buckets), unused_zero))
    dist.destroy_process_group()


def test_gradient_arena_bucketed_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, early, order, nb, unused = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5                                            # mean of shard gradients == full-batch gradient
    # buckets were reduced WHILE backward was running, in arena order (= reverse execution order); only the bucket
    # holding the never-used parameter (registered last -> arena bucket 0) had to wait for finish()
    assert len(early) == nb - 1 and early == sorted(early) and 0 not in early and order[-1] == 0
    assert sorted(order) == list(range(nb))                      # every bucket reduced exactly once, incl. the unused one
    assert unused == 0.0


def _trainstep_worker(rank, world, port, out):
    """Full TrainStep on 2 ranks (gloo), native entry points replaced by tests/kernel_model.py: each rank trains on its own
    shard; after every step all ranks must hold identical weights, equal to one process training on the whole batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import kernel_model as KM
    import uformer_b200 as U
    from uformer_b200.training import TrainStep
    from paramgen import randomize_state
    cfg = dict(img_size=128, embed_dim=16, depths=[1] * 9, win_size=8, modulator=True, drop_path_rate=0.0)

    def build(seed=9):
        net = U.Uformer(**cfg)
        net.load_state_dict(randomize_state(net.state_dict(), seed), strict=True)
        return net
    torch.manual_seed(0)
    clean = torch.rand(2, 3, 128, 128)
    noisy = (clean + 0.1 * torch.randn_like(clean)).clamp(0, 1)
    with KM.patched():
        net = build(9 + rank)                     # ranks initialise DIFFERENTLY: TrainStep must start everyone from rank 0's weights
        step = TrainStep(net, lr=1e-4, bucket_bytes=1 << 20)
        assert step.world == world and len(step.reducer.buckets) > 2
        for _ in range(2):
            step(noisy[rank:rank + 1], clean[rank:rank + 1])
        flat = step.arena.flat.clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        err = None
        if rank == 0:
            # single-process reference: same two steps on the full batch, no process group involved
            ref = build()
            rs = TrainStep(ref, lr=1e-4, data_parallel=False)
            assert rs.world == 1
            for _ in range(2):
                rs(noisy, clean)
            err = ((rs.arena.flat - flat).norm() / rs.arena.flat.norm()).item()
    if rank == 0:
        out.put((same, err))
    dist.destroy_process_group()


def test_trainstep_data_parallel_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainstep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same                      # rank 0's weights were broadcast at construction and every rank applied the same update
    assert err < 2e-3                # == training on the whole batch (Adam's 1/sqrt(v) amplifies fp32 summation-order noise)

"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: view partition + metric reduce."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from saro_gs_b200.sharding import (allreduce_batch_gradients, mean_metrics, reduce_metrics, render_shard,
                                   shard_indices)


def test_partition_covers_every_view_once():
    for n in (0, 1, 7, 20, 6000):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert seen == list(range(n))


def _worker(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    views = list(range(n_views))
    local = render_shard(views, rank, world, render_fn=lambda v: {"img": torch.full((2, 2), float(v))},
                         metric_fn=lambda v, out: [out["img"].mean().item(), float(v * v)])
    total = reduce_metrics(local, n_metrics=2)
    q.put((rank, total.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n_views, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_reduce_matches_serial():
    n = 11
    res = _run(2, n, 29541)
    want = [sum(range(n)), sum(v * v for v in range(n)), n]
    for _, tot in res:
        assert tot == [float(w) for w in want]
    assert mean_metrics(torch.tensor(res[0][1])) == [want[0] / n, want[1] / n]


def test_rank_without_views_contributes_zeros():
    res = _run(2, 1, 29543)   # rank 1 has nothing to render
    for _, tot in res:
        assert tot == [0.0, 0.0, 1.0]


def _make_params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1000, 3), (1000, 1, 3), (1000, 15, 3), (1000, 1), (128, 41), (128,), (48, 128), (48,)]
    return [torch.nn.Parameter(torch.zeros(s)) for s in shapes], [torch.randn(s, generator=g) for s in shapes]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params, grads = _make_params(100 + rank)               # this rank's view
    for p, g in zip(params, grads):
        p.grad = g.clone()
    if rank == 1:
        params[3].grad = None                               # no gradient on ONE rank only: must contribute zeros
    n = allreduce_batch_gradients(params, batch=world, bucket_bytes=40_000)      # small buckets: several collectives
    q.put((rank, n, [None if p.grad is None else p.grad.clone() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_equals_the_reference_batch_cache():
    """One view per rank + allreduce_batch_gradients == cache_gradient over the batch followed by
    set_batch_gradient (scene/saro_gaussian.py:224-294): sum over views, times 1 / batch."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, 29549, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    views = [_make_params(100 + r)[1] for r in range(world)]
    ratio = 1 / world
    for _, n_coll, got in res:
        assert n_coll > 1
        for i, g in enumerate(got):
            cache = torch.zeros_like(views[0][i])
            for r, v in enumerate(views):
                if not (i == 3 and r == 1):                 # rank 1 had no gradient for parameter 3
                    cache += v[i].clone()                   # cache_gradient
            assert torch.equal(g, cache * ratio)            # set_batch_gradient


def _mismatch_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1000 if rank == 0 else 1001                         # the replicated Gaussians have diverged
    p = torch.nn.Parameter(torch.zeros(n, 3))
    p.grad = torch.ones(n, 3)
    try:
        allreduce_batch_gradients([p], batch=world)
        q.put((rank, "no error"))
    except RuntimeError as e:
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_diverged_layouts_raise_instead_of_summing_misaligned_gradients():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mismatch_worker, args=(r, world, 29551, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, msg in res:
        assert "disagree on the gradient layout" in msg


def test_gradient_scale_without_process_group():
    params, grads = _make_params(7)
    for p, g in zip(params, grads):
        p.grad = g.clone()
    assert allreduce_batch_gradients(params, batch=4) == 0
    for p, g in zip(params, grads):
        assert torch.equal(p.grad, g * (1 / 4))

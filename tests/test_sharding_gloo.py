"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: view partition + metric reduce."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from saro_gs_b200.sharding import mean_metrics, reduce_metrics, render_shard, shard_indices


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

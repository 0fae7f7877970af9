"""Multi-GPU plumbing for the one place this path shards: independent views.

Frames/cameras of one scene are independent units: the Gaussians are replicated
on every rank, rank r renders views r, r+world, ... and the only collective is one all-reduce of a
small metrics vector at the end (NCCL on GPUs, gloo in the CPU tests).  There is deliberately no
data-path collective in rendering.  The reference has no distributed code at all.

Data-parallel TRAINING (SURVEY.md §8(e) "training extension", §8(f) rank 4) maps the reference's batch of views
(train.py:198-226) to one view per rank; its two exchange steps are `allreduce_batch_gradients` below (the
distributed form of cache_gradient / set_batch_gradient, scene/saro_gaussian.py:224-294) and
`saro_gs_b200.densify.BatchDensifyStats.all_reduce`.
"""
from typing import Callable, Iterable, List, Sequence

import torch


def shard_indices(n_items: int, rank: int, world: int) -> range:
    """Round-robin partition: every index in [0, n_items) belongs to exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return range(rank, n_items, world)


def render_shard(views: Sequence, rank: int, world: int, render_fn: Callable, metric_fn: Callable) -> torch.Tensor:
    """Render this rank's share of `views` and return local sums [sum_metric..., count] (float64).

    render_fn(view) -> outputs ; metric_fn(view, outputs) -> 1-D tensor/list of per-view metrics."""
    acc = None
    count = 0
    for i in shard_indices(len(views), rank, world):
        m = torch.as_tensor(metric_fn(views[i], render_fn(views[i])), dtype=torch.float64).flatten().cpu()
        acc = m.clone() if acc is None else acc + m
        count += 1
    if acc is None:
        acc = torch.zeros(0, dtype=torch.float64)
    return torch.cat([acc, torch.tensor([float(count)], dtype=torch.float64)])


def reduce_metrics(local: torch.Tensor, n_metrics: int, device=None) -> torch.Tensor:
    """SUM all-reduce of [metrics..., count] across ranks (no-op when not distributed).
    Ranks with no views contribute zeros of the right length."""
    import torch.distributed as dist
    vec = torch.zeros(n_metrics + 1, dtype=torch.float64)
    if local.numel() == n_metrics + 1:
        vec = local.clone()
    elif local.numel() == 1:
        vec[-1] = local[0]
    else:
        raise ValueError("metric vector length mismatch")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            vec = vec.to(device)
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        vec = vec.cpu()
    return vec


def mean_metrics(total: torch.Tensor) -> List[float]:
    """[sum..., count] -> per-view means."""
    n = max(float(total[-1]), 1.0)
    return [float(v) / n for v in total[:-1]]


def allreduce_batch_gradients(params: Iterable[torch.Tensor], batch: int, group=None, bucket_bytes: int = 64 << 20,
                              check_layout: bool = True) -> int:
    """Distributed form of the reference's batch-gradient cache (scene/saro_gaussian.py:224-294): the reference sums
    every parameter's gradient over the views of a batch (`cache_gradient`, :224-245) and hands the optimizer
    `sum * (1 / batch)` (`set_batch_gradient`, :263-294).  With one view per rank the sum over views is a SUM
    all-reduce; gradients are packed into flat buckets (per dtype, <= bucket_bytes) so that ~25 small tensors and the
    75 MB of per-Gaussian gradients cost a handful of collectives, then scaled by 1 / batch exactly as the reference
    does (a multiplication by the reciprocal, not a division).

    The bucket layout is derived from the PARAMETERS, never from which of them happen to hold a gradient on this rank:
    a parameter that requires a gradient but received none (its view saw nothing of it) contributes zeros, so every
    rank builds the same buckets.  With `check_layout` one extra 2-element MAX all-reduce verifies that all ranks
    agree on the number of gradient elements (replicated Gaussians can drift apart if ranks densify differently) and
    raises instead of summing misaligned gradients.  Returns the number of data collectives issued; a no-op (scale
    only) when not distributed."""
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad or p.grad is not None]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    ratio = 1 / batch
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    collectives = 0
    if distributed and check_layout and grads:
        n = float(sum(g.numel() for g in grads))
        sig = torch.tensor([n, -n], dtype=torch.float64, device=grads[0].device)
        dist.all_reduce(sig, op=dist.ReduceOp.MAX, group=group)
        if float(sig[0]) != n or float(sig[1]) != -n:
            raise RuntimeError("allreduce_batch_gradients: ranks disagree on the gradient layout "
                               f"({int(n)} elements here, between {int(-float(sig[1]))} and {int(float(sig[0]))} elsewhere): "
                               "the replicated Gaussians have diverged")
    by_type = {}
    for g in grads:
        by_type.setdefault((g.dtype, g.device), []).append(g)
    for (dtype, device), gs in by_type.items():
        bucket, size = [], 0
        buckets = []
        for g in gs:
            nbytes = g.numel() * g.element_size()
            if bucket and size + nbytes > bucket_bytes:
                buckets.append(bucket)
                bucket, size = [], 0
            bucket.append(g)
            size += nbytes
        if bucket:
            buckets.append(bucket)
        for bucket in buckets:
            if distributed:
                flat = torch.cat([g.reshape(-1) for g in bucket])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                collectives += 1
                flat *= ratio
                offset = 0
                for g in bucket:
                    g.copy_(flat[offset:offset + g.numel()].view_as(g))
                    offset += g.numel()
            else:
                for g in bucket:
                    g *= ratio
    return collectives

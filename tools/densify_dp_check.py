#!/usr/bin/env python
"""Developer tool (GPU box, 2+ GPUs): the data-parallel form of the densification statistics over NCCL.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/densify_dp_check.py
Each rank renders nothing here: it feeds one synthetic view (gradient + radii) to BatchDensifyStats, all-reduces and
commits; every rank must end with exactly what one process gets from the whole batch."""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from saro_gs_b200.densify import BatchDensifyStats

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = 300_000
g = torch.Generator().manual_seed(123)
grads = [torch.randn(P, 3, generator=g) * 1e-4 for _ in range(world)]
radii = [torch.where(torch.rand(P, generator=g) < 0.4, 0, torch.randint(1, 60, (P,), generator=g)).to(torch.int32) for _ in range(world)]


def model():
    return types.SimpleNamespace(max_radii2D=torch.zeros(P, device=dev), xyz_gradient_accum=torch.zeros(P, 1, device=dev),
                                 denom=torch.zeros(P, 1, device=dev))


dp, serial = model(), model()
stats = BatchDensifyStats(P, dev)
stats.add_view(grads[rank].to(dev), radii[rank].to(dev))          # this rank's view only
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier()
e0.record()
stats.all_reduce()
e1.record()
stats.commit(dp)
whole = BatchDensifyStats(P, dev)                                   # the whole batch on one rank
for a, b in zip(grads, radii):
    whole.add_view(a.to(dev), b.to(dev))
whole.commit(serial)
torch.cuda.synchronize()
ok = torch.equal(dp.max_radii2D, serial.max_radii2D) and torch.equal(dp.denom, serial.denom) and \
    torch.allclose(dp.xyz_gradient_accum, serial.xyz_gradient_accum, rtol=1e-6, atol=0)
flags = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"what": "densification statistics, one view per rank, NCCL all-reduce (SUM, SUM, MAX) then commit",
                      "world": world, "P": P, "all_ranks_equal_single_process_batch": bool(flags.item()),
                      "all_reduce_ms": e0.elapsed_time(e1)}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flags.item() else 1)

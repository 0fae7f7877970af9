"""CPU: the numpy restatement of the per-iteration densification statistics against the result of executing the
reference's own statements (tests/golden/densify_*.npz), and the data-parallel exchange step (gloo, world_size 2)."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import densify_oracle
from saro_gs_b200.densify import reduce_running_buffers

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "densify_*.npz")))


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_statements(path):
    z = np.load(path)
    mr, acc, den = densify_oracle.batch_statistics(list(z["grads"]), list(z["radii"]), z["start_max_radii2D"],
                                                   z["start_xyz_gradient_accum"], z["start_denom"], dtype=np.float32)
    np.testing.assert_array_equal(mr, z["out_max_radii2D"])
    np.testing.assert_array_equal(den, z["out_denom"])
    np.testing.assert_allclose(acc, z["out_xyz_gradient_accum"], rtol=2e-6, atol=0)
    never = (z["radii"] > 0).sum(0) == 0
    assert never.any()
    np.testing.assert_array_equal(acc[never], z["start_xyz_gradient_accum"][never])      # untouched where never visible


def _worker(rank, world, port, path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(path)
    g, r = z["grads"][rank], z["radii"][rank]                      # one view per rank
    grad_sum = torch.from_numpy(np.sqrt((g[:, :2] ** 2).sum(-1)))
    vis = torch.from_numpy((r > 0).astype(np.int32))
    rmax = torch.from_numpy(np.maximum(r, 0).astype(np.int32))
    reduce_running_buffers(grad_sum, vis, rmax)
    q.put((rank, grad_sum.numpy(), vis.numpy(), rmax.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_equals_batch_of_two():
    path = [p for p in GOLDEN if p.endswith("densify_batch2_sparse.npz")][0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29547, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = np.load(path)
    for _, grad_sum, vis, rmax in res:
        n = vis
        seen = n > 0
        mr = z["start_max_radii2D"].copy()
        mr[seen] = np.maximum(mr[seen], rmax[seen])
        acc = z["start_xyz_gradient_accum"].copy().reshape(-1)
        acc[seen] += grad_sum[seen] / n[seen]
        np.testing.assert_array_equal(mr, z["out_max_radii2D"])
        np.testing.assert_allclose(acc, z["out_xyz_gradient_accum"].reshape(-1), rtol=2e-6)

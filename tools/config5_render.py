#!/usr/bin/env python
"""BASELINE.json configs[4] at its stated size: 300 frames x 20 test views = 6 000 renders of the configs[1] cloud
(alive set changing per frame, synthetic.DeviceSequence; cameras on the arc of SURVEY.md section 8(d) config 5),
sharded one view per rank through saro_gs_b200.sharding.render_shard and reduced with sharding.reduce_metrics
(one NCCL all-reduce of a small float64 vector at the end — the path's only collective).

Metrics per render (all deterministic: the forward pass is bit-reproducible): mean colour, mean depth, num visible
Gaussians, a PSNR-like score against the frame's first-camera image statistics.  The reduced sums must not depend on
the number of ranks: run at N = 1 and N = 8 and compare `metric_sums` (tools/gpu_config5.sh does, and stores both lines
under profiles/).

    python tools/config5_render.py                                          # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \\
        tools/config5_render.py
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--views", type=int, default=20)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic, sharding

    scene, cam0 = synthetic.config2_scene()
    H, W = cam0.height, cam0.width
    seq = synthetic.DeviceSequence(scene, dev)
    yaws = [(-0.3 + 0.6 * k / (args.views - 1)) if args.views > 1 else 0.0 for k in range(args.views)]
    cams = [synthetic.yaw_camera(W, H, 729.0, yaw=y, pivot=(0.0, 0.0, 10.0)) for y in yaws]
    bg = torch.zeros(3, device=dev)
    rasts = [sgs.GaussianRasterizer(sgs.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0,
                                                                       c.viewmatrix.to(dev), c.projmatrix.to(dev), 3,
                                                                       c.campos.to(dev), False)) for c in cams]
    views = [(f, v) for f in range(args.frames) for v in range(args.views)]      # 6 000 units of work
    cache = {"frame": -1, "scene": None}

    def render_fn(view):
        f, v = view
        if cache["frame"] != f:                      # consecutive units of a rank mostly share the frame
            cache["frame"], cache["scene"] = f, seq.frame(f / args.frames)
        sc = cache["scene"]
        with torch.no_grad():
            return rasts[v](means3D=sc.means3D, means2D=torch.zeros_like(sc.means3D), opacities=sc.opacities, shs=sc.shs,
                            scales=sc.scales, rotations=sc.rotations)

    def metric_fn(view, out):
        color, radii, depth = out
        mse = ((color - 0.25) ** 2).mean()
        psnr_like = -10.0 * torch.log10(mse + 1e-12)
        return torch.stack([color.double().mean(), depth.double().mean(), (radii > 0).sum().double(), psnr_like.double()])

    for v in views[rank:rank + 3 * world:world]:     # warm-up
        render_fn(v)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    local_sums = sharding.render_shard(views, rank, world, render_fn, metric_fn)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    total = sharding.reduce_metrics(local_sums, 4, device=dev)
    if rank == 0:
        print(json.dumps({"what": "configs[4]: %d frames x %d views = %d renders @%dx%d, one view per rank" %
                                  (args.frames, args.views, len(views), W, H),
                          "n_gpus": world, "renders": int(total[-1]), "wall_s": wall, "renders_per_s": len(views) / wall,
                          "metric_sums": [float(x) for x in total[:-1]],
                          "metric_means": sharding.mean_metrics(total),
                          "metrics": ["mean colour", "mean depth", "visible Gaussians", "PSNR-like score"]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

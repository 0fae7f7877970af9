#!/usr/bin/env python
"""Data-parallel training iterations on N GPUs (SURVEY.md section 8(e) "training extension", 8(f) rank 4; VERDICT round 1
item 7): the reference's batch of views (train.py:198-226, opt.batch views per optimizer step) mapped to one process per
GPU.  Global batch = 8 views of the configs[1] cloud (300 k Gaussians @1352x1014); rank r renders views r, r + N, ...;
per view: rasterizer forward -> fused L1 + D-SSIM loss -> backward -> densification statistics (sgs_densify_add_view).
After the last local view the two exchange steps run over NCCL / NVLink:
    gradients : saro_gs_b200.sharding-style bucketed SUM all-reduce, launched bucket by bucket as async collectives;
                the fused Adam step of bucket k runs while bucket k + 1 is still in flight (the overlap available when
                every rank holds one view: there is no "next view's forward" inside the iteration);
    statistics: BatchDensifyStats.all_reduce (SUM, SUM, MAX) + commit, also while the gradient buckets fly.
Prints one JSON line (rank 0): iterations/s, ms per iteration, the render / exchange / optimizer shares, and the
exposed (non-overlapped) collective time.  All times are CUDA events on the launching stream, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \\
        tools/dp_train.py --iters 20
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-overlap", action="store_true", help="one blocking all-reduce, then Adam (the baseline to beat)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic, loss_utils
    from saro_gs_b200.densify import BatchDensifyStats

    assert args.batch % world == 0, "the global batch must divide over the ranks"
    scene, cam0 = synthetic.config2_scene()
    H, W = cam0.height, cam0.width
    cams = [synthetic.yaw_camera(W, H, 729.0, yaw=0.01 * (k - args.batch / 2), pivot=(0.0, 0.0, 22.0))
            for k in range(args.batch)]
    mine = list(range(rank, args.batch, world))
    bg = torch.zeros(3, device=dev)
    P = scene.means3D.shape[0]
    params = {"means3D": scene.means3D, "log_scales": scene.scales.log(), "rotations": scene.rotations,
              "opacity_logit": torch.logit(scene.opacities.clamp(1e-4, 1 - 1e-4)), "shs": scene.shs}
    params = {k: v.to(dev).clone().requires_grad_(True) for k, v in params.items()}
    lrs = {"means3D": 1.6e-4, "log_scales": 5e-3, "rotations": 1e-3, "opacity_logit": 5e-2, "shs": 2.5e-3}
    # one fused Adam per bucket so that a bucket's step can run as soon as its all-reduce has landed
    buckets = [["means3D", "log_scales", "rotations", "opacity_logit"], ["shs"]]
    opts = [torch.optim.Adam([{"params": [params[n]], "lr": lrs[n]} for n in b], eps=1e-15, fused=True) for b in buckets]
    settings = [sgs.GaussianRasterizationSettings(H, W, c.tanfovx, c.tanfovy, bg, 1.0, c.viewmatrix.to(dev),
                                                  c.projmatrix.to(dev), 3, c.campos.to(dev), False) for c in cams]

    def render(k):
        m = params["means3D"]
        m2d = torch.zeros_like(m, requires_grad=True)
        color, radii, depth = sgs.GaussianRasterizer(settings[k])(
            means3D=m, means2D=m2d, opacities=torch.sigmoid(params["opacity_logit"]), shs=params["shs"],
            scales=torch.exp(params["log_scales"]), rotations=torch.nn.functional.normalize(params["rotations"]))
        return color, radii, m2d

    with torch.no_grad():
        hidden = {k: (v + 0.01 * torch.randn_like(v)) for k, v in params.items()}
        saved = {k: v.detach().clone() for k, v in params.items()}
        for k, v in hidden.items():
            params[k].data.copy_(v)
        targets = {k: render(k)[0].detach().clone() for k in mine}
        for k, v in saved.items():
            params[k].data.copy_(v)

    stats = BatchDensifyStats(P, dev)
    model = type("M", (), {})()
    model.max_radii2D = torch.zeros(P, device=dev)
    model.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
    model.denom = torch.zeros(P, 1, device=dev)
    ratio = 1.0 / args.batch
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def iteration(timing):
        e = [ev() for _ in range(4)]
        e[0].record()
        stats.reset()
        for p in params.values():
            p.grad = None
        for k in mine:
            color, radii, m2d = render(k)
            loss = loss_utils.l1_dssim_loss(color, targets[k], 0.2)
            stats.attach_next_backward()                      # the view's statistics ride in the backward's last kernel
            loss.backward()                                   # gradients of the local views accumulate in .grad
        e[1].record()
        flats = []
        for b in buckets:
            flats.append(torch.cat([params[n].grad.reshape(-1) for n in b]))
        if world > 1 and args.no_overlap:
            for f in flats:
                dist.all_reduce(f)
            works = [None] * len(flats)
        elif world > 1:
            works = [dist.all_reduce(f, async_op=True) for f in flats]
        else:
            works = [None] * len(flats)
        stats.all_reduce()
        stats.commit(model)
        e[2].record()
        for b, f, w, opt in zip(buckets, flats, works, opts):
            if w is not None:
                w.wait()                                      # stream-level wait: the host does not block
            f.mul_(ratio)                                     # set_batch_gradient: sum * (1 / batch)
            off = 0
            for n in b:
                g = params[n].grad
                g.copy_(f[off:off + g.numel()].view_as(g))
                off += g.numel()
            opt.step()
        e[3].record()
        if timing is not None:
            timing.append(e)

    for _ in range(args.warmup):
        iteration(None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    timing = []
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.iters):
        iteration(timing)
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1)
    render_ms = sum(e[0].elapsed_time(e[1]) for e in timing) / len(timing)
    exch_ms = sum(e[1].elapsed_time(e[2]) for e in timing) / len(timing)
    opt_ms = sum(e[2].elapsed_time(e[3]) for e in timing) / len(timing)
    vec = torch.tensor([total, render_ms, exch_ms, opt_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
    if rank == 0:
        total, render_ms, exch_ms, opt_ms = (float(v) for v in vec)
        ms_it = total / args.iters
        print(json.dumps({
            "what": "data-parallel training iterations, global batch %d views of the configs[1] cloud "
                    "(300k Gaussians @1352x1014), one process per GPU" % args.batch,
            "n_gpus": world, "views_per_rank": len(mine), "iters": args.iters, "overlap": not args.no_overlap,
            "ms_per_iteration": ms_it, "iterations_per_s": 1e3 / ms_it, "views_per_s": 1e3 * args.batch / ms_it,
            "render_fwd_loss_bwd_ms": render_ms,
            "pack_and_stats_exchange_ms": exch_ms,
            "gradient_collective_wait_plus_adam_ms": opt_ms,
            "gradient_bytes_per_rank": int(sum(p.numel() for p in params.values()) * 4),
            "collective_share_of_iteration": (exch_ms + opt_ms) / ms_it if world > 1 else 0.0,
            "note": "gradient buckets are all-reduced asynchronously; the fused Adam step of bucket k overlaps the "
                    "all-reduce of bucket k + 1 and the densification-statistics exchange"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

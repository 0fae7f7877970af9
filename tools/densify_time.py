#!/usr/bin/env python
"""Developer tool (GPU box): densification statistics of a 4-view batch on 300k Gaussians, fused kernels vs the
reference's list / stack / boolean-index PyTorch ops (train.py:211-215, :281-291)."""
import json
import sys
import types
import torch
sys.path.insert(0, '.')
from saro_gs_b200.densify import BatchDensifyStats

dev = torch.device('cuda:0')
P, V = 300_000, 4
g = torch.Generator().manual_seed(0)
grads = [(torch.randn(P, 3, generator=g) * 1e-4).to(dev) for _ in range(V)]
radii = [torch.where(torch.rand(P, generator=g) < 0.3, 0, torch.randint(1, 60, (P,), generator=g)).to(torch.int32).to(dev) for _ in range(V)]


def model():
    return types.SimpleNamespace(max_radii2D=torch.zeros(P, device=dev), xyz_gradient_accum=torch.zeros(P, 1, device=dev),
                                 denom=torch.zeros(P, 1, device=dev))


def fused(m, stats):
    stats.reset()
    for a, b in zip(grads, radii):
        stats.add_view(a, b)
    stats.commit(m)


def torch_ops(m, _):
    batch_point_grad, batch_radii, batch_vis = [], [], []
    for a, b in zip(grads, radii):
        batch_point_grad.append(torch.norm(a[:, :2], dim=-1))
        batch_radii.append(b)
        batch_vis.append(b > 0)
    visibility_count = torch.stack(batch_vis, 1).sum(1)
    visibility_filter = visibility_count > 0
    r = torch.stack(batch_radii, 1).max(1)[0]
    gr = torch.stack(batch_point_grad, 1).sum(1)
    gr[visibility_filter] = gr[visibility_filter] / visibility_count[visibility_filter]
    gr = gr.unsqueeze(1)
    m.max_radii2D[visibility_filter] = torch.max(m.max_radii2D[visibility_filter], r[visibility_filter])
    m.xyz_gradient_accum[visibility_filter] += gr[visibility_filter]
    m.denom[visibility_filter] += 1


def run(fn):
    m, stats, ms = model(), BatchDensifyStats(P, dev), []
    for i in range(23):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(m, stats)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    return sum(ms) / len(ms), m


a, ma = run(fused)
b, mb = run(torch_ops)
ok = torch.equal(ma.max_radii2D, mb.max_radii2D) and torch.equal(ma.denom, mb.denom) and \
    torch.allclose(ma.xyz_gradient_accum, mb.xyz_gradient_accum, rtol=1e-5)
print(json.dumps({"what": "densification statistics, 4 views x 300k Gaussians", "fused_ms": a, "pytorch_ops_ms": b, "speedup": b / a,
                  "fused_GBps": (V * 32 + 32) * P / (a * 1e-3) / 1e9, "agree": bool(ok)}))

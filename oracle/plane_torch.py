"""TEST / BENCH INFRASTRUCTURE — the scale-aware plane field as plain PyTorch ops on the GPU: what SaRO-GS executes
per call (scene/hexplane.py:26-60, 91-137, 231-274: permute to channels-last, mip construction, sampling, sum over
planes), with torch.nn.functional.grid_sample / avg_pool2d standing in for the un-vendored nvdiffrast op (same stand-in
as tests/golden/make_golden_plane.py).  Used by bench.py's `plane_path` leg as the PyTorch baseline and as an
on-device cross-check; never imported by the product."""
import itertools

import torch
import torch.nn.functional as F

COO_COMBS = list(itertools.combinations(range(4), 2))


def texture(t, uv, bias, max_mip_level):
    """t [1, C, H, W], uv [N, 2] in [0, 1], bias [N] -> [N, C]"""
    mips = [t]
    while (mips[-1].shape[2] > 1 or mips[-1].shape[3] > 1) and len(mips) - 1 < max_mip_level:
        h, w = mips[-1].shape[2:]
        mips.append(F.avg_pool2d(mips[-1], (2 if h > 1 else 1, 2 if w > 1 else 1)))
    top = len(mips) - 1
    level = bias.clamp(0.0, float(top))
    l0 = level.floor()
    f = level - l0
    l1 = (l0 + 1).clamp(max=float(top))
    grid = (uv * 2.0 - 1.0)[None, None]
    out = 0.0
    for lv, m in enumerate(mips):
        s = F.grid_sample(m, grid, mode="bilinear", padding_mode="border", align_corners=False)[0, :, 0].t()
        w = (l0 == lv).to(s.dtype) * (1 - f) + ((l1 == lv) & (level > l0)).to(s.dtype) * f
        out = out + w[:, None] * s
    return out


def field_forward(field, pts, timestamps, scales):
    """Same arithmetic as saro_gs_b200.hexplane.ScaleAwareResField.forward, in PyTorch ops (differentiable in the planes)."""
    p = (pts - field.aabb[0]) / (field.aabb[1] - field.aabb[0])
    dur = float(field.duration.item())
    t = timestamps.reshape(-1, 1) * dur / (dur - 1)
    p4 = torch.cat([p, t], dim=1)
    level = field.get_level(scales)
    outs = []
    for planes in field.grids:
        acc = 0.0
        for ci, comb in enumerate(COO_COMBS):
            bias = level[:, list(comb)].min(dim=1).values
            acc = acc + texture(planes[ci], p4[:, list(comb)], bias, 0 if 3 in comb else 7)
        outs.append(acc)
    return torch.cat(outs, dim=1)

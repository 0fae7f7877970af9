#!/usr/bin/env python
"""Generates tests/golden/plane_*.npz by running the reference's OWN scene/hexplane.py (ScaleAwareResField and
everything it calls: normalize_aabb, normalize_time, get_level, grid_sample_wrapper, interpolate_ms_features,
init_grid_param) on the CPU in float64 and float32, with autograd gradients of the planes.

The module cannot be imported as it is: it imports `nvdiffrast.torch`, an un-vendored third-party CUDA package that is
neither in /root/reference nor installed.  So (1) the source is parsed with `ast` at generation time and executed as
it is except for `device="cuda"` -> `device="cpu"` in set_aabb (a device literal, no arithmetic) — nothing is copied
into this repository — and (2) `nvdiffrast.torch.texture` is provided by the stand-in below, a plain torch composition
(avg_pool2d mip stack + grid_sample(align_corners=False, padding_mode="border") + level blend) of the library's
PUBLISHED algorithm (see oracle/plane_oracle.py for the statement and the citation).  The parity pin of row f2 is
therefore: the reference's own Python around an independently written implementation of the published op.

    python tests/golden/make_golden_plane.py          # build container only (needs /root/reference)
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/scene/hexplane.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def texture_standin(tex, uv, mip_level_bias=None, boundary_mode="wrap", max_mip_level=None, filter_mode="auto"):
    """tex [B, H, W, C], uv [B, 1, N, 2], mip_level_bias [B, 1, N] -> [B, 1, N, C]; linear-mipmap-linear, clamp."""
    assert boundary_mode == "clamp" and filter_mode == "auto" and mip_level_bias is not None
    t = tex.permute(0, 3, 1, 2)
    mips = [t]
    while (mips[-1].shape[2] > 1 or mips[-1].shape[3] > 1) and len(mips) - 1 < max_mip_level:
        h, w = mips[-1].shape[2:]
        assert not ((h > 1 and h & 1) or (w > 1 and w & 1))
        mips.append(F.avg_pool2d(mips[-1], (2 if h > 1 else 1, 2 if w > 1 else 1)))
    top = len(mips) - 1
    level = mip_level_bias.clamp(0.0, float(top))                 # [B, 1, N]
    l0 = level.floor()
    f = (level - l0).unsqueeze(-1)                                # [B, 1, N, 1]
    l1 = (l0 + 1).clamp(max=float(top))
    grid = uv * 2.0 - 1.0
    out = 0.0
    for lv, m in enumerate(mips):
        s = F.grid_sample(m, grid, mode="bilinear", padding_mode="border", align_corners=False)   # [B, C, 1, N]
        s = s.permute(0, 2, 3, 1)
        w0 = (l0 == lv).unsqueeze(-1).to(s.dtype) * (1 - f)
        w1 = ((l1 == lv) & (level > l0)).unsqueeze(-1).to(s.dtype) * f
        out = out + (w0 + w1) * s
    return out


def load_reference_module():
    src = open(REF).read()
    tree = ast.parse(src)
    body = [n for n in tree.body if not (isinstance(n, ast.Import) and any(a.name.startswith("nvdiffrast") for a in n.names))]

    class CpuDevice(ast.NodeTransformer):
        def visit_Constant(self, node):
            return ast.copy_location(ast.Constant("cpu"), node) if node.value == "cuda" else node

    mod = CpuDevice().visit(ast.Module(body=body, type_ignores=[]))
    ast.fix_missing_locations(mod)
    nvd = types.ModuleType("nvdiffrast")
    nvd.torch = types.ModuleType("nvdiffrast.torch")
    nvd.torch.texture = texture_standin
    ns = {"nvdiffrast": nvd, "__name__": "reference_hexplane"}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def make_case(name, reso, out_dim, multires, n, seed, duration=50, spread=1.3):
    ns = load_reference_module()
    g = torch.Generator().manual_seed(seed)
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": out_dim, "resolution": list(reso)}
    xyz_max, xyz_min = [1.0, 0.8, 1.2], [-1.1, -0.9, -0.7]
    ext = torch.tensor(xyz_max) - torch.tensor(xyz_min)
    centre = 0.5 * (torch.tensor(xyz_max) + torch.tensor(xyz_min))
    # positions spill over the box (clamp mode), scales span below the finest and above the coarsest level
    pts = centre + (torch.rand(n, 3, generator=g) - 0.5) * ext * spread
    t = torch.rand(n, 1, generator=g) * (duration - 1) / duration
    scales = torch.exp(torch.rand(n, 3, generator=g) * 9.0 - 7.5)
    out = dict(pts=pts.numpy(), timestamps=t.numpy(), scales=scales.numpy(), aabb=np.array([xyz_max, xyz_min], np.float32),
               duration=np.int64(duration), reso=np.array(reso), multires=np.array(multires), out_dim=np.int64(out_dim))
    grids_f64 = None
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        torch.set_default_dtype(dt)
        try:
            field = ns["ScaleAwareResField"](cfg, list(multires))
            field.set_aabb(xyz_max, xyz_min, duration)
            with torch.no_grad():
                if grids_f64 is None:
                    gi = torch.Generator().manual_seed(seed + 1)
                    grids_f64 = [[(torch.randn(p.shape, generator=gi, dtype=torch.float64) * 0.5).float().double() for p in level]
                                 for level in field.grids]   # float32-representable values
                for level, vals in zip(field.grids, grids_f64):
                    for p, v in zip(level, vals):
                        p.copy_(v.to(dt))
            field.aabb = field.aabb.to(dt)
            field.base_scale = field.base_scale.to(dt)
            feats = field(pts.to(dt), t.to(dt), scales.to(dt))
            gd = torch.Generator().manual_seed(seed + 2)
            dout = torch.randn(feats.shape, generator=gd, dtype=torch.float64)
            feats.backward(dout.to(dt))
        finally:
            torch.set_default_dtype(torch.float32)
        out[f"features_{tag}"] = feats.detach().numpy()
        out["base_scale"] = field.base_scale.detach().numpy().astype(np.float32)
        out["dout"] = dout.numpy()
        for li, level in enumerate(field.grids):
            for ci, p in enumerate(level):
                if tag == "f64":
                    out[f"dgrid_{tag}_{li}_{ci}"] = p.grad.numpy().astype(np.float32)   # float64 run, stored as float32
                    out[f"grid_{li}_{ci}"] = grids_f64[li][ci].numpy().astype(np.float32)
    path = os.path.join(HERE, f"plane_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; features", out["features_f64"].shape)


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("needs /root/reference")
    make_case("small", reso=(16, 16, 16, 8), out_dim=4, multires=(1, 2), n=600, seed=0)
    make_case("ragged", reso=(32, 16, 8, 6), out_dim=8, multires=(1,), n=257, seed=1)
    make_case("wide", reso=(32, 32, 32, 10), out_dim=32, multires=(1,), n=400, seed=2)

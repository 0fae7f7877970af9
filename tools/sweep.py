#!/usr/bin/env python
"""Developer tool (GPU box): forward+backward ms/frame of the native library and of the compiled reference over a
few workload sizes (same synthetic generator as configs[1], different P / resolution / splat size)."""
import json
import sys
import torch
sys.path.insert(0, '.')
import saro_gs_b200 as sgs
from saro_gs_b200 import synthetic
from oracle import ref_loader

dev = torch.device('cuda:0')
Ref = ref_loader.ref_api()[1] if ref_loader.available() else None
CASES = [
    ("P=100k 800x800", dict(P=100_000, width=800, height=800, fx=700.0)),
    ("P=300k 1352x1014 (headline)", dict()),
    ("P=300k 1352x1014 big splats", dict(log_scale_mean=-2.4)),
    ("P=1M 1352x1014", dict(P=1_000_000, log_scale_mean=-3.4)),
    ("P=300k 1920x1080", dict(width=1920, height=1080, fx=1040.0)),
    ("P=2M 1920x1080", dict(P=2_000_000, width=1920, height=1080, fx=1040.0, log_scale_mean=-3.6)),
]
rows = []
for name, kw in CASES:
    scene, cam = synthetic.config2_scene(**kw)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    params = {k: getattr(scene, k).to(dev).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(params["means3D"], requires_grad=True)
    cot = synthetic.cotangent(cam.height, cam.width).to(dev)
    res = {}
    from saro_gs_b200 import _lib
    for tag, Rast in (("native", sgs.GaussianRasterizer), ("native_sort_in_supertile", sgs.GaussianRasterizer),
                      ("native_global_depth_sort", sgs.GaussianRasterizer), ("reference", Ref)):
        if Rast is None:
            continue
        # where the binning stage sorts by depth: automatic choice (what users get) and both forced modes
        _lib.load().sgs_debug_set_binning_mode({"native_sort_in_supertile": 1, "native_global_depth_sort": 0}.get(tag, -1))
        ts = []
        for i in range(13):
            for p in list(params.values()) + [m2d]:
                p.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            color, radii, depth = Rast(rs)(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"],
                                           shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
            color.backward(cot)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        res[tag] = sorted(ts)[len(ts) // 2]
        res["visible"] = int((radii > 0).sum())
    e = torch.Tensor([])
    R = sgs._C.rasterize_gaussians(rs.bg, params["means3D"].detach(), e, params["opacities"].detach(), params["scales"].detach(),
                                   params["rotations"].detach(), 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                                   cam.height, cam.width, params["shs"].detach(), 3, rs.campos, False)[0]
    row = dict(case=name, num_rendered=R, **res)
    if "reference" in res:
        row["speedup"] = res["reference"] / res["native"]
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open("gpurun_out/sweep.json", "w"), indent=1)

#!/usr/bin/env python
"""Developer tool (GPU box): full-size (configs[1]) gradients of the native kernels and of the compiled reference
against the float64 CPU oracle run live on the box: max |diff| / max |f64| and norm-relative error per tensor,
two runs each.  Output: one JSON line (also written to gpurun_out/grad_vs_f64.json)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import saro_gs_b200 as sgs
from saro_gs_b200 import synthetic
from oracle import oracle, ref_loader

dev = torch.device('cuda:0')
scene, cam = synthetic.config2_scene()
rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                       cam.campos.to(dev), False)
cot_cpu = synthetic.cotangent(cam.height, cam.width)
cot = cot_cpu.to(dev)


def grads(Rast):
    leaves = {k: getattr(scene, k).to(dev).clone().requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = Rast(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                   shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(cot)
    g = {k: v.grad.double().cpu().numpy() for k, v in leaves.items()}
    g["means2D"] = m2d.grad.double().cpu().numpy()
    return g


oracle.build()
ref64 = oracle.forward_scene(scene, cam, torch.zeros(3), precision="f64").backward(cot_cpu)
out = {}
arms = {"native": sgs.GaussianRasterizer}
if ref_loader.available():
    arms["reference"] = ref_loader.ref_api()[1]
for name, R in arms.items():
    for run in range(2):
        g = grads(R)
        for k, a in g.items():
            b = np.asarray(ref64[k], dtype=np.float64).reshape(a.shape)
            mx = float(np.abs(a - b).max() / np.abs(b).max())
            nr = float(np.linalg.norm(a - b) / np.linalg.norm(b))
            out.setdefault(name, {}).setdefault(k, []).append({"max": mx, "norm": nr})
for name in out:
    for k in out[name]:
        print(f"{name:10s} {k:10s} " + " | ".join(f"max {r['max']:.2e} nrm {r['norm']:.2e}" for r in out[name][k]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/grad_vs_f64.json", "w"), indent=1)

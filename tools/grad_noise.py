#!/usr/bin/env python
"""Developer tool (GPU box): run-to-run noise of the full-size (configs[1]) gradients — reference vs itself,
native vs itself, native vs reference — as max |diff| / max |ref| and norm-relative error per tensor."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import saro_gs_b200 as sgs
from saro_gs_b200 import synthetic
from oracle import ref_loader

dev = torch.device('cuda:0')
scene, cam = synthetic.config2_scene()
rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                       cam.campos.to(dev), False)
cot = synthetic.cotangent(cam.height, cam.width).to(dev)


def grads(Rast):
    leaves = {k: getattr(scene, k).to(dev).clone().requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = Rast(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                   shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(cot)
    g = {k: v.grad.double() for k, v in leaves.items()}
    g["means2D"] = m2d.grad.double()
    return g


Ref = ref_loader.ref_api()[1]
r1, r2, n1, n2 = grads(Ref), grads(Ref), grads(sgs.GaussianRasterizer), grads(sgs.GaussianRasterizer)
for k in r1:
    sc = r1[k].abs().max().item()
    f = lambda a, b: ((a - b).abs().max().item() / sc, ((a - b).norm() / b.norm()).item())
    print(f"{k:10s} ref-ref max {f(r2[k], r1[k])[0]:.2e} nrm {f(r2[k], r1[k])[1]:.2e} | nat-nat max {f(n2[k], n1[k])[0]:.2e} "
          f"nrm {f(n2[k], n1[k])[1]:.2e} | nat-ref max {f(n1[k], r1[k])[0]:.2e} nrm {f(n1[k], r1[k])[1]:.2e}")

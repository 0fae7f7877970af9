#!/usr/bin/env python
"""Developer tool (GPU box): sizes of the supertile buckets (distinct Gaussians per 4x4-tile supertile) at configs[1],
i.e. what one block of tile_fill_sorted_kernel sorts in shared memory (capacity SGS_FS_CAP = 4096 entries)."""
import json
import sys

import torch

sys.path.insert(0, '.')
import saro_gs_b200 as sgs
from saro_gs_b200 import synthetic

dev = torch.device('cuda:0')
scene, cam = synthetic.config2_scene()
e = torch.Tensor([])
args = (torch.zeros(3, device=dev), scene.means3D.to(dev), e, scene.opacities.to(dev), scene.scales.to(dev),
        scene.rotations.to(dev), 1.0, e, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.tanfovx, cam.tanfovy,
        cam.height, cam.width, scene.shs.to(dev), scene.sh_degree, cam.campos.to(dev), False)
R, color, radii, gb, bb, ib, depth = sgs._C.rasterize_gaussians(*args)
st = sgs._C.debug_export(scene.means3D.shape[0], cam.width, cam.height, R, gb, bb, ib)
tx, ty = (cam.width + 15) // 16, (cam.height + 15) // 16
rng = st["ranges"].long()
cnt = rng[:, 1] - rng[:, 0]
tile_of = torch.repeat_interleave(torch.arange(tx * ty, device=dev), cnt)
sx = (tx + 3) // 4
sup = (tile_of // tx // 4) * sx + (tile_of % tx) // 4
pair = sup * (1 << 32) + st["point_list"].long()
usup = torch.unique(pair) >> 32
sizes = torch.bincount(usup, minlength=sx * ((ty + 3) // 4)).float()
print(json.dumps({"supertiles": int(sizes.numel()), "mean": sizes.mean().item(), "median": sizes.median().item(),
                  "p90": sizes.quantile(0.9).item(), "p99": sizes.quantile(0.99).item(), "max": sizes.max().item(),
                  "over_4096": int((sizes > 4096).sum()), "tile_list_mean": cnt.float().mean().item(),
                  "tile_list_max": cnt.max().item()}))

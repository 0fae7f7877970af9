#!/usr/bin/env python
"""Developer tool (GPU box): phase timestamps of the two persistent binning kernels at configs[1]
(sgs_debug_binning_profile) — where the microseconds go between the grid-wide barriers."""
import ctypes
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import saro_gs_b200 as sgs
from saro_gs_b200 import synthetic, _lib

dev = torch.device('cuda:0')
scene, cam = synthetic.config2_scene()
rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                       cam.campos.to(dev), False)
t = {k: getattr(scene, k).to(dev) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
lib = _lib.load()
lib.sgs_debug_binning_profile(1, None)
rast = sgs.GaussianRasterizer(rs)
acc = []
with torch.no_grad():
    for i in range(12):
        rast(means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"]), opacities=t["opacities"], shs=t["shs"],
             scales=t["scales"], rotations=t["rotations"])
        buf = (ctypes.c_uint64 * 128)()
        lib.sgs_debug_binning_profile(1, buf)
        if i >= 2:
            acc.append(np.array(buf[:], dtype=np.int64))
a = np.stack(acc)
for name, lo, w in (("depth_sort_kernel", 0, 64), ("tile_sort_kernel", 64, 32), ("tile_fill_sorted block 0", 96, 16),
                    ("tile_fill_sorted block 200", 112, 16)):
    seg = a[:, lo:lo + w]
    n = int((seg[0] > 0).sum())
    rel = (seg[:, :n] - seg[:, :1]) / 1e3
    print(name, "phase marks (us from kernel start, median over runs):")
    print("  ", np.round(np.median(rel, 0), 1).tolist())
    print("   deltas:", np.round(np.diff(np.median(rel, 0)), 1).tolist())
print("gap depth end -> tile start (us):", float(np.median(a[:, 64] - a[:, :64].max(1))) / 1e3)
print("fill block 0 start - coarse end (us):", float(np.median(a[:, 96] - a[:, 64:96].max(1))) / 1e3,
      " block 200 start - block 0 start (us):", float(np.median(a[:, 112] - a[:, 96])) / 1e3)

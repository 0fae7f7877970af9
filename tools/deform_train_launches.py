#!/usr/bin/env python
"""Developer tool (GPU box, run under ncu): ONE native get_deformation forward + backward at 300k Gaussians after two
warm-up iterations, bracketed by cudaProfilerStart/Stop so that `ncu --profile-from-start off` lists only its kernels.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/deform_train_launches.py
tools/deform_train_launches.py --summarise out.csv   prints the per-kernel totals."""
import csv
import sys

sys.path.insert(0, ".")

if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10 and r[0].isdigit()]
    tot, agg, order = 0.0, {}, []
    for r in rows:
        name = r[4].split("(")[0][-60:]
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += float(r[14]) / 1e3
        tot += float(r[14]) / 1e3
    for name in sorted(order, key=lambda k: -agg[k][1]):
        print(f"{agg[name][1]:9.1f} us  x{agg[name][0]:<3d} {name}")
    print(f"{tot:9.1f} us  total, {len(rows)} launches")
    sys.exit(0)

import torch
from oracle import deform_torch
from saro_gs_b200 import deformation

dev = torch.device("cuda:0")
N = 300_000
g = torch.Generator().manual_seed(5)
rn = lambda *s: torch.randn(*s, generator=g)
t = dict(xyz=rn(N, 3) * 2, rotation=rn(N, 4), scaling=rn(N, 3) * 0.5 - 3.5, opacity=rn(N, 1) * 2, features_dc=rn(N, 1, 3) * 0.5,
         features_rest=rn(N, 15, 3) * 0.1, temporal_pos=torch.rand(N, 1, generator=g), hexplane_feature=rn(N, 32) * 0.5)
leaves = {k: v.to(dev).requires_grad_(True) for k, v in t.items()}
mlps = deform_torch.make_train_mlps(32, device=dev, seed=6)
pc = deform_torch.TrainModelStandIn(leaves, mlps, (1, 0, 0), 6.0, 300.0)
w = [rn(N, 3).to(dev), rn(N, 4).to(dev), (rn(N, 3) * 20).to(dev), rn(N, 1).to(dev), rn(N, 16, 3).to(dev)]
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    outs = deformation.get_deformation(pc, 0.3 + 0.05 * i)
    deform_torch.train_objective(pc, outs, w, (8e-6, 0.0, 0.0)).backward()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()

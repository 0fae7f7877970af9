"""Developer tool (GPU box): the training-time deformation path alone, for `compute-sanitizer --tool initcheck`
(every byte the weight-gradient kernel streams must have been written by the forward / data-gradient kernels)."""
import sys
sys.path.insert(0, ".")
import torch
from oracle import deform_torch
from saro_gs_b200 import deformation

dev = torch.device("cuda:0")
for P, F, flags in ((700, 32, (1, 0, 0)), (129, 16, (1, 1, 1)), (31, 8, (0, 0, 0))):
    g = torch.Generator().manual_seed(P)
    rn = lambda *s: torch.randn(*s, generator=g)
    t = dict(xyz=rn(P, 3) * 2, rotation=rn(P, 4), scaling=rn(P, 3) * 0.5 - 3.5, opacity=rn(P, 1) * 2, features_dc=rn(P, 1, 3) * 0.5,
             features_rest=rn(P, 15, 3) * 0.1, temporal_pos=torch.rand(P, 1, generator=g), hexplane_feature=rn(P, F) * 0.5)
    lv = {k: v.to(dev).requires_grad_(True) for k, v in t.items()}
    mlps = deform_torch.make_train_mlps(F, device=dev, seed=P)
    pc = deform_torch.TrainModelStandIn(lv, mlps, flags, 6.0, 300.0)
    outs = deformation.get_deformation(pc, 0.4)
    deform_torch.train_objective(pc, outs, [torch.ones_like(o) for o in outs], (0.3, 0.2, 0.1)).backward()
    torch.cuda.synchronize()
    print("ok", P, F, float(lv["hexplane_feature"].grad.abs().sum()))

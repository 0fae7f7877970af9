#!/usr/bin/env python
"""Developer tool (GPU box): one small forward + backward (+ fused loss) — run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool synccheck python tools/sanitize_small.py
"""
import sys
import torch
sys.path.insert(0, '.')
import saro_gs_b200 as sgs
from saro_gs_b200 import synthetic, loss_utils

from saro_gs_b200 import _lib
from saro_gs_b200.densify import BatchDensifyStats

dev = torch.device('cuda:0')
# both places the depth sort can happen (round 2: per-supertile sort / global sort); P = 9000 on 160 x 112 pixels gives
# supertile buckets beyond the 4096 entries a block sorts in shared memory (chunked path through global scratch)
for (P, seed, kw, mode) in [(3000, 0, dict(width=160, height=112, fx=120.0, log_scale_mean=-1.6), 1),
                            (3000, 0, dict(width=160, height=112, fx=120.0, log_scale_mean=-1.6), 0),
                            (9000, 2, dict(width=160, height=112, fx=120.0, log_scale_mean=-1.6), 1), (64, 1, {}, 1)]:
    _lib.load().sgs_debug_set_binning_mode(mode)
    scene, cam = synthetic.small_scene(P=P, seed=seed, **kw)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    image, radii, depth = sgs.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=torch.zeros_like(leaves["means3D"]),
                                                    opacities=leaves["opacities"], shs=leaves["shs"],
                                                    scales=leaves["scales"], rotations=leaves["rotations"])
    gt = torch.rand_like(image)
    loss = loss_utils.l1_dssim_loss(image, gt, 0.2)
    sink = BatchDensifyStats(P, dev)
    sink.attach_next_backward()          # densification statistics in the epilogue of the backward-preprocess kernel
    loss.backward()
    torch.cuda.synchronize()
    print("ok", P, "mode", mode, float(loss), int((radii > 0).sum()), float(leaves["means3D"].grad.abs().sum()),
          int(sink.vis_count.sum()))
_lib.load().sgs_debug_set_binning_mode(-1)

# deformation hand-off (tcgen05 / TMEM kernel) and densification statistics on small clouds
from saro_gs_b200 import deformation
import types
for P in (700, 129):
    scene, _ = synthetic.small_scene(P=P, seed=3)
    pc = synthetic.model_to(synthetic.dynamic_model(scene, feat_dim=32 if P == 700 else 16), dev)
    with torch.no_grad():
        out = deformation.get_deformation_eval(pc, 0.4)
    torch.cuda.synchronize()
    print("deform ok", P, out[0].shape[0], float(out[4].abs().sum()))
    stats = BatchDensifyStats(P, dev)
    stats.add_view(torch.randn(P, 3, device=dev), torch.randint(0, 9, (P,), device=dev, dtype=torch.int32))
    m = types.SimpleNamespace(max_radii2D=torch.zeros(P, device=dev), xyz_gradient_accum=torch.zeros(P, 1, device=dev),
                              denom=torch.zeros(P, 1, device=dev))
    stats.commit(m)
    torch.cuda.synchronize()
    print("densify ok", P, float(m.denom.sum()))

# scale-aware plane sampler (round 2): forward + backward on a small field with two resolution levels
from saro_gs_b200.hexplane import ScaleAwareResField
cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 8, "resolution": [16, 16, 8, 6]}
field = ScaleAwareResField(cfg, [1, 2]).to(dev)
with torch.no_grad():
    for level in field.grids:
        for p in level:
            p.normal_()
field.set_aabb([1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], 10)
pts = torch.rand(777, 3, device=dev) * 2.4 - 1.2
feats = field(pts, torch.rand(777, 1, device=dev) * 0.9, torch.exp(torch.randn(777, 3, device=dev) * 2 - 3))
feats.sum().backward()
torch.cuda.synchronize()
print("plane ok", tuple(feats.shape), float(field.grids[0][0].grad.abs().sum()))

# the capacity re-launch path of the binning (prediction forced to a few instances)
_lib.load().sgs_debug_set_capacity(64)
scene, cam = synthetic.small_scene(P=900, seed=5)
rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
with torch.no_grad():
    img = sgs.GaussianRasterizer(rs)(means3D=scene.means3D.to(dev), means2D=torch.zeros(900, 3, device=dev),
                                     opacities=scene.opacities.to(dev), shs=scene.shs.to(dev), scales=scene.scales.to(dev),
                                     rotations=scene.rotations.to(dev))[0]
_lib.load().sgs_debug_set_capacity(-1)
torch.cuda.synchronize()
print("relaunch ok", float(img.sum()))

# training-time deformation (round 2c): tcgen05 job kernels forward + data gradients, TMA -> tcgen05 weight gradients,
# fused epilogues, on small clouds (ragged tail, all regularisers / none)
from oracle import deform_torch
for P, F, flags in ((700, 32, (1, 0, 0)), (129, 16, (1, 1, 1)), (31, 8, (0, 0, 0))):
    g = torch.Generator().manual_seed(P)
    rn = lambda *s: torch.randn(*s, generator=g)
    t = dict(xyz=rn(P, 3) * 2, rotation=rn(P, 4), scaling=rn(P, 3) * 0.5 - 3.5, opacity=rn(P, 1) * 2, features_dc=rn(P, 1, 3) * 0.5,
             features_rest=rn(P, 15, 3) * 0.1, temporal_pos=torch.rand(P, 1, generator=g), hexplane_feature=rn(P, F) * 0.5)
    lv = {k: v.to(dev).requires_grad_(True) for k, v in t.items()}
    mlps = deform_torch.make_train_mlps(F, device=dev, seed=P)
    tpc = deform_torch.TrainModelStandIn(lv, mlps, flags, 6.0, 300.0)
    outs = deformation.get_deformation(tpc, 0.4)
    w = [torch.ones_like(o) for o in outs]
    deform_torch.train_objective(tpc, outs, w, (0.3, 0.2, 0.1)).backward()
    torch.cuda.synchronize()
    print("deform train ok", P, F, float(lv["hexplane_feature"].grad.abs().sum()), float(mlps["shs"][2].weight.grad.abs().sum()))

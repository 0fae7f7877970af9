"""GPU parity tests of the scale-aware plane sampler (csrc/sgs_plane.cu behind
saro_gs_b200.hexplane.ScaleAwareResField): against the golden outputs of the reference's own hexplane.py
(tests/golden/plane_*.npz, see make_golden_plane.py for how the missing third-party op is stood in for) and against the
float64 numpy oracle at sizes the goldens do not reach.  Tolerance: 1e-4 of the largest entry (north_star), forward and
plane gradients."""
import numpy as np
import pytest
import torch

from golden_util import load, maxrel
from oracle import plane_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev(native_lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def make_field(d, dev):
    from saro_gs_b200.hexplane import ScaleAwareResField
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": int(d["out_dim"]),
           "resolution": [int(r) for r in d["reso"]]}
    field = ScaleAwareResField(cfg, [int(m) for m in d["multires"]]).to(dev)
    with torch.no_grad():
        for li, level in enumerate(field.grids):
            for ci, p in enumerate(level):
                p.copy_(torch.from_numpy(d[f"grid_{li}_{ci}"]))
    field.set_aabb([float(v) for v in d["aabb"][0]], [float(v) for v in d["aabb"][1]], int(d["duration"]))
    with torch.no_grad():
        field.base_scale.copy_(torch.from_numpy(d["base_scale"]))   # the golden run's buffer (float32 of the double quotient)
    return field


@pytest.mark.parametrize("name", ["plane_small", "plane_ragged", "plane_wide"])
def test_field_vs_reference_golden(dev, name):
    d = load(name)
    field = make_field(d, dev)
    t = lambda k: torch.from_numpy(d[k]).to(dev)
    feats = field(t("pts"), t("timestamps"), t("scales"))
    assert tuple(feats.shape) == d["features_f64"].shape
    assert maxrel(feats.detach().cpu().numpy(), d["features_f64"]) < TOL
    feats.backward(torch.from_numpy(d["dout"]).float().to(dev))
    for li, level in enumerate(field.grids):
        for ci, p in enumerate(level):
            want = d[f"dgrid_f64_{li}_{ci}"]
            assert tuple(p.grad.shape) == want.shape
            assert np.abs(p.grad.cpu().numpy() - want).max() <= TOL * max(np.abs(want).max(), 1e-30), (li, ci)


@pytest.mark.parametrize("reso,C,multires,N", [((64, 64, 64, 12), 32, (1,), 50_000), ((32, 64, 16, 7), 8, (1, 2), 20_001),
                                               ((128, 128, 128, 32), 16, (1,), 100_000)])
def test_field_vs_float64_oracle(dev, reso, C, multires, N):
    from saro_gs_b200.hexplane import ScaleAwareResField
    g = torch.Generator().manual_seed(N)
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": C, "resolution": list(reso)}
    field = ScaleAwareResField(cfg, list(multires)).to(dev)
    with torch.no_grad():
        for level in field.grids:
            for p in level:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
    xyz_max, xyz_min, duration = [2.0, 1.5, 3.0], [-2.5, -1.0, -0.5], 30
    field.set_aabb(xyz_max, xyz_min, duration)
    ext = torch.tensor(xyz_max) - torch.tensor(xyz_min)
    pts = 0.5 * (torch.tensor(xyz_max) + torch.tensor(xyz_min)) + (torch.rand(N, 3, generator=g) - 0.5) * ext * 1.2
    ts = torch.rand(N, 1, generator=g) * (duration - 1) / duration
    scales = torch.exp(torch.rand(N, 3, generator=g) * 10.0 - 8.0)
    feats = field(pts.to(dev), ts.to(dev), scales.to(dev))
    dout = torch.randn(feats.shape, generator=g)
    feats.backward(dout.to(dev))
    grids = [[p.detach().cpu().numpy() for p in level] for level in field.grids]
    args = (pts.numpy(), ts.numpy(), scales.numpy())
    want = plane_oracle.field_forward(*args, grids, field.aabb.cpu().numpy(), duration, field.base_scale.cpu().numpy(),
                                      field.reso_list)
    assert maxrel(feats.detach().cpu().numpy(), want) < TOL
    shapes = [[g_.shape for g_ in level] for level in grids]
    dwant = plane_oracle.field_backward(*args, shapes, field.aabb.cpu().numpy(), duration,
                                        field.base_scale.cpu().numpy(), field.reso_list, dout.numpy().astype(np.float64))
    for li, level in enumerate(field.grids):
        for ci, p in enumerate(level):
            w = dwant[li][ci]
            assert np.abs(p.grad.cpu().numpy() - w).max() <= TOL * max(np.abs(w).max(), 1e-30), (li, ci)


def test_pyramid_cache_follows_in_place_updates_and_edge_cases(dev):
    from saro_gs_b200.hexplane import ScaleAwareResField, UnsupportedPlaneConfig
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 4, "resolution": [8, 8, 8, 4]}
    field = ScaleAwareResField(cfg, [1]).to(dev)
    field.set_aabb([1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], 10)
    pts = torch.zeros(5, 3, device=dev)
    ts = torch.full((5, 1), 0.3, device=dev)
    sc = torch.full((5, 3), 0.01, device=dev)
    assert float(field(pts, ts, sc).abs().max()) == 0.0            # zero-initialised planes (hexplane.py:80-86)
    with torch.no_grad():
        for p in field.grids[0]:
            p.add_(1.0)                                             # in-place, like an optimizer step
    out = field(pts, ts, sc)
    assert torch.allclose(out, torch.full_like(out, 6.0))           # six planes of ones, any level, any position
    assert tuple(field(pts[:0], ts[:0], sc[:0]).shape) == (0, 1)    # the reference's empty-input convention
    with pytest.raises(UnsupportedPlaneConfig):
        field(pts.clone().requires_grad_(True), ts, sc)
    with pytest.raises(RuntimeError):
        field(pts.cpu(), ts.cpu(), sc.cpu())
    names = sorted(field.state_dict().keys())
    assert names == sorted([f"grids.0.{i}" for i in range(6)] + ["aabb", "duration", "max_level", "base_scale"])

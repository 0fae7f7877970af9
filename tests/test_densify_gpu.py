"""GPU parity of the densification-statistics kernels (csrc/sgs_densify.cu) through the C ABI: against the result of
the reference's own statements (tests/golden/densify_*.npz), the numpy oracle, and on real rasterizer outputs."""
import glob
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import densify_oracle
from saro_gs_b200.densify import BatchDensifyStats

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "densify_*.npz")))
DEV = torch.device("cuda:0")


def model_from(mr, acc, den):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return types.SimpleNamespace(max_radii2D=t(mr), xyz_gradient_accum=t(acc), denom=t(den))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_against_reference_golden(path):
    z = np.load(path)
    P = z["radii"].shape[1]
    g = model_from(z["start_max_radii2D"], z["start_xyz_gradient_accum"], z["start_denom"])
    stats = BatchDensifyStats(P, DEV)
    for grad, radii in zip(z["grads"], z["radii"]):
        stats.add_view(torch.from_numpy(grad).to(DEV), torch.from_numpy(radii).to(DEV))
    stats.commit(g)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(g.max_radii2D.cpu().numpy(), z["out_max_radii2D"])        # integer-valued: exact
    np.testing.assert_array_equal(g.denom.cpu().numpy(), z["out_denom"])
    np.testing.assert_allclose(g.xyz_gradient_accum.cpu().numpy(), z["out_xyz_gradient_accum"], rtol=2e-6, atol=0)


def test_two_iterations_and_reset_vs_oracle():
    rng = np.random.default_rng(0)
    P = 100_003
    mr, acc, den = np.zeros(P, np.float32), np.zeros((P, 1), np.float32), np.zeros((P, 1), np.float32)
    g = model_from(mr, acc, den)
    stats = BatchDensifyStats(P, DEV)
    for it in range(2):
        grads = [rng.standard_normal((P, 3)).astype(np.float32) * 1e-4 for _ in range(4)]
        radii = [np.where(rng.random(P) < 0.4, 0, rng.integers(1, 60, P)).astype(np.int32) for _ in range(4)]
        stats.reset()
        for a, b in zip(grads, radii):
            stats.add_view(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
        stats.commit(g)
        mr, acc, den = densify_oracle.batch_statistics(grads, radii, mr, acc, den, dtype=np.float64)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(g.max_radii2D.cpu().numpy(), mr)
    np.testing.assert_array_equal(g.denom.cpu().numpy(), den)
    np.testing.assert_allclose(g.xyz_gradient_accum.cpu().numpy(), acc, rtol=1e-5, atol=1e-12)


def test_on_rasterizer_outputs():
    """The statistics of a two-view batch computed from the rasterizer's own radii / means2D gradient equal what the
    reference's list-and-stack code (train.py:211-215, :281-291, restated with torch ops here) gives."""
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.config2_scene(P=30_000, width=320, height=240, fx=250.0)
    P = scene.means3D.shape[0]
    params = {k: getattr(scene, k).to(DEV).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    g = model_from(np.zeros(P, np.float32), np.zeros((P, 1), np.float32), np.zeros((P, 1), np.float32))
    stats = BatchDensifyStats(P, DEV)
    norms, rads = [], []
    for view in range(2):
        c = cam if view == 0 else synthetic.yaw_camera(cam.width, cam.height, 250.0, 0.35)      # second camera on an arc
        rs = sgs.GaussianRasterizationSettings(c.height, c.width, c.tanfovx, c.tanfovy, torch.zeros(3, device=DEV), 1.0,
                                               c.viewmatrix.to(DEV), c.projmatrix.to(DEV), 3, c.campos.to(DEV), False)
        m2d = torch.zeros(P, 3, device=DEV, requires_grad=True)
        color, radii, _ = sgs.GaussianRasterizer(rs)(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"],
                                                     shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        color.square().mean().backward()
        stats.add_view(m2d.grad, radii)
        norms.append(torch.norm(m2d.grad[:, :2], dim=-1))
        rads.append(radii)
    stats.commit(g)
    count = torch.stack([r > 0 for r in rads], 1).sum(1)
    seen = count > 0
    want_acc = torch.zeros(P, device=DEV)
    want_acc[seen] = torch.stack(norms, 1).sum(1)[seen] / count[seen]
    want_mr = torch.zeros(P, device=DEV)
    want_mr[seen] = torch.stack(rads, 1).max(1)[0][seen].float()
    assert seen.any() and (~seen).any()
    assert torch.equal(g.max_radii2D, want_mr)
    assert torch.equal(g.denom.reshape(-1), seen.float())
    assert torch.allclose(g.xyz_gradient_accum.reshape(-1), want_acc, rtol=1e-5, atol=0)


def test_fused_sink_in_rasterizer_backward_equals_add_view():
    """stats.attach_next_backward(): the rasterizer's backward applies the view's update in the epilogue of its last
    kernel (sgs_densify_attach) — bit-identical running buffers to the stand-alone add_view kernel, one-shot."""
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.config2_scene(P=30_000, width=320, height=240, fx=250.0)
    P = scene.means3D.shape[0]
    params = {k: getattr(scene, k).to(DEV).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    fused, plain = BatchDensifyStats(P, DEV), BatchDensifyStats(P, DEV)
    for view in range(3):
        c = cam if view == 0 else synthetic.yaw_camera(cam.width, cam.height, 250.0, 0.2 * view)
        rs = sgs.GaussianRasterizationSettings(c.height, c.width, c.tanfovx, c.tanfovy, torch.zeros(3, device=DEV), 1.0,
                                               c.viewmatrix.to(DEV), c.projmatrix.to(DEV), 3, c.campos.to(DEV), False)
        m2d = torch.zeros(P, 3, device=DEV, requires_grad=True)
        color, radii, _ = sgs.GaussianRasterizer(rs)(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"],
                                                     shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        if view < 2:
            fused.attach_next_backward()
        color.square().mean().backward()
        if view < 2:
            plain.add_view(m2d.grad, radii)
    torch.cuda.synchronize()
    assert fused.views == 2                                   # the third backward found no armed sink (one-shot)
    assert (plain.vis_count > 0).any() and (plain.vis_count == 0).any()
    assert torch.equal(fused.vis_count, plain.vis_count)
    assert torch.equal(fused.radii_max, plain.radii_max)
    assert torch.equal(fused.grad_sum, plain.grad_sum)
    # a sink armed for another cloud size is refused and disarmed
    other = BatchDensifyStats(P + 1, DEV)
    m2d = torch.zeros(P, 3, device=DEV, requires_grad=True)
    color, radii, _ = sgs.GaussianRasterizer(rs)(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"],
                                                 shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
    other.attach_next_backward()
    with pytest.raises(RuntimeError):
        color.square().mean().backward()
    assert other.views == 0


def test_rejects_bad_inputs():
    stats = BatchDensifyStats(10, DEV)
    with pytest.raises(RuntimeError):
        stats.add_view(torch.zeros(10, 3), torch.zeros(10, dtype=torch.int32, device=DEV))          # CPU gradient
    with pytest.raises(RuntimeError):
        stats.add_view(torch.zeros(10, 3, device=DEV), torch.zeros(10, dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError):
        BatchDensifyStats(10, "cpu")
    BatchDensifyStats(0, DEV).commit(model_from(np.zeros(0, np.float32), np.zeros((0, 1), np.float32), np.zeros((0, 1), np.float32)))

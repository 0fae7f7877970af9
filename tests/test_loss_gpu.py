"""GPU parity tests of the fused L1 + D-SSIM loss (csrc/sgs_loss.cu through the C ABI) against (1) golden vectors
from the reference's own Python functions and (2) the float64 numpy oracle."""
import numpy as np
import pytest
import torch

from golden_util import load

pytestmark = pytest.mark.gpu
CASES = ["loss_chw_ragged", "loss_chw_tiles", "loss_batched", "loss_identical"]
VAL_TOL = 5e-6      # float32 separable convolution vs the reference's float32/float64 2-D convolution
GRAD_TOL = 1e-4     # of the largest gradient entry (same bar as the rasterizer)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def lu(native_lib):
    from saro_gs_b200 import loss_utils
    return loss_utils


@pytest.mark.parametrize("name", CASES)
def test_loss_and_gradient_vs_reference_golden(lu, dev, name):
    d = load(name)
    x = torch.from_numpy(d["img"]).to(dev).requires_grad_(True)
    y = torch.from_numpy(d["gt"]).to(dev)
    l1 = lu.l1_loss(x, y)
    s = lu.ssim(x, y)
    loss = 0.8 * l1 + 0.2 * (1.0 - s)
    loss.backward()
    assert abs(l1.item() - float(d["l1_f64"])) < VAL_TOL
    assert abs(s.item() - float(d["ssim_f64"])) < VAL_TOL
    assert abs(loss.item() - float(d["loss_f64"])) < VAL_TOL
    g = x.grad.cpu().numpy().astype(np.float64)
    for tag in ("f32", "f64"):
        ref = d[f"grad_{tag}"]
        # (absolute floor 1e-8: for img == gt the true gradient is 0 and both sides hold only rounding noise)
        assert np.abs(g - ref).max() <= GRAD_TOL * np.abs(ref).max() + 1e-8, tag
    if "ssim_per_image_f64" in d.files:
        per = lu.ssim(x.detach(), y, size_average=False).cpu().numpy()
        assert np.abs(per - d["ssim_per_image_f64"]).max() < VAL_TOL
    # the one-call form used by a training step
    x2 = x.detach().clone().requires_grad_(True)
    lu.l1_dssim_loss(x2, y, 0.2).backward()
    assert torch.allclose(x2.grad, x.grad, rtol=1e-5, atol=1e-9)


def test_full_size_properties_and_oracle_spot_check(lu, dev):
    """1352x1014 (config 2 image size): ssim(x, x) = 1, loss symmetric bounds, gradient vs oracle on a crop."""
    from oracle import ssim_oracle as so
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(3, 1014, 1352, generator=g)
    img = (gt + 0.05 * torch.randn(3, 1014, 1352, generator=g)).clamp(0, 1)
    x = img.to(dev).requires_grad_(True)
    y = gt.to(dev)
    assert abs(lu.ssim(y, y).item() - 1.0) < 1e-6
    assert lu.l1_loss(y, y).item() == 0.0
    loss = lu.l1_dssim_loss(x, y, 0.2)
    loss.backward()
    assert 0.0 < loss.item() < 1.0 and torch.isfinite(x.grad).all()
    # deterministic
    x2 = img.to(dev).requires_grad_(True)
    loss2 = lu.l1_dssim_loss(x2, y, 0.2)
    loss2.backward()
    assert loss2.item() == loss.item() and torch.equal(x2.grad, x.grad)
    # oracle on a 96x128 crop taken far enough from the crop border (the window only sees 5 pixels)
    r0, c0, h, w = 300, 500, 96, 128
    crop_x = img[:, r0:r0 + h, c0:c0 + w].numpy()
    crop_y = gt[:, r0:r0 + h, c0:c0 + w].numpy()
    _, og = so.loss_and_grad(crop_x, crop_y, 0.2)
    n_full, n_crop = img.numel(), crop_x.size
    got = x.grad[:, r0 + 10:r0 + h - 10, c0 + 10:c0 + w - 10].cpu().numpy().astype(np.float64) * n_full
    want = og[:, 10:h - 10, 10:w - 10] * n_crop
    assert np.abs(got - want).max() <= GRAD_TOL * np.abs(want).max()


def test_loss_feeds_the_rasterizer_backward(lu, dev, native_lib):
    """render -> fused loss -> backward through the rasterizer: the loss gradient is the dL/dcolor the rasterizer
    consumes (train.py:199-211 of the reference)."""
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.small_scene(P=512, seed=0)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    image, radii, depth = sgs.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=torch.zeros_like(leaves["means3D"]),
                                                    opacities=leaves["opacities"], shs=leaves["shs"],
                                                    scales=leaves["scales"], rotations=leaves["rotations"])
    gt = torch.rand_like(image)
    Ll1 = lu.l1_loss(image, gt)
    loss = 0.8 * Ll1 + 0.2 * (1.0 - lu.ssim(image, gt))
    loss.backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() and v.grad.abs().sum() > 0 for v in leaves.values())

"""CPU tests: the numpy loss oracle (oracle/ssim_oracle.py) against golden vectors produced by the reference's own
Python functions (tests/golden/make_golden_loss.py), and the host-side argument checks of saro_gs_b200.loss_utils."""
import numpy as np
import pytest
import torch

from golden_util import load

CASES = ["loss_chw_ragged", "loss_chw_tiles", "loss_batched", "loss_identical"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_python(name):
    from oracle import ssim_oracle as so
    d = load(name)
    x, y = d["img"].astype(np.float64), d["gt"].astype(np.float64)
    assert abs(so.l1_loss(x, y) - float(d["l1_f64"])) < 1e-12
    assert abs(so.ssim(x, y) - float(d["ssim_f64"])) < 1e-10
    loss, grad = so.loss_and_grad(x, y, 0.2)
    assert abs(loss - float(d["loss_f64"])) < 1e-10
    ref = d["grad_f64"]
    assert np.abs(grad - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-30) + 1e-15
    if "ssim_per_image_f64" in d.files:
        per = so.ssim_map(x, y).mean(axis=(1, 2, 3))
        assert np.abs(per - d["ssim_per_image_f64"]).max() < 1e-10
    # the reference's own float32 run is within float32 noise of its float64 run: the tolerance used on the GPU
    assert abs(float(d["ssim_f32"]) - float(d["ssim_f64"])) < 2e-6


def test_window_is_the_reference_window():
    from oracle import ssim_oracle as so
    g = so.window_1d()
    assert g.dtype == np.float32 and g.shape == (11,) and abs(float(g.sum()) - 1.0) < 1e-6
    assert g[5] == np.float32(2.660117149e-01) and g[0] == np.float32(1.028380124e-03)
    assert np.abs(g - so.window_1d_formula()).max() < 1e-7          # the formula, up to float32 summation order
    # and bit-equal to what torch (the reference's arithmetic) produces
    from math import exp
    t = torch.Tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    assert np.array_equal((t / t.sum()).numpy(), g)


def test_host_checks_fail_loudly_without_gpu():
    from saro_gs_b200 import loss_utils
    a = torch.rand(3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        loss_utils.l1_loss(a, a)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        loss_utils.ssim(a, a)
    with pytest.raises(NotImplementedError):
        loss_utils.ssim(a, a, window_size=7)

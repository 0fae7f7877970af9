"""CPU: the PyTorch restatement of the training-time deformation (oracle/deform_torch.py::torch_get_deformation — the
checker of the GPU tests and the bench baseline) against outputs AND autograd gradients of the reference's own
get_deformation source (tests/golden/deformtrain_*.npz, made by tests/golden/make_golden_deform_train.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import deform_torch
from deform_train_util import GOLDEN, OUTS, build, gradients, maxrel


def test_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_restatement_float64_matches_reference(path):
    z = np.load(path)
    pc, leaves, mlps, weights = build(z, "cpu", torch.float64)
    outs = deform_torch.torch_get_deformation(pc, float(z["timestamp"]))
    for k, o in zip(OUTS, outs):
        np.testing.assert_allclose(o.detach().numpy(), z[f"f64_{k}"], rtol=1e-11, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(pc._lifespan.detach().numpy(), z["f64_lifespan"], rtol=1e-12)
    np.testing.assert_allclose(pc.real_xyz.numpy(), z["f64_real_xyz"], rtol=1e-11, atol=1e-12)
    loss = deform_torch.train_objective(pc, outs, weights, z["lambdas"])
    assert abs(loss.item() - float(z["f64_loss"])) <= 1e-10 * abs(float(z["f64_loss"]))
    loss.backward()
    for k, g in gradients(leaves, mlps).items():
        assert g is not None, k
        assert maxrel(g.numpy(), z[f"f64_{k}"]) <= 1e-6, k            # fixture gradients are stored as float32


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_restatement_float32_outputs(path):
    z = np.load(path)
    pc, _, _, _ = build(z, "cpu", torch.float32)
    with torch.no_grad():
        outs = deform_torch.torch_get_deformation(pc, float(z["timestamp"]))
    for k, o in zip(OUTS, outs):
        assert maxrel(o.numpy(), z[f"f32_{k}"]) <= 2e-6, k


def test_time_embedding_of_zero_is_the_base_feature():
    e = deform_torch.time_embedding(torch.zeros(5, 1))
    assert torch.equal(e, torch.tensor([0.0, 0, 1, 0, 1, 0, 1, 0, 1]).repeat(5, 1))


def test_relu_kink_sensitivity_of_weight_gradients():
    """Why the at-scale GPU test compares weight gradients on rows away from the ReLU kink (tests/test_deform_train_gpu.py):
    in pure float64, moving every hidden pre-activation by 1e-7 (far below float32 resolution of these values) flips the
    ReLU mask of a handful of the 10^7 units, and each flip changes dL/dW by that row's whole term — the weight gradient
    of a 60 000-row batch moves by more than the 1e-4 parity bar.  No float32 implementation can therefore agree with a
    float64 reference (or with another float32 implementation) to 1e-4 on such rows; on the rows that keep a margin
    from 0 the same perturbation changes nothing."""
    torch.manual_seed(0)
    n, F = 60_000, 32
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, F + 9, generator=g, dtype=torch.float64) * 0.5
    mlp = deform_torch.make_train_mlps(F, dtype=torch.float64, seed=2)["motion"]
    W1, b1, W2, b2, W3 = mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, mlp[4].weight
    dy = torch.randn(n, 3, generator=g, dtype=torch.float64)

    def grad_w1(eps, rows=None):
        with torch.no_grad():
            y1 = x @ W1.t() + b1
            m1 = (y1 + eps * torch.randn(y1.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)) > 0
            h1 = y1 * m1
            y2 = h1 @ W2.t() + b2
            m2 = (y2 + eps * torch.randn(y2.shape, generator=torch.Generator().manual_seed(4), dtype=torch.float64)) > 0
            dh1 = (((dy @ W3) * m2) @ W2) * m1
            if rows is not None:
                return dh1[rows].t() @ x[rows], torch.minimum(y1.abs().min(1).values, y2.abs().min(1).values)
            return dh1.t() @ x, torch.minimum(y1.abs().min(1).values, y2.abs().min(1).values)

    exact, margin = grad_w1(0.0)
    moved, _ = grad_w1(1e-7)
    rel = float((exact - moved).abs().max() / exact.abs().max())
    assert rel > 1e-4, rel                                   # the bar is unattainable on rows at the kink ...
    smooth = margin > 1e-5
    assert 0.5 < float(smooth.double().mean()) < 1.0
    a, _ = grad_w1(0.0, smooth)
    b, _ = grad_w1(1e-7, smooth)
    assert torch.equal(a, b)                                 # ... and trivially met on the rows away from it

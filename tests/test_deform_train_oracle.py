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

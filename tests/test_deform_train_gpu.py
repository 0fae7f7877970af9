"""GPU parity of the training-time deformation path (saro_gs_b200.deformation.get_deformation -> the tcgen05 job
kernels of csrc/sgs_deform.cu through the C ABI): outputs, side effects and EVERY gradient against the reference's own
get_deformation source (tests/golden/deformtrain_*.npz, float64 autograd), and at scale against the float64 PyTorch
restatement pinned on those fixtures (tests/test_deform_train_oracle.py).
Tolerance: 1e-4 of the largest entry of each tensor (BASELINE north_star's floating-point bar)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from deform_train_util import GOLDEN, LEAVES, MLPS, OUTS, build, gradients, maxrel
from oracle import deform_torch
from saro_gs_b200 import deformation

TOL = 1e-4
DEV = "cuda:0"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_outputs_and_gradients_against_reference_golden(path):
    z = np.load(path)
    pc, leaves, mlps, weights = build(z, DEV, torch.float32)
    outs = deformation.get_deformation(pc, float(z["timestamp"]))
    for k, o in zip(OUTS, outs):
        assert o.shape == z[f"f64_{k}"].shape, k
        assert maxrel(o.detach().cpu().numpy(), z[f"f64_{k}"]) <= TOL, k
    assert maxrel(pc._lifespan.detach().cpu().numpy(), z["f64_lifespan"]) <= TOL
    assert maxrel(pc.real_xyz.cpu().numpy(), z["f64_real_xyz"]) <= TOL
    assert not pc.real_xyz.requires_grad
    if z["flags"][0]:
        assert maxrel(pc.scale_residual.detach().cpu().numpy(), z["f64_scale_residual"]) <= TOL
    loss = deform_torch.train_objective(pc, outs, weights, z["lambdas"])
    # the objective is a signed sum with heavy cancellation: the bar is relative to the sum of the terms' magnitudes
    magnitude = sum(float((w * o.detach()).abs().sum()) for w, o in zip(weights, outs))
    assert abs(loss.item() - float(z["f64_loss"])) <= 1e-5 * magnitude
    loss.backward()
    torch.cuda.synchronize()
    for k, g in gradients(leaves, mlps).items():
        assert g is not None, k
        assert maxrel(g.cpu().numpy(), z[f"f64_{k}"]) <= TOL, k


def _random_case(n, feat_dim, flags, seed):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    t = dict(xyz=rn(n, 3) * 2, rotation=rn(n, 4), scaling=rn(n, 3) * 0.5 - 3.5, opacity=rn(n, 1) * 2, features_dc=rn(n, 1, 3) * 0.5,
             features_rest=rn(n, 15, 3) * 0.1, temporal_pos=torch.rand(n, 1, generator=g), hexplane_feature=rn(n, feat_dim) * 0.5)
    w = [rn(n, 3), rn(n, 4), rn(n, 3) * 20, rn(n, 1), rn(n, 16, 3)]
    return t, w


def _run(fn, tensors, weights, mlps_src, flags, dtype, timestamp, kink=None):
    leaves = {k: v.to(device=DEV, dtype=dtype).requires_grad_(True) for k, v in tensors.items()}
    mlps = {k: __import__("copy").deepcopy(m).to(device=DEV, dtype=dtype) for k, m in mlps_src.items()}
    if kink is not None:       # smallest |pre-activation| of every row over all hidden layers of all evaluations
        def hook(_, __, out):
            kink[0] = out.abs().min(dim=1).values if kink[0] is None else torch.minimum(kink[0], out.abs().min(dim=1).values)
        for m in mlps.values():
            m[0].register_forward_hook(hook)
            m[2].register_forward_hook(hook)
    pc = deform_torch.TrainModelStandIn(leaves, mlps, flags, 6.0, 300.0)
    outs = fn(pc, timestamp)
    loss = deform_torch.train_objective(pc, outs, [w.to(device=DEV, dtype=dtype) for w in weights], (0.3, 0.2, 0.1))
    loss.backward()
    return [o.detach() for o in outs], gradients(leaves, mlps)


def _smooth_rows(tensors, mlps, flags, timestamp, margin=2e-4):
    """Indices of the rows on which every hidden pre-activation of every evaluation is at least `margin` from 0
    (float64 restatement)."""
    kink = [None]
    _run(deform_torch.torch_get_deformation, tensors, [torch.zeros_like(tensors[k]) for k in
         ("xyz", "rotation", "scaling", "opacity")] + [torch.zeros(tensors["xyz"].shape[0], 16, 3)], mlps, flags, torch.float64,
         timestamp, kink)
    return torch.nonzero(kink[0].cpu() > margin).reshape(-1)


@pytest.mark.parametrize("n,feat_dim,flags", [(100_000, 32, (1, 0, 0)), (40_001, 16, (1, 1, 1)), (127, 8, (0, 0, 0)), (129, 24, (1, 0, 1)),
                                               (1, 32, (1, 1, 0)), (4_097, 24, (0, 1, 0)), (33, 16, (0, 0, 1))])
def test_at_scale_against_float64_restatement(n, feat_dim, flags):
    """ReLU is not differentiable at 0: on a row where some hidden pre-activation is within rounding of 0, two
    arithmetics may take different sides of the kink, that row's feature gradient then differs by one hidden unit's
    whole contribution and every weight gradient (a sum over rows) by that row's term — for ANY pair of
    implementations (a 1e-7 perturbation of the pre-activations moves dL/dW1 of a 100 000-row batch by 2e-3 of its
    largest entry in float64; each row has ~1 800 hidden units over the seven evaluations).  The parity bar is
    therefore checked where the function is differentiable: the batch is drawn from rows whose pre-activations all
    keep a 2e-4 margin from 0, repeated with fresh objective weights up to n rows."""
    base, _ = _random_case(4000, feat_dim, flags, seed=n)
    mlps = deform_torch.make_train_mlps(feat_dim, seed=n + 1)
    keep = _smooth_rows(base, mlps, flags, 0.37)
    assert keep.numel() > 1000
    idx = keep[torch.arange(n) % keep.numel()]
    tensors, weights = _random_case(n, feat_dim, flags, seed=n + 2)
    tensors["hexplane_feature"] = base["hexplane_feature"][idx].clone()
    tensors["temporal_pos"] = base["temporal_pos"][idx].clone()
    o_n, g_n = _run(deformation.get_deformation, tensors, weights, mlps, flags, torch.float32, 0.37)
    o_r, g_r = _run(deform_torch.torch_get_deformation, tensors, weights, mlps, flags, torch.float64, 0.37)
    for k, a, b in zip(OUTS, o_n, o_r):
        assert maxrel(a.cpu().numpy(), b.cpu().numpy()) <= TOL, k
    report = {k: maxrel(g_n[k].cpu().numpy(), g_r[k].cpu().numpy()) for k in g_r}
    bad = {k: v for k, v in report.items() if not v <= TOL}
    assert not bad, bad


def test_no_grad_and_repacking_after_an_optimizer_step():
    z = np.load(GOLDEN[0])
    pc, leaves, mlps, weights = build(z, DEV, torch.float32)
    t = float(z["timestamp"])
    with torch.no_grad():
        a = deformation.get_deformation(pc, t)
    assert all(not o.requires_grad for o in a)
    with torch.no_grad():
        for m in mlps.values():
            for p in m.parameters():
                p.mul_(1.01)                      # in-place write bumps the version: images must be re-packed
        b = deformation.get_deformation(pc, t)
        c = deform_torch.torch_get_deformation(pc, t)
    assert float((a[0] - b[0]).abs().max()) > 0
    for k, x, y in zip(OUTS, b, c):
        assert maxrel(x.cpu().numpy(), y.cpu().numpy()) <= TOL, k


def test_loud_failures():
    z = np.load(GOLDEN[0])
    pc, *_ = build(z, "cpu", torch.float32)
    with pytest.raises(RuntimeError):
        deformation.get_deformation(pc, 0.4)          # CPU tensors: there is no CPU path
    pc, *_ = build(z, DEV, torch.float32)
    pc.args.dsh = False
    with pytest.raises(deformation.UnsupportedDeformationConfig):
        deformation.get_deformation(pc, 0.4)


def test_get_deformfeature_and_get_intergral_mirrors():
    """The two small callers around the lifespan MLP (scene/saro_gaussian.py:761-777, :863-869) against their PyTorch
    restatement in float64 (same statements; the MLP is what differs: tcgen05 job vs nn.Sequential)."""
    import math
    z = np.load(GOLDEN[0])
    pc, leaves, mlps, _ = build(z, DEV, torch.float32)
    with torch.no_grad():
        deformation.get_deformfeature(pc)
        got_int = deformation.get_intergral(pc, 0.1, 0.8)
    assert torch.equal(pc.hexplane_feature, leaves["hexplane_feature"])
    pc64, leaves64, mlps64, _ = build(z, DEV, torch.float64)
    with torch.no_grad():
        ms = pc64.args.min_interval / pc64.duration
        life = (1 - ms) * (1 - mlps64["opacity"](leaves64["hexplane_feature"])) + ms
        Q = lambda x: 1 - 1 / (1 + torch.exp(0.070565902 * x ** 3 + 1.5976 * x))
        tp = leaves64["temporal_pos"]
        want_int = life * math.sqrt(math.pi) / 2 * (Q(2 * math.sqrt(2) * (0.8 - tp) / life) - Q(2 * math.sqrt(2) * (0.1 - tp) / life))
    assert maxrel(pc._lifespan.cpu().numpy(), life.cpu().numpy()) <= TOL
    assert maxrel(pc._lifespan.cpu().numpy(), z["f64_lifespan"]) <= TOL          # and the reference's own get_deformation value
    assert maxrel(got_int.cpu().numpy(), want_int.cpu().numpy()) <= TOL
    # the cached pair feeds the test-time hand-off
    pc.hexplane_feature = pc.hexplane_feature.detach()
    pc._lifespan = pc._lifespan.detach()
    out = deformation.get_deformation_eval(pc, float(z["timestamp"]))
    assert out[0].shape[0] > 0


def test_empty_cloud():
    """No Gaussians (everything pruned): empty outputs of the right shapes, backward is a no-op that still reaches
    the parameters with zero / absent gradients — the reference's PyTorch ops behave the same way."""
    tensors, weights = _random_case(0, 32, (1, 0, 0), seed=3)
    mlps = deform_torch.make_train_mlps(32, seed=4)
    leaves = {k: v.to(DEV).requires_grad_(True) for k, v in tensors.items()}
    dm = {k: m.to(DEV) for k, m in mlps.items()}
    pc = deform_torch.TrainModelStandIn(leaves, dm, (1, 0, 0), 6.0, 300.0)
    outs = deformation.get_deformation(pc, 0.5)
    assert [tuple(o.shape) for o in outs] == [(0, 3), (0, 4), (0, 3), (0, 1), (0, 16, 3)]
    assert tuple(pc._lifespan.shape) == (0, 1) and tuple(pc.real_xyz.shape) == (0, 3) and tuple(pc.scale_residual.shape) == (0, 3)
    sum(o.sum() for o in outs).backward()
    torch.cuda.synchronize()
    g = dm["shs"][0].weight.grad
    assert g is None or float(g.abs().sum()) == 0.0

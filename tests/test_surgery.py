"""CPU tests (the surgery is device-agnostic host logic over torch tensors): saro_gs_b200.surgery — clone / split /
prune planned as ONE gather per tensor, Adam moments included — must reproduce, bit for bit, what the reference's own
statements produce (tests/golden/surgery_*.npz: scene/saro_gaussian.py:451-454, 540-739 executed via `ast`,
tests/golden/make_golden_surgery.py), from the same inputs and the same RNG seed."""
import types

import numpy as np
import pytest
import torch

from golden_util import load
from saro_gs_b200 import surgery

GROUPS = [("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"),
          ("scaling", "_scaling"), ("rotation", "_rotation"), ("temporal_pos", "_temporal_pos")]


class Model:
    """A GaussianModel stand-in with the reference's attribute names (what surgery.install targets)."""
    scaling_activation, scaling_inverse_activation = staticmethod(torch.exp), staticmethod(torch.log)
    opacity_activation = staticmethod(torch.sigmoid)
    get_scaling = property(lambda self: self.scaling_activation(self._scaling))
    get_opacity = property(lambda self: self.opacity_activation(self._opacity))
    get_xyz = property(lambda self: self._xyz)
    get_temporalpos = property(lambda self: self._temporal_pos)

    def get_intergral(self, start=0.0, end=1.0):      # the stand-in of make_golden_surgery.py
        return 0.02 + 0.5 * torch.sigmoid(3.0 * self.get_temporalpos - 0.2 * self._xyz[:, 2:3])


surgery.install(Model)


def model_from(d):
    m = Model()
    m.args = types.SimpleNamespace(loader=str(d["loader"]), pw=bool(d["pw"]), sigmoid_tcenter=False, rgbdecoder=False)
    m.percent_dense, m.min_intergral = 0.01, 0.1
    for name, attr in GROUPS:
        setattr(m, attr, torch.nn.Parameter(torch.from_numpy(d[f"in_{name}"]).clone()))
    m.optimizer = torch.optim.Adam([{"params": [getattr(m, attr)], "lr": 1e-3, "name": name} for name, attr in GROUPS] +
                                   [{"params": [torch.nn.Parameter(torch.zeros(4, 4))], "lr": 1e-3, "name": "motion_mlp"}],
                                   lr=0.0, eps=1e-15)
    for name, attr in GROUPS:
        m.optimizer.state[getattr(m, attr)] = {"step": torch.tensor(2.0),
                                               "exp_avg": torch.from_numpy(d[f"in_{name}_exp_avg"]).clone(),
                                               "exp_avg_sq": torch.from_numpy(d[f"in_{name}_exp_avg_sq"]).clone()}
    for k in ("xyz_gradient_accum", "t_gradient_accum", "denom", "max_radii2D"):
        setattr(m, k, torch.from_numpy(d[f"in_{k}"]).clone())
    m.inv_intergral_fordensify = torch.from_numpy(d["inv_intergral"]).clone()
    return m


def check(m, d):
    for name, attr in GROUPS:
        p = getattr(m, attr)
        assert isinstance(p, torch.nn.Parameter) and p.requires_grad and p.is_leaf
        assert np.array_equal(p.detach().numpy(), d[f"out_{name}"]), name
        group = [g for g in m.optimizer.param_groups if g["name"] == name][0]
        assert group["params"][0] is p                                   # the optimizer holds the new leaf
        st = m.optimizer.state[p]
        assert np.array_equal(st["exp_avg"].numpy(), d[f"out_{name}_exp_avg"]), name
        assert np.array_equal(st["exp_avg_sq"].numpy(), d[f"out_{name}_exp_avg_sq"]), name
    assert len(m.optimizer.state) == len(GROUPS)                        # no stale state left behind
    for k in ("xyz_gradient_accum", "t_gradient_accum", "denom", "max_radii2D"):
        assert np.array_equal(getattr(m, k).numpy(), d[f"out_{k}"]), k


@pytest.mark.parametrize("name", ["surgery_colmap", "surgery_blender_pw", "surgery_first_rounds"])
def test_densify_pruneclone_bit_exact_vs_reference_statements(name):
    d = load(name)
    m = model_from(d)
    torch.manual_seed(1000 + int(d["seed"]))
    mss = float(d["max_screen_size"]) or None
    m.densify_pruneclone(float(d["max_grad"]), float(d["min_opacity"]), float(d["extent"]), mss)
    if bool(d["then_reset_and_prune"]):
        m.reset_opacity()
        m.prune_points(m._xyz[:, 2] < 8.0)
    check(m, d)
    # the optimizer still steps on the new leaves
    for _, attr in GROUPS:
        getattr(m, attr).grad = torch.ones_like(getattr(m, attr))
    m.optimizer.step()


def test_step_by_step_methods_equal_the_planned_gather():
    """densify_and_clone + densify_and_splitv2 + prune_points (the reference's three rounds) == densify_pruneclone."""
    d = load("surgery_colmap")
    a, b = model_from(d), model_from(d)
    torch.manual_seed(5)
    a.densify_pruneclone(2e-4, 0.005, 5.0, 20)
    torch.manual_seed(5)
    grads = b.xyz_gradient_accum / b.denom
    grads[grads.isnan()] = 0.0
    grads = grads * b.inv_intergral_fordensify
    b.densify_and_clone(grads, 2e-4, 5.0)
    b.densify_and_splitv2(grads, 2e-4, 5.0, 2)
    prune = (b.get_opacity < 0.005).squeeze() | (b.get_intergral() < b.min_intergral).squeeze() | (b._xyz[:, 2] < 4.5)
    b.prune_points(prune)
    for _, attr in GROUPS:
        assert torch.equal(getattr(a, attr), getattr(b, attr)), attr
        assert torch.equal(a.optimizer.state[getattr(a, attr)]["exp_avg"], b.optimizer.state[getattr(b, attr)]["exp_avg"])


def test_prune_everything_and_nothing():
    d = load("surgery_first_rounds")
    m = model_from(d)
    P = m._xyz.shape[0]
    m.prune_points(torch.zeros(P, dtype=torch.bool))
    assert m._xyz.shape[0] == P
    m.prune_points(torch.ones(P, dtype=torch.bool))
    assert m._xyz.shape[0] == 0 and m.denom.shape == (0, 1) and m._features_rest.shape == (0, 15, 3)

#!/usr/bin/env python
"""Generates tests/golden/surgery_*.npz by executing the reference's OWN clone / split / prune statements
(GaussianModel._prune_optimizer, prune_points, cat_tensors_to_optimizer, densification_postfix, densify_and_splitv2,
densify_and_clone, densify_pruneclone, replace_tensor_to_optimizer, reset_opacity — scene/saro_gaussian.py:451-454,
540-739 — plus utils/general_utils.build_rotation / inverse_sigmoid) on the CPU.

saro_gaussian.py cannot be imported here (nvdiffrast / simple_knn are not installed), so the methods are cut out with
`ast` at generation time and executed as they are, except for the device literal "cuda" -> "cpu"; nothing is copied
into this repository.  get_intergral (which needs the plane field) is replaced ON BOTH SIDES by the same closed-form
stand-in, stored by name in the fixture.  Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_surgery.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
MODEL = "/root/reference/scene/saro_gaussian.py"
UTILS = "/root/reference/utils/general_utils.py"
HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = {"_prune_optimizer", "prune_points", "cat_tensors_to_optimizer", "densification_postfix", "densify_and_splitv2",
           "densify_and_clone", "densify_pruneclone", "replace_tensor_to_optimizer", "reset_opacity",
           "get_scaling", "get_opacity", "get_xyz", "get_temporalpos"}
GROUPS = [("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"),
          ("scaling", "_scaling"), ("rotation", "_rotation"), ("temporal_pos", "_temporal_pos")]


class CpuDevice(ast.NodeTransformer):
    def visit_Constant(self, node):
        return ast.copy_location(ast.Constant("cpu"), node) if node.value == "cuda" else node


def reference_class():
    utils = ast.parse(open(UTILS).read())
    helpers = [n for n in utils.body if isinstance(n, ast.FunctionDef) and n.name in ("build_rotation", "inverse_sigmoid")]
    assert len(helpers) == 2
    tree = ast.parse(open(MODEL).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianModel"][0]
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    assert {n.name for n in keep} == METHODS, METHODS - {n.name for n in keep}
    body = helpers + [ast.ClassDef(name="GaussianModel", bases=[], keywords=[], body=keep, decorator_list=[])]
    mod = CpuDevice().visit(ast.Module(body=body, type_ignores=[]))
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "nn": torch.nn, "np": np}
    exec(compile(mod, MODEL, "exec"), ns)
    return ns["GaussianModel"]


def standin_intergral(self, start=0.0, end=1.0):
    """closed-form stand-in for get_intergral (the real one needs the plane field): same on both sides"""
    return 0.02 + 0.5 * torch.sigmoid(3.0 * self.get_temporalpos - 0.2 * self._xyz[:, 2:3])


def make_model(cls, P, seed, loader, pw, percent_dense=0.01):
    g = torch.Generator().manual_seed(seed)
    m = cls.__new__(cls)
    m.args = types.SimpleNamespace(loader=loader, pw=pw, sigmoid_tcenter=False, rgbdecoder=False)
    m.scaling_activation, m.scaling_inverse_activation = torch.exp, torch.log
    m.opacity_activation = torch.sigmoid
    m.percent_dense = percent_dense
    m.min_intergral = 0.1
    z = torch.rand(P, 1, generator=g) * 30 + 3.5
    vals = {"_xyz": torch.cat([torch.randn(P, 2, generator=g) * 3, z], 1),
            "_features_dc": torch.randn(P, 1, 3, generator=g), "_features_rest": torch.randn(P, 15, 3, generator=g) * 0.1,
            "_opacity": torch.randn(P, 1, generator=g) * 2.5, "_scaling": torch.randn(P, 3, generator=g) * 1.2 - 3.0,
            "_rotation": torch.randn(P, 4, generator=g), "_temporal_pos": torch.rand(P, 1, generator=g)}
    for k, v in vals.items():
        setattr(m, k, torch.nn.Parameter(v.clone()))
    m.optimizer = torch.optim.Adam([{"params": [getattr(m, attr)], "lr": 1e-3, "name": name} for name, attr in GROUPS] +
                                   [{"params": [torch.nn.Parameter(torch.randn(4, 4, generator=g))], "lr": 1e-3, "name": "motion_mlp"}],
                                   lr=0.0, eps=1e-15)
    # two Adam steps with random gradients so that every moment is populated
    for _ in range(2):
        for group in m.optimizer.param_groups:
            p = group["params"][0]
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
        m.optimizer.step()
    m.xyz_gradient_accum = torch.rand(P, 1, generator=g) * 4e-4
    m.t_gradient_accum = torch.rand(P, 1, generator=g) * 1e-4
    m.denom = torch.randint(0, 4, (P, 1), generator=g).float()          # zeros -> NaN gradients -> 0 (:706-707)
    m.max_radii2D = torch.rand(P, generator=g) * 40
    m.inv_intergral_fordensify = 1.0 + torch.rand(P, 1, generator=g)
    return m


def snapshot(m, prefix, out):
    for name, attr in GROUPS:
        p = getattr(m, attr)
        out[f"{prefix}_{name}"] = p.detach().numpy().copy()
        st = m.optimizer.state.get(p, None)
        assert st is not None
        out[f"{prefix}_{name}_exp_avg"] = st["exp_avg"].numpy().copy()
        out[f"{prefix}_{name}_exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    for k in ("xyz_gradient_accum", "t_gradient_accum", "denom", "max_radii2D"):
        out[f"{prefix}_{k}"] = getattr(m, k).numpy().copy()


def make_case(name, P, seed, loader, pw, max_screen_size, then_reset_and_prune=False):
    cls = reference_class()
    cls.get_intergral = standin_intergral
    m = make_model(cls, P, seed, loader, pw)
    out = {"seed": np.int64(seed), "P": np.int64(P), "loader": loader, "pw": np.bool_(pw),
           "max_screen_size": np.float64(max_screen_size or 0), "extent": np.float64(5.0), "max_grad": np.float64(2e-4),
           "min_opacity": np.float64(0.005), "inv_intergral": m.inv_intergral_fordensify.numpy().copy(),
           "then_reset_and_prune": np.bool_(then_reset_and_prune)}
    snapshot(m, "in", out)
    torch.manual_seed(1000 + seed)                       # the device RNG the split samples come from
    m.densify_pruneclone(2e-4, 0.005, 5.0, max_screen_size)
    if then_reset_and_prune:
        m.reset_opacity()
        m.prune_points(m._xyz[:, 2] < 8.0)
    snapshot(m, "out", out)
    path = os.path.join(HERE, f"surgery_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", P, "->", m._xyz.shape[0], "Gaussians")


if __name__ == "__main__":
    if not os.path.exists(MODEL):
        sys.exit("needs /root/reference")
    make_case("colmap", P=500, seed=0, loader="colmap", pw=False, max_screen_size=20)
    make_case("blender_pw", P=400, seed=1, loader="blender", pw=True, max_screen_size=20, then_reset_and_prune=True)
    make_case("first_rounds", P=300, seed=2, loader="colmap", pw=False, max_screen_size=None)

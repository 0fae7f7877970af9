#!/usr/bin/env python
"""Generates tests/golden/densify_*.npz by executing the reference's OWN statements for the per-iteration densification
statistics: the per-view `batch_point_grad.append(...)` (train.py:211), the batch reduction block under
`if opt.batch>1:` (train.py:280-292) and GaussianModel.add_densification_stats_grad (scene/saro_gaussian.py:745-747).
train.py is a script and saro_gaussian.py cannot be imported here (nvdiffrast / simple_knn are not installed), so the
statements are cut out with `ast` at generation time and executed as they are, on CPU — nothing is copied into this
repository.  Run in the build container only (needs /root/reference):  python tests/golden/make_golden_densify.py
"""
import ast
import os
import types

import numpy as np
import torch

TRAIN = "/root/reference/train.py"
MODEL = "/root/reference/scene/saro_gaussian.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_statements():
    tree = ast.parse(open(TRAIN).read())
    per_view, block = None, None
    for node in ast.walk(tree):
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Call) and ast.unparse(node.value.func) == "batch_point_grad.append":
            per_view = node
        if isinstance(node, ast.If) and ast.unparse(node.test) == "opt.batch > 1" and node.body and \
                isinstance(node.body[0], ast.Assign) and ast.unparse(node.body[0].targets[0]) == "visibility_count":
            block = node.body
    assert per_view is not None and block is not None
    mtree = ast.parse(open(MODEL).read())
    cls = [n for n in mtree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianModel"][0]
    meth = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "add_densification_stats_grad"]
    assert len(meth) == 1

    def compiled(stmts, name):
        mod = ast.Module(body=stmts, type_ignores=[])
        ast.fix_missing_locations(mod)
        return compile(mod, name, "exec")

    model_cls = ast.ClassDef(name="GaussianModel", bases=[], keywords=[], body=meth, decorator_list=[])
    return compiled([per_view], TRAIN), compiled(block, TRAIN), compiled([model_cls], MODEL)


def make_case(name, P, views, seed, hidden_frac=0.3):
    per_view, block, model_src = reference_statements()
    ns_model = {"torch": torch}
    exec(model_src, ns_model)
    g = torch.Generator().manual_seed(seed)
    gaussians = ns_model["GaussianModel"].__new__(ns_model["GaussianModel"])
    gaussians.gaussian_dim = 3
    gaussians.max_radii2D = (torch.rand(P, generator=g) * 30).floor()
    gaussians.xyz_gradient_accum = torch.rand(P, 1, generator=g) * 1e-3
    gaussians.denom = torch.randint(0, 5, (P, 1), generator=g).float()
    start = {k: getattr(gaussians, k).clone().numpy() for k in ("max_radii2D", "xyz_gradient_accum", "denom")}
    grads, radii_list = [], []
    ns = {"torch": torch, "gaussians": gaussians, "batch_point_grad": [], "batch_radii": [], "batch_visibility_filter": [],
          "opt": types.SimpleNamespace(batch=views)}
    for v in range(views):
        grad = torch.randn(P, 3, generator=g) * 1e-4
        radii = (torch.rand(P, generator=g) * 40).floor().to(torch.int32)
        radii[torch.rand(P, generator=g) < hidden_frac] = 0
        grad[radii == 0] = 0                                  # culled Gaussians receive no gradient
        ns["viewspace_point_tensor"] = types.SimpleNamespace(grad=grad)
        exec(per_view, ns)                                    # train.py:211
        ns["batch_radii"].append(radii)                       # train.py:214
        ns["batch_visibility_filter"].append(radii > 0)       # train.py:215 with renderer/__init__.py:131
        grads.append(grad.numpy())
        radii_list.append(radii.numpy())
    exec(block, ns)                                           # train.py:281-291
    path = os.path.join(HERE, f"densify_{name}.npz")
    np.savez_compressed(path, grads=np.stack(grads), radii=np.stack(radii_list), **{f"start_{k}": v for k, v in start.items()},
                        out_max_radii2D=gaussians.max_radii2D.numpy(), out_xyz_gradient_accum=gaussians.xyz_gradient_accum.numpy(),
                        out_denom=gaussians.denom.numpy())
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; never visible:",
          int((np.stack(radii_list) > 0).sum(0).__eq__(0).sum()))


if __name__ == "__main__":
    make_case("batch4", 5000, 4, 1)
    make_case("batch2_sparse", 1231, 2, 2, hidden_frac=0.8)
    make_case("single_view", 257, 1, 3)

#!/usr/bin/env python
"""Generates tests/golden/deform_*.npz by running the reference's OWN source of GaussianModel.get_deformation_eval
(plus get_survival_state, the get_temporalpos property, get_embedder and Embedder) on CPU, float32 and float64.

The reference module cannot be imported here (it pulls in nvdiffrast / simple_knn, which are not installed), so the
needed definitions are cut out of /root/reference/scene/saro_gaussian.py with `ast` at generation time and executed
as they are — nothing is copied into this repository.  Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_deform.py
"""
import ast
import os
import types

import numpy as np
import torch
from torch import nn

REF = "/root/reference/scene/saro_gaussian.py"
HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = {"get_deformation_eval", "get_survival_state", "get_temporalpos"}


def load_reference_definitions():
    tree = ast.parse(open(REF).read())
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "get_embedder":
            body.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "Embedder":
            body.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "GaussianModel":
            keep = [n for n in node.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
            assert {n.name for n in keep} == METHODS
            body.append(ast.ClassDef(name="GaussianModel", bases=[], keywords=[], body=keep, decorator_list=[]))
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "nn": nn, "np": np}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def make_mlp(in_dim, out_dim, gain, gen):
    m = nn.Sequential(nn.Linear(in_dim, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, out_dim))
    with torch.no_grad():
        for layer in m:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight, gain=gain, generator=gen)
                layer.bias.uniform_(-0.1, 0.1, generator=gen)
    return m


def make_case(name, n, feat_dim, timestamp, seed, life_lo=0.05, life_hi=1.0):
    ns = load_reference_definitions()
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    rn = lambda *s: torch.randn(*s, generator=g)
    embed, out_dim = ns["get_embedder"](4)
    assert out_dim == 9
    inp = dict(xyz=rn(n, 3) * 2, rotation=rn(n, 4), scaling=rn(n, 3) * 0.5 - 3.5, opacity=rn(n, 1) * 2,
               features_dc=rn(n, 1, 3) * 0.5, features_rest=rn(n, 15, 3) * 0.1, temporal_pos=r(n, 1),
               lifespan=r(n, 1) * (life_hi - life_lo) + life_lo, hexplane_feature=rn(n, feat_dim) * 0.5)
    mlps = dict(motion=make_mlp(feat_dim + 9, 3, 1.0, g), rot=make_mlp(feat_dim + 9, 7, 1.0, g),
                shs=make_mlp(feat_dim + 9, 48, 1.5, g))
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        pc = ns["GaussianModel"].__new__(ns["GaussianModel"])
        pc.args = types.SimpleNamespace(dx=True, drot=True, dopacity=True, dsh=True, sigmoid_tcenter=False)
        pc.time_emb = embed
        pc.rotation_activation = torch.nn.functional.normalize
        pc.scaling_activation = torch.exp
        pc.opacity_activation = torch.sigmoid
        pc._xyz, pc._rotation, pc._scaling, pc._opacity = (inp[k].to(dt) for k in ("xyz", "rotation", "scaling", "opacity"))
        pc._features_dc, pc._features_rest = inp["features_dc"].to(dt), inp["features_rest"].to(dt)
        pc._temporal_pos, pc._lifespan = inp["temporal_pos"].to(dt), inp["lifespan"].to(dt)
        pc.hexplane_feature = inp["hexplane_feature"].to(dt)
        pc.motion_mlp, pc.rot_mlp, pc.shs_mlp = (mlps[k].to(dt) for k in ("motion", "rot", "shs"))
        with torch.no_grad():
            motion, rot, scale, opacity, shs = pc.get_deformation_eval(timestamp)
            state = pc.get_survival_state((timestamp - pc.get_temporalpos) / pc._lifespan)
        for k, v in (("means3D", motion), ("rotations", rot), ("scales", scale), ("opacity", opacity), ("shs", shs),
                     ("state", state.reshape(-1))):
            out[f"{tag}_{k}"] = v.numpy()
        for k in mlps:
            mlps[k].to(torch.float32)
    arrays = {f"in_{k}": v.numpy() for k, v in inp.items()}
    for k, m in mlps.items():
        layers = [l for l in m if isinstance(l, nn.Linear)]
        for i, l in enumerate(layers):
            arrays[f"mlp_{k}_W{i + 1}"] = l.weight.detach().numpy()
            arrays[f"mlp_{k}_b{i + 1}"] = l.bias.detach().numpy()
    arrays.update(out)
    arrays["timestamp"] = np.float64(timestamp)
    path = os.path.join(HERE, f"deform_{name}.npz")
    np.savez_compressed(path, **arrays)
    print(name, "selected", out["f32_means3D"].shape[0], "of", n, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    make_case("feat32", 1500, 32, 0.4, 1)
    make_case("feat16_ragged", 333, 16, 0.73, 2)
    make_case("all_alive", 257, 32, 0.5, 3, life_lo=2.0, life_hi=3.0)

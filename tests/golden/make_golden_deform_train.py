#!/usr/bin/env python
"""Generates tests/golden/deformtrain_*.npz by running the reference's OWN source of GaussianModel.get_deformation
(scene/saro_gaussian.py:779-847, plus get_survival_state, the activations' properties, get_embedder / Embedder) on CPU
in float32 and float64, forward AND autograd backward of a fixed scalar objective:

    L = sum_k <w_k, out_k>  (k over means3D, rotations, scales, opacity, shs; w_k stored in the fixture)
        + lambda_scale * ||scale_residual||_2            (helper_train.py:68-69, when scale_reg)
        + lambda_shs * ||shs_residual||_F                (helper_train.py:76, when shs_reg)
        + lambda_motion * ||motion_residual||_F          (helper_train.py:80, when motion_reg)

The reference module cannot be imported here (it pulls in nvdiffrast / simple_knn), so the needed definitions are cut
out of the reference file with `ast` at generation time and executed as they are — nothing is copied into this
repository.  The plane field is replaced by a differentiable table (a leaf tensor returned by `pc.hexplane(...)`): the
sampler has its own fixtures (make_golden_plane.py); what is pinned here is everything downstream of it.
Run in the build container only (needs /root/reference):   python tests/golden/make_golden_deform_train.py
"""
import ast
import os
import types

import numpy as np
import torch
from torch import nn

REF = "/root/reference/scene/saro_gaussian.py"
HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = {"get_deformation", "get_survival_state", "get_temporalpos", "get_scaling", "get_rotation", "get_opacity"}
OUTS = ("means3D", "rotations", "scales", "opacity", "shs")
LEAVES = ("xyz", "rotation", "scaling", "opacity", "features_dc", "features_rest", "temporal_pos", "hexplane_feature")
MLPS = ("motion", "rot", "shs", "opacity")


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
            assert {n.name for n in keep} == METHODS, {n.name for n in keep}
            body.append(ast.ClassDef(name="GaussianModel", bases=[], keywords=[], body=keep, decorator_list=[]))
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "nn": nn, "np": np}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def make_mlp(in_dim, hid2, out_dim, gain, gen, sigmoid=False):
    layers = [nn.Linear(in_dim, 128), nn.ReLU(), nn.Linear(128, hid2), nn.ReLU(), nn.Linear(hid2, out_dim)]
    if sigmoid:
        layers.append(nn.Sigmoid())
    m = nn.Sequential(*layers)
    with torch.no_grad():
        for layer in m:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight, gain=gain, generator=gen)
                layer.bias.uniform_(-0.1, 0.1, generator=gen)
    return m


def make_case(name, n, feat_dim, timestamp, seed, scale_reg, shs_reg, motion_reg, min_interval=6.0, duration=300.0):
    ns = load_reference_definitions()
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    rn = lambda *s: torch.randn(*s, generator=g)
    embed, out_dim = ns["get_embedder"](4)
    assert out_dim == 9
    inp = dict(xyz=rn(n, 3) * 2, rotation=rn(n, 4), scaling=rn(n, 3) * 0.5 - 3.5, opacity=rn(n, 1) * 2,
               features_dc=rn(n, 1, 3) * 0.5, features_rest=rn(n, 15, 3) * 0.1, temporal_pos=r(n, 1),
               hexplane_feature=rn(n, feat_dim) * 0.5)
    mlps = dict(motion=make_mlp(feat_dim + 9, 128, 3, 1.0, g), rot=make_mlp(feat_dim + 9, 128, 7, 1.0, g),
                shs=make_mlp(feat_dim + 9, 128, 48, 1.5, g), opacity=make_mlp(feat_dim, 64, 1, 2.0, g, sigmoid=True))
    weights = dict(means3D=rn(n, 3), rotations=rn(n, 4), scales=rn(n, 3) * 20, opacity=rn(n, 1), shs=rn(n, 16, 3))
    lambdas = dict(scale=0.3, shs=0.2, motion=0.1)
    arrays = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        pc = ns["GaussianModel"].__new__(ns["GaussianModel"])
        pc.args = types.SimpleNamespace(dx=True, drot=True, dopacity=True, dsh=True, sigmoid_tcenter=False, scale_reg=scale_reg,
                                        shs_reg=shs_reg, motion_reg=motion_reg, min_interval=min_interval)
        pc.duration = duration
        pc.time_emb = embed
        pc.rotation_activation = torch.nn.functional.normalize
        pc.scaling_activation = torch.exp
        pc.opacity_activation = torch.sigmoid
        leaves = {k: inp[k].to(dt).clone().requires_grad_(True) for k in LEAVES}
        pc._xyz, pc._rotation, pc._scaling, pc._opacity = (leaves[k] for k in ("xyz", "rotation", "scaling", "opacity"))
        pc._features_dc, pc._features_rest, pc._temporal_pos = leaves["features_dc"], leaves["features_rest"], leaves["temporal_pos"]
        pc.hexplane = lambda xyz, t, s: leaves["hexplane_feature"]
        for k in MLPS:
            mlps[k].to(dt)
            mlps[k].zero_grad()
        pc.motion_mlp, pc.rot_mlp, pc.shs_mlp, pc.opacity_mlp = (mlps[k] for k in MLPS)
        outs = pc.get_deformation(timestamp)
        loss = sum((weights[k].to(dt) * o).sum() for k, o in zip(OUTS, outs))
        if scale_reg:
            loss = loss + lambdas["scale"] * torch.linalg.vector_norm(pc.scale_residual, ord=2)
        if shs_reg:
            loss = loss + lambdas["shs"] * torch.linalg.matrix_norm(pc.shs_residual.reshape(n, -1))
        if motion_reg:
            loss = loss + lambdas["motion"] * torch.linalg.matrix_norm(pc.motion_residual)
        loss.backward()
        for k, o in zip(OUTS, outs):
            arrays[f"{tag}_{k}"] = o.detach().numpy()
        arrays[f"{tag}_lifespan"] = pc._lifespan.detach().numpy()
        arrays[f"{tag}_real_xyz"] = pc.real_xyz.detach().numpy()
        arrays[f"{tag}_loss"] = np.float64(loss.item())
        if scale_reg:
            arrays[f"{tag}_scale_residual"] = pc.scale_residual.detach().numpy()
        if tag == "f64":          # gradients: the float64 run only (stored as float32 values of the float64 result)
            for k in LEAVES:
                arrays[f"{tag}_grad_{k}"] = leaves[k].grad.numpy().astype(np.float32)
            for k in MLPS:
                for i, l in enumerate([l for l in mlps[k] if isinstance(l, nn.Linear)]):
                    arrays[f"{tag}_grad_{k}_W{i + 1}"] = l.weight.grad.numpy().astype(np.float32)
                    arrays[f"{tag}_grad_{k}_b{i + 1}"] = l.bias.grad.numpy().astype(np.float32)
        for k in MLPS:
            mlps[k].to(torch.float32)
    arrays.update({f"in_{k}": v.numpy() for k, v in inp.items()})
    arrays.update({f"w_{k}": v.numpy() for k, v in weights.items()})
    for k in MLPS:
        for i, l in enumerate([l for l in mlps[k] if isinstance(l, nn.Linear)]):
            arrays[f"mlp_{k}_W{i + 1}"] = l.weight.detach().numpy()
            arrays[f"mlp_{k}_b{i + 1}"] = l.bias.detach().numpy()
    arrays["timestamp"] = np.float64(timestamp)
    arrays["flags"] = np.array([scale_reg, shs_reg, motion_reg], dtype=np.int32)
    arrays["lambdas"] = np.array([lambdas["scale"], lambdas["shs"], lambdas["motion"]])
    arrays["min_interval"], arrays["duration"] = np.float64(min_interval), np.float64(duration)
    path = os.path.join(HERE, f"deformtrain_{name}.npz")
    np.savez_compressed(path, **arrays)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    make_case("feat32_scale_reg", 700, 32, 0.4, 11, True, False, False)
    make_case("feat16_all_regs", 333, 16, 0.73, 12, True, True, True)
    make_case("feat32_no_regs", 257, 32, 0.5, 13, False, False, False)

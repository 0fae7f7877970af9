#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the UNMODIFIED reference rasterizer
(compiled by oracle/build_ref.py into oracle/_ref/) on a B200:

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/             # then commit

The reference ships no tests or golden vectors; these files are the pin for the
CPU oracle (tests/test_oracle_golden.py, runs without a GPU) and for the CUDA path
(tests/test_parity_gpu.py).  Inputs are stored next to the outputs so the fixtures do not depend
on torch's RNG streams.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from saro_gs_b200 import synthetic  # noqa: E402
from saro_gs_b200.rasterizer import GaussianRasterizationSettings  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def cov3d_from_scale_rot(scales, rotations, mod=1.0):
    """float64 CPU: Sigma = R S^2 R^T, upper triangle (for the cov3D_precomp case)."""
    q = rotations.double()
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
    S = torch.diag_embed((scales.double() * mod) ** 2)
    Sig = R @ S @ R.transpose(1, 2)
    return torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], dim=1).float()


CASES = {
    # name: (scene kwargs, bg, sh_degree, mode, scale_modifier)
    "small_sh3": (dict(P=512, seed=0), (0.3, 0.5, 0.7), 3, "sh", 1.0),
    "small_sh1_black": (dict(P=384, seed=1, width=80, height=64), (0.0, 0.0, 0.0), 1, "sh", 1.0),
    "small_sh0_white": (dict(P=300, seed=4, width=64, height=48), (1.0, 1.0, 1.0), 0, "sh", 1.0),
    "small_precomp_color": (dict(P=400, seed=2), (0.1, 0.1, 0.1), 0, "color", 0.7),
    "small_precomp_cov": (dict(P=400, seed=3, width=70, height=50), (0.0, 0.2, 0.0), 2, "cov", 1.0),
    "small_big_splats": (dict(P=256, seed=5, width=64, height=64, log_scale_mean=-0.8), (0.0, 0.0, 0.0), 3, "sh", 1.0),
}


def run_case(RefRast, dev, name, spec):
    kw, bg, deg, mode, mod = spec
    scene, cam = synthetic.small_scene(**kw)
    bg_t = torch.tensor(bg, dtype=torch.float32)
    rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, bg_t.to(dev), mod,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), deg, cam.campos.to(dev), False)
    gen = torch.Generator().manual_seed(100 + kw["seed"])
    cot = torch.randn(3, cam.height, cam.width, generator=gen).float()
    inputs = dict(means3D=scene.means3D, opacities=scene.opacities)
    if mode == "color":
        inputs["colors_precomp"] = torch.rand(scene.means3D.shape[0], 3, generator=gen).float()
    else:
        inputs["shs"] = scene.shs
    if mode == "cov":
        inputs["cov3D_precomp"] = cov3d_from_scale_rot(scene.scales, scene.rotations)
    else:
        inputs["scales"] = scene.scales
        inputs["rotations"] = scene.rotations
    leaves = {k: v.to(dev).clone().requires_grad_(True) for k, v in inputs.items()}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = RefRast(rs)(means2D=means2D, **leaves)
    color.backward(cot.to(dev))
    torch.cuda.synchronize()
    # integer state from the reference's opaque buffers
    refC = ref_loader.load_ref_C()
    e = torch.Tensor([])
    with torch.no_grad():
        args = (rs.bg, leaves["means3D"], leaves.get("colors_precomp", e), leaves["opacities"],
                leaves.get("scales", e), leaves.get("rotations", e), mod, leaves.get("cov3D_precomp", e),
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, cam.height, cam.width, leaves.get("shs", e),
                deg, rs.campos, False)
        R, _c, _r, gb, bb, ib, _d = refC.rasterize_gaussians(*args)
        g = ref_loader.parse_ref_geom(gb, scene.means3D.shape[0])
        im = ref_loader.parse_ref_img(ib, cam.width * cam.height)
        pl = ref_loader.parse_ref_binning(bb, R)["point_list"]
    tiles = ((cam.width + 15) // 16) * ((cam.height + 15) // 16)
    out = dict(
        width=cam.width, height=cam.height, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.array(bg, np.float32),
        sh_degree=deg, scale_modifier=mod, viewmatrix=cam.viewmatrix.numpy(), projmatrix=cam.projmatrix.numpy(),
        campos=cam.campos.numpy(), cotangent=cot.numpy(),
        out_color=color.detach().cpu().numpy(), out_depth=depth.detach().cpu().numpy(),
        out_radii=radii.cpu().numpy(), num_rendered=np.int64(R),
        tiles_touched=g["tiles_touched"].cpu().numpy(), n_contrib=im["n_contrib"].cpu().numpy(),
        final_T=im["accum_alpha"].cpu().numpy(), ranges=im["ranges"][:tiles].cpu().numpy(),
        point_list=pl.cpu().numpy(), grad_means2D=means2D.grad.cpu().numpy())
    for k, v in inputs.items():
        out["in_" + k] = v.numpy()
        out["grad_" + k] = leaves[k].grad.cpu().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "R =", R, "visible =", int((radii > 0).sum()), flush=True)


def run_big(RefRast, dev, name, scene, cam, crop=64):
    """Full-size configs: hashes of the integer/float outputs + a crop + per-channel sums."""
    rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy,
                                       torch.zeros(3, device=dev), 1.0, cam.viewmatrix.to(dev),
                                       cam.projmatrix.to(dev), scene.sh_degree, cam.campos.to(dev), False)
    refC = ref_loader.load_ref_C()
    e = torch.Tensor([])
    with torch.no_grad():
        args = (rs.bg, scene.means3D.to(dev), e, scene.opacities.to(dev), scene.scales.to(dev),
                scene.rotations.to(dev), 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, cam.height,
                cam.width, scene.shs.to(dev), scene.sh_degree, rs.campos, False)
        R, color, radii, gb, bb, ib, depth = refC.rasterize_gaussians(*args)
        P = scene.means3D.shape[0]
        g = ref_loader.parse_ref_geom(gb, P)
        im = ref_loader.parse_ref_img(ib, cam.width * cam.height)
    h0, w0 = cam.height // 2 - crop // 2, cam.width // 2 - crop // 2
    out = dict(num_rendered=np.int64(R), P=np.int64(P), width=cam.width, height=cam.height,
               sha_radii=sha(radii), sha_tiles_touched=sha(g["tiles_touched"]), sha_n_contrib=sha(im["n_contrib"]),
               sha_color=sha(color), sha_depth=sha(depth),
               visible=np.int64((radii > 0).sum().item()),
               color_sum=color.double().sum(dim=(1, 2)).cpu().numpy(),
               crop_origin=np.array([h0, w0]), color_crop=color[:, h0:h0 + crop, w0:w0 + crop].cpu().numpy(),
               depth_crop=depth[:, h0:h0 + crop, w0:w0 + crop].cpu().numpy(),
               n_contrib_sum=np.int64(im["n_contrib"].long().sum().item()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "R =", R, "visible =", int(out["visible"]), flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    dev = torch.device("cuda:0")
    RefRast = ref_loader.ref_api()[1]
    for name, spec in CASES.items():
        run_case(RefRast, dev, name, spec)
    run_big(RefRast, dev, "config1_fwd", *synthetic.config1_scene())
    run_big(RefRast, dev, "config2_fwd", *synthetic.config2_scene())


if __name__ == "__main__":
    main()

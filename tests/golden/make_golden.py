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


SEQ_TIMES = (0.0, 0.37, 0.74, 0.99)


def run_seq(RefRast, dev, name):
    """BASELINE.json configs[2] stand-in: frames of the config-2 cloud under per-Gaussian temporal survival
    (P varies per frame) — colour pass + the test-time second pass with precomputed colours
    (renderer/__init__.py:212-226 of the reference renders `lifespan.expand(-1, 3)` with shs=None)."""
    base, cam = synthetic.config2_scene()
    rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), base.sh_degree,
                                       cam.campos.to(dev), False)
    out = dict(times=np.array(SEQ_TIMES, np.float64))
    refC = ref_loader.load_ref_C()
    e = torch.Tensor([])
    with torch.no_grad():
        for k, t in enumerate(SEQ_TIMES):
            sc = synthetic.temporal_frame(base, t)
            P = sc.means3D.shape[0]
            m3, op, scl, rot, sh = (x.to(dev) for x in (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.shs))
            R, color, radii, gb, bb, ib, depth = refC.rasterize_gaussians(
                rs.bg, m3, e, op, scl, rot, 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, cam.height,
                cam.width, sh, sc.sh_degree, rs.campos, False)
            life = op.expand(-1, 3).contiguous()
            R2, color2, radii2, _, _, _, depth2 = refC.rasterize_gaussians(
                rs.bg, m3, life, op, scl, rot, 1.0, e, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                cam.height, cam.width, e, 0, rs.campos, False)
            out[f"P_{k}"] = np.int64(P)
            out[f"R_{k}"] = np.int64(R)
            out[f"sha_color_{k}"] = sha(color)
            out[f"sha_depth_{k}"] = sha(depth)
            out[f"sha_radii_{k}"] = sha(radii)
            out[f"sha_color_life_{k}"] = sha(color2)
            print(name, "t =", t, "P =", P, "R =", R, flush=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def run_seq300(RefRast, dev, name, frames=300):
    """BASELINE.json configs[2] at its stated size: a 300-frame sequence (timestamp = frame / 300, like the reference's
    Camera.timestamp for a duration of 300), inputs from synthetic.DeviceSequence; per frame the alive count, the
    reference's num_rendered and SHA-256 of colour, depth and radii."""
    base, cam = synthetic.config2_scene()
    rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), base.sh_degree,
                                       cam.campos.to(dev), False)
    seq = synthetic.DeviceSequence(base, dev)
    refC = ref_loader.load_ref_C()
    e = torch.Tensor([])
    P, R, sc_, sd_, sr_ = [], [], [], [], []
    with torch.no_grad():
        for k in range(frames):
            sc = seq.frame(k / frames)
            r, color, radii, gb, bb, ib, depth = refC.rasterize_gaussians(
                rs.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, cam.height, cam.width, sc.shs, sc.sh_degree, rs.campos, False)
            P.append(sc.means3D.shape[0])
            R.append(r)
            sc_.append(sha(color))
            sd_.append(sha(depth))
            sr_.append(sha(radii))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), frames=np.int64(frames), P=np.array(P, np.int64),
                        R=np.array(R, np.int64), sha_color=np.array(sc_), sha_depth=np.array(sd_), sha_radii=np.array(sr_))
    print(name, "frames", frames, "P", min(P), "..", max(P), "R", min(R), "..", max(R), flush=True)


def run_big_grad(RefRast, dev, name, scene, cam, n_samples=4096):
    """Full-size backward of the reference (configs[1], standard cotangent): per-tensor norms, max |.| and a
    seeded sample of entries.  The reference sums with float atomics, so these are reproducible to ~1e-6 only."""
    rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                       cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                       cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).clone().requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = RefRast(rs)(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                      shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(synthetic.cotangent(cam.height, cam.width).to(dev))
    torch.cuda.synchronize()
    grads = {k: v.grad for k, v in leaves.items()}
    grads["means2D"] = means2D.grad
    out = {}
    gen = torch.Generator().manual_seed(1234)
    for k, g in grads.items():
        flat = g.detach().reshape(-1).cpu()
        idx = torch.randint(0, flat.numel(), (n_samples,), generator=gen)
        # half of the sample among the largest entries, so the check is not dominated by zeros
        top = torch.topk(flat.abs(), n_samples // 2).indices
        idx[: n_samples // 2] = top
        out[f"idx_{k}"] = idx.numpy()
        out[f"val_{k}"] = flat[idx].numpy()
        out[f"norm_{k}"] = np.float64(flat.double().norm().item())
        out[f"maxabs_{k}"] = np.float64(flat.abs().max().item())
        out[f"sum_{k}"] = np.float64(flat.double().sum().item())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: float(out[f"norm_{k}"]) for k in grads}, flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    dev = torch.device("cuda:0")
    RefRast = ref_loader.ref_api()[1]
    if len(sys.argv) > 1 and sys.argv[1] == "seq300":       # only the 300-frame sequence (round 2 addition)
        run_seq300(RefRast, dev, "config3_seq300")
        return
    for name, spec in CASES.items():
        run_case(RefRast, dev, name, spec)
    run_big(RefRast, dev, "config1_fwd", *synthetic.config1_scene())
    run_big(RefRast, dev, "config2_fwd", *synthetic.config2_scene())
    run_big_grad(RefRast, dev, "config2_bwd", *synthetic.config2_scene())
    run_seq(RefRast, dev, "config3_seq")
    run_seq300(RefRast, dev, "config3_seq300")


if __name__ == "__main__":
    main()

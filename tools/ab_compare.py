#!/usr/bin/env python
"""Developer tool (GPU box): A/B the native rasterizer against the compiled reference
(oracle/_ref) on the synthetic configs — integer state bit-equality, float parity, timings.
Writes a JSON summary to gpurun_out/ab_compare.json.  Not part of the product path."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import saro_gs_b200 as sgs  # noqa: E402
from saro_gs_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402


def settings_for(cam, bg, sh_degree, dev):
    return sgs.GaussianRasterizationSettings(
        image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=bg.to(dev), scale_modifier=1.0, viewmatrix=cam.viewmatrix.to(dev), projmatrix=cam.projmatrix.to(dev),
        sh_degree=sh_degree, campos=cam.campos.to(dev), prefiltered=False)


def run(Rast, scene, rs, cot, dev):
    leaves = {}
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        leaves[k] = getattr(scene, k).to(dev).clone().requires_grad_(True)
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    rast = Rast(rs)
    color, radii, depth = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                               shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(cot)
    grads = {k: v.grad for k, v in leaves.items()}
    grads["means2D"] = means2D.grad
    return color.detach(), radii, depth.detach(), grads


def relerr(a, b):
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    return d, d / max(s, 1e-30)


def normrel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def timeit(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return {"median_ms": ts[len(ts) // 2], "min_ms": ts[0], "mean_ms": sum(ts) / len(ts)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="small,c1,c2")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--oracle", action="store_true", help="arbitrate gradients with the float64 CPU oracle")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    have_ref = ref_loader.available() and not args.no_ref
    RefRast = ref_loader.ref_api()[1] if have_ref else None
    refC = ref_loader.load_ref_C() if have_ref else None
    out = {}
    for name in args.configs.split(","):
        if name == "small":
            scene, cam = synthetic.small_scene()
        elif name == "c1":
            scene, cam = synthetic.config1_scene()
        else:
            scene, cam = synthetic.config2_scene()
        bg = torch.tensor([0.0, 0.0, 0.0]) if name != "small" else torch.tensor([0.3, 0.5, 0.7])
        rs = settings_for(cam, bg, scene.sh_degree, dev)
        cot = synthetic.cotangent(cam.height, cam.width).to(dev)
        res = {"P": int(scene.means3D.shape[0]), "W": cam.width, "H": cam.height}

        c_n, r_n, d_n, g_n = run(sgs.GaussianRasterizer, scene, rs, cot, dev)
        torch.cuda.synchronize()
        res["native_finite"] = bool(torch.isfinite(c_n).all().item())
        # exactness of tile-cull: compare against native run with tile culling disabled
        with torch.no_grad():
            args_fw = (rs.bg, scene.means3D.to(dev), torch.Tensor([]), scene.opacities.to(dev), scene.scales.to(dev),
                       scene.rotations.to(dev), 1.0, torch.Tensor([]), rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                       rs.tanfovy, cam.height, cam.width, scene.shs.to(dev), scene.sh_degree, rs.campos, False)
            nR, col_nc, rad_nc, gb, bb, ib, dep_nc = sgs._C.rasterize_gaussians(*args_fw, keep_for_backward=True,
                                                                              _no_tile_cull=True)
            nR2, col_c, rad_c, gb2, bb2, ib2, dep_c = sgs._C.rasterize_gaussians(*args_fw, keep_for_backward=True)
            st_nc = sgs._C.debug_export(res["P"], cam.width, cam.height, nR, gb, bb, ib)
            st_c = sgs._C.debug_export(res["P"], cam.width, cam.height, nR2, gb2, bb2, ib2)
            res["num_rendered"] = nR
            res["cull_exact_color"] = bool(torch.equal(col_nc, col_c))
            res["cull_exact_depth"] = bool(torch.equal(dep_nc, dep_c))
            res["cull_exact_ncontrib"] = bool(torch.equal(st_nc["n_contrib"], st_c["n_contrib"]))
            res["packed_fraction"] = float(st_c["tile_count"].sum().item()) / max(1, nR)
            res["packed_fraction_nocull"] = float(st_nc["tile_count"].sum().item()) / max(1, nR)

        if have_ref:
            c_r, r_r, d_r, g_r = run(RefRast, scene, rs, cot, dev)
            torch.cuda.synchronize()
            with torch.no_grad():
                rR, rcol, rrad, rgb_, rbb, rib, rdep = refC.rasterize_gaussians(*args_fw)
                rg = ref_loader.parse_ref_geom(rgb_, res["P"])
                ri = ref_loader.parse_ref_img(rib, cam.width * cam.height)
                rb = ref_loader.parse_ref_binning(rbb, rR)
                vis = rrad > 0
                res["ref_num_rendered"] = rR
                res["eq_num_rendered"] = bool(rR == nR2)
                res["eq_radii"] = bool(torch.equal(rrad, rad_c))
                res["eq_tiles_touched"] = bool(torch.equal(rg["tiles_touched"], st_c["tiles_touched"]))
                ntiles = st_c["ranges"].shape[0]
                res["eq_ranges"] = bool(torch.equal(ri["ranges"][:ntiles], st_c["ranges"]))
                res["eq_point_list"] = bool(rR == nR2 and torch.equal(rb["point_list"], st_c["point_list"]))
                res["eq_n_contrib"] = bool(torch.equal(ri["n_contrib"], st_c["n_contrib"]))
                res["n_contrib_mismatch"] = int((ri["n_contrib"] != st_c["n_contrib"]).sum().item())
                res["eq_final_T"] = bool(torch.equal(ri["accum_alpha"], st_c["final_T"]))
                res["biteq_means2D"] = bool(torch.equal(rg["means2D"][vis], st_c["means2D"][vis]))
                res["biteq_conic_opacity"] = bool(torch.equal(rg["conic_opacity"][vis], st_c["conic_opacity"][vis]))
                res["biteq_cov3D"] = bool(torch.equal(rg["cov3D"][vis], st_c["cov3D"][vis]))
                res["biteq_rgb"] = bool(torch.equal(rg["rgb"][vis], st_c["rgbd"][vis][:, :3]))
                res["rgb_maxdiff"] = (rg["rgb"][vis] - st_c["rgbd"][vis][:, :3]).abs().max().item()
                res["biteq_color"] = bool(torch.equal(rcol, col_c))
                res["biteq_depth"] = bool(torch.equal(rdep, dep_c))
            res["color_maxabs_rel"] = relerr(c_n, c_r)
            res["depth_mismatch_px"] = int((d_n != d_r).sum().item())
            res["grads"] = {k: {"normrel": normrel(g_n[k], g_r[k]), "maxabs_rel": relerr(g_n[k], g_r[k])}
                            for k in g_n}

        if args.oracle:
            import numpy as np
            from oracle import oracle
            t0 = time.time()
            orc = oracle.forward_scene(scene, cam, bg, precision="f64")
            og = orc.backward(cot.cpu())
            res["oracle_seconds"] = time.time() - t0
            res["oracle_num_rendered"] = orc.num_rendered
            res["oracle_radii_mismatch"] = int((orc.radii != r_n.cpu().numpy()).sum())
            res["oracle_color_maxabs"] = float(np.abs(orc.color - c_n.cpu().numpy()).max())
            res["oracle_depth_mismatch_px"] = int((orc.depth.astype(np.float32) != d_n.cpu().numpy()).sum())

            def nrel(t, o):
                t = t.cpu().numpy().astype(np.float64).reshape(o.shape)
                return float(np.linalg.norm(t - o) / max(np.linalg.norm(o), 1e-300))
            res["grads_vs_oracle_native"] = {k: nrel(g_n[k], og[k]) for k in g_n}
            if have_ref:
                res["grads_vs_oracle_ref"] = {k: nrel(g_r[k], og[k]) for k in g_r}

        # timings (fwd+bwd through the Python API)
        leaves = {k: getattr(scene, k).to(dev).clone().requires_grad_(True)
                  for k in ("means3D", "scales", "rotations", "opacities", "shs")}
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

        def step(Rast):
            rast = Rast(rs)
            color, radii, depth = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                       shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
            color.backward(cot)

        def fwd_only(Rast):
            with torch.no_grad():
                rast = Rast(rs)
                rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                     shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])

        res["t_native_fwdbwd"] = timeit(lambda: step(sgs.GaussianRasterizer), 5, args.iters)
        res["t_native_fwd"] = timeit(lambda: fwd_only(sgs.GaussianRasterizer), 5, args.iters)
        if have_ref:
            res["t_ref_fwdbwd"] = timeit(lambda: step(RefRast), 5, args.iters)
            res["t_ref_fwd"] = timeit(lambda: fwd_only(RefRast), 5, args.iters)
            res["speedup_fwdbwd"] = res["t_ref_fwdbwd"]["median_ms"] / res["t_native_fwdbwd"]["median_ms"]
            res["speedup_fwd"] = res["t_ref_fwd"]["median_ms"] / res["t_native_fwd"]["median_ms"]
        out[name] = res
        print(name, json.dumps(res, indent=1), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ab_compare.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Float64 arbiter for BASELINE.json configs[1] at FULL size (P = 300 000, 1352x1014): runs the C oracle
(oracle/splat_oracle.c, -DORACLE_REAL=double) forward + backward on the CPU and stores, per gradient tensor,

    norm_<k>, maxabs_<k>           float64 2-norm and largest |entry| of the whole tensor
    idx_<k>, val_<k>               a 32 768-entry stratified sample: the 16 384 largest-magnitude entries plus
                                   16 384 entries drawn uniformly (seeded) from the rest
    colsum_<k>                     column sums of the tensor viewed as [P, C] (a whole-tensor checksum)
    referr_max_<k>, referr_norm_<k>  error of the compiled REFERENCE against this arbiter, measured on a B200
                                   (merged from gpurun_out/grad_vs_f64.json, produced by tools/grad_vs_f64.py)

    python tests/golden/make_golden_f64.py            # CPU only, ~1-2 minutes; writes tests/golden/config2_bwd_f64.npz

The VERDICT of round 1 asked for this pin: the compiled reference does not reproduce its own scale / rotation
gradients to 1e-4 at this size (float atomics in arbitrary order), so the float64 restatement is the arbiter for
`tests/test_parity_gpu.py::test_full_size_backward_vs_f64_oracle`.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from saro_gs_b200 import synthetic  # noqa: E402

N_TOP = 16384
N_RAND = 16384


def main():
    oracle.build()
    scene, cam = synthetic.config2_scene()
    cot = synthetic.cotangent(cam.height, cam.width)
    t0 = time.time()
    r = oracle.forward_scene(scene, cam, torch.zeros(3), precision="f64")
    g = r.backward(cot)
    print(f"f64 oracle forward+backward: {time.time() - t0:.1f} s, num_rendered {r.num_rendered}")
    out = {"num_rendered": np.int64(r.num_rendered), "visible": np.int64((r.radii > 0).sum())}
    rng = np.random.default_rng(2024)
    for k in ("means3D", "means2D", "scales", "rotations", "opacities", "shs"):
        a = np.asarray(g[k], dtype=np.float64)
        cols = a.reshape(a.shape[0], -1)
        flat = a.reshape(-1)
        top = np.argpartition(np.abs(flat), -N_TOP)[-N_TOP:]
        mask = np.ones(flat.size, dtype=bool)
        mask[top] = False
        rest = np.flatnonzero(mask)
        rnd = rng.choice(rest, size=min(N_RAND, rest.size), replace=False)
        idx = np.sort(np.concatenate([top, rnd])).astype(np.int64)
        out[f"idx_{k}"] = idx.astype(np.int32)
        out[f"val_{k}"] = flat[idx]
        out[f"norm_{k}"] = np.float64(np.linalg.norm(flat))
        out[f"maxabs_{k}"] = np.float64(np.abs(flat).max())
        out[f"colsum_{k}"] = cols.sum(axis=0)
        print(k, a.shape, "norm", out[f"norm_{k}"], "max", out[f"maxabs_{k}"])
    # How far the compiled, unmodified REFERENCE sits from this float64 arbiter on a B200 (tools/grad_vs_f64.py run
    # through gpurun, two runs, worst of the two): the bar the native kernels are held to where float32 itself
    # cannot reach 1e-4 (the reference's own expression trees lose that much on ill-conditioned Gaussians).
    ref_json = os.path.join(ROOT, "gpurun_out", "grad_vs_f64.json")
    if os.path.exists(ref_json):
        import json
        ref = json.load(open(ref_json)).get("reference", {})
        for k, runs in ref.items():
            out[f"referr_max_{k}"] = np.float64(max(r["max"] for r in runs))
            out[f"referr_norm_{k}"] = np.float64(max(r["norm"] for r in runs))
            print("reference vs f64", k, out[f"referr_max_{k}"], out[f"referr_norm_{k}"])
    # forward image statistics of the float64 run (the forward image itself is pinned bit-exactly elsewhere)
    out["color_mean"] = np.float64(r.color.mean())
    path = os.path.join(ROOT, "tests", "golden", "config2_bwd_f64.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/loss_*.npz by running the reference's OWN Python loss functions
(/root/reference/utils/loss_utils.py: l1_loss, ssim) on the CPU, in float32 and float64, with autograd gradients.

    python tests/golden/make_golden_loss.py            # in the build container (needs /root/reference)

The reference module imports torchmetrics at load time (for an unrelated MS-SSIM helper); the package is absent in
this image, so an empty stand-in module is registered before the import — l1_loss / ssim do not touch it.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SARO_REFERENCE_ROOT", "/root/reference")

stub = types.ModuleType("torchmetrics")
stub.MultiScaleStructuralSimilarityIndexMeasure = lambda **kw: None
sys.modules.setdefault("torchmetrics", stub)
sys.path.insert(0, REF)
from utils import loss_utils as ref  # noqa: E402

CASES = {
    # name: (shape, seed, kind)
    "loss_chw_ragged": ((3, 37, 53), 0, "noise"),          # smaller than two tiles, ragged edges
    "loss_chw_tiles": ((3, 70, 96), 1, "smooth"),          # several 32x32 tiles, exact multiple in W
    "loss_batched": ((2, 3, 33, 45), 2, "noise"),          # [B,C,H,W] with size_average=False
    "loss_identical": ((3, 40, 40), 3, "identical"),       # img == gt: ssim = 1, |x-y| gradient = 0
}


def make(shape, seed, kind):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(shape, generator=g, dtype=torch.float64)
    if kind == "smooth":
        yy, xx = torch.meshgrid(torch.linspace(0, 3, shape[-2], dtype=torch.float64),
                                torch.linspace(0, 4, shape[-1], dtype=torch.float64), indexing="ij")
        gt = 0.5 + 0.4 * torch.sin(xx * 2.0 + yy)[None].expand(shape).clone()
    if kind == "identical":
        img = gt.clone()
    else:
        img = (gt + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)).clamp(0, 1)
    return img, gt


def main():
    for name, (shape, seed, kind) in CASES.items():
        img64, gt64 = make(shape, seed, kind)
        out = {"img": img64.float().numpy(), "gt": gt64.float().numpy()}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            x = img64.float().to(dt).clone().requires_grad_(True)      # identical float32-representable inputs
            y = gt64.float().to(dt)
            l1 = ref.l1_loss(x, y)
            s = ref.ssim(x, y)
            loss = 0.8 * l1 + 0.2 * (1.0 - s)
            loss.backward()
            out[f"l1_{tag}"] = np.float64(l1.item())
            out[f"ssim_{tag}"] = np.float64(s.item())
            out[f"loss_{tag}"] = np.float64(loss.item())
            out[f"grad_{tag}"] = x.grad.numpy().astype(np.float64)
            if len(shape) == 4:
                out[f"ssim_per_image_{tag}"] = ref.ssim(x.detach(), y, size_average=False).numpy().astype(np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, shape, "l1", out["l1_f64"], "ssim", out["ssim_f64"])


if __name__ == "__main__":
    main()

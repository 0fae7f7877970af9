"""TEST / BASELINE INFRASTRUCTURE — the reference's photometric loss as plain PyTorch ops (what SaRO-GS runs today:
five depthwise conv2d calls + elementwise ops + autograd), restated from the maths of utils/loss_utils.py:18-68 so
that it can be TIMED on the GPU box, where /root/reference does not exist.  Used only by bench.py's `loss_path`
baseline leg and tests; never by the product path."""
import torch
import torch.nn.functional as F

from .ssim_oracle import window_1d


def torch_l1_dssim_loss(image, gt, lambda_dssim=0.2):
    ch = image.shape[-3]
    g = torch.from_numpy(window_1d()).to(image.device)
    win = (g[:, None] @ g[None, :]).to(image.dtype)[None, None].expand(ch, 1, 11, 11).contiguous()
    blur = lambda t: F.conv2d(t, win, padding=5, groups=ch)
    m1, m2 = blur(image), blur(gt)
    v1 = blur(image * image) - m1 * m1
    v2 = blur(gt * gt) - m2 * m2
    v12 = blur(image * gt) - m1 * m2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    smap = ((2 * m1 * m2 + c1) * (2 * v12 + c2)) / ((m1 * m1 + m2 * m2 + c1) * (v1 + v2 + c2))
    return (1.0 - lambda_dssim) * (image - gt).abs().mean() + lambda_dssim * (1.0 - smap.mean())

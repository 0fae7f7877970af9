"""TEST INFRASTRUCTURE — numpy (float64) restatement of the reference's photometric loss path.

Follows utils/loss_utils.py of the reference:
    l1_loss                         :18-19
    gaussian / create_window        :27-35   (11 taps, sigma 1.5; the 1-D window is built in float32, normalised in
                                              float32, and the 2-D window is its float32 outer product)
    _ssim                           :49-68   (zero padding 5, C1 = 0.01^2, C2 = 0.03^2)
and the closed-form gradient of  sum(ssim_map)  with respect to img1 (what autograd computes for the reference).

PARITY PIN: the reference has no tests for this path; tests/golden/loss_*.npz hold outputs of the reference's OWN
Python functions (imported from /root/reference by tests/golden/make_golden_loss.py, float32 and float64, CPU).
Only tests/ and bench.py's baseline legs may import this module.
"""
from math import exp

import numpy as np


# The 1-D window as torch builds it in float32 (torch.Tensor([...]) / its float32 sum).  numpy's float32 sum uses a
# different summation order and lands one ulp away, so the eleven float32 values are pinned here (they are
# re-derived and compared in tests/test_loss_oracle.py::test_window_is_the_reference_window).
_G = np.array([1.028380124e-03, 7.598758209e-03, 3.600077331e-02, 1.093606874e-01, 2.130055279e-01, 2.660117149e-01,
               2.130055279e-01, 1.093606874e-01, 3.600077331e-02, 7.598758209e-03, 1.028380124e-03], dtype=np.float32)


def window_1d():
    return _G.copy()


def window_1d_formula():
    g = np.array([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=np.float32)
    return (g / g.sum(dtype=np.float32)).astype(np.float32)


def window_2d():
    g = window_1d()
    return np.outer(g, g).astype(np.float32).astype(np.float64)      # float32 product, as _1D_window.mm(...)


def _conv(img, w):
    """zero-padded 11x11 correlation of every plane of img [..., H, W] with w [11, 11]."""
    H, W = img.shape[-2:]
    pad = np.zeros(img.shape[:-2] + (H + 10, W + 10), dtype=np.float64)
    pad[..., 5:5 + H, 5:5 + W] = img
    out = np.zeros(img.shape, dtype=np.float64)
    for i in range(11):
        for j in range(11):
            out += w[i, j] * pad[..., i:i + H, j:j + W]
    return out


def l1_loss(x, y):
    return float(np.abs(np.asarray(x, np.float64) - np.asarray(y, np.float64)).mean())


def ssim_map(x, y):
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    w = window_2d()
    mu1, mu2 = _conv(x, w), _conv(y, w)
    s1 = _conv(x * x, w) - mu1 * mu1
    s2 = _conv(y * y, w) - mu2 * mu2
    s12 = _conv(x * y, w) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))


def ssim(x, y):
    return float(ssim_map(x, y).mean())


def loss_and_grad(x, y, lambda_dssim=0.2):
    """loss = (1 - l) * mean|x - y| + l * (1 - mean ssim_map)  and  d loss / d x  (helper_train.py:50-53)."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    w = window_2d()
    mu1, mu2 = _conv(x, w), _conv(y, w)
    e11, e22, e12 = _conv(x * x, w), _conv(y * y, w), _conv(x * y, w)
    s1, s2, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    A1, A2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    B1, B2 = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    S = A1 * A2 / (B1 * B2)
    n = x.size
    loss = (1 - lambda_dssim) * np.abs(x - y).mean() + lambda_dssim * (1 - S.mean())
    dS_dmu1 = 2 * (mu2 * (A2 - A1) - mu1 * S * (B2 - B1)) / (B1 * B2)
    dS_de11 = -S / B2
    dS_de12 = 2 * A1 / (B1 * B2)
    wt = w[::-1, ::-1]      # adjoint of a correlation = correlation with the flipped window (symmetric here)
    dsum = _conv(dS_dmu1, wt) + 2 * x * _conv(dS_de11, wt) + y * _conv(dS_de12, wt)
    grad = (1 - lambda_dssim) * np.sign(x - y) / n - lambda_dssim * dsum / n
    return float(loss), grad

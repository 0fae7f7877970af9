"""TEST INFRASTRUCTURE — numpy restatement of the reference's per-iteration densification statistics.

Follows train.py of the reference:
    :211        per view: norm of the means2D gradient's x, y
    :214-215    per view: radii, visibility filter (radii > 0, renderer/__init__.py:131)
    :281-287    batch reduction: visibility count, max radii, summed gradient norm / count where visible
    :290        max_radii2D[visible] = max(max_radii2D[visible], radii[visible])
    :291        GaussianModel.add_densification_stats_grad (scene/saro_gaussian.py:745-747)

PARITY PIN: tests/golden/densify_*.npz hold the result of executing the reference's OWN statements (cut out of
train.py and saro_gaussian.py with `ast` by tests/golden/make_golden_densify.py) on CPU.  Only tests/ and bench.py's
baseline legs may import this module.
"""
import numpy as np


def batch_statistics(view_grads, view_radii, max_radii2D, xyz_gradient_accum, denom, dtype=np.float64):
    """view_grads: list of [P,3]; view_radii: list of int [P].  Returns the three updated statistics (copies)."""
    norms = [np.sqrt((g[:, :2].astype(dtype) ** 2).sum(-1)) for g in view_grads]                     # :211
    vis = [r > 0 for r in view_radii]                                                                # :215
    visibility_count = np.stack(vis, 1).sum(1)                                                       # :281
    visibility_filter = visibility_count > 0                                                         # :282
    radii = np.stack(view_radii, 1).max(1)                                                           # :283
    grad = np.stack(norms, 1).sum(1)                                                                 # :284
    grad[visibility_filter] = grad[visibility_filter] / visibility_count[visibility_filter]          # :285
    mr, acc, den = (np.array(a, dtype=dtype).copy() for a in (max_radii2D, xyz_gradient_accum, denom))
    mr[visibility_filter] = np.maximum(mr[visibility_filter], radii[visibility_filter])              # :290
    acc.reshape(-1)[visibility_filter] += grad[visibility_filter]                                    # saro_gaussian.py:746
    den.reshape(-1)[visibility_filter] += 1                                                          # saro_gaussian.py:747
    return mr, acc, den

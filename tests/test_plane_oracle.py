"""CPU tests: pins oracle/plane_oracle.py (numpy restatement of ScaleAwareResField.forward, scene/hexplane.py of the
reference, and of the un-vendored nvdiffrast `texture` op underneath it) on
  (a) tests/golden/plane_*.npz — outputs of the reference's own hexplane.py executed via `ast`
      (tests/golden/make_golden_plane.py) around a torch stand-in of the published texture algorithm, and
  (b) torch.nn.functional.grid_sample / avg_pool2d directly, at integer and fractional mip levels."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_util import load, maxrel
from oracle import plane_oracle

CASES = ["plane_small", "plane_ragged", "plane_wide"]


def grids_of(d):
    n_res = len(d["multires"])
    return [[d[f"grid_{li}_{ci}"] for ci in range(6)] for li in range(n_res)]


def reso_list(d):
    return [[int(r) * int(m) for r in d["reso"][:3]] + [int(d["reso"][3])] for m in d["multires"]]


@pytest.mark.parametrize("name", CASES)
def test_field_forward_matches_reference_python(name):
    d = load(name)
    got = plane_oracle.field_forward(d["pts"], d["timestamps"], d["scales"], grids_of(d), d["aabb"], int(d["duration"]),
                                     d["base_scale"], reso_list(d))
    assert got.shape == d["features_f64"].shape
    assert maxrel(got, d["features_f64"]) < 1e-12
    got32 = plane_oracle.field_forward(d["pts"], d["timestamps"], d["scales"], grids_of(d), d["aabb"],
                                       int(d["duration"]), d["base_scale"], reso_list(d), dtype=np.float32)
    assert maxrel(got32, d["features_f32"]) < 2e-5


@pytest.mark.parametrize("name", CASES)
def test_field_backward_matches_reference_autograd(name):
    d = load(name)
    shapes = [[g.shape for g in level] for level in grids_of(d)]
    got = plane_oracle.field_backward(d["pts"], d["timestamps"], d["scales"], shapes, d["aabb"], int(d["duration"]),
                                      d["base_scale"], reso_list(d), d["dout"])
    for li, level in enumerate(got):
        for ci, g in enumerate(level):
            want = d[f"dgrid_f64_{li}_{ci}"]
            assert g.shape == want.shape
            assert np.abs(g - want).max() <= 2e-7 * max(np.abs(want).max(), 1e-30), (li, ci)   # stored as float32


@pytest.mark.parametrize("h,w,max_level", [(16, 16, 7), (8, 32, 7), (6, 10, 0), (2, 64, 3)])
def test_texture_vs_grid_sample(h, w, max_level):
    """Independent check of the texel-space convention, the clamp rule and the mip stack."""
    rng = np.random.default_rng(h * 100 + w)
    c, n = 5, 300
    tex = rng.standard_normal((h, w, c))
    uv = rng.uniform(-0.2, 1.2, size=(n, 2))
    t = torch.from_numpy(tex).permute(2, 0, 1)[None]
    mips = [t]
    while (mips[-1].shape[2] > 1 or mips[-1].shape[3] > 1) and len(mips) - 1 < max_level:
        hh, ww = mips[-1].shape[2:]
        mips.append(F.avg_pool2d(mips[-1], (2 if hh > 1 else 1, 2 if ww > 1 else 1)))
    grid = torch.from_numpy(uv)[None, None] * 2 - 1
    for lv, m in enumerate(mips):
        want = F.grid_sample(m, grid, mode="bilinear", padding_mode="border", align_corners=False)[0, :, 0].t().numpy()
        got = plane_oracle.texture(tex, uv, np.full(n, float(lv)), max_level)
        assert np.abs(got - want).max() < 1e-12
        if lv + 1 < len(mips):
            nxt = F.grid_sample(mips[lv + 1], grid, mode="bilinear", padding_mode="border",
                                align_corners=False)[0, :, 0].t().numpy()
            got = plane_oracle.texture(tex, uv, np.full(n, lv + 0.3), max_level)
            assert np.abs(got - (0.7 * want + 0.3 * nxt)).max() < 1e-12
    # bias outside [0, top] clamps
    top = len(mips) - 1
    lo = plane_oracle.texture(tex, uv, np.full(n, -3.0), max_level)
    hi = plane_oracle.texture(tex, uv, np.full(n, 99.0), max_level)
    assert np.array_equal(lo, plane_oracle.texture(tex, uv, np.zeros(n), max_level))
    assert np.array_equal(hi, plane_oracle.texture(tex, uv, np.full(n, float(top)), max_level))


def test_texture_backward_is_the_adjoint():
    """<texture(T), D> == <T, texture_backward(D)> for random T, D (linearity in the texture)."""
    rng = np.random.default_rng(7)
    h, w, c, n = 16, 8, 3, 200
    uv = rng.uniform(-0.1, 1.1, size=(n, 2))
    bias = rng.uniform(-1, 5, size=n)
    tex = rng.standard_normal((h, w, c))
    dout = rng.standard_normal((n, c))
    lhs = (plane_oracle.texture(tex, uv, bias, 7) * dout).sum()
    rhs = (tex * plane_oracle.texture_backward((h, w, c), uv, bias, 7, dout)).sum()
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0)

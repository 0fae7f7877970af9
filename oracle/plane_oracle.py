"""TEST INFRASTRUCTURE — numpy restatement of the scale-aware plane sampler:
`ScaleAwareResField.forward` of the reference (scene/hexplane.py:258-286) and the third-party op underneath it.

What is restated, and from where:
  * normalize_aabb / normalize_time / get_level            scene/hexplane.py:19-23, 231-242
  * the six coordinate pairs, `levels[..., pair].min(-1)`   scene/hexplane.py:26-45, 102-117 (itertools.combinations(range(4), 2))
  * sum over planes, concat over resolutions                scene/hexplane.py:119-136
  * `nvdiffrast.torch.texture(tex, uv, mip_level_bias=..., boundary_mode="clamp", max_mip_level=7 | 0)`
                                                            scene/hexplane.py:49-56
    nvdiffrast is an UN-VENDORED third-party dependency (not in /root/reference, not installed, no version pinned by
    the reference: it ships no requirements file).  Its published algorithm (Laine et al. 2020, "Modular Primitives for
    High-Performance Differentiable Rendering", section 3.4 + the library documentation of `texture`):
      - mip stack: level l+1 = 2x2 box average of level l (extents must be even while > 1), built while any extent
        is > 1 and l < max_mip_level;
      - with `mip_level_bias` and no screen-space derivatives the level is the bias alone, clamped to [0, top level];
        filter_mode 'auto' then means linear-mipmap-linear: bilinear taps on floor(level) and floor(level) + 1, blended
        by the fractional part;
      - a bilinear tap at uv in [0, 1]^2: texel space u = uv.x * W - 0.5, v = uv.y * H - 0.5 (texel centres at
        half-integers), boundary "clamp": u, v clamped to [0, W-1] x [0, H-1].
    PARITY PIN: the third-party source is absent, so this restatement is pinned (tests/test_plane_oracle.py) on
    (a) torch.nn.functional.grid_sample(align_corners=False, padding_mode="border") + avg_pool2d, an independent
    implementation of the same published rule, and (b) golden vectors made by executing the reference's OWN
    scene/hexplane.py (cut out with `ast`, tests/golden/make_golden_plane.py) with exactly that torch composition
    standing in for the missing third-party op.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import itertools

import numpy as np


def build_mips(tex, max_mip_level):
    """tex [H, W, C] -> list of levels; 2x2 box filter while any extent > 1 and level < max_mip_level."""
    levels = [np.asarray(tex)]
    while (levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1) and len(levels) - 1 < max_mip_level:
        t = levels[-1]
        h, w = t.shape[:2]
        if (h > 1 and h & 1) or (w > 1 and w & 1):
            raise ValueError("texture extents must be even at every mip level that is built")
        if h > 1:
            t = 0.5 * (t[0::2] + t[1::2])
        if w > 1:
            t = 0.5 * (t[:, 0::2] + t[:, 1::2])
        levels.append(t)
    return levels


def _taps(uv, h, w):
    """texel indices + weights of the clamped bilinear footprint; uv [N, 2] (x, y) in texture coordinates."""
    u = np.clip(uv[:, 0] * w - 0.5, 0.0, w - 1.0)
    v = np.clip(uv[:, 1] * h - 0.5, 0.0, h - 1.0)
    iu0 = np.floor(u).astype(np.int64)
    iv0 = np.floor(v).astype(np.int64)
    fu, fv = u - iu0, v - iv0
    iu1 = np.minimum(iu0 + 1, w - 1)
    iv1 = np.minimum(iv0 + 1, h - 1)
    return (iu0, iu1, iv0, iv1), ((1 - fu) * (1 - fv), fu * (1 - fv), (1 - fu) * fv, fu * fv)


def _bilinear(t, uv):
    h, w = t.shape[:2]
    (iu0, iu1, iv0, iv1), (w00, w10, w01, w11) = _taps(uv, h, w)
    return (w00[:, None] * t[iv0, iu0] + w10[:, None] * t[iv0, iu1] + w01[:, None] * t[iv1, iu0] +
            w11[:, None] * t[iv1, iu1])


def _level_split(bias, top):
    level = np.clip(bias, 0.0, float(top))
    l0 = np.floor(level).astype(np.int64)
    f = level - l0
    l1 = np.minimum(l0 + 1, top)
    return l0, l1, f


def texture(tex, uv, bias, max_mip_level):
    """tex [H, W, C], uv [N, 2], bias [N] -> [N, C]: linear-mipmap-linear, boundary clamp."""
    mips = build_mips(tex, max_mip_level)
    l0, l1, f = _level_split(np.asarray(bias, dtype=tex.dtype), len(mips) - 1)
    out = np.zeros((uv.shape[0], tex.shape[2]), dtype=tex.dtype)
    for lv, t in enumerate(mips):
        m0 = l0 == lv
        if m0.any():
            out[m0] += (1 - f[m0])[:, None] * _bilinear(t, uv[m0])
        m1 = (l1 == lv) & (f > 0)
        if m1.any():
            out[m1] += f[m1][:, None] * _bilinear(t, uv[m1])
    return out


def texture_backward(tex_shape, uv, bias, max_mip_level, dout, dtype=np.float64):
    """Gradient of `texture` with respect to the base texture: taps scatter into their mip level, the mip levels
    fold down through the transpose of the 2x2 box filter."""
    h, w, c = tex_shape
    shapes = [(h, w)]
    while (shapes[-1][0] > 1 or shapes[-1][1] > 1) and len(shapes) - 1 < max_mip_level:
        hh, ww = shapes[-1]
        shapes.append((max(hh // 2, 1), max(ww // 2, 1)))
    grads = [np.zeros((hh, ww, c), dtype=dtype) for hh, ww in shapes]
    l0, l1, f = _level_split(np.asarray(bias, dtype=dtype), len(shapes) - 1)
    for lv, (hh, ww) in enumerate(shapes):
        for sel, wt in ((l0 == lv, 1 - f), ((l1 == lv) & (f > 0), f)):
            if not sel.any():
                continue
            (iu0, iu1, iv0, iv1), ws = _taps(uv[sel], hh, ww)
            g = dout[sel] * wt[sel][:, None]
            for (iv, iu), wgt in zip(((iv0, iu0), (iv0, iu1), (iv1, iu0), (iv1, iu1)), ws):
                np.add.at(grads[lv], (iv, iu), wgt[:, None] * g)
    for lv in range(len(shapes) - 1, 0, -1):
        g = grads[lv]
        hh, ww = shapes[lv - 1]
        if ww > 1:
            g = 0.5 * np.repeat(g, 2, axis=1)
        if hh > 1:
            g = 0.5 * np.repeat(g, 2, axis=0)
        grads[lv - 1] += g
    return grads[0]


COO_COMBS = list(itertools.combinations(range(4), 2))   # (0,1) (0,2) (0,3) (1,2) (1,3) (2,3); 3 = time


def get_level(scales, base_scale, reso0):
    """scene/hexplane.py:231-242: per-axis mip level from the Gaussian's scale; the time axis gets level 0."""
    min_scale = base_scale / 2
    max_scale = min_scale * np.asarray(reso0[:3], dtype=scales.dtype)
    s = np.clip(scales, min_scale, max_scale)
    level = np.log2(2 * s / base_scale[None, :])
    return np.concatenate([level, np.zeros((level.shape[0], 1), dtype=level.dtype)], axis=1)


def field_forward(pts, timestamps, scales, grids, aabb, duration, base_scale, reso_list, dtype=np.float64):
    """ScaleAwareResField.forward.  grids: list (per resolution) of 6 arrays [1, C, H, W] (the reference's parameter
    layout); aabb [2, 3] (row 0 = xyz_max, row 1 = xyz_min) and base_scale [3] are the module's float32 buffers as
    set_aabb registers them (scene/hexplane.py:202-229: base_scale = float32((max - min) / resolution) of the coarsest
    grid); returns [N, C * len(grids)]."""
    pts = np.asarray(pts, dtype=dtype)
    aabb = np.asarray(aabb, dtype=dtype)
    p = (pts - aabb[0]) / (aabb[1] - aabb[0])
    t = np.asarray(timestamps, dtype=dtype).reshape(-1, 1) * duration / (duration - 1)
    p4 = np.concatenate([p, t], axis=1)
    base_scale = np.asarray(base_scale, dtype=dtype)
    level = get_level(np.asarray(scales, dtype=dtype), base_scale, reso_list[0])
    outs = []
    for planes in grids:
        acc = 0.0
        for ci, comb in enumerate(COO_COMBS):
            tex = np.transpose(np.asarray(planes[ci], dtype=dtype)[0], (1, 2, 0))   # [H, W, C]
            spatio_only = 3 not in comb
            bias = level[:, list(comb)].min(axis=1)
            acc = acc + texture(tex, p4[:, list(comb)], bias, 7 if spatio_only else 0)
        outs.append(acc)
    return np.concatenate(outs, axis=1)


def field_backward(pts, timestamps, scales, grid_shapes, aabb, duration, base_scale, reso_list, dout, dtype=np.float64):
    """Gradients of the planes (the only inputs of the field that require one: the reference calls it on detached
    positions / scales, scene/saro_gaussian.py:765,780,865).  Returns a list (per resolution) of 6 arrays [1, C, H, W]."""
    pts = np.asarray(pts, dtype=dtype)
    aabb = np.asarray(aabb, dtype=dtype)
    p = (pts - aabb[0]) / (aabb[1] - aabb[0])
    t = np.asarray(timestamps, dtype=dtype).reshape(-1, 1) * duration / (duration - 1)
    p4 = np.concatenate([p, t], axis=1)
    base_scale = np.asarray(base_scale, dtype=dtype)
    level = get_level(np.asarray(scales, dtype=dtype), base_scale, reso_list[0])
    out, c0 = [], 0
    for shapes in grid_shapes:
        res = []
        C = shapes[0][1]
        d = np.asarray(dout, dtype=dtype)[:, c0:c0 + C]
        for ci, comb in enumerate(COO_COMBS):
            _, _, H, W = shapes[ci]
            spatio_only = 3 not in comb
            bias = level[:, list(comb)].min(axis=1)
            g = texture_backward((H, W, C), p4[:, list(comb)], bias, 7 if spatio_only else 0, d, dtype)
            res.append(np.transpose(g, (2, 0, 1))[None])
        out.append(res)
        c0 += C
    return out

"""CPU test of the round-2 binning ALGORITHM (csrc/sgs_binning.cu), restated in numpy: sorting the P Gaussians by
depth bits, bucketing the depth-ordered (supertile, Gaussian) stream with one stable pass and expanding every supertile
list into its 4x4 tile lists in order must give exactly the lists the reference builds with one stable sort of
`key = tile << 32 | float_bits(depth)` over all tile instances ($R/cuda_rasterizer/rasterizer_impl.cu:70-111, 299-309).
The CUDA kernels themselves are checked against the same reference construction on the GPU (tests/test_binning_gpu.py);
this file pins the design claim on machines without one."""
import numpy as np
import pytest

ST = 4   # supertile edge in tiles


def reference_lists(rects, depth_bits, tiles_x, tiles_y):
    gid, tile = [], []
    for g, (x0, x1, y0, y1) in enumerate(rects):
        for y in range(y0, y1):
            for x in range(x0, x1):
                gid.append(g)
                tile.append(y * tiles_x + x)
    gid, tile = np.array(gid, np.int64), np.array(tile, np.int64)
    key = (tile << 32) | depth_bits[gid].astype(np.int64)
    order = np.argsort(key, kind="stable")
    point_list = gid[order]
    counts = np.bincount(tile, minlength=tiles_x * tiles_y)
    ends = np.cumsum(counts)
    return point_list, np.stack([ends - counts, ends], 1)


def supertile_lists(rects, depth_bits, tiles_x, tiles_y):
    P = len(rects)
    sx_n = (tiles_x + ST - 1) // ST
    # 1. stable depth sort of the Gaussians (culled ones have an empty rect and emit nothing)
    order = np.argsort(depth_bits, kind="stable")
    # 2. depth-ordered coarse stream, one stable bucketing pass by supertile
    cg, cs = [], []
    for g in order:
        x0, x1, y0, y1 = rects[g]
        if x1 <= x0 or y1 <= y0:
            continue
        for sy in range(y0 // ST, (y1 + ST - 1) // ST):
            for sx in range(x0 // ST, (x1 + ST - 1) // ST):
                cg.append(g)
                cs.append(sy * sx_n + sx)
    cg, cs = np.array(cg, np.int64), np.array(cs, np.int64)
    bucket = np.argsort(cs, kind="stable")
    cg, cs = cg[bucket], cs[bucket]
    # 3. expansion: every supertile streams its list in order and appends to the lists of the tiles each rect covers
    lists = [[] for _ in range(tiles_x * tiles_y)]
    for g, s in zip(cg, cs):
        tx0, ty0 = (s % sx_n) * ST, (s // sx_n) * ST
        x0, x1, y0, y1 = rects[g]
        for ly in range(ST):
            for lx in range(ST):
                tx, ty = tx0 + lx, ty0 + ly
                if x0 <= tx < x1 and y0 <= ty < y1:
                    lists[ty * tiles_x + tx].append(g)
    counts = np.array([len(l) for l in lists])
    ends = np.cumsum(counts)
    point_list = np.array([g for l in lists for g in l], np.int64)
    return point_list, np.stack([ends - counts, ends], 1)


@pytest.mark.parametrize("seed,P,tiles_x,tiles_y,max_extent", [(0, 400, 9, 7, 3), (1, 300, 25, 25, 12), (2, 50, 3, 2, 3),
                                                              (3, 600, 17, 5, 20)])
def test_supertile_binning_reproduces_the_reference_order(seed, P, tiles_x, tiles_y, max_extent):
    rng = np.random.default_rng(seed)
    x0 = rng.integers(0, tiles_x, P)
    y0 = rng.integers(0, tiles_y, P)
    x1 = np.minimum(tiles_x, x0 + rng.integers(0, max_extent + 1, P))      # some empty rects (culled Gaussians)
    y1 = np.minimum(tiles_y, y0 + rng.integers(0, max_extent + 1, P))
    rects = list(zip(x0.tolist(), x1.tolist(), y0.tolist(), y1.tolist()))
    rects = [(a, b, c, d) if (b > a and d > c) else (0, 0, 0, 0) for a, b, c, d in rects]
    depth = rng.uniform(0.2, 40.0, P).astype(np.float32)
    depth[rng.integers(0, P, P // 5)] = depth[0]                             # ties: order must fall back to the index
    bits = depth.view(np.uint32)
    ref_pl, ref_rng = reference_lists(rects, bits, tiles_x, tiles_y)
    got_pl, got_rng = supertile_lists(rects, bits, tiles_x, tiles_y)
    assert np.array_equal(ref_pl, got_pl)
    assert np.array_equal(ref_rng, got_rng)

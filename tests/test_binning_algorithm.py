"""CPU test of the round-2 binning ALGORITHMS (csrc/sgs_binning.cu), restated in numpy.  Both modes must give exactly
the lists the reference builds with one stable sort of `key = tile << 32 | float_bits(depth)` over all tile instances
($R/cuda_rasterizer/rasterizer_impl.cu:70-111, 299-309):
  * global mode: sort the P Gaussians by depth bits, bucket the depth-ordered (supertile, Gaussian) stream with one
    stable pass, expand every supertile list into its 4x4 tile lists in order;
  * per-supertile mode (default): bucket the INDEX-ordered stream, then sort every bucket by depth bits — one stable
    pass on the top digit of the frame-normalised key followed by exact placement inside the digit bucket by counting
    the entries that precede in (key, position) order (tile_fill_sorted_kernel) — then expand.
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


def msd_then_count_sort(keys, top_bits=9):
    """Order of `keys` (stable) as tile_fill_sorted_kernel's fast path computes it."""
    keys = np.asarray(keys, np.int64)
    n = len(keys)
    if n == 0:
        return np.zeros(0, np.int64)
    span = int(keys.max()) + 1
    nbits = max(1, (span - 1).bit_length())
    mbits = min(top_bits, nbits)
    digit = keys >> (nbits - mbits)
    first = np.argsort(digit, kind="stable")                 # ONE stable radix pass on the top digit
    k1, d1 = keys[first], digit[first]
    out = np.empty(n, np.int64)
    starts = np.searchsorted(d1, np.arange(1 << mbits), side="left")
    ends = np.searchsorted(d1, np.arange(1 << mbits), side="right")
    for j in range(n):                                        # exact placement inside the digit bucket
        b0, b1 = starts[d1[j]], ends[d1[j]]
        seg = k1[b0:b1]
        c = int(np.sum(seg < k1[j])) + int(np.sum(seg[:j - b0] == k1[j]))
        out[b0 + c] = first[j]
    return out


def supertile_lists(rects, depth_bits, tiles_x, tiles_y, sort_in_supertile=False):
    P = len(rects)
    sx_n = (tiles_x + ST - 1) // ST
    # 1. stable depth sort of the Gaussians (culled ones have an empty rect and emit nothing) — or, in the
    #    per-supertile mode, index order here and the depth sort inside every bucket below
    order = np.arange(P) if sort_in_supertile else np.argsort(depth_bits, kind="stable")
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
    if sort_in_supertile:
        vis = [g for g in range(P) if rects[g][1] > rects[g][0] and rects[g][3] > rects[g][2]]
        kmin = int(depth_bits[vis].min()) if vis else 0       # the frame's key range (visible Gaussians)
        for sidx in np.unique(cs):
            sel = np.nonzero(cs == sidx)[0]
            o = msd_then_count_sort(depth_bits[cg[sel]].astype(np.int64) - kmin)
            cg[sel] = cg[sel][o]
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
    for mode in (False, True):
        got_pl, got_rng = supertile_lists(rects, bits, tiles_x, tiles_y, sort_in_supertile=mode)
        assert np.array_equal(ref_pl, got_pl), mode
        assert np.array_equal(ref_rng, got_rng), mode


def test_msd_then_count_sort_is_a_stable_sort():
    rng = np.random.default_rng(5)
    for n, hi in ((0, 10), (1, 10), (500, 7), (700, 1 << 26), (300, 1 << 9)):
        keys = rng.integers(0, hi, n)
        assert np.array_equal(msd_then_count_sort(keys), np.argsort(keys, kind="stable"))

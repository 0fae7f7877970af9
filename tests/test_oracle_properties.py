"""CPU tests of the oracle itself (it is the checker, so it gets checked): finite differences of
its float64 forward against its analytic backward, and the behavioural fine print of the
reference that needs no golden data."""
import numpy as np
import torch

from saro_gs_b200 import synthetic


def _tiny():
    scene, cam = synthetic.small_scene(P=64, seed=3, width=48, height=32, fx=40.0, log_scale_mean=-1.5)
    return scene, cam, torch.tensor([0.3, 0.5, 0.7])


def test_backward_matches_finite_differences(oracle_mod):
    scene, cam, bg = _tiny()
    cot = torch.randn(3, cam.height, cam.width, generator=torch.Generator().manual_seed(5)).double().numpy()

    def loss(sc):
        r = oracle_mod.forward_scene(sc, cam, bg)
        return float((r.color * cot).sum()), r

    _, r0 = loss(scene)
    g = r0.backward(cot.astype(np.float32))
    vis = np.nonzero(r0.radii > 0)[0]
    rng = np.random.default_rng(0)
    for name in ("means3D", "scales", "rotations", "opacities", "shs"):
        base = getattr(scene, name)
        rels = []
        for _ in range(16):
            i = int(rng.choice(vis))
            idx = (i,) + tuple(int(rng.integers(0, s)) for s in base.shape[1:])
            h = max(1e-3 * abs(float(base[idx])), 2e-4)
            p, m = base.clone(), base.clone()
            p[idx] += h
            m[idx] -= h
            lp, _ = loss(scene._replace(**{name: p}))
            lm, _ = loss(scene._replace(**{name: m}))
            fd = (lp - lm) / float(p[idx] - m[idx])
            an = float(g[name][idx])
            rels.append(abs(fd - an) / max(abs(fd), abs(an), 1e-7))
        # the forward is piecewise smooth (alpha / transmittance thresholds, 0.99 cap, tile rects): a few
        # samples straddle a discontinuity, the bulk must agree tightly
        assert np.median(rels) < 1e-4, (name, rels)
        assert np.mean(np.array(rels) < 1e-3) >= 0.6, (name, rels)


def test_backward_is_linear_in_cotangent(oracle_mod):
    scene, cam, bg = _tiny()
    r = oracle_mod.forward_scene(scene, cam, bg)
    gen = torch.Generator().manual_seed(1)
    a = torch.randn(3, cam.height, cam.width, generator=gen).numpy()
    b = torch.randn(3, cam.height, cam.width, generator=gen).numpy()
    ga, gb, gab = r.backward(a), r.backward(b), r.backward(2.0 * a - 0.5 * b)
    for k in ga:
        want = 2.0 * ga[k] - 0.5 * gb[k]
        assert np.allclose(gab[k], want, rtol=1e-5, atol=1e-6 * max(1.0, np.abs(want).max())), k


def test_empty_input_gives_zeros_not_background(oracle_mod):
    _, cam, bg = _tiny()
    r = oracle_mod.forward(torch.zeros(0, 3), torch.zeros(0, 1), cam.viewmatrix, cam.projmatrix, cam.campos, bg,
                           cam.width, cam.height, cam.tanfovx, cam.tanfovy, sh_degree=0,
                           colors_precomp=torch.zeros(0, 3), scales=torch.zeros(0, 3), rotations=torch.zeros(0, 4))
    assert r.num_rendered == 0 and not r.color.any() and not r.depth.any()


def test_all_culled_gives_background_and_default_depth(oracle_mod):
    _, cam, bg = _tiny()
    P = 5
    means = torch.zeros(P, 3)
    means[:, 2] = 0.1                       # z_view <= 0.2 => culled
    r = oracle_mod.forward(means, torch.ones(P, 1), cam.viewmatrix, cam.projmatrix, cam.campos, bg, cam.width,
                           cam.height, cam.tanfovx, cam.tanfovy, colors_precomp=torch.ones(P, 3),
                           scales=torch.ones(P, 3) * 0.1, rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(P, 1))
    assert r.num_rendered == 0 and (r.radii == 0).all()
    assert np.allclose(r.color, bg.numpy()[:, None, None])
    assert (r.depth == 15.0).all() and (r.n_contrib == 0).all() and (r.final_T == 1.0).all()


def test_single_opaque_splat_semantics(oracle_mod):
    """One big opaque Gaussian in front of the camera: alpha capped at 0.99, T never < 1e-4 after one
    splat, median depth crossed at its centre, rect-clipped outside 3 sigma."""
    _, cam, _ = _tiny()
    bg = torch.zeros(3)
    means = torch.tensor([[0.0, 0.0, 2.0]])
    r = oracle_mod.forward(means, torch.ones(1, 1), cam.viewmatrix, cam.projmatrix, cam.campos, bg, cam.width,
                           cam.height, cam.tanfovx, cam.tanfovy, colors_precomp=torch.tensor([[1.0, 0.5, 0.25]]),
                           scales=torch.ones(1, 3) * 0.2, rotations=torch.tensor([[1.0, 0, 0, 0]]))
    cy, cx = cam.height // 2, cam.width // 2
    # pixel centres sit at integer coordinates: the projected centre is at (W-1)/2, (H-1)/2
    # (so the nearest pixel is half a pixel off-centre: alpha = exp(-0.25/16.3) = 0.985 < 0.99 cap)
    T = r.final_T[cy, cx]
    assert 0.01 <= T < 0.02
    assert abs(r.color[0, cy, cx] - (1 - T)) < 1e-6 and abs(r.color[1, cy, cx] - 0.5 * (1 - T)) < 1e-6
    assert r.depth[0, cy, cx] == 2.0 and r.n_contrib[cy, cx] == 1
    assert r.radii[0] > 0 and r.tiles_touched[0] == r.num_rendered


def test_point_list_is_sorted_by_tile_then_depth_then_index(oracle_mod):
    scene, cam, bg = _tiny()
    r = oracle_mod.forward_scene(scene, cam, bg, precision="f32")
    depth_bits = {}
    view = cam.viewmatrix.numpy().astype(np.float32)
    m = scene.means3D.numpy()
    z = (view[0, 2] * m[:, 0] + view[1, 2] * m[:, 1] + view[2, 2] * m[:, 2] + view[3, 2]).astype(np.float32)
    bits = z.view(np.uint32)
    for t, (lo, hi) in enumerate(r.ranges):
        ids = r.point_list[lo:hi]
        keys = [(int(bits[g]), int(g)) for g in ids]
        assert keys == sorted(keys), t
    assert int((r.ranges[:, 1] - r.ranges[:, 0]).sum()) == r.num_rendered == int(r.tiles_touched.sum())


def test_f32_and_f64_oracles_agree(oracle_mod):
    scene, cam, bg = _tiny()
    a = oracle_mod.forward_scene(scene, cam, bg, precision="f32")
    b = oracle_mod.forward_scene(scene, cam, bg, precision="f64")
    assert np.array_equal(a.radii, b.radii) and a.num_rendered == b.num_rendered
    assert np.abs(a.color - b.color).max() < 1e-5
    cot = synthetic.cotangent(cam.height, cam.width)
    ga, gb = a.backward(cot), b.backward(cot)
    for k in ga:
        assert np.linalg.norm(ga[k] - gb[k]) <= 1e-3 * max(np.linalg.norm(gb[k]), 1e-30), k


def test_precomputed_colour_and_covariance_paths(oracle_mod):
    scene, cam, bg = _tiny()
    base = oracle_mod.forward_scene(scene, cam, bg)
    # feeding the oracle's own rgb / cov3D back as precomputed inputs must reproduce the image
    r2 = oracle_mod.forward(scene.means3D, scene.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, bg, cam.width,
                            cam.height, cam.tanfovx, cam.tanfovy, colors_precomp=base.rgb.astype(np.float32),
                            cov3D_precomp=base.cov3D.astype(np.float32))
    assert np.array_equal(r2.radii, base.radii)
    assert np.abs(r2.color - base.color).max() < 1e-5

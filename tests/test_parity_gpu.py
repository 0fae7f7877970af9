"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI
(libsaro_gs_b200.so via saro_gs_b200.backend); three independent checkers:
  1. golden fixtures = outputs of the compiled unmodified reference (tests/golden/*.npz);
  2. the compiled reference itself, live, when oracle/_ref travelled to the box;
  3. the float64 CPU oracle (oracle/splat_oracle.c).
Bars: integer state (radii, tiles_touched, ranges, point_list, n_contrib) bit-exact; forward
colour/depth/final_T bit-exact against the reference (the kernels reproduce its float expression
trees); gradients within 1e-4 of the max entry (float atomics in the reference make element-wise
bit equality impossible — the float64 oracle arbitrates)."""
import hashlib

import numpy as np
import pytest
import torch

from golden_util import SMALL_CASES, inputs_of, load, maxrel, normrel

pytestmark = pytest.mark.gpu
GRAD_TOL = 1e-4          # north_star: "within 1e-4 relative"
# largest-entry difference between two runs of the compiled reference at configs[1] (tools/grad_noise.py on a B200)
REF_RUN_TO_RUN_MAX = {"scales": 8.6e-4, "rotations": 1.1e-3}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def sgs(native_lib):
    import saro_gs_b200
    return saro_gs_b200


def settings_from(sgs, d, dev, prefiltered=False):
    t = lambda k: torch.from_numpy(np.asarray(d[k])).to(dev)
    return sgs.GaussianRasterizationSettings(int(d["height"]), int(d["width"]), float(d["tanfovx"]),
                                             float(d["tanfovy"]), t("bg"), float(d["scale_modifier"]),
                                             t("viewmatrix"), t("projmatrix"), int(d["sh_degree"]), t("campos"),
                                             prefiltered)


def run_native(sgs, d, dev, Rast=None):
    ins = inputs_of(d)
    leaves = {k: v.to(dev).clone().requires_grad_(True) for k, v in ins.items()}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    rs = settings_from(sgs, d, dev)
    kw = {("cov3D_precomp" if k == "cov3D_precomp" else k): v for k, v in leaves.items()}
    color, radii, depth = (Rast or sgs.GaussianRasterizer)(rs)(means2D=means2D, **kw)
    color.backward(torch.from_numpy(d["cotangent"]).to(dev))
    torch.cuda.synchronize()
    grads = {k: v.grad.cpu().numpy() for k, v in leaves.items()}
    grads["means2D"] = means2D.grad.cpu().numpy()
    return color.detach().cpu().numpy(), radii.cpu().numpy(), depth.detach().cpu().numpy(), grads, leaves, rs


def native_state(sgs, d, dev, **kw):
    ins = {k: v.to(dev) for k, v in inputs_of(d).items()}
    rs = settings_from(sgs, d, dev)
    e = torch.Tensor([])
    out = sgs._C.rasterize_gaussians(rs.bg, ins["means3D"], ins.get("colors_precomp", e), ins["opacities"],
                                     ins.get("scales", e), ins.get("rotations", e), rs.scale_modifier,
                                     ins.get("cov3D_precomp", e), rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                                     rs.tanfovy, rs.image_height, rs.image_width, ins.get("shs", e), rs.sh_degree,
                                     rs.campos, False, **kw)
    R, color, radii, gb, bb, ib, depth = out
    st = sgs._C.debug_export(ins["means3D"].shape[0], rs.image_width, rs.image_height, R, gb, bb, ib)
    return R, color, radii, depth, {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in st.items()}


@pytest.mark.parametrize("name", SMALL_CASES)
def test_forward_bit_exact_vs_reference_golden(sgs, dev, name):
    d = load(name)
    # product mode (exact tile/quadrant culling): everything API-visible + per-pixel transmittance
    R, color, radii, depth, st = native_state(sgs, d, dev)
    assert R == int(d["num_rendered"])                                # bit-exact tile counts
    assert np.array_equal(radii.cpu().numpy(), d["out_radii"])
    assert np.array_equal(st["tiles_touched"], d["tiles_touched"])
    assert np.array_equal(st["final_T"], d["final_T"])
    assert np.array_equal(color.cpu().numpy(), d["out_color"])       # bit-exact image
    assert np.array_equal(depth.cpu().numpy(), d["out_depth"])
    assert st["kept"] <= R
    # validation mode (no culling): the internal per-tile lists are the reference's, bit for bit
    R, color, radii, depth, st = native_state(sgs, d, dev, _no_tile_cull=True)
    assert R == int(d["num_rendered"]) == st["kept"]
    assert np.array_equal(st["ranges"], d["ranges"])
    assert np.array_equal(st["point_list"], d["point_list"])
    assert np.array_equal(st["n_contrib"], d["n_contrib"])
    assert np.array_equal(st["final_T"], d["final_T"])
    assert np.array_equal(color.cpu().numpy(), d["out_color"])
    assert np.array_equal(depth.cpu().numpy(), d["out_depth"])


@pytest.mark.parametrize("name", SMALL_CASES)
def test_backward_vs_reference_golden_and_oracle(sgs, dev, oracle_mod, name):
    d = load(name)
    color, radii, depth, grads, _, _ = run_native(sgs, d, dev)
    assert np.array_equal(color, d["out_color"])
    ins = inputs_of(d)
    orc = oracle_mod.forward(ins["means3D"], ins["opacities"], d["viewmatrix"], d["projmatrix"], d["campos"],
                             d["bg"], int(d["width"]), int(d["height"]), float(d["tanfovx"]), float(d["tanfovy"]),
                             sh_degree=int(d["sh_degree"]), shs=ins.get("shs"),
                             colors_precomp=ins.get("colors_precomp"), scales=ins.get("scales"),
                             rotations=ins.get("rotations"), cov3D_precomp=ins.get("cov3D_precomp"),
                             scale_modifier=float(d["scale_modifier"]), precision="f64")
    og = orc.backward(d["cotangent"])
    okey = {"colors_precomp": "colors", "cov3D_precomp": "cov3D"}
    for k, got in grads.items():
        ref = d["grad_" + k]
        assert maxrel(got, ref) < GRAD_TOL, (k, "vs reference golden", maxrel(got, ref))
        o = og[okey.get(k, k)].reshape(ref.shape)
        assert maxrel(got, o) < GRAD_TOL, (k, "vs float64 oracle", maxrel(got, o))
        assert normrel(got, o) < 1e-3, (k, normrel(got, o))


@pytest.mark.parametrize("name", ["small_sh3", "small_big_splats", "small_precomp_color"])
def test_tile_culling_is_exact(sgs, dev, name):
    """The staged ellipse/tile cull and the exp() short-circuit must not change a single bit."""
    d = load(name)
    _, c0, r0, d0, s0 = native_state(sgs, d, dev)
    _, c1, r1, d1, s1 = native_state(sgs, d, dev, _no_tile_cull=True)
    assert torch.equal(c0, c1) and torch.equal(d0, d1) and torch.equal(r0, r1)
    assert np.array_equal(s0["final_T"], s1["final_T"]) and np.array_equal(s0["tiles_touched"], s1["tiles_touched"])
    assert (s0["n_contrib"] <= s1["n_contrib"]).all()      # list positions in the culled lists can only shrink
    assert s0["kept"] <= s1["kept"] and s0["tile_count"].sum() <= s1["tile_count"].sum()


def test_backward_on_inference_state_fails_loudly(sgs, dev):
    d = load("small_sh3")
    ins = {k: v.to(dev) for k, v in inputs_of(d).items()}
    rs = settings_from(sgs, d, dev)
    e = torch.Tensor([])
    R, color, radii, gb, bb, ib, depth = sgs._C.rasterize_gaussians(
        rs.bg, ins["means3D"], e, ins["opacities"], ins["scales"], ins["rotations"], 1.0, e, rs.viewmatrix, rs.projmatrix,
        rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, ins["shs"], rs.sh_degree, rs.campos, False,
        keep_for_backward=False)
    with pytest.raises(RuntimeError, match="keep_for_backward=False"):
        sgs._C.rasterize_gaussians_backward(rs.bg, ins["means3D"], radii, e, ins["scales"], ins["rotations"], 1.0, e,
                                            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, torch.zeros_like(color),
                                            ins["shs"], rs.sh_degree, rs.campos, gb, R, bb, ib)


def test_inference_forward_equals_training_forward(sgs, dev):
    d = load("small_sh3")
    _, c0, _, d0, _ = native_state(sgs, d, dev, keep_for_backward=True)
    _, c1, _, d1, _ = native_state(sgs, d, dev, keep_for_backward=False)
    assert torch.equal(c0, c1) and torch.equal(d0, d1)
    with torch.no_grad():
        ins = {k: v.to(dev) for k, v in inputs_of(d).items()}
        rs = settings_from(sgs, d, dev)
        c2, _, d2 = sgs.GaussianRasterizer(rs)(means2D=torch.zeros_like(ins["means3D"]), **ins)
    assert torch.equal(c0, c2) and torch.equal(d0, d2)


def test_live_reference_ab(sgs, dev):
    """Same inputs through the compiled unmodified reference, when it travelled to this box."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not present")
    RefRast = ref_loader.ref_api()[1]
    for name in ("small_sh3", "small_precomp_cov"):
        d = load(name)
        cn, rn, dn, gn, _, _ = run_native(sgs, d, dev)
        cr, rr, dr, gr, _, _ = run_native(sgs, d, dev, Rast=RefRast)
        assert np.array_equal(cn, cr) and np.array_equal(rn, rr) and np.array_equal(dn, dr)
        for k in gn:
            assert maxrel(gn[k], gr[k]) < GRAD_TOL, (name, k, maxrel(gn[k], gr[k]))


def _sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


@pytest.mark.parametrize("cfg", ["config1_fwd", "config2_fwd"])
def test_full_size_bit_exact_tile_counts(sgs, dev, cfg):
    """BASELINE.json configs[0]/[1] at full size: radii, tile counts, n_contrib, image and depth hash-equal
    to the reference (north_star: 'bit-exact tile counts')."""
    from saro_gs_b200 import synthetic
    d = load(cfg)
    scene, cam = synthetic.config1_scene() if cfg == "config1_fwd" else synthetic.config2_scene()
    e = torch.Tensor([])
    args = (torch.zeros(3, device=dev), scene.means3D.to(dev), e, scene.opacities.to(dev), scene.scales.to(dev),
            scene.rotations.to(dev), 1.0, e, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.tanfovx, cam.tanfovy,
            cam.height, cam.width, scene.shs.to(dev), scene.sh_degree, cam.campos.to(dev), False)
    for no_cull in (False, True):
        R, color, radii, gb, bb, ib, depth = sgs._C.rasterize_gaussians(*args, _no_tile_cull=no_cull)
        st = sgs._C.debug_export(scene.means3D.shape[0], cam.width, cam.height, R, gb, bb, ib)
        assert R == int(d["num_rendered"])
        assert _sha(radii) == str(d["sha_radii"])
        assert _sha(st["tiles_touched"]) == str(d["sha_tiles_touched"])
        assert _sha(color) == str(d["sha_color"])
        assert _sha(depth) == str(d["sha_depth"])
        # size-independent properties
        rng = st["ranges"].long()
        assert R == int(st["tiles_touched"].long().sum())
        assert int((rng[:, 1] - rng[:, 0]).sum()) == st["kept"] == st["point_list"].numel()
        assert bool((st["tile_count"].long() <= (rng[:, 1] - rng[:, 0])).all())
        if no_cull:
            assert st["kept"] == R
            assert _sha(st["n_contrib"]) == str(d["sha_n_contrib"])     # the reference's per-pixel list positions
        else:
            assert st["kept"] < R


def test_full_size_backward_vs_reference_golden(sgs, dev):
    """configs[1] backward at full size against the compiled reference's gradients (tests/golden/config2_bwd.npz:
    per-tensor norms + a sample of entries, half of them the largest ones)."""
    from saro_gs_b200 import synthetic
    d = load("config2_bwd")
    scene, cam = synthetic.config2_scene()
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                           cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d,
                                                    opacities=leaves["opacities"], shs=leaves["shs"],
                                                    scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(synthetic.cotangent(cam.height, cam.width).to(dev))
    grads = {k: v.grad for k, v in leaves.items()}
    grads["means2D"] = m2d.grad
    # At this size the reference does not reproduce ITSELF to 1e-4 on the cancellation-prone tensors: two runs of
    # the compiled reference differ by 2.5e-4 (scales) and 1.1e-3 (rotations) of the largest entry, because its
    # float atomics sum in a different order every run (measured on B200 with tools/grad_noise.py; the native
    # kernels' own run-to-run spread is smaller: 2.3e-4 / 7e-4).  Those two tensors get a bar of a few times the
    # reference's own noise; everything else meets the 1e-4 bar.
    tol = {"scales": 1e-3, "rotations": 3e-3}
    for k, g in grads.items():
        flat = g.detach().reshape(-1).cpu().numpy().astype(np.float64)
        ref_val = d[f"val_{k}"].astype(np.float64)
        got_val = flat[d[f"idx_{k}"]]
        scale = float(d[f"maxabs_{k}"])
        err = np.abs(got_val - ref_val).max() / scale
        assert err < tol.get(k, GRAD_TOL), (k, err)
        assert normrel(got_val, ref_val) < 5e-4, (k, normrel(got_val, ref_val))
        assert abs(np.linalg.norm(flat) - float(d[f"norm_{k}"])) / float(d[f"norm_{k}"]) < GRAD_TOL, k


def test_full_size_backward_vs_f64_oracle(sgs, dev, oracle_mod):
    """configs[1] backward at FULL size (P = 300 000, 1352x1014), every entry of every gradient tensor, against the
    float64 arbiter: the C oracle in double precision run live on the host (about 6 s) and its committed pin
    tests/golden/config2_bwd_f64.npz (norms, column sums, a 32 768-entry stratified sample per tensor).

    Bars.  Measured on a B200 (tools/grad_vs_f64.py, profiles/r2a_grad_vs_f64.json): the compiled REFERENCE itself sits
    3.1e-4 (means3D), 8.2e-4 (scales), 6.3e-4 (rotations), 1.9e-4 (opacities), 2.4e-4 (means2D), 7.4e-5 (shs) of the
    largest entry away from float64 — float32 rounding inside its per-Gaussian expression trees on ill-conditioned
    Gaussians, identical to three digits for the native kernels, which evaluate the same trees.  So per tensor the
    native error must be <= 1e-4 (north_star) OR no worse than the reference's own distance to float64 (the fixture's
    referr_* values, 10 % slack for atomic-order noise); the vector as a whole must agree to the same rule in norm.
    The LARGEST-entry error of `scales` / `rotations` is set by a handful of needle-shaped Gaussians on which the order
    of the float atomics moves the result: two runs of the compiled reference differ from each other by 8.6e-4 /
    1.1e-3 there (tools/grad_noise.py), two native runs sit 6.3e-4 and 8.4e-4 from float64
    (profiles/r2a_grad_vs_f64.json) — so for these two tensors the largest-entry bar is the reference's own run-to-run
    spread, and in exchange 99.99 % of the entries must individually meet 1e-4."""
    from saro_gs_b200 import synthetic
    d = load("config2_bwd_f64")
    scene, cam = synthetic.config2_scene()
    cot_cpu = synthetic.cotangent(cam.height, cam.width)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                           cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d,
                                                    opacities=leaves["opacities"], shs=leaves["shs"],
                                                    scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(cot_cpu.to(dev))
    grads = {k: v.grad.detach().double().cpu().numpy() for k, v in leaves.items()}
    grads["means2D"] = m2d.grad.detach().double().cpu().numpy()

    live = oracle_mod.forward_scene(scene, cam, torch.zeros(3), precision="f64")
    assert live.num_rendered == int(d["num_rendered"])
    g64 = live.backward(cot_cpu)
    for k, got in grads.items():
        want = np.asarray(g64[k], dtype=np.float64).reshape(got.shape)
        # the live oracle reproduces its committed pin (sample, norm, column sums)
        flat = want.reshape(-1)
        # (to 1e-7 of the largest entry: the host's core count changes the OpenMP summation order of the oracle)
        assert np.abs(flat[d[f"idx_{k}"]] - d[f"val_{k}"]).max() <= 1e-7 * float(d[f"maxabs_{k}"]), k
        assert abs(np.linalg.norm(flat) - float(d[f"norm_{k}"])) <= 1e-7 * float(d[f"norm_{k}"]), k
        # native vs float64: the whole tensor
        err_max = maxrel(got, want)
        err_nrm = normrel(got, want)
        bar_max = max(GRAD_TOL, 1.10 * float(d[f"referr_max_{k}"]), REF_RUN_TO_RUN_MAX.get(k, 0.0))
        # 25 % slack on the norm: the same few Gaussians make it bimodal from run to run (rotations: native 1.5e-4 or
        # 2.3e-4, reference 1.96e-4 .. 2.06e-4 in profiles/r2a_grad_vs_f64.json)
        bar_nrm = max(GRAD_TOL, 1.25 * float(d[f"referr_norm_{k}"]))
        q9999 = float(np.quantile(np.abs(got - want).reshape(-1), 0.9999)) / max(float(np.abs(want).max()), 1e-30)
        print(f"{k:10s} max {err_max:.2e} (bar {bar_max:.2e})  norm {err_nrm:.2e} (bar {bar_nrm:.2e})  q99.99 {q9999:.2e}")
        assert err_max <= bar_max, (k, err_max, bar_max)
        assert err_nrm <= bar_nrm, (k, err_nrm, bar_nrm)
        assert q9999 <= GRAD_TOL, (k, q9999)
        # no systematic bias: the column sums of the error stay within 2e-5 of the column sums of |gradient| plus the
        # one-entry noise of the ill-conditioned Gaussians discussed above (2e-3 of the largest entry).  (Round 2a
        # compared the column sums themselves with 2e-3 of the largest column sum: where a column nearly cancels, the
        # run-to-run flip of a single needle-shaped Gaussian exceeded that and the test failed about one run in five.)
        G2, W2 = got.reshape(got.shape[0], -1), want.reshape(want.shape[0], -1)
        ref_colsum = d[f"colsum_{k}"]
        assert np.abs(W2.sum(axis=0) - ref_colsum).max() <= 1e-7 * np.abs(W2).sum(axis=0).max() + 1e-12, k   # pin
        bias = np.abs((G2 - W2).sum(axis=0))
        assert (bias <= 2e-5 * np.abs(W2).sum(axis=0) + 2e-3 * float(np.abs(W2).max())).all(), (k, bias)


def test_config3_sequence_bit_exact(sgs, dev):
    """BASELINE.json configs[2] stand-in: frames with a varying number of live Gaussians; colour, depth and
    radii hash-equal to the reference, for the SH pass and for the precomputed-colour ('lifespan') pass."""
    from saro_gs_b200 import synthetic
    d = load("config3_seq")
    base, cam = synthetic.config2_scene()
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), base.sh_degree,
                                           cam.campos.to(dev), False)
    rast = sgs.GaussianRasterizer(rs)
    with torch.no_grad():
        for k, t in enumerate(d["times"]):
            sc = synthetic.temporal_frame(base, float(t))
            assert sc.means3D.shape[0] == int(d[f"P_{k}"])
            m3, op, scl, rot, sh = (x.to(dev) for x in (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.shs))
            color, radii, depth = rast(means3D=m3, means2D=torch.zeros_like(m3), opacities=op, shs=sh, scales=scl,
                                       rotations=rot)
            assert _sha(color) == str(d[f"sha_color_{k}"]) and _sha(depth) == str(d[f"sha_depth_{k}"])
            assert _sha(radii) == str(d[f"sha_radii_{k}"])
            color2, _, _ = rast(means3D=m3, means2D=torch.zeros_like(m3), opacities=op, shs=None,
                                colors_precomp=op.expand(-1, 3), scales=scl, rotations=rot)
            assert _sha(color2) == str(d[f"sha_color_life_{k}"])


def test_config3_sequence_300_frames_bit_exact(sgs, dev):
    """BASELINE.json configs[2] at its stated size: 300 frames of the 300 k cloud with a varying number of live
    Gaussians (synthetic.DeviceSequence); colour, depth and radii hash-equal to the compiled reference on every frame
    (tests/golden/config3_seq300.npz, made on a B200 by tests/golden/make_golden.py seq300), and the API-visible
    num_rendered equal."""
    from saro_gs_b200 import synthetic
    d = load("config3_seq300")
    base, cam = synthetic.config2_scene()
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), base.sh_degree,
                                           cam.campos.to(dev), False)
    seq = synthetic.DeviceSequence(base, dev)
    frames = int(d["frames"])
    e = torch.Tensor([])
    bad = []
    with torch.no_grad():
        for k in range(frames):
            sc = seq.frame(k / frames)
            assert sc.means3D.shape[0] == int(d["P"][k]), \
                f"frame {k}: synthetic inputs are not reproducible on this platform (alive count differs from the fixture)"
            R, color, radii, gb, bb, ib, depth = sgs._C.rasterize_gaussians(
                rs.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, cam.height, cam.width, sc.shs, sc.sh_degree, rs.campos, False,
                keep_for_backward=False)
            ok = (R == int(d["R"][k]) and _sha(color) == str(d["sha_color"][k]) and _sha(depth) == str(d["sha_depth"][k])
                  and _sha(radii) == str(d["sha_radii"][k]))
            if not ok:
                bad.append(k)
    assert not bad, f"frames differing from the reference: {bad[:10]} ({len(bad)} of {frames})"


def test_full_size_backward_properties(sgs, dev):
    """configs[1] backward at full size: linear in the cotangent, zero for culled Gaussians,
    deterministic forward."""
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.config2_scene()
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), scene.sh_degree,
                                           cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)

    def grads(cot):
        for p in list(leaves.values()) + [m2d]:
            p.grad = None
        color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d,
                                                        opacities=leaves["opacities"], shs=leaves["shs"],
                                                        scales=leaves["scales"], rotations=leaves["rotations"])
        color.backward(cot)
        return color.detach(), radii, {k: v.grad.clone() for k, v in leaves.items()}

    gen = torch.Generator().manual_seed(3)
    a = (torch.randn(3, cam.height, cam.width, generator=gen) / (3 * cam.height * cam.width)).to(dev)
    b = (torch.randn(3, cam.height, cam.width, generator=gen) / (3 * cam.height * cam.width)).to(dev)
    c1, radii, ga = grads(a)
    c2, _, gb = grads(b)
    _, _, gab = grads(2.0 * a - 0.5 * b)
    assert torch.equal(c1, c2)                                   # forward is deterministic
    culled = radii == 0
    assert int(culled.sum()) > 0
    for k in ga:
        want = 2.0 * ga[k] - 0.5 * gb[k]
        # property check, NOT the parity bar: three independent float-atomic sums + a float32 linear
        # combination; the scale/rotation gradients are differences of large cov3D terms (cancellation),
        # so a single worst entry can sit near 1e-3 of the max while the vector as a whole agrees to ~1e-5
        # (run-to-run spread of these two tensors at this size, tools/grad_noise.py: ~2e-4 norm-relative, ~1e-3 max)
        noisy = k in ("scales", "rotations")
        assert normrel(gab[k].cpu().numpy(), want.cpu().numpy()) < (2e-3 if noisy else 1e-4), k
        assert maxrel(gab[k].cpu().numpy(), want.cpu().numpy()) < (1e-2 if noisy else 5e-4), k
        assert not gab[k][culled].any(), k                       # culled Gaussians get exact zeros
    assert not m2d.grad[:, 2].any()                              # dL/dmean2D.z is always 0


def test_edge_cases(sgs, dev):
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.small_scene(P=64, seed=9)
    bg = torch.tensor([0.2, 0.4, 0.6], device=dev)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, bg, 1.0,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    rast = sgs.GaussianRasterizer(rs)
    # P == 0: zeros (not background), empty radii — SURVEY.md Appendix A.2
    z = lambda *s: torch.zeros(*s, device=dev)
    color, radii, depth = rast(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), shs=z(0, 16, 3), scales=z(0, 3),
                               rotations=z(0, 4))
    assert color.shape == (3, cam.height, cam.width) and not color.any() and not depth.any() and radii.numel() == 0
    # everything culled: background, depth 15, radii 0, zero grads
    m = z(8, 3).requires_grad_(True)
    color, radii, depth = rast(means3D=m, means2D=z(8, 3), opacities=torch.ones(8, 1, device=dev),
                               colors_precomp=torch.ones(8, 3, device=dev), scales=torch.ones(8, 3, device=dev) * .1,
                               rotations=torch.tensor([[1.0, 0, 0, 0]], device=dev).repeat(8, 1))
    color.sum().backward()
    assert torch.equal(color, bg[:, None, None].expand_as(color)) and (depth == 15.0).all() and not radii.any()
    assert not m.grad.any()
    # markVisible == (z_view > 0.2)
    vis = rast.markVisible(scene.means3D.to(dev))
    assert vis.dtype == torch.bool and torch.equal(vis.cpu(), scene.means3D[:, 2] > 0.2)
    # non-contiguous / strided inputs are accepted like the reference's .contiguous()
    big = torch.randn(64, 6, device=dev)
    color2, _, _ = rast(means3D=scene.means3D.to(dev), means2D=z(64, 3), opacities=scene.opacities.to(dev),
                        colors_precomp=big[:, ::2].abs(), scales=scene.scales.to(dev), rotations=scene.rotations.to(dev))
    color3, _, _ = rast(means3D=scene.means3D.to(dev), means2D=z(64, 3), opacities=scene.opacities.to(dev),
                        colors_precomp=big[:, ::2].abs().contiguous(), scales=scene.scales.to(dev),
                        rotations=scene.rotations.to(dev))
    assert torch.equal(color2, color3)


def test_unmodified_reference_style_caller(sgs, dev):
    """The call pattern of renderer/__init__.py:119-127,191-226 of the reference: keyword call,
    screenspace_points.retain_grad(), second pass with colors_precomp, through the drop-in module name."""
    from diff_gaussian_rasterization_ch3 import GaussianRasterizationSettings, GaussianRasterizer
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.small_scene(P=256, seed=11)
    rs = GaussianRasterizationSettings(image_height=cam.height, image_width=cam.width, tanfovx=cam.tanfovx,
                                       tanfovy=cam.tanfovy, bg=torch.zeros(3, device=dev), scale_modifier=1.0,
                                       viewmatrix=cam.viewmatrix.to(dev), projmatrix=cam.projmatrix.to(dev),
                                       sh_degree=3, campos=cam.campos.to(dev), prefiltered=False)
    rasterizer = GaussianRasterizer(raster_settings=rs)
    means3D = scene.means3D.to(dev).requires_grad_(True)
    screenspace_points = torch.zeros_like(means3D, requires_grad=True) + 0
    screenspace_points.retain_grad()
    rendered_image, radii, depth = rasterizer(means3D=means3D, means2D=screenspace_points, shs=scene.shs.to(dev),
                                              colors_precomp=None, opacities=scene.opacities.to(dev),
                                              scales=scene.scales.to(dev), rotations=scene.rotations.to(dev),
                                              cov3D_precomp=None)
    rendered_image.mean().backward()
    assert screenspace_points.grad is not None and screenspace_points.grad[radii > 0].abs().sum() > 0
    lifespan = torch.rand(256, 1, device=dev)
    img2, _, _ = rasterizer(means3D=means3D, means2D=screenspace_points, shs=None,
                            colors_precomp=lifespan.expand(-1, 3), opacities=scene.opacities.to(dev),
                            scales=scene.scales.to(dev), rotations=scene.rotations.to(dev), cov3D_precomp=None)
    assert img2.shape == rendered_image.shape and torch.isfinite(img2).all()


@pytest.mark.parametrize("seed", range(8))
def test_culling_fuzz_extreme_shapes(sgs, dev, seed):
    """The three conservative culls (alpha-box rect clip, quadrant masks, power threshold) against the unculled
    path on deliberately nasty inputs: needle-like and pancake Gaussians, opacities straddling 1/255, huge splats,
    centres far off screen.  Image, depth, radii and final_T must be bit-identical."""
    gen = torch.Generator().manual_seed(1000 + seed)
    P, W, H, fx = 600, 112, 96, 100.0
    z = torch.rand(P, generator=gen) * 8.0 + 0.25
    spread = 3.0 if seed % 2 else 1.2                      # odd seeds: many centres off screen
    x = (torch.rand(P, generator=gen) * 2 - 1) * spread * (W / (2 * fx)) * z
    y = (torch.rand(P, generator=gen) * 2 - 1) * spread * (H / (2 * fx)) * z
    means = torch.stack([x, y, z], 1)
    scales = torch.exp(torch.randn(P, 3, generator=gen) * 2.0 - 2.5)       # 3 decades of anisotropy
    q = torch.randn(P, 4, generator=gen)
    rots = q / q.norm(dim=1, keepdim=True)
    op = torch.sigmoid(torch.randn(P, 1, generator=gen) * 3.0)
    op[: P // 6] = (1.0 / 255.0) * (1.0 + (torch.rand(P // 6, 1, generator=gen) - 0.5) * 0.02)   # at the threshold
    op[P // 6: P // 5] = 1.0
    cols = torch.rand(P, 3, generator=gen)
    from saro_gs_b200 import synthetic
    cam = synthetic.make_camera(W, H, fx)
    args = lambda: (torch.tensor([0.1, 0.2, 0.3], device=dev), means.to(dev), cols.to(dev), op.to(dev), scales.to(dev),
                    rots.to(dev), 1.0, torch.Tensor([]), cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.tanfovx,
                    cam.tanfovy, H, W, torch.Tensor([]), 0, cam.campos.to(dev), False)
    R0, c0, r0, g0, b0, i0, d0 = sgs._C.rasterize_gaussians(*args())
    R1, c1, r1, g1, b1, i1, d1 = sgs._C.rasterize_gaussians(*args(), _no_tile_cull=True)
    s0 = sgs._C.debug_export(P, W, H, R0, g0, b0, i0)
    s1 = sgs._C.debug_export(P, W, H, R1, g1, b1, i1)
    assert R0 == R1 and torch.equal(r0, r1)
    assert torch.equal(c0, c1) and torch.equal(d0, d1)
    assert torch.equal(s0["final_T"], s1["final_T"]) and torch.equal(s0["tiles_touched"], s1["tiles_touched"])
    assert s0["kept"] <= s1["kept"] == R1


def test_more_than_65536_tiles_uses_32_bit_keys(sgs, dev, oracle_mod):
    """4112 x 4112 = 257 x 257 = 66049 tiles: the binning falls back from 16-bit to 32-bit tile keys."""
    from saro_gs_b200 import synthetic
    W = H = 4112
    gen = torch.Generator().manual_seed(77)
    P = 48
    cam = synthetic.make_camera(W, H, 3000.0)
    z = torch.rand(P, generator=gen) * 3 + 2
    x = (torch.rand(P, generator=gen) * 2 - 1) * (W / 6000.0) * z
    y = (torch.rand(P, generator=gen) * 2 - 1) * (H / 6000.0) * z
    means = torch.stack([x, y, z], 1)
    scales = torch.exp(torch.randn(P, 3, generator=gen) * 0.5 - 3.0)
    q = torch.randn(P, 4, generator=gen)
    rots = q / q.norm(dim=1, keepdim=True)
    op = torch.rand(P, 1, generator=gen) * 0.9 + 0.05
    cols = torch.rand(P, 3, generator=gen)
    bg = torch.tensor([0.0, 0.0, 0.0])
    rs = sgs.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg.to(dev), 1.0, cam.viewmatrix.to(dev),
                                           cam.projmatrix.to(dev), 0, cam.campos.to(dev), False)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=means.to(dev), means2D=torch.zeros(P, 3, device=dev),
                                                    opacities=op.to(dev), colors_precomp=cols.to(dev),
                                                    scales=scales.to(dev), rotations=rots.to(dev))
    ref = oracle_mod.forward(means, op, cam.viewmatrix, cam.projmatrix, cam.campos, bg, W, H, cam.tanfovx, cam.tanfovy,
                             sh_degree=0, colors_precomp=cols, scales=scales, rotations=rots, precision="f32")
    assert np.array_equal(radii.cpu().numpy(), ref.radii)
    # the float32 C oracle is not bit-compatible with the GPU (no FMA contraction) and pixel coordinates reach 4111,
    # so single pixels can flip the alpha >= 1/255 test (measured: 3.5e-3 = one threshold flip, identical for the
    # compiled reference): a loose value check here, bit-exactness against the live reference below
    err = np.abs(color.cpu().numpy() - ref.color).max()
    assert err < 5e-3, err
    assert np.abs(color.cpu().numpy() - ref.color).mean() < 1e-6
    assert int((radii > 0).sum()) > 10 and float(color.max()) > 0.05
    from oracle import ref_loader
    if ref_loader.available():
        RefRast = ref_loader.ref_api()[1]
        c2, r2, d2 = RefRast(rs)(means3D=means.to(dev), means2D=torch.zeros(P, 3, device=dev), opacities=op.to(dev),
                                 colors_precomp=cols.to(dev), scales=scales.to(dev), rotations=rots.to(dev))
        assert torch.equal(color, c2) and torch.equal(depth, d2) and torch.equal(radii, r2)


def test_non_finite_inputs_do_not_hang_or_crash(sgs, dev):
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.small_scene(P=128, seed=21)
    means = scene.means3D.clone()
    scales = scene.scales.clone()
    op = scene.opacities.clone()
    means[3] = float("nan")
    means[7, 0] = float("inf")
    scales[11] = float("inf")
    scales[13] = 0.0
    op[17] = float("nan")
    op[19] = -1.0
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev),
                                           1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    m = means.to(dev).requires_grad_(True)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=m, means2D=torch.zeros_like(m), opacities=op.to(dev),
                                                    shs=scene.shs.to(dev), scales=scales.to(dev),
                                                    rotations=scene.rotations.to(dev))
    color.nan_to_num().sum().backward()
    torch.cuda.synchronize()
    assert color.shape == (3, cam.height, cam.width) and m.grad is not None


@pytest.mark.parametrize("deg,M,mod", [(2, 9, 1.0), (1, 4, 0.6), (0, 1, 1.7), (3, 16, 0.8)])
def test_sh_layouts_and_scale_modifier_vs_oracle(sgs, dev, oracle_mod, deg, M, mod):
    """SH tensors with M != 16 coefficients take the generic (non-vectorised) SH path; scale_modifier != 1 flows
    through forward and backward.  Checked against the float64 oracle."""
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.small_scene(P=400, seed=50 + M)
    shs = scene.shs[:, :M, :].contiguous()
    bg = torch.tensor([0.1, 0.3, 0.2])
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, bg.to(dev), mod,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), deg, cam.campos.to(dev), False)
    leaves = dict(means3D=scene.means3D, scales=scene.scales, rotations=scene.rotations, opacities=scene.opacities, shs=shs)
    leaves = {k: v.to(dev).clone().requires_grad_(True) for k, v in leaves.items()}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means2D=m2d, **leaves)
    cot = synthetic.cotangent(cam.height, cam.width, seed=9)
    color.backward(cot.to(dev))
    orc = oracle_mod.forward(scene.means3D, scene.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, bg, cam.width,
                             cam.height, cam.tanfovx, cam.tanfovy, sh_degree=deg, shs=shs, scales=scene.scales,
                             rotations=scene.rotations, scale_modifier=mod, precision="f64")
    og = orc.backward(cot)
    assert np.array_equal(radii.cpu().numpy(), orc.radii)
    assert np.abs(color.detach().cpu().numpy() - orc.color).max() < 1e-4
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert maxrel(leaves[k].grad.cpu().numpy(), og[k].reshape(leaves[k].shape)) < GRAD_TOL, k


def test_runs_on_a_non_default_stream(sgs, dev):
    """All work is enqueued on the caller's current stream (the reference uses the legacy default stream only)."""
    d = load("small_sh3")
    c0, r0, d0, g0, _, _ = run_native(sgs, d, dev)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        c1, r1, d1, g1, _, _ = run_native(sgs, d, dev)
    side.synchronize()
    assert np.array_equal(c0, c1) and np.array_equal(r0, r1) and np.array_equal(d0, d1)
    for k in g0:
        assert maxrel(g1[k], g0[k]) < GRAD_TOL, k


def test_backward_twice_on_the_same_forward_state(sgs, dev):
    """retain_graph=True: the second backward pass over the same saved state must find the moment accumulator of the
    geometry buffer zeroed again (the backward-preprocess kernel re-zeroes what it consumes; there is no memset)."""
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.config2_scene(P=20_000, width=320, height=240, fx=250.0)
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means2D=m2d, **leaves)
    cot = synthetic.cotangent(cam.height, cam.width).to(dev)
    color.backward(cot, retain_graph=True)
    first = {k: v.grad.clone() for k, v in leaves.items()}
    for v in leaves.values():
        v.grad = None
    color.backward(cot)
    for k, v in leaves.items():
        scale = float(first[k].abs().max())
        # only the order of the float atomics differs (a stale accumulator would double the gradient)
        assert float((v.grad - first[k]).abs().max()) <= 1e-3 * scale, k


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_vs_live_reference(sgs, dev, seed):
    """Random scenes, image sizes, ROTATED cameras, SH degrees, scale modifiers and backgrounds through both the
    native library and the compiled unmodified reference: forward bit-exact, gradients within tolerance."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not present")
    from saro_gs_b200 import synthetic
    import math
    RefRast = ref_loader.ref_api()[1]
    gen = torch.Generator().manual_seed(7000 + seed)
    r = lambda lo, hi: lo + (hi - lo) * float(torch.rand(1, generator=gen))
    W, H = int(r(40, 300)), int(r(40, 220))
    P = int(r(50, 4000))
    fx = r(0.6, 1.6) * W
    yaw, pitch = r(-0.5, 0.5), r(-0.3, 0.3)
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    Rx = torch.tensor([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=torch.float64)
    cam = synthetic.make_camera(W, H, fx, fy=fx * r(0.8, 1.25), R=(Rx @ Ry), t=(r(-0.5, 0.5), r(-0.5, 0.5), r(0.0, 2.0)))
    means = (torch.rand(P, 3, generator=gen) * 2 - 1) * torch.tensor([3.0, 3.0, 3.0]) + torch.tensor([0.0, 0.0, 3.0])
    scales = torch.exp(torch.randn(P, 3, generator=gen) * r(0.3, 1.2) + r(-3.5, -1.5))
    q = torch.randn(P, 4, generator=gen)
    rots = q / q.norm(dim=1, keepdim=True)
    op = torch.sigmoid(torch.randn(P, 1, generator=gen) * 2.0)
    deg = int(r(0, 3.999))
    shs = torch.randn(P, 16, 3, generator=gen) * 0.3
    bg = torch.rand(3, generator=gen)
    mod = r(0.5, 1.5)
    rs = sgs.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg.to(dev), mod, cam.viewmatrix.to(dev),
                                           cam.projmatrix.to(dev), deg, cam.campos.to(dev), False)
    cot = torch.randn(3, H, W, generator=gen).to(dev)
    outs = []
    for Rast in (sgs.GaussianRasterizer, RefRast, RefRast, RefRast):
        leaves = {k: v.to(dev).clone().requires_grad_(True)
                  for k, v in dict(means3D=means, scales=scales, rotations=rots, opacities=op, shs=shs).items()}
        m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii, depth = Rast(rs)(means2D=m2d, **leaves)
        color.backward(cot)
        g = {k: v.grad for k, v in leaves.items()}
        g["means2D"] = m2d.grad
        outs.append((color.detach(), radii, depth.detach(), g))
    (c0, r0, d0, g0), (c1, r1, d1, g1), (_, _, _, g2), (_, _, _, g3) = outs
    assert torch.equal(r0, r1) and torch.equal(c0, c1) and torch.equal(d0, d1)
    g64 = None      # float64 oracle gradients, computed only if some row needs the arbiter
    for k in g0:
        scale = float(g1[k].abs().max())
        if scale == 0.0:
            assert not g0[k].any(), k
            continue
        # The reference sums with float atomics in a different order every run; for ill-conditioned Gaussians
        # (needle / pancake shapes: the cov3D -> scale, rotation, mean chain cancels by 1e4 and more) two runs of the
        # REFERENCE differ by percents of the largest entry (seed 8: 5e-2 on means3D; the native kernels' own spread
        # is 3e-5 there and they sit closer to the float64 oracle than the reference does).  Rows on which the
        # reference does not reproduce itself to 1e-5 cannot arbitrate and are left out; every other row must agree
        # to 2e-4 of the largest entry, and such rows must be the overwhelming majority.
        # Three reference runs decide which rows are stable (two runs can agree by chance: the full suite then failed
        # about once in ten runs on a "stable" row).  A stable row that still differs by more than 2e-4 goes to the
        # float64 oracle: it must find the native value at least as close to the truth as the reference's.
        a, b1, b2, b3 = (t.reshape(t.shape[0], -1) for t in (g0[k], g1[k], g2[k], g3[k]))
        stable = ((b1 - b2).abs().amax(dim=1) <= 1e-5 * scale) & ((b1 - b3).abs().amax(dim=1) <= 1e-5 * scale)
        assert float(stable.float().mean()) > 0.95, (k, float(stable.float().mean()))
        row_err = (a - b1).abs().amax(dim=1) / scale
        doubtful = stable & (row_err >= 2e-4)
        if bool(doubtful.any()):
            if g64 is None:
                from oracle import oracle as oracle_mod
                ctx = oracle_mod.forward(means, op, cam.viewmatrix, cam.projmatrix, cam.campos, bg, W, H, cam.tanfovx,
                                         cam.tanfovy, sh_degree=deg, shs=shs, scales=scales, rotations=rots,
                                         scale_modifier=mod, precision="f64")
                g64 = ctx.backward(cot.cpu())
            o = torch.as_tensor(np.asarray(g64[k], dtype=np.float64)).reshape(a.shape[0], -1).to(dev)
            nat = (a.double() - o).abs().amax(dim=1)[doubtful]
            ref = (b1.double() - o).abs().amax(dim=1)[doubtful]
            assert bool((nat <= 1.05 * ref + 1e-6 * scale).all()), (k, int(doubtful.sum()), float(row_err[doubtful].max()))

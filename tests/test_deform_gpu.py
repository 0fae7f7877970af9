"""GPU parity of the fused deformation -> rasterizer hand-off (csrc/sgs_deform.cu) through the C ABI:
against the reference's own outputs (tests/golden/deform_*.npz) and the numpy oracle.
Tolerance: 1e-4 of the largest entry of each output tensor (BASELINE north_star's floating-point bar)."""
import glob
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import deform_oracle
from saro_gs_b200 import deformation

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "deform_*.npz")))
OUTS = ("means3D", "rotations", "scales", "opacity", "shs")
TOL = 1e-4


def load_case(path):
    z = np.load(path)
    inputs = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    mlps = {name: tuple(z[f"mlp_{name}_{p}{i}"] for i in (1, 2, 3) for p in ("W", "b")) for name in deform_oracle.MLP_NAMES}
    return z, inputs, mlps, float(z["timestamp"])


def to_dev(inputs, mlps):
    dev = torch.device("cuda:0")
    ti = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in inputs.items()}
    tm = {k: tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in v) for k, v in mlps.items()}
    return ti, tm


def run_native(t, inputs, mlps):
    ti, tm = to_dev(inputs, mlps)
    packed = deformation.PackedMLPs(tm["motion"], tm["rot"], tm["shs"])
    out = deformation.deformation_eval(t, ti["xyz"], ti["rotation"], ti["scaling"], ti["opacity"], ti["features_dc"],
                                       ti["features_rest"], ti["temporal_pos"], ti["lifespan"], ti["hexplane_feature"], packed)
    torch.cuda.synchronize()
    return dict(zip(OUTS, (o.cpu().numpy() for o in out)))


def assert_close(got, want, what):
    for k in OUTS:
        assert got[k].shape == want[k].shape, (what, k, got[k].shape, want[k].shape)
        if want[k].size == 0:
            continue
        err = np.abs(got[k].astype(np.float64) - want[k]).max() / max(np.abs(want[k]).max(), 1e-12)
        assert err <= TOL, (what, k, err)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_against_reference_golden(path):
    z, inputs, mlps, t = load_case(path)
    got = run_native(t, inputs, mlps)
    assert_close(got, {k: z[f"f32_{k}"] for k in OUTS}, "reference float32")
    assert_close(got, {k: z[f"f64_{k}"] for k in OUTS}, "reference float64")


def random_case(n, feat_dim, seed, life=(0.05, 1.0)):
    rng = np.random.default_rng(seed)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    inputs = dict(xyz=f(n, 3) * 2, rotation=f(n, 4), scaling=f(n, 3) * 0.5 - 3.5, opacity=f(n, 1) * 2,
                  features_dc=f(n, 1, 3) * 0.5, features_rest=f(n, 15, 3) * 0.1,
                  temporal_pos=rng.random((n, 1), dtype=np.float32),
                  lifespan=(rng.random((n, 1), dtype=np.float32) * (life[1] - life[0]) + life[0]).astype(np.float32),
                  hexplane_feature=f(n, feat_dim) * 0.5)
    mlps = {}
    for name, out in zip(deform_oracle.MLP_NAMES, (3, 7, 48)):
        dims = [(128, feat_dim + 9), (128, 128), (out, 128)]
        ps = []
        for (o, i) in dims:
            lim = np.sqrt(6.0 / (o + i))
            ps += [rng.uniform(-lim, lim, (o, i)).astype(np.float32), rng.uniform(-0.1, 0.1, (o,)).astype(np.float32)]
        mlps[name] = tuple(ps)
    return inputs, mlps


def exact_mask_rows(t, inputs):
    """Rows whose survival state is not within float32 rounding of the 0.001 threshold (selection is exact there)."""
    st, _ = deform_oracle.survival_state(t, inputs["temporal_pos"].astype(np.float64), inputs["lifespan"].astype(np.float64))
    return np.abs(st.reshape(-1) - 0.001) > 1e-6


@pytest.mark.parametrize("n,feat_dim,t", [(70_001, 32, 0.37), (128, 32, 0.5), (129, 16, 0.2), (1, 32, 0.5), (5000, 24, 0.9),
                                          (4097, 8, 0.1), (40_000, 16, 0.6)])
def test_against_oracle_random(n, feat_dim, t):
    inputs, mlps = random_case(n, feat_dim, seed=n + feat_dim)
    inputs["temporal_pos"][~exact_mask_rows(t, inputs)] = t      # rows on the threshold: make them unambiguous
    assert exact_mask_rows(t, inputs).all()
    want = deform_oracle.deformation_eval(t, mlps=mlps, dtype=np.float64, **inputs)
    got = run_native(t, inputs, mlps)
    assert_close(got, want, f"oracle n={n}")


def test_nothing_selected_and_empty_cloud():
    inputs, mlps = random_case(300, 32, seed=5)
    got = run_native(50.0, inputs, mlps)          # far outside every lifespan
    assert got["means3D"].shape == (0, 3) and got["shs"].shape == (0, 16, 3) and got["opacity"].shape == (0, 1)
    empty = {k: v[:0] for k, v in inputs.items()}
    got = run_native(0.5, empty, mlps)
    assert got["rotations"].shape == (0, 4)


def test_drop_in_method_and_weight_refresh():
    """The GaussianModel-method form (scene/saro_gaussian.py:871) on a stand-in model; weights changed in place are
    re-packed on the next call."""
    inputs, mlps = random_case(2000, 32, seed=11)
    ti, tm = to_dev(inputs, mlps)

    def seq(ps):
        W1, b1, W2, b2, W3, b3 = ps
        m = torch.nn.Sequential(torch.nn.Linear(W1.shape[1], 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                                torch.nn.Linear(128, W3.shape[0])).cuda()
        with torch.no_grad():
            for layer, (W, b) in zip([m[0], m[2], m[4]], ((W1, b1), (W2, b2), (W3, b3))):
                layer.weight.copy_(W)
                layer.bias.copy_(b)
        return m

    pc = types.SimpleNamespace(
        args=types.SimpleNamespace(dx=True, drot=True, dopacity=True, dsh=True),
        _xyz=ti["xyz"], _rotation=ti["rotation"], _scaling=ti["scaling"], _opacity=ti["opacity"],
        _features_dc=ti["features_dc"], _features_rest=ti["features_rest"], get_temporalpos=ti["temporal_pos"],
        _lifespan=ti["lifespan"], hexplane_feature=ti["hexplane_feature"],
        motion_mlp=seq(tm["motion"]), rot_mlp=seq(tm["rot"]), shs_mlp=seq(tm["shs"]))
    with torch.no_grad():
        out = deformation.get_deformation_eval(pc, 0.45)
    want = deform_oracle.deformation_eval(0.45, mlps=mlps, dtype=np.float64, **inputs)
    assert_close(dict(zip(OUTS, (o.cpu().numpy() for o in out))), want, "method form")
    with torch.no_grad():
        pc.motion_mlp[4].bias.add_(1.0)
        out2 = deformation.get_deformation_eval(pc, 0.45)
    assert torch.allclose(out2[0], out[0] + 1.0, atol=1e-5)
    pc.args.dsh = False
    with pytest.raises(deformation.UnsupportedDeformationConfig):
        deformation.get_deformation_eval(pc, 0.45)


def test_feeds_the_rasterizer():
    """End of the hand-off: the outputs go straight into the rasterizer (renderer/__init__.py:190-199)."""
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic
    scene, cam = synthetic.config2_scene(P=20_000, width=320, height=240, fx=250.0)
    n = scene.means3D.shape[0]
    inputs, mlps = random_case(n, 32, seed=3, life=(0.3, 2.0))
    inputs["xyz"] = scene.means3D.numpy()
    inputs["scaling"] = np.log(scene.scales.numpy())
    inputs["rotation"] = scene.rotations.numpy()
    ti, tm = to_dev(inputs, mlps)
    packed = deformation.PackedMLPs(tm["motion"], tm["rot"], tm["shs"])
    m3, rot, sc, op, shs = deformation.deformation_eval(0.5, ti["xyz"], ti["rotation"], ti["scaling"], ti["opacity"],
                                                        ti["features_dc"], ti["features_rest"], ti["temporal_pos"],
                                                        ti["lifespan"], ti["hexplane_feature"], packed)
    dev = m3.device
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 3, cam.campos.to(dev), False)
    with torch.no_grad():
        color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=m3, means2D=torch.zeros_like(m3), opacities=op, shs=shs,
                                                         scales=sc, rotations=rot)
    assert color.shape == (3, cam.height, cam.width) and torch.isfinite(color).all()
    assert radii.shape[0] == m3.shape[0] and (radii > 0).any()


def test_rejects_cpu_tensors_and_wide_features():
    inputs, mlps = random_case(10, 32, seed=1)
    ti, tm = to_dev(inputs, mlps)
    packed = deformation.PackedMLPs(tm["motion"], tm["rot"], tm["shs"])
    with pytest.raises(RuntimeError):
        deformation.deformation_eval(0.5, ti["xyz"].cpu(), ti["rotation"], ti["scaling"], ti["opacity"], ti["features_dc"],
                                     ti["features_rest"], ti["temporal_pos"], ti["lifespan"], ti["hexplane_feature"], packed)
    _, wide = random_case(10, 64, seed=1)
    _, tw = to_dev({}, wide)
    with pytest.raises(deformation.UnsupportedDeformationConfig):
        deformation.PackedMLPs(tw["motion"], tw["rot"], tw["shs"])

"""The C ABI driven from plain C (tests/abi/abi_driver.c: gcc, cudaMalloc, no torch, no C++) must give exactly what
the Python host layer gives on the same inputs — the drop-in boundary is the C library, not the Python package."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi", "abi_driver.c")
EXE = os.path.join(ROOT, "tests", "abi", "_build", "abi_driver")


SRC_DEFORM = os.path.join(ROOT, "tests", "abi", "deform_driver.c")
EXE_DEFORM = os.path.join(ROOT, "tests", "abi", "_build", "deform_driver")


SRC_TRAIN = os.path.join(ROOT, "tests", "abi", "deform_train_driver.c")
EXE_TRAIN = os.path.join(ROOT, "tests", "abi", "_build", "deform_train_driver")


def build_driver(SRC=SRC, EXE=EXE):
    from saro_gs_b200 import build as native_build
    lib = native_build.build(verbose=False)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(SRC), os.path.getmtime(lib)):
        return EXE
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    libdir = os.path.dirname(lib)
    cmd = [cc, "-O2", "-std=c99", "-Wall", SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           "-L", libdir, "-lsaro_gs_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + libdir,
           "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def test_c_driver_compiles_against_the_header():
    """CPU: the header is valid C99 and every entry point the drivers use links against the built library."""
    assert os.path.exists(build_driver())
    assert os.path.exists(build_driver(SRC_DEFORM, EXE_DEFORM))
    assert os.path.exists(build_driver(SRC_TRAIN, EXE_TRAIN))


@pytest.mark.gpu
def test_c_driver_matches_python_host_layer(tmp_path):
    import saro_gs_b200 as sgs
    from saro_gs_b200 import synthetic
    exe = build_driver()
    dev = torch.device("cuda:0")
    scene, cam = synthetic.small_scene(P=700, seed=41, width=112, height=80, fx=100.0)
    bg = torch.tensor([0.2, 0.1, 0.4])
    cot = synthetic.cotangent(cam.height, cam.width, seed=5) * 1000.0
    P, M, D = scene.means3D.shape[0], 16, 3
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<5i3f", P, D, M, cam.width, cam.height, cam.tanfovx, cam.tanfovy, 1.0))
        for t in (bg, cam.viewmatrix, cam.projmatrix, cam.campos, scene.means3D, scene.shs, scene.opacities, scene.scales,
                  scene.rotations, cot):
            f.write(t.contiguous().numpy().astype("<f4").tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = open(fout, "rb").read()
    R = struct.unpack_from("<q", raw, 0)[0]
    off = 8

    def take(n, dt="<f4"):
        nonlocal off
        a = np.frombuffer(raw, dtype=dt, count=n, offset=off)
        off += a.nbytes
        return a

    HW = cam.height * cam.width
    c_color, c_depth, c_radii = take(3 * HW), take(HW), take(P, "<i4")
    c = dict(means3D=take(P * 3), opacities=take(P), scales=take(P * 3), rotations=take(P * 4), shs=take(P * M * 3),
             means2D=take(P * 3))
    # the same inputs through the Python host layer
    rs = sgs.GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, bg.to(dev), 1.0,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), D, cam.campos.to(dev), False)
    leaves = {k: getattr(scene, k).to(dev).requires_grad_(True)
              for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth = sgs.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d,
                                                    opacities=leaves["opacities"], shs=leaves["shs"],
                                                    scales=leaves["scales"], rotations=leaves["rotations"])
    color.backward(cot.to(dev))
    assert np.array_equal(c_color, color.detach().cpu().numpy().ravel())      # forward is deterministic: bit-equal
    assert np.array_equal(c_depth, depth.detach().cpu().numpy().ravel())
    assert np.array_equal(c_radii, radii.cpu().numpy())
    assert R > 0
    grads = dict(leaves, means2D=m2d)
    for k, v in c.items():
        ref = grads[k].grad.cpu().numpy().ravel()
        assert np.abs(v - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-12, k     # float atomics: order-dependent


@pytest.mark.gpu
def test_c_driver_widened_rows_match_python_host_layer(tmp_path):
    """Deformation hand-off and densification statistics from plain C == through the Python host layer (bit-equal:
    both are deterministic)."""
    import types
    from saro_gs_b200 import deformation, synthetic
    from saro_gs_b200.densify import BatchDensifyStats
    exe = build_driver(SRC_DEFORM, EXE_DEFORM)
    dev = torch.device("cuda:0")
    scene, _ = synthetic.small_scene(P=1000, seed=9)
    pc = synthetic.dynamic_model(scene, feat_dim=32, seed=4)
    N, t = 1000, 0.35
    g = torch.Generator().manual_seed(2)
    dm2 = torch.randn(N, 3, generator=g) * 1e-3
    radii = torch.where(torch.rand(N, generator=g) < 0.3, 0, torch.randint(1, 50, (N,), generator=g)).to(torch.int32)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<2if", N, 32, t))
        for a in (pc._xyz, pc._rotation, pc._scaling, pc._opacity, pc._features_dc, pc._features_rest, pc.get_temporalpos,
                  pc._lifespan, pc.hexplane_feature):
            f.write(a.contiguous().numpy().astype("<f4").tobytes())
        for m in (pc.motion_mlp, pc.rot_mlp, pc.shs_mlp):
            for layer in (m[0], m[2], m[4]):
                f.write(layer.weight.detach().contiguous().numpy().astype("<f4").tobytes())
                f.write(layer.bias.detach().contiguous().numpy().astype("<f4").tobytes())
        f.write(dm2.numpy().astype("<f4").tobytes())
        f.write(radii.numpy().astype("<i4").tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = open(fout, "rb").read()
    S = struct.unpack_from("<q", raw, 0)[0]
    off = 8

    def take(n):
        nonlocal off
        a = np.frombuffer(raw, dtype="<f4", count=n, offset=off)
        off += a.nbytes
        return a

    c_out = [take(S * 3), take(S * 4), take(S * 3), take(S), take(S * 48)]
    c_stats = [take(N), take(N), take(N)]
    with torch.no_grad():
        py = deformation.get_deformation_eval(synthetic.model_to(pc, dev), t)
    assert py[0].shape[0] == S and 0 < S < N
    for a, b in zip(c_out, py):
        assert np.array_equal(a, b.cpu().numpy().ravel())
    stats = BatchDensifyStats(N, dev)
    stats.add_view(dm2.to(dev), radii.to(dev))
    m = types.SimpleNamespace(max_radii2D=torch.zeros(N, device=dev), xyz_gradient_accum=torch.zeros(N, 1, device=dev),
                              denom=torch.zeros(N, 1, device=dev))
    stats.commit(m)
    for a, b in zip(c_stats, (m.max_radii2D, m.xyz_gradient_accum, m.denom)):
        assert np.array_equal(a, b.cpu().numpy().ravel())


@pytest.mark.gpu
def test_c_driver_training_deformation_matches_python_host_layer(tmp_path):
    """One MLP evaluation of the training-time deformation from plain C — forward, data gradients, weight gradients,
    with the operand planes and sign bits carried between the calls — == the Python host layer's autograd Function on
    the same inputs (bit-equal: every kernel on this path is deterministic)."""
    from oracle import deform_torch
    from saro_gs_b200 import deformation
    exe = build_driver(SRC_TRAIN, EXE_TRAIN)
    dev = torch.device("cuda:0")
    N, F, t = 1111, 32, 0.45
    g = torch.Generator().manual_seed(17)
    tpos = torch.rand(N, 1, generator=g)
    feat = torch.randn(N, F, generator=g) * 0.5
    dy = torch.randn(N, 7, generator=g)
    mlps = deform_torch.make_train_mlps(F, seed=18)
    rot = mlps["rot"]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<2if", N, F, t))
        f.write(tpos.numpy().astype("<f4").tobytes())
        f.write(feat.numpy().astype("<f4").tobytes())
        for layer in (rot[0], rot[2], rot[4]):
            f.write(layer.weight.detach().contiguous().numpy().astype("<f4").tobytes())
            f.write(layer.bias.detach().contiguous().numpy().astype("<f4").tobytes())
        f.write(dy.numpy().astype("<f4").tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = open(fout, "rb").read()
    off = 0

    def take(*shape):
        nonlocal off
        a = np.frombuffer(raw, dtype="<f4", count=int(np.prod(shape)), offset=off).reshape(shape)
        off += a.nbytes
        return a

    c = dict(out=take(N, 7), dfeat=take(N, F), W1=take(128, F + 9), b1=take(128), W2=take(128, 128), b2=take(128), W3=take(7, 128))
    assert off == len(raw)
    # the same evaluation through the Python host layer
    dmlps = {k: m.to(dev) for k, m in mlps.items()}
    images = deformation.TrainImages(dmlps["motion"], dmlps["rot"], dmlps["shs"], dmlps["opacity"])
    feat_d = feat.to(dev).requires_grad_(True)
    (out,) = deformation._TrainMLPs.apply(feat_d, tpos.to(dev), t, ((1, False, True),), images, *images.params())
    (out * dy.to(dev)).sum().backward()
    layers = (dmlps["rot"][0], dmlps["rot"][2], dmlps["rot"][4])
    py = dict(out=out.detach(), dfeat=feat_d.grad, W1=layers[0].weight.grad, b1=layers[0].bias.grad, W2=layers[1].weight.grad,
              b2=layers[1].bias.grad, W3=layers[2].weight.grad)
    for k, v in c.items():
        assert np.array_equal(v, py[k].cpu().numpy()), k
    # and against float64 autograd of the same MLP: outputs and the feature gradient (rows away from a ReLU kink)
    rot64 = __import__("copy").deepcopy(mlps["rot"]).double().cpu()
    x = torch.cat([feat.double(), deform_torch.time_embedding((t - tpos).double())], dim=1).requires_grad_(True)
    ref = rot64(x)
    assert np.abs(c["out"] - ref.detach().numpy()).max() <= 1e-4 * float(ref.abs().max())

"""Seeded synthetic scenes/cameras for the BASELINE.json configs (no dataset is on disk).

Camera conventions restate the reference's camera builders:
  world_view_transform = W2C^T, full_proj = view @ proj, campos = inverse(view)[3,:3]
      (scene/cameras.py:90-101 of the reference)
  projection with (zfar+znear)/(zfar-znear) in [2,2], znear=0.01, zfar=100
      (utils/graphics_utils.py:53-74, scene/cameras.py:84-85)
Everything is generated on the CPU with torch.Generator so that inputs are identical on
every machine; callers move the tensors to the GPU.
"""
import math
from typing import NamedTuple

import torch


class Camera(NamedTuple):
    width: int
    height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # [4,4] row-vector convention (W2C^T)
    projmatrix: torch.Tensor   # [4,4] full projection (view @ proj)
    campos: torch.Tensor       # [3]


class Scene(NamedTuple):
    means3D: torch.Tensor      # [P,3]
    scales: torch.Tensor       # [P,3]  (already exp-activated)
    rotations: torch.Tensor    # [P,4]  (normalised, (r,x,y,z))
    opacities: torch.Tensor    # [P,1]  (already sigmoid-activated)
    shs: torch.Tensor          # [P,16,3]
    sh_degree: int


def projection_matrix(znear, zfar, tanfovx, tanfovy):
    P = torch.zeros(4, 4, dtype=torch.float64)
    P[0, 0] = 1.0 / tanfovx
    P[1, 1] = 1.0 / tanfovy
    P[3, 2] = 1.0
    P[2, 2] = (zfar + znear) / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(width, height, fx, fy=None, R=None, t=None, znear=0.01, zfar=100.0):
    """Pinhole camera; R (3x3, world->camera rotation) and t (3,) give x_cam = R x_world + t."""
    fy = fx if fy is None else fy
    tanfovx = width / (2.0 * fx)
    tanfovy = height / (2.0 * fy)
    W2C = torch.eye(4, dtype=torch.float64)
    if R is not None:
        W2C[:3, :3] = torch.as_tensor(R, dtype=torch.float64)
    if t is not None:
        W2C[:3, 3] = torch.as_tensor(t, dtype=torch.float64)
    view = W2C.t().contiguous()
    proj = projection_matrix(znear, zfar, tanfovx, tanfovy).t().contiguous()
    full = view @ proj
    campos = torch.linalg.inv(view)[3, :3]
    return Camera(width, height, float(tanfovx), float(tanfovy), view.float().contiguous(),
                  full.float().contiguous(), campos.float().contiguous())


def yaw_camera(width, height, fx, yaw, pivot=(0.0, 0.0, 10.0)):
    """Camera rotated by `yaw` (rad) about the vertical axis through `pivot` (config 5 arc)."""
    c, s = math.cos(yaw), math.sin(yaw)
    Rw = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)  # camera-to-world rotation
    pv = torch.tensor(pivot, dtype=torch.float64)
    centre = pv - Rw @ pv          # camera centre so that the pivot stays on the optical axis at the same distance
    R = Rw.t()
    t = -R @ centre
    return make_camera(width, height, fx, R=R, t=t)


def _common_attributes(P, gen, log_scale_mean, log_scale_std, sh_degree):
    scales = torch.exp(torch.randn(P, 3, generator=gen) * log_scale_std + log_scale_mean)
    q = torch.randn(P, 4, generator=gen)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(torch.randn(P, 1, generator=gen) * 2.0)
    C0 = 0.28209479177387814
    shs = torch.zeros(P, 16, 3)
    shs[:, 0, :] = (torch.rand(P, 3, generator=gen) - 0.5) / C0
    shs[:, 1:, :] = torch.randn(P, 15, 3, generator=gen) * 0.05
    return scales.float(), rotations.float(), opacities.float(), shs.float().contiguous()


def config1_scene(P=10_000, seed=0):
    """BASELINE.json configs[0]: 10k random Gaussians, one pinhole camera @400x400 (D-NeRF-like)."""
    gen = torch.Generator().manual_seed(seed)
    means = (torch.rand(P, 3, generator=gen) * 2.0 - 1.0) * 1.3
    scales, rotations, opacities, shs = _common_attributes(P, gen, -3.5, 0.6, 3)
    cam = make_camera(400, 400, 555.56, t=(0.0, 0.0, 4.0))
    return Scene(means.float(), scales, rotations, opacities, shs, 3), cam


def config2_scene(P=300_000, seed=0, width=1352, height=1014, fx=729.0, log_scale_mean=-3.0, log_scale_std=0.9):
    """BASELINE.json configs[1] (headline): ~300k Gaussians @1352x1014, N3D-like, identity pose.
    95% of the means uniform in the view frustum (z in [4.5, 40]), 5% behind/near the camera.
    log_scale_mean was tuned once (R/P ~ 10-15) and is frozen at -3.0:
    num_rendered R = 3,927,052 (R/P = 13.1), 253,700 visible, mean n_contrib 359/pixel."""
    gen = torch.Generator().manual_seed(seed)
    tanx, tany = width / (2.0 * fx), height / (2.0 * fx)
    n_front = int(P * 0.95)
    z = torch.rand(n_front, generator=gen) * (40.0 - 4.5) + 4.5
    x = (torch.rand(n_front, generator=gen) * 2.0 - 1.0) * 1.1 * tanx * z
    y = (torch.rand(n_front, generator=gen) * 2.0 - 1.0) * 1.1 * tany * z
    front = torch.stack([x, y, z], dim=1)
    n_back = P - n_front
    zb = torch.rand(n_back, generator=gen) * 5.2 - 5.0
    xb = (torch.rand(n_back, generator=gen) * 2.0 - 1.0) * 3.0
    yb = (torch.rand(n_back, generator=gen) * 2.0 - 1.0) * 3.0
    back = torch.stack([xb, yb, zb], dim=1)
    means = torch.cat([front, back], dim=0)
    perm = torch.randperm(P, generator=gen)
    means = means[perm].contiguous()
    scales, rotations, opacities, shs = _common_attributes(P, gen, log_scale_mean, log_scale_std, 3)
    cam = make_camera(width, height, fx)
    return Scene(means.float(), scales, rotations, opacities, shs, 3), cam


def small_scene(P=512, seed=0, width=96, height=80, fx=90.0, log_scale_mean=-2.2):
    """Tiny scene for oracle-sized parity tests (ragged image size: 96x80 -> 6x5 tiles)."""
    gen = torch.Generator().manual_seed(seed)
    z = torch.rand(P, generator=gen) * 6.0 + 0.05          # some inside the near cull (z <= 0.2)
    tanx, tany = width / (2.0 * fx), height / (2.0 * fx)
    x = (torch.rand(P, generator=gen) * 2.0 - 1.0) * 1.3 * tanx * z
    y = (torch.rand(P, generator=gen) * 2.0 - 1.0) * 1.3 * tany * z
    means = torch.stack([x, y, z], dim=1)
    scales, rotations, opacities, shs = _common_attributes(P, gen, log_scale_mean, 0.7, 3)
    cam = make_camera(width, height, fx)
    return Scene(means.float(), scales, rotations, opacities, shs, 3), cam


def cotangent(height, width, seed=1):
    """dL/dcolor used by the fwd+bwd benchmark and parity tests: N(0,1)/(3HW)."""
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(3, height, width, generator=gen) / (3.0 * height * width)).float()


def temporal_frame(scene, t, seed=7):
    """Config 3: per-Gaussian temporal survival + small sinusoidal motion; returns the Scene of the
    Gaussians alive at time t (survival > 0.001, like saro_gaussian.py:878-881 of the reference)."""
    gen = torch.Generator().manual_seed(seed)
    P = scene.means3D.shape[0]
    centre = torch.rand(P, generator=gen)
    life = torch.rand(P, generator=gen) * 0.9 + 0.1
    phase = torch.rand(P, generator=gen) * 2 * math.pi
    direction = torch.randn(P, 3, generator=gen)
    direction = direction / direction.norm(dim=1, keepdim=True)
    survival = torch.exp(-4.0 * ((t - centre) / life) ** 2)
    keep = survival > 0.001
    means = scene.means3D + 0.05 * torch.sin(2 * math.pi * t + phase)[:, None] * direction
    return Scene(means[keep].contiguous(), scene.scales[keep].contiguous(), scene.rotations[keep].contiguous(),
                 (scene.opacities * survival[:, None])[keep].contiguous(), scene.shs[keep].contiguous(),
                 scene.sh_degree)


class DeviceSequence:
    """BASELINE.json configs[2] at its stated size (300 frames): `temporal_frame` evaluated with torch ops ON THE
    DEVICE, so that a 300-frame sequence of a 300 k cloud does not have to be built and uploaded from the host.
    The per-Gaussian temporal parameters are the CPU-generated ones of `temporal_frame` (same seed); the elementwise
    arithmetic runs on the GPU, so fixtures made from it (tests/golden/config3_seq300.npz) are tied to the platform
    that made them (B200, this image) — every frame's alive count is stored as a guard."""

    def __init__(self, scene, device, seed=7):
        gen = torch.Generator().manual_seed(seed)
        P = scene.means3D.shape[0]
        self.centre = torch.rand(P, generator=gen).to(device)
        self.life = (torch.rand(P, generator=gen) * 0.9 + 0.1).to(device)
        self.phase = (torch.rand(P, generator=gen) * 2 * math.pi).to(device)
        direction = torch.randn(P, 3, generator=gen)
        self.direction = (direction / direction.norm(dim=1, keepdim=True)).to(device)
        self.scene = Scene(*[x.to(device) if torch.is_tensor(x) else x for x in scene])

    def frame(self, t):
        sc = self.scene
        survival = torch.exp(-4.0 * ((t - self.centre) / self.life) ** 2)
        keep = survival > 0.001
        means = sc.means3D + 0.05 * torch.sin(2 * math.pi * t + self.phase)[:, None] * self.direction
        return Scene(means[keep].contiguous(), sc.scales[keep].contiguous(), sc.rotations[keep].contiguous(),
                     (sc.opacities * survival[:, None])[keep].contiguous(), sc.shs[keep].contiguous(), sc.sh_degree)


def dynamic_model(scene, feat_dim=32, seed=0, lifespan=(0.15, 1.2), residual_gain=0.05):
    """A stand-in for the reference's dynamic GaussianModel after get_deformfeature() (scene/saro_gaussian.py:863-869):
    the attributes get_deformation_eval reads (:871-921) — raw (pre-activation) parameters derived from `scene`, a
    temporal centre and lifespan per Gaussian, cached plane features and the three 3-layer MLPs (:104,:108,:110).
    CPU tensors / modules; move with `.to(device)` per attribute (see tests and bench.py)."""
    import types
    gen = torch.Generator().manual_seed(seed + 1234)
    P = scene.means3D.shape[0]

    def mlp(out_dim):
        m = torch.nn.Sequential(torch.nn.Linear(feat_dim + 9, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128),
                                torch.nn.ReLU(), torch.nn.Linear(128, out_dim))
        with torch.no_grad():
            for i, layer in enumerate(l for l in m if isinstance(l, torch.nn.Linear)):
                torch.nn.init.xavier_uniform_(layer.weight, generator=gen)
                layer.bias.uniform_(-0.05, 0.05, generator=gen)
                if i == 2:
                    layer.weight.mul_(residual_gain)
                    layer.bias.mul_(residual_gain)
        return m

    op = scene.opacities.clamp(1e-4, 1 - 1e-4)
    pc = types.SimpleNamespace(
        args=types.SimpleNamespace(dx=True, drot=True, dopacity=True, dsh=True, sigmoid_tcenter=False),
        _xyz=scene.means3D.clone(), _rotation=scene.rotations.clone(), _scaling=torch.log(scene.scales),
        _opacity=torch.log(op / (1 - op)).reshape(P, 1),
        _features_dc=scene.shs[:, :1, :].contiguous(), _features_rest=scene.shs[:, 1:, :].contiguous(),
        get_temporalpos=torch.rand(P, 1, generator=gen),
        _lifespan=torch.rand(P, 1, generator=gen) * (lifespan[1] - lifespan[0]) + lifespan[0],
        hexplane_feature=torch.randn(P, feat_dim, generator=gen) * 0.3,
        motion_mlp=mlp(3), rot_mlp=mlp(7), shs_mlp=mlp(48))
    return pc


def model_to(pc, device):
    """Move every tensor / module of a dynamic_model() to `device` (returns a new namespace)."""
    import copy
    import types
    out = types.SimpleNamespace(args=pc.args)
    for k, v in vars(pc).items():
        if k != "args":
            # nn.Module.to() moves in place: copy first so that `pc` keeps its own (CPU) parameters
            setattr(out, k, copy.deepcopy(v).to(device) if isinstance(v, torch.nn.Module) else v.to(device))
    return out

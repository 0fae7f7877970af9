"""TEST / BENCH INFRASTRUCTURE — the deformation hand-off as the PyTorch ops SaRO-GS runs (timing baseline on the GPU
box, where /root/reference does not exist).  Restates scene/saro_gaussian.py:871-921 (get_deformation_eval), :757-759
(get_survival_state) and :939-969 (Embedder, 4 frequencies) op for op: boolean-mask gathers, three nn.Sequential MLPs,
elementwise epilogues, cat.  Checked against the numpy oracle in tests/test_deform_gpu.py.  Never imported by the
product path."""
import torch


def time_embedding(x):
    outs = [x]
    for freq in 2.0 ** torch.linspace(0.0, 3.0, steps=4):
        outs.append(torch.sin(x * freq))
        outs.append(torch.cos(x * freq))
    return torch.cat(outs, -1)


def torch_get_deformation_eval(pc, timestamp):
    distance = timestamp - pc.get_temporalpos
    state = torch.exp(-4 * ((distance / pc._lifespan) ** 2))
    time_embbed = time_embedding(distance)
    deform_feature = torch.cat((pc.hexplane_feature, time_embbed.detach()), dim=1)
    select_mask = (state > 0.001).squeeze()
    deform_feature = deform_feature[select_mask]
    state = state[select_mask]
    motion = pc._xyz[select_mask] + pc.motion_mlp(deform_feature)
    rot_residual = pc.rot_mlp(deform_feature)
    rot = torch.nn.functional.normalize(pc._rotation[select_mask] + rot_residual[:, :4])
    scale = torch.exp(pc._scaling[select_mask] + rot_residual[:, 4:])
    opacity = torch.sigmoid(pc._opacity[select_mask]) * state
    shs_residual = pc.shs_mlp(deform_feature).reshape(-1, 16, 3)
    shs = torch.cat((pc._features_dc[select_mask], pc._features_rest[select_mask]), dim=1) + shs_residual
    return motion, rot, scale, opacity, shs


def torch_get_deformation(pc, timestamp):
    """Restates scene/saro_gaussian.py:779-847 (get_deformation, the training-time path) op for op, with
    dx = drot = dopacity = dsh = True; same side effects on `pc` (_lifespan, scale_residual, shs_residual,
    motion_residual, real_xyz).  Pinned (outputs and autograd gradients, float64) on tests/golden/deformtrain_*.npz,
    which come from the reference's own method source (tests/test_deform_train_oracle.py)."""
    hexplane_feature = pc.hexplane(pc._xyz.detach(), pc.get_temporalpos.detach(), pc.get_scaling.detach())      # :780
    lifespan = 1 - pc.opacity_mlp(hexplane_feature)                                                             # :782
    min_scale = pc.args.min_interval / pc.duration
    lifespan = (1 - min_scale) * lifespan + min_scale
    pc._lifespan = lifespan
    distance = timestamp - pc.get_temporalpos                                                                   # :788
    trbfoutput = torch.exp(-4 * ((distance / lifespan) ** 2))                                                   # :789, :757-759
    deform_feature = torch.cat((hexplane_feature, time_embedding(distance).detach()), dim=1)                    # :791-792
    base_deform_feature = torch.cat((hexplane_feature, time_embedding(torch.zeros_like(distance)).detach()), dim=1)
    if pc.args.scale_reg:
        pc.scale_residual = pc.rot_mlp(base_deform_feature)[:, 4:]                                              # :796-797
    if pc.args.shs_reg:
        pc.shs_residual = pc.shs_mlp(base_deform_feature).reshape(-1, 16, 3)
    if pc.args.motion_reg:
        pc.motion_residual = pc.motion_mlp(base_deform_feature)
    with torch.no_grad():
        pc.real_xyz = pc._xyz + pc.motion_mlp(base_deform_feature)                                              # :803-804
    motion = pc._xyz + pc.motion_mlp(deform_feature)                                                            # :807-809
    rot_residual = pc.rot_mlp(deform_feature)
    rot = torch.nn.functional.normalize(pc._rotation + rot_residual[:, :4])                                     # :813-817
    scale = torch.exp(pc._scaling + rot_residual[:, 4:])                                                        # :819-821
    opacity = torch.sigmoid(pc._opacity) * trbfoutput                                                           # :830-831
    shs = torch.cat((pc._features_dc, pc._features_rest), dim=1) + pc.shs_mlp(deform_feature).reshape(-1, 16, 3)
    return motion, rot, scale, opacity, shs


class TrainModelStandIn:
    """The attributes of GaussianModel that get_deformation touches (scene/saro_gaussian.py:39-47,:126-147,:757-759),
    built from explicit tensors — test / bench scaffolding for both the native path and the restatement above."""

    def __init__(self, tensors, mlps, flags, min_interval, duration, hexplane=None):
        import types
        scale_reg, shs_reg, motion_reg = (bool(f) for f in flags)
        self.args = types.SimpleNamespace(dx=True, drot=True, dopacity=True, dsh=True, sigmoid_tcenter=False, scale_reg=scale_reg,
                                          shs_reg=shs_reg, motion_reg=motion_reg, min_interval=float(min_interval))
        self.duration = float(duration)
        self._xyz, self._rotation, self._scaling, self._opacity = (tensors[k] for k in ("xyz", "rotation", "scaling", "opacity"))
        self._features_dc, self._features_rest = tensors["features_dc"], tensors["features_rest"]
        self._temporal_pos = tensors["temporal_pos"]
        self.motion_mlp, self.rot_mlp, self.shs_mlp, self.opacity_mlp = (mlps[k] for k in ("motion", "rot", "shs", "opacity"))
        self.rotation_activation = torch.nn.functional.normalize
        self.scaling_activation = torch.exp
        self.opacity_activation = torch.sigmoid
        self.hexplane = hexplane if hexplane is not None else (lambda xyz, t, s: tensors["hexplane_feature"])
        self.scale_residual = self.shs_residual = self.motion_residual = None

    @property
    def get_temporalpos(self):
        return self._temporal_pos

    @property
    def get_scaling(self):
        return self.scaling_activation(self._scaling)

    def get_survival_state(self, trbfdistance):
        return torch.exp(-4 * (trbfdistance ** 2))


def make_train_mlps(feat_dim, arrays=None, dtype=torch.float32, device="cpu", seed=0):
    """motion / rot / shs / opacity MLPs with the reference's shapes (scene/saro_gaussian.py:102-108); parameters from a
    fixture's `mlp_<name>_{W,b}{1,2,3}` arrays or seeded random."""
    from torch import nn
    g = torch.Generator().manual_seed(seed)
    spec = dict(motion=(feat_dim + 9, 128, 3, False), rot=(feat_dim + 9, 128, 7, False), shs=(feat_dim + 9, 128, 48, False),
                opacity=(feat_dim, 64, 1, True))
    out = {}
    for name, (i, h2, o, sig) in spec.items():
        layers = [nn.Linear(i, 128), nn.ReLU(), nn.Linear(128, h2), nn.ReLU(), nn.Linear(h2, o)] + ([nn.Sigmoid()] if sig else [])
        m = nn.Sequential(*layers)
        with torch.no_grad():
            for q, l in enumerate([l for l in m if isinstance(l, nn.Linear)]):
                if arrays is not None:
                    l.weight.copy_(torch.from_numpy(arrays[f"mlp_{name}_W{q + 1}"]))
                    l.bias.copy_(torch.from_numpy(arrays[f"mlp_{name}_b{q + 1}"]))
                else:
                    nn.init.xavier_uniform_(l.weight, gain=1.0, generator=g)
                    l.bias.uniform_(-0.1, 0.1, generator=g)
        out[name] = m.to(device=device, dtype=dtype)
    return out


def train_objective(pc, outs, weights, lambdas):
    """The scalar the deformtrain fixtures differentiate (tests/golden/make_golden_deform_train.py)."""
    n = outs[0].shape[0]
    loss = sum((w * o).sum() for w, o in zip(weights, outs))
    if pc.args.scale_reg:
        loss = loss + lambdas[0] * torch.linalg.vector_norm(pc.scale_residual, ord=2)
    if pc.args.shs_reg:
        loss = loss + lambdas[1] * torch.linalg.matrix_norm(pc.shs_residual.reshape(n, -1))
    if pc.args.motion_reg:
        loss = loss + lambdas[2] * torch.linalg.matrix_norm(pc.motion_residual)
    return loss

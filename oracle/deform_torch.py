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

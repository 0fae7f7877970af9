"""TEST INFRASTRUCTURE — numpy restatement of the reference's per-frame deformation -> rasterizer hand-off.

Follows scene/saro_gaussian.py of the reference:
    get_survival_state                :757-759   state = exp(-4 d^2)
    get_deformation_eval              :871-921   selection (state > 0.001), three MLPs, residual + activation epilogues
    get_embedder / Embedder           :922-969   [x, sin(f x), cos(f x) for f in 2**linspace(0, L-1, L)], L = 4
    motion_mlp / rot_mlp / shs_mlp    :104,:108,:110   Linear-ReLU-Linear-ReLU-Linear
    activations                       :39,:44,:47      exp, sigmoid, F.normalize (eps 1e-12)

PARITY PIN: the reference has no tests for this path; tests/golden/deform_*.npz hold outputs of the reference's OWN
method source (loaded at generation time from /root/reference/scene/saro_gaussian.py and executed on CPU in float32
and float64 by tests/golden/make_golden_deform.py).  Only tests/ and bench.py's baseline legs may import this module.
"""
import numpy as np

MLP_NAMES = ("motion", "rot", "shs")


def time_embedding(d, num_freqs=4):
    """d: [N,1] -> [N, 1 + 2*num_freqs]   (saro_gaussian.py:939-969)"""
    dt = d.dtype
    cols = [d]
    for f in (2.0 ** np.linspace(0.0, num_freqs - 1, num_freqs)).astype(np.float32):
        x = d * dt.type(f)
        cols += [np.sin(x), np.cos(x)]
    return np.concatenate(cols, axis=-1)


def mlp(x, params):
    """params = (W1, b1, W2, b2, W3, b3), nn.Linear convention y = x W^T + b   (saro_gaussian.py:104-110)"""
    W1, b1, W2, b2, W3, b3 = params
    h = np.maximum(x @ W1.T + b1, 0)
    h = np.maximum(h @ W2.T + b2, 0)
    return h @ W3.T + b3


def survival_state(timestamp, temporal_pos, lifespan):
    """saro_gaussian.py:872-873 with :757-759"""
    dt = temporal_pos.dtype
    distance = dt.type(timestamp) - temporal_pos
    q = distance / lifespan
    return np.exp(dt.type(-4) * (q * q)), distance


def deformation_eval(timestamp, xyz, rotation, scaling, opacity, features_dc, features_rest, temporal_pos, lifespan,
                     hexplane_feature, mlps, dtype=np.float64):
    """Returns dict(mask, state, means3D, rotations, scales, opacity, shs) — saro_gaussian.py:871-921 with
    dx = drot = dopacity = dsh = True."""
    c = lambda a: np.asarray(a, dtype=dtype)
    xyz, rotation, scaling, opacity = c(xyz), c(rotation), c(scaling), c(opacity).reshape(-1, 1)
    features_dc, features_rest = c(features_dc).reshape(-1, 1, 3), c(features_rest).reshape(-1, 15, 3)
    temporal_pos, lifespan, hexplane_feature = c(temporal_pos).reshape(-1, 1), c(lifespan).reshape(-1, 1), c(hexplane_feature)
    mlps = {k: tuple(c(t) for t in v) for k, v in mlps.items()}

    state, distance = survival_state(timestamp, temporal_pos, lifespan)
    feature = np.concatenate([hexplane_feature, time_embedding(distance)], axis=1)            # :875-876
    mask = (state > dtype(0.001)).reshape(-1)                                                  # :878
    feature, st = feature[mask], state[mask]                                                   # :880-881

    means3D = xyz[mask] + mlp(feature, mlps["motion"])                                         # :883-885
    rr = mlp(feature, mlps["rot"])                                                             # :890
    rot = rotation[mask] + rr[:, :4]                                                           # :891
    rot = rot / np.maximum(np.sqrt((rot * rot).sum(axis=1, keepdims=True)), dtype(1e-12))      # :893 F.normalize
    scale = np.exp(scaling[mask] + rr[:, 4:])                                                  # :896-897
    opa = (1 / (1 + np.exp(-opacity[mask]))) * st                                              # :904-905
    shs = np.concatenate([features_dc[mask], features_rest[mask]], axis=1) + mlp(feature, mlps["shs"]).reshape(-1, 16, 3)
    return dict(mask=mask, state=state.reshape(-1), means3D=means3D, rotations=rot, scales=scale, opacity=opa, shs=shs)

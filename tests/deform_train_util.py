"""Shared by the CPU pin test and the GPU parity test of the training-time deformation path."""
import glob
import os

import numpy as np
import torch

from oracle import deform_torch

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "deformtrain_*.npz")))
OUTS = ("means3D", "rotations", "scales", "opacity", "shs")
LEAVES = ("xyz", "rotation", "scaling", "opacity", "features_dc", "features_rest", "temporal_pos", "hexplane_feature")
MLPS = ("motion", "rot", "shs", "opacity")


def build(z, device, dtype):
    """(model stand-in, leaf tensors, mlps, objective weights) from a fixture."""
    feat_dim = z["in_hexplane_feature"].shape[1]
    leaves = {k: torch.from_numpy(z[f"in_{k}"]).to(device=device, dtype=dtype).requires_grad_(True) for k in LEAVES}
    mlps = deform_torch.make_train_mlps(feat_dim, arrays=z, dtype=dtype, device=device)
    pc = deform_torch.TrainModelStandIn(leaves, mlps, z["flags"], float(z["min_interval"]), float(z["duration"]))
    weights = [torch.from_numpy(z[f"w_{k}"]).to(device=device, dtype=dtype) for k in OUTS]
    return pc, leaves, mlps, weights


def gradients(leaves, mlps):
    out = {f"grad_{k}": v.grad for k, v in leaves.items()}
    for name in MLPS:
        for q, l in enumerate([l for l in mlps[name] if isinstance(l, torch.nn.Linear)]):
            out[f"grad_{name}_W{q + 1}"] = l.weight.grad
            out[f"grad_{name}_b{q + 1}"] = l.bias.grad
    return out


def maxrel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

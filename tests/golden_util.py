"""Helpers shared by the golden-fixture tests (fixtures = outputs of the compiled, unmodified
reference on a B200; see tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["small_sh3", "small_sh1_black", "small_sh0_white", "small_precomp_color", "small_precomp_cov",
               "small_big_splats"]
INPUT_KEYS = ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def inputs_of(d):
    return {k: torch.from_numpy(d["in_" + k]) for k in INPUT_KEYS if "in_" + k in d.files}


def normrel(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def maxrel(a, b):
    """max |a-b| relative to max |b|: the '1e-4 relative' bar of BASELINE.json north_star."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

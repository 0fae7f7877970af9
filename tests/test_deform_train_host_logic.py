"""CPU: host-side logic of the training-time deformation path that needs no GPU — the ctypes mirrors of the C structs
(layout checked against the header with gcc), the shape / configuration checks, and the loud failures (no CPU path)."""
import ctypes
import os
import shutil
import subprocess
import types

import pytest
import torch

from oracle import deform_torch
from saro_gs_b200 import _lib
from saro_gs_b200 import deformation as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of sgs_mlp_job_t and sgs_wgrad_task_t as gcc sees them in include/saro_gs_b200.h == the ctypes
    Structures the Python host layer passes (a silent mismatch would shift every pointer after the first)."""
    fields = {"sgs_mlp_job_t": ["packed", "in", "out", "save_a", "save_b", "save_in", "mask_a", "mask_b", "n_io", "zero_time"],
              "sgs_wgrad_task_t": ["A", "B", "groups_b", "dW", "ldw", "rows", "cols", "transposed", "db", "accumulate"]}
    src = ['#include <stddef.h>', '#include <stdio.h>', '#include "saro_gs_b200.h"', "int main(void) {"]
    for name, fs in fields.items():
        src.append(f'printf("%zu", sizeof({name}));')
        src += [f'printf(" %zu", offsetof({name}, {f}));' for f in fs]
        src.append('printf("\\n");')
    src.append("return 0; }")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    cc = shutil.which("gcc") or "/usr/bin/gcc"
    subprocess.run([cc, "-std=c99", str(c), "-I", os.path.join(ROOT, "include"), "-o", exe], check=True, capture_output=True)
    lines = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line, (cls, names) in zip(lines, ((_lib.MLPJob, ["packed", "inp", "out", "save_a", "save_b", "save_in", "mask_a", "mask_b", "n_io", "zero_time"]),
                                          (_lib.WgradTask, fields["sgs_wgrad_task_t"]))):
        nums = [int(x) for x in line.split()]
        assert ctypes.sizeof(cls) == nums[0], cls
        assert [getattr(cls, n).offset for n in names] == nums[1:], cls


def _model(feat_dim=32, n=10, **flags):
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g)
    t = dict(xyz=rn(n, 3), rotation=rn(n, 4), scaling=rn(n, 3), opacity=rn(n, 1), features_dc=rn(n, 1, 3), features_rest=rn(n, 15, 3),
             temporal_pos=torch.rand(n, 1, generator=g), hexplane_feature=rn(n, feat_dim))
    mlps = deform_torch.make_train_mlps(feat_dim, seed=1)
    return deform_torch.TrainModelStandIn(t, mlps, (1, 0, 0), 6.0, 300.0), mlps


def test_cpu_models_fail_loudly():
    pc, _ = _model()
    for fn in (lambda: D.get_deformation(pc, 0.3), lambda: D.get_deformfeature(pc), lambda: D.get_intergral(pc)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            fn()


def test_unsupported_switches_raise_before_any_work():
    pc, _ = _model()
    pc._xyz = types.SimpleNamespace(is_cuda=True)           # get past the device check without a GPU
    for switch in ("dx", "drot", "dopacity", "dsh"):
        setattr(pc.args, switch, False)
        with pytest.raises(D.UnsupportedDeformationConfig):
            D.get_deformation(pc, 0.3)
        setattr(pc.args, switch, True)


@pytest.mark.parametrize("mutate,what", [
    (lambda m: m["motion"].__setitem__(4, torch.nn.Linear(128, 9)), "n_out 9 is neither <= 8 nor 48"),
    (lambda m: m["rot"].__setitem__(0, torch.nn.Linear(41, 64)), "first hidden width must be 128"),
    (lambda m: m["opacity"].__setitem__(0, torch.nn.Linear(41, 128)), "opacity_mlp takes the plane feature only"),
    (lambda m: m["shs"].__setitem__(2, torch.nn.Linear(64, 128)), "second layer must take 128 inputs"),
])
def test_mlp_shape_checks(mutate, what, monkeypatch):
    _, mlps = _model()
    mutate(mlps)
    monkeypatch.setattr(D._lib, "load", lambda: types.SimpleNamespace(sgs_deform_image_bytes=lambda: 115904))
    with pytest.raises(D.UnsupportedDeformationConfig):
        D.TrainImages(mlps["motion"], mlps["rot"], mlps["shs"], mlps["opacity"])


def test_feature_width_check(monkeypatch):
    monkeypatch.setattr(D._lib, "load", lambda: types.SimpleNamespace(sgs_deform_image_bytes=lambda: 115904))
    mlps = deform_torch.make_train_mlps(12)
    with pytest.raises(D.UnsupportedDeformationConfig, match="8, 16, 24 and 32"):
        D.TrainImages(mlps["motion"], mlps["rot"], mlps["shs"], mlps["opacity"])
    ok = deform_torch.make_train_mlps(24)
    images = D.TrainImages(ok["motion"], ok["rot"], ok["shs"], ok["opacity"])
    assert images.feat_dim == 24 and images.shapes == [(33, 128, 3), (33, 128, 7), (33, 128, 48), (24, 64, 1)]
    assert len(images.params()) == 24

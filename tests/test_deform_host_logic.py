"""CPU: host-side logic of the deformation hand-off and the densification statistics that needs no GPU — the
validated-input cache of the drop-in method, the weight-version key that triggers re-packing, configuration checks
and the loud failures on CPU tensors (there is no CPU path)."""
import types

import pytest
import torch

from saro_gs_b200 import deformation as D
from saro_gs_b200 import synthetic
from saro_gs_b200.densify import BatchDensifyStats


@pytest.fixture
def stubbed(monkeypatch):
    """Replace the CUDA-only pieces: validation accepts CPU tensors, packing and the launch are recorded."""
    calls = []

    class FakePacked:
        feat_dim = 32

        def __init__(self, *mlps):
            self.mlps = mlps

        def refresh(self):
            pass

    def fake_run(timestamp, n, inputs, packed, workspace):
        calls.append((timestamp, n, tuple(t.data_ptr() for t in inputs)))
        return tuple(inputs[:5])

    monkeypatch.setattr(D, "_check", lambda t, name, tail, n=None: t.detach().contiguous())
    monkeypatch.setattr(D, "PackedMLPs", FakePacked)
    monkeypatch.setattr(D, "_run", fake_run)
    monkeypatch.setattr(D._lib, "load", lambda: types.SimpleNamespace(sgs_deform_workspace_bytes=lambda n: 1024))
    return calls


def model(P=100):
    scene, _ = synthetic.small_scene(P=P)
    return synthetic.dynamic_model(scene)


def test_validated_inputs_are_cached_until_the_model_changes(stubbed):
    pc = model()
    D.get_deformation_eval(pc, 0.3)
    first = pc._sgs_deform_cache["inputs"]
    D.get_deformation_eval(pc, torch.tensor(0.4))                    # tensor timestamps are accepted
    assert pc._sgs_deform_cache["inputs"] is first
    assert [c[0] for c in stubbed] == [0.3, pytest.approx(0.4)]
    pc._xyz = pc._xyz.clone()                                        # attribute replaced (densification / load)
    D.get_deformation_eval(pc, 0.5)
    second = pc._sgs_deform_cache["inputs"]
    assert second is not first and second[0].data_ptr() == pc._xyz.data_ptr()
    pc._opacity.data = pc._opacity.data.clone()                      # storage swapped under the same tensor object
    D.get_deformation_eval(pc, 0.6)
    assert pc._sgs_deform_cache["inputs"] is not second
    assert stubbed[-1][2][3] == pc._opacity.data_ptr()


def test_non_contiguous_inputs_are_revalidated_every_call(stubbed):
    pc = model()
    pc.hexplane_feature = pc.hexplane_feature.t().contiguous().t()   # a view: validation makes a contiguous copy
    D.get_deformation_eval(pc, 0.1)
    a = pc._sgs_deform_cache["inputs"]
    D.get_deformation_eval(pc, 0.2)
    assert pc._sgs_deform_cache["inputs"] is not a                   # a stale copy is never re-used


def test_new_mlp_modules_get_a_new_packed_image(stubbed):
    pc = model()
    D.get_deformation_eval(pc, 0.3)
    packed = pc._sgs_deform_cache["packed"]
    D.get_deformation_eval(pc, 0.3)
    assert pc._sgs_deform_cache["packed"] is packed
    pc.shs_mlp = model().shs_mlp
    D.get_deformation_eval(pc, 0.3)
    assert pc._sgs_deform_cache["packed"] is not packed


def test_unsupported_switches_raise(stubbed):
    pc = model()
    for off in ("dx", "drot", "dopacity", "dsh"):
        setattr(pc.args, off, False)
        with pytest.raises(D.UnsupportedDeformationConfig):
            D.get_deformation_eval(pc, 0.3)
        setattr(pc.args, off, True)


def test_weight_version_key_sees_in_place_updates_and_replaced_parameters():
    pc = model(10)
    mlps = (pc.motion_mlp, pc.rot_mlp, pc.shs_mlp)
    p = object.__new__(D.PackedMLPs)                                  # key logic only; packing needs the GPU
    p._layers = [[l for l in m if hasattr(l, "weight")] for m in mlps]
    p._sources = [D._linears(m) for m in mlps]
    k1 = p._current_versions()
    assert len(k1) == 18
    with torch.no_grad():
        pc.rot_mlp[2].bias.add_(1.0)                                  # optimizer step / load_state_dict: version bump
    k2 = p._current_versions()
    assert sum(a != b for a, b in zip(k1, k2)) == 1
    pc.shs_mlp[4].weight = torch.nn.Parameter(pc.shs_mlp[4].weight.detach().clone())     # .to() / parameter swap
    k3 = p._current_versions()
    assert k3 != k2 and p._sources[2][4] is pc.shs_mlp[4].weight
    fixed = tuple(D._linears(pc.motion_mlp))                           # plain tuples of tensors are taken as they are
    q = object.__new__(D.PackedMLPs)
    q._layers, q._sources = [None, None, None], [fixed, fixed, fixed]
    assert len(q._current_versions()) == 18


def test_mlp_shape_and_feature_width_checks():
    pc = model(10)
    with pytest.raises(D.UnsupportedDeformationConfig):
        D._linears(torch.nn.Sequential(torch.nn.Linear(41, 128), torch.nn.ReLU(), torch.nn.Linear(128, 3)))
    with pytest.raises(D.UnsupportedDeformationConfig):
        D._linears((pc.motion_mlp[0].weight,))
    assert len(D._linears(pc.motion_mlp)) == 6


def test_cpu_tensors_fail_loudly():
    pc = model(10)
    with pytest.raises(RuntimeError, match="CUDA"):
        D._check(pc._xyz, "xyz", [(3,)], 10)
    with pytest.raises(RuntimeError, match="CUDA"):
        BatchDensifyStats(10, "cpu")
